// Stand-in for the Arcade Learning Environment's <ale_interface.hpp>, exactly as wide as environment/atari/atari.{h,cpp}
// need it (atari.h:6,47,60,85-87; atari.cpp:13-15,50-56,68-70,136-141): ALE and its ROMs are not in this image.
// The emulator behind it is the repo's deterministic synthetic frame source (minizero_b200/host/synth_atari.h), so that the
// compiled reference and the worker's Atari mode see the same frames, rewards and lives for the same seed and actions.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../../minizero_b200/host/synth_atari.h"
#include <string>
#include <vector>

namespace ale {

enum Action { PLAYER_A_NOOP = 0 };
typedef int reward_t;
typedef std::vector<Action> ActionVect;

inline std::string action_to_string(Action a)
{
    static const char* names[18] = {"NOOP", "FIRE", "UP", "RIGHT", "LEFT", "DOWN", "UPRIGHT", "UPLEFT", "DOWNRIGHT", "DOWNLEFT", "UPFIRE", "RIGHTFIRE", "LEFTFIRE",
                                    "DOWNFIRE", "UPRIGHTFIRE", "UPLEFTFIRE", "DOWNRIGHTFIRE", "DOWNLEFTFIRE"};
    const int i = static_cast<int>(a);
    return std::string("PLAYER_A_") + (i >= 0 && i < 18 ? names[i] : "NOOP");
}

struct Logger {
    enum mode { Info, Warning, Error };
    static void setMode(mode) {}
};

struct ALEScreen {
    int height() const { return mzhost::SynthAtari::kRes; }
    int width() const { return mzhost::SynthAtari::kRes; }
};

class ALEInterface {
public:
    void setInt(const std::string& key, int value)
    {
        if (key == "random_seed") { seed_ = value; }
    }
    void setFloat(const std::string&, float) {}
    void loadROM(const std::string&) {}
    void reset_game() { emu_.reset(seed_); }
    ActionVect getMinimalActionSet() const
    {
        ActionVect v;
        for (int a : mzhost::SynthAtari::minimalActionSet()) { v.push_back(static_cast<Action>(a)); }
        return v;
    }
    int lives() const { return emu_.lives(); }
    reward_t act(Action a) { return emu_.act(static_cast<int>(a)); }
    bool game_over(bool = true) const { return emu_.gameOver(); }
    void getScreenRGB(std::vector<unsigned char>& out) const
    {
        out.resize(3 * mzhost::SynthAtari::kRes * mzhost::SynthAtari::kRes);
        emu_.screenRGB(out.data());
    }
    const ALEScreen& getScreen() const { return screen_; }
    int getFrameNumber() const { return emu_.frameNumber(); }
    int getEpisodeFrameNumber() const { return emu_.frameNumber(); }

private:
    int seed_ = 0;
    mzhost::SynthAtari emu_;
    ALEScreen screen_;
};

} // namespace ale
