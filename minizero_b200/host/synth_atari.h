// Synthetic Atari frame source. The Arcade Learning Environment and its ROMs are not part of this image, so the worker's Atari
// mode (BASELINE configs[4]) and the compiled reference (built against oracle/shim/ale_interface.hpp, which wraps THIS class)
// both run on the same deterministic stand-in: a small arcade world whose screen, rewards, lives and episode end are pure
// functions of (seed, action sequence). It is a data source, not part of the search: ZeroActor never touches the emulator
// below the root (zero_actor.cpp:59-67,238). A real ALE build would replace this class behind the same five calls
// (reset / act / lives / gameOver / screenRGB), see INTEGRATION.md.
//
// World: a 96 x 96 screen (already the network's resolution, atari.h:24, so the reference's INTER_AREA resize is the
// identity), a player sprite moved by the 9 actions of ms_pacman's minimal action set (ALE action ids 0, 2..9), pellets on a
// grid that score when eaten, and two chasers that cost a life on contact. All arithmetic is integer.
#pragma once
#include <cstdint>
#include <vector>

namespace mzhost {

class SynthAtari {
public:
    static constexpr int kRes = 96;
    static constexpr int kGrid = 12;               // pellets sit on a kGrid x kGrid lattice of 8-pixel cells
    static constexpr int kCell = kRes / kGrid;

    void reset(int seed)
    {
        seed_ = seed;
        rng_ = 0x9E3779B97F4A7C15ull ^ (static_cast<uint64_t>(static_cast<uint32_t>(seed)) * 0xBF58476D1CE4E5B9ull);
        frame_ = 0, lives_ = 3, over_ = false;
        px_ = 5, py_ = 6;
        cx_[0] = 0, cy_[0] = 0, cx_[1] = kGrid - 1, cy_[1] = kGrid - 1;
        pellets_.assign(kGrid * kGrid, 1);
        pellets_[py_ * kGrid + px_] = 0;
        left_ = kGrid * kGrid - 1;
    }

    // one emulator frame with `action` held (the environment calls this kAtariFrameSkip times per move, atari.cpp:68); returns the reward
    int act(int action)
    {
        if (over_) { return 0; }
        ++frame_;
        int reward = 0;
        if (frame_ % 4 == 0) { // the world advances once per four frames
            static const int dx[18] = {0, 0, 0, 1, -1, 0, 1, -1, 1, -1, 0, 1, -1, 0, 1, -1, 1, -1};
            static const int dy[18] = {0, 0, -1, 0, 0, 1, -1, -1, 1, 1, -1, 0, 0, 1, -1, -1, 1, 1};
            const int a = (action >= 0 && action < 18 ? action : 0);
            px_ = clampi(px_ + dx[a]), py_ = clampi(py_ + dy[a]);
            if (pellets_[py_ * kGrid + px_]) {
                pellets_[py_ * kGrid + px_] = 0;
                --left_;
                reward += ((px_ + py_) % 5 == 0 ? 50 : 10);
            }
            const uint64_t r = next();
            for (int c = 0; c < 2; ++c) { // chasers: towards the player every second world step, else a random step
                const bool chase = ((frame_ / 4 + c) % 2 == 0);
                int sx = (chase ? sign(px_ - cx_[c]) : static_cast<int>((r >> (8 * c)) % 3) - 1);
                int sy = (chase ? sign(py_ - cy_[c]) : static_cast<int>((r >> (8 * c + 4)) % 3) - 1);
                cx_[c] = clampi(cx_[c] + sx), cy_[c] = clampi(cy_[c] + sy);
                if (cx_[c] == px_ && cy_[c] == py_) {
                    --lives_;
                    cx_[c] = (c == 0 ? 0 : kGrid - 1), cy_[c] = cx_[c];
                }
            }
            if (left_ == 0) { // board cleared: bonus and a fresh board
                reward += 200;
                pellets_.assign(kGrid * kGrid, 1);
                pellets_[py_ * kGrid + px_] = 0;
                left_ = kGrid * kGrid - 1;
            }
            if (lives_ <= 0) { lives_ = 0, over_ = true; }
        }
        return reward;
    }

    int lives() const { return lives_; }
    bool gameOver() const { return over_; }
    int frameNumber() const { return frame_; }
    int seed() const { return seed_; }
    static const std::vector<int>& minimalActionSet()
    {
        static const std::vector<int> set = {0, 2, 3, 4, 5, 6, 7, 8, 9};
        return set;
    }

    // screen as interleaved RGB bytes [96][96][3] (ALEInterface::getScreenRGB order)
    void screenRGB(uint8_t* out) const
    {
        for (int y = 0; y < kRes; ++y) {
            for (int x = 0; x < kRes; ++x) {
                uint8_t* p = out + (y * kRes + x) * 3;
                const int gx = x / kCell, gy = y / kCell, ox = x % kCell, oy = y % kCell;
                p[0] = 0, p[1] = 0, p[2] = static_cast<uint8_t>(24 + 2 * gy); // background: a vertical gradient
                if (pellets_[gy * kGrid + gx] && ox >= 3 && ox <= 4 && oy >= 3 && oy <= 4) { p[0] = 228, p[1] = 200, p[2] = 160; }
                if (gx == px_ && gy == py_ && ox >= 1 && ox <= 6 && oy >= 1 && oy <= 6) { p[0] = 252, p[1] = 224, p[2] = 0; }
                for (int c = 0; c < 2; ++c) {
                    if (gx == cx_[c] && gy == cy_[c] && ox >= 1 && ox <= 6 && oy >= 2 && oy <= 7) { p[0] = (c ? 0 : 224), p[1] = (c ? 200 : 40), p[2] = (c ? 224 : 40); }
                }
                if (y < 2 && x < 8 * lives_) { p[0] = 200, p[1] = 72, p[2] = 72; } // lives bar
            }
        }
    }

private:
    static int clampi(int v) { return v < 0 ? 0 : (v >= kGrid ? kGrid - 1 : v); }
    static int sign(int v) { return (v > 0) - (v < 0); }
    uint64_t next()
    { // splitmix64
        uint64_t z = (rng_ += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }

    int seed_ = 0, frame_ = 0, lives_ = 3, px_ = 0, py_ = 0, cx_[2] = {0, 0}, cy_[2] = {0, 0}, left_ = 0;
    bool over_ = false;
    uint64_t rng_ = 0;
    std::vector<uint8_t> pellets_;
};

} // namespace mzhost
