// The self-play worker: same outside behaviour as the reference's ActorGroup (actor/actor_group.{h,cpp}) — stdin
// commands, one `SelfPlay ... #` line per finished game on stdout, diagnostics on stderr — with every simulation of
// every game executed on the GPU(s) through the C ABI of include/mz_b200.h. The host is touched once per MOVE: it draws
// the search's randomness in the reference's order, decides the moves from the root tables, keeps the records.
#pragma once
#include "../../include/mz_b200.h"
#include "config.h"
#include "net_loader.h"
#include "record.h"
#include "rng.h"
#include "synth_atari.h"
#include "../csrc/killallgo_rules.h"
#include <atomic>
#include <deque>
#include <istream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mzhost {

struct Game {                    // what BaseActor / ZeroActor keep per game on the host
    std::vector<MoveRecord> moves; // action_info_history_ + action history
    bool enable_resign = true;     // zero_actor.cpp:23-27
    int turn = 1;
    int num_legal = 0;             // root children of the next search (= Dirichlet draws)
    uint8_t ttt[9] = {0};          // tictactoe board, only to know the end of the game in RNG order
    std::vector<uint8_t> stones;   // NoGo board (stones are never removed), for the same purpose
    uint64_t ka[2] = {0, 0};       // KillAllGo: Black / White stones as 64-bit boards (csrc/killallgo_rules.h), captures applied: the game ends by Benson
    // Atari (environment/atari/atari.{h,cpp}): the emulator and what AtariEnv keeps beside it
    SynthAtari emu;
    int seed = 0;                          // AtariEnv::seed_ (SD tag)
    float reward = 0.0f, total_reward = 0.0f;
    std::vector<int> lives_history;        // lives before every action, then the current ones
    std::vector<std::string> observations; // one screen per position (3 x 96 x 96 bytes, channel-major); old ones emptied (atari.cpp:72-79)
};

struct RootView { // one game's root child table, children in stored order (what MCTS hands to the move decision and the record)
    int num_children = 0;
    float root_mean = 0.0f, root_value = 0.0f;
    const int32_t* acts = nullptr;
    const float *cnt = nullptr, *mean = nullptr, *policy = nullptr, *logit = nullptr, *noise = nullptr;
    const float* reward = nullptr; // MCTSNode::getReward of the children (null: all zero)
    int bound_size = 0;            // MCTS::getTreeValueBound: number of distinct values, smallest, largest (actor_mcts_value_rescale)
    float bound_lo = 0.0f, bound_hi = 0.0f;
    int gumbel_best = -1;
};

class Worker {
public:
    Worker(Config& cfg, int wire_fd) : cfg_(cfg), wire_fd_(wire_fd) {}
    ~Worker();
    int run(); // ActorGroup::run (actor_group.cpp:136-148)
    int rngTest(std::istream& in); // `-mode rng_test`: the host's draw sequence without devices (CPU test hook)

private:
    bool initialize();                       // actor_group.cpp:150-187
    bool loadModel(const std::string& path); // load_model command (actor_group.cpp:227-232): rank-0 GPU reads, NCCL broadcast
    void handleIO();                         // actor_group.cpp:189-198
    void handleCommands();                   // actor_group.cpp:200-252
    bool playOneMove(int e0, int e1);        // S + 1 cycles of actor_group.cpp:81-134 for every game of engines e0 .. e1-1
    int runThreaded();                       // one host thread per engine (zero_num_threads > 1 and more than one GPU)
    void reportTiming();
    void startGames();                       // createActors + first rotation draws, in the reference's draw order
    void drawSearchRandomness(int e0, int e1); // root noise + rotations of cycles 1 .. S of one search
    int advanceGame(int g, const RootView& r, bool& resign, bool& end); // decide / act / end / next-game draws of one actor
    void restartGameHost(int g);
    int decideAction(int g, const RootView& r, bool& resign, int& child_index);
    float normalizedMean(const RootView& r, int child, int player) const; // MCTSNode::getNormalizedMean of a root child (child < 0: the root)
    void atariReset(Game& game, int seed);          // AtariEnv::reset(seed), atari.cpp:36-59
    void atariAct(Game& game, int action);          // AtariEnv::act, atari.cpp:61-92
    bool atariTerminal(const Game& game) const;     // AtariEnv::isTerminal, atari.h:58
    bool hostTerminal(const Game& game) const;
    bool nogoHasLegalMove(const Game& game) const; // NoGoEnv::isTerminal needs the legal set (environment/nogo/nogo.h:27-68)
    void emitGame(int g, bool terminal, float eval_score);
    void resetGameHost(int g);

    Config& cfg_;
    int wire_fd_; // the zero server's end of the pipe (the process's original stdout)
    Random rng_;                            // the main thread's generator
    std::vector<Random> engine_rng_;        // runThreaded: one per engine thread (program_seed + thread id, actor_group.cpp:66-70)
    static thread_local Random* tl_rng_;    // generator of the calling thread (null: the main thread's)
    Random& rng() { return tl_rng_ ? *tl_rng_ : rng_; }
    std::mutex emit_mutex_, stat_mutex_;    // the wire + games_finished_; the timing sums
    NetInfo net_;
    GameHeader header_;
    int game_type_ = MZ_GAME_GO, board_ = 9, actions_ = 82, sims_ = 0, num_games_ = 0;
    bool muzero_ = false, gumbel_ = false, atari_ = false;
    SequenceConfig sequenceConfig() const
    {
        SequenceConfig c;
        c.sequence_length = cfg_.getInt("zero_actor_intermediate_sequence_length"), c.unrolling_step = cfg_.getInt("learner_muzero_unrolling_step");
        c.n_step_return = cfg_.getInt("learner_n_step_return");
        return c;
    }
    int initialNumLegal() const
    {
        if (game_type_ == MZ_GAME_HEX) { return board_ * board_; }
        if (game_type_ == MZ_GAME_GOMOKU) { // every point, or the two outer lines under env_gomoku_rule=outer_open (gomoku.cpp:53-56)
            const int inner = (board_ > 4 ? board_ - 4 : 0);
            return cfg_.getString("env_gomoku_rule") == "outer_open" ? board_ * board_ - inner * inner : board_ * board_;
        }
        if (game_type_ == MZ_GAME_ATARI) { return static_cast<int>(SynthAtari::minimalActionSet().size()); }
        if (game_type_ == MZ_GAME_KILLALLGO) { return board_ * board_; } // Black's first move must be a stone (killallgo.cpp:29-31)
        return game_type_ == MZ_GAME_GO ? board_ * board_ + 1 : (game_type_ == MZ_GAME_NOGO ? board_ * board_ : (game_type_ == MZ_GAME_OTHELLO ? 4 : 9));
    }
    std::vector<mz_engine*> engines_;  // one per visible GPU (actor_group.cpp:168-177)
    std::vector<int> engine_games_;    // games handled by each engine: game g -> engine g % n, slot g / n (actor_group.cpp:184-186)
    std::vector<Game> games_;
    std::vector<int> next_seed_;                  // Atari: emulator seed of every game's NEXT episode, drawn in order when the current one ends
    std::vector<std::vector<uint8_t>> rotations_; // per engine [(S+1)][games]
    std::vector<std::vector<float>> noise_;       // per engine [games][A]
    void* nccl_comms_ = nullptr;
    std::atomic<bool> running_{false}, quit_{false};
    std::deque<std::string> commands_;
    std::mutex mutex_;
    std::thread io_thread_;
    long long moves_played_ = 0, games_finished_ = 0;
    double t_draw_ = 0, t_search_ = 0, t_decide_ = 0, t_play_ = 0, t_emit_ = 0; // host wall seconds per phase (diagnostics on stderr at exit)
};

} // namespace mzhost
