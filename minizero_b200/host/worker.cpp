#include "worker.h"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>
#include <iostream>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include <random>
#include <sstream>
#include <unistd.h>

namespace mzhost {

Worker::~Worker()
{
    for (mz_engine* e : engines_) { mz_destroy(e); }
    if (io_thread_.joinable()) { io_thread_.detach(); }
}

// ---- start-up: createNeuralNetworks + createActors (actor_group.cpp:150-187) ---------------------------------
thread_local Random* Worker::tl_rng_ = nullptr;

bool Worker::initialize()
{
    const std::string model = cfg_.getString("nn_file_name");
    std::string err;
    if (!readNetInfo(model, net_, err)) {
        std::cerr << "Failed to load model \"" << model << "\": " << err << std::endl;
        return false;
    }
    if (net_.type_name != "alphazero" && net_.type_name != "muzero" && net_.type_name != "muzero_atari") {
        std::cerr << "nn_type_name \"" << net_.type_name << "\" is not implemented by this worker" << std::endl;
        return false;
    }
    muzero_ = (net_.type_name != "alphazero"); // the actor follows the loaded network's type (zero_actor.cpp:100-114)
    gumbel_ = cfg_.getBool("actor_use_gumbel");
    // the reference binary is compiled per game (environment/environment.h:5-110); here the model names its game
    if (net_.game_name == "tictactoe") {
        game_type_ = MZ_GAME_TICTACTOE, board_ = 3;
    } else if (net_.game_name.rfind("go_", 0) == 0) {
        game_type_ = MZ_GAME_GO, board_ = net_.dims.input_height;
        if (cfg_.getInt("env_board_size") != 0 && cfg_.getInt("env_board_size") != board_) {
            std::cerr << "env_board_size does not match the model's board" << std::endl;
            return false;
        }
    } else if (net_.game_name.rfind("killallgo_", 0) == 0) { // environment/killallgo: 7 x 7 only (killallgo.h:12,23)
        game_type_ = MZ_GAME_KILLALLGO, board_ = net_.dims.input_height;
        if (board_ != 7 || muzero_ || cfg_.getBool("env_killallgo_use_seki")) {
            std::cerr << "KillAllGo is built for 7x7 AlphaZero networks without the seki table (env_killallgo_use_seki=false)" << std::endl;
            return false;
        }
    } else if (net_.game_name.rfind("hex_", 0) == 0) {
        game_type_ = MZ_GAME_HEX, board_ = net_.dims.input_height;
    } else if (net_.game_name.rfind("gomoku_", 0) == 0) { // "gomoku_15x15" or, with env_gomoku_rule=outer_open, "gomoku_oo_15x15" (gomoku.h:36)
        game_type_ = MZ_GAME_GOMOKU, board_ = net_.dims.input_height;
    } else if (net_.game_name.rfind("nogo_", 0) == 0) {
        game_type_ = MZ_GAME_NOGO, board_ = net_.dims.input_height;
    } else if (net_.game_name.rfind("othello_", 0) == 0) {
        game_type_ = MZ_GAME_OTHELLO, board_ = net_.dims.input_height;
    } else if (net_.game_name.rfind("atari_", 0) == 0 && net_.type_name == "muzero_atari") {
        // the emulator is a host-side plug-in; this image has no ALE, so the deterministic synthetic frame source stands in (synth_atari.h)
        game_type_ = MZ_GAME_ATARI, board_ = 6, atari_ = true;
    } else {
        std::cerr << "game \"" << net_.game_name << "\" is not implemented by this worker" << std::endl;
        return false;
    }
    actions_ = net_.dims.action_size;
    sims_ = cfg_.getInt("actor_num_simulation");
    num_games_ = cfg_.getInt("zero_num_parallel_games");
    header_.game_name = net_.game_name, header_.board_size = board_, header_.has_komi = (game_type_ == MZ_GAME_GO || game_type_ == MZ_GAME_NOGO || game_type_ == MZ_GAME_KILLALLGO), header_.komi = cfg_.getFloat("env_go_komi");
    header_.model_file = model;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::cerr << "No CUDA device: this worker has no CPU path" << std::endl;
        return false;
    }
    ndev = std::min(ndev, num_games_);
    engine_games_.assign(ndev, 0);
    for (int g = 0; g < num_games_; ++g) { engine_games_[g % ndev]++; }
    for (int dev = 0; dev < ndev; ++dev) {
        mz_config c{};
        c.device = dev, c.game = game_type_, c.board_size = board_, c.num_games = engine_games_[dev], c.num_simulation = sims_;
        c.puct_base = cfg_.getFloat("actor_mcts_puct_base"), c.puct_init = cfg_.getFloat("actor_mcts_puct_init");
        c.reward_discount = cfg_.getFloat("actor_mcts_reward_discount"), c.komi = cfg_.getFloat("env_go_komi");
        // env_killallgo_ko_rule is a second name of the same setting in the reference (configuration.cpp:178,187: both bind env_go_ko_rule)
        c.ko_situational = (cfg_.getString("env_go_ko_rule") == "situational" || (game_type_ == MZ_GAME_KILLALLGO && cfg_.getString("env_killallgo_ko_rule") == "situational"));
        c.dirichlet_epsilon = cfg_.getFloat("actor_dirichlet_noise_epsilon");
        c.muzero = muzero_, c.use_gumbel = gumbel_;
        c.value_rescale = cfg_.getBool("actor_mcts_value_rescale");
        if (atari_) {
            for (int a : SynthAtari::minimalActionSet()) { c.atari_legal_mask |= 1u << a; }
        }
        c.hex_swap_rule = cfg_.getBool("env_hex_use_swap_rule");
        c.gomoku_exactly_five = cfg_.getBool("env_gomoku_exactly_five_stones"), c.gomoku_outer_open = (cfg_.getString("env_gomoku_rule") == "outer_open");
        c.gumbel_noise = (!cfg_.getBool("actor_use_dirichlet_noise") && cfg_.getBool("actor_use_gumbel_noise")); // zero_actor.cpp:197,205
        c.gumbel_sample_size = cfg_.getInt("actor_gumbel_sample_size");
        c.gumbel_sigma_visit_c = cfg_.getFloat("actor_gumbel_sigma_visit_c"), c.gumbel_sigma_scale_c = cfg_.getFloat("actor_gumbel_sigma_scale_c");
        mz_engine* e = nullptr;
        if (mz_create(&c, &e) != MZ_OK) {
            std::cerr << "mz_create failed on device " << dev << ": " << mz_last_error() << std::endl;
            return false;
        }
        engines_.push_back(e);
        rotations_.emplace_back(static_cast<size_t>(sims_ + 1) * engine_games_[dev], 0);
        noise_.emplace_back(static_cast<size_t>(engine_games_[dev]) * actions_, 0.0f);
    }
    if (ndev > 1) { // one communicator per GPU in this process: the packed weights travel over NVLink
        ncclComm_t* comms = new ncclComm_t[ndev];
        std::vector<int> devs(ndev);
        for (int i = 0; i < ndev; ++i) { devs[i] = i; }
        if (ncclCommInitAll(comms, ndev, devs.data()) != ncclSuccess) {
            std::cerr << "ncclCommInitAll failed" << std::endl;
            return false;
        }
        nccl_comms_ = comms;
    }
    if (!loadModel(model)) { return false; }

    startGames();
    io_thread_ = std::thread(&Worker::handleIO, this);
    return true;
}

// the host-side start of a run, in the reference's draw order
void Worker::startGames()
{
    // createActors: ZeroActor::reset draws the resign switch of every game on the main thread's generator,
    // seeded with program_seed (console/mode_handler.cpp:62, create_actor.h:12-14, zero_actor.cpp:23-27)
    games_.assign(num_games_, Game());
    next_seed_.assign(num_games_, 0);
    for (int g = 0; g < num_games_; ++g) {
        if (atari_) { (void)rng().randInt(); } // the actor's AtariEnv member resets itself when it is constructed (atari.h:45-49): one seed drawn and dropped
        resetGameHost(g);
    }
    if (atari_) { // the first screens
        const int ne0 = static_cast<int>(engine_games_.size());
        for (int e = 0; e < ne0; ++e) {
            std::vector<int32_t> acts(engine_games_[e], -1);
            std::vector<uint8_t> frames(static_cast<size_t>(engine_games_[e]) * 3 * 96 * 96);
            for (int slot = 0; slot < engine_games_[e]; ++slot) {
                const std::string& obs = games_[slot * ne0 + e].observations.back();
                std::copy(obs.begin(), obs.end(), frames.begin() + static_cast<size_t>(slot) * obs.size());
            }
            if (!engines_.empty() && mz_atari_observe(engines_[e], acts.data(), frames.data()) != MZ_OK) { std::cerr << "mz_atari_observe failed: " << mz_last_error() << std::endl; }
        }
    }
    // slave thread 0 re-seeds: program_seed + thread id, or a random device (actor_group.cpp:66-70)
    rng().seed(cfg_.getBool("program_auto_seed") ? static_cast<int>(std::random_device()()) : cfg_.getInt("program_seed") + 0);
    // first beforeNNEvaluation of every game: rotation draw of cycle 0 (zero_actor.cpp:56)
    const bool random_rotation = cfg_.getBool("actor_use_random_rotation_features") && !muzero_; // only the AlphaZero branch draws one (zero_actor.cpp:54-57)
    const int ne = static_cast<int>(engine_games_.size());
    for (int g = 0; g < num_games_; ++g) {
        const int e = g % ne, slot = g / ne;
        rotations_[e][slot] = (random_rotation ? static_cast<uint8_t>(rng().randInt() % 8) : 0);
    }
}

void Worker::resetGameHost(int g)
{
    Game& game = games_[g];
    game.moves.clear();
    game.turn = 1;
    game.num_legal = initialNumLegal();
    std::fill(game.ttt, game.ttt + 9, 0);
    game.ka[0] = game.ka[1] = 0;
    game.stones.assign((game_type_ == MZ_GAME_NOGO || game_type_ == MZ_GAME_GOMOKU || game_type_ == MZ_GAME_HEX) ? board_ * board_ : 0, 0);
    if (atari_) { atariReset(game, rng().randInt()); } // BaseActor::reset -> AtariEnv::reset() draws the emulator seed (atari.h:54) before the resign switch
    game.enable_resign = (rng().randReal() < cfg_.getFloat("zero_disable_resign_ratio") ? false : true);
}

namespace {
std::string atariObservation(const SynthAtari& emu) // AtariEnv::getObservationString (atari.cpp:162-172): the 96 x 96 screen, channel-major bytes
{
    std::vector<uint8_t> rgb(3 * 96 * 96);
    emu.screenRGB(rgb.data());
    std::string obs(rgb.size(), '\0');
    for (int c = 0; c < 3; ++c) {
        for (int j = 0; j < 96 * 96; ++j) { obs[static_cast<size_t>(c) * 96 * 96 + j] = static_cast<char>(rgb[static_cast<size_t>(j) * 3 + c]); }
    }
    return obs;
}
} // namespace

void Worker::atariReset(Game& game, int seed)
{
    game.seed = seed, game.reward = 0.0f, game.total_reward = 0.0f;
    game.emu.reset(seed);
    game.lives_history.assign(1, game.emu.lives());
    game.observations.assign(1, atariObservation(game.emu));
}

void Worker::atariAct(Game& game, int action)
{
    game.reward = 0.0f;
    for (int i = 0; i < 4; ++i) { game.reward += static_cast<float>(game.emu.act(action)); } // kAtariFrameSkip, atari.cpp:68
    game.total_reward += game.reward;
    game.lives_history.push_back(game.emu.lives());
    game.observations.push_back(atariObservation(game.emu));
    // "only keep the most recent N observations" (atari.cpp:72-79)
    const int seq = cfg_.getInt("zero_actor_intermediate_sequence_length");
    const size_t recent = (seq == 0 ? 108000 : seq + 8 + cfg_.getInt("learner_n_step_return") + cfg_.getInt("learner_muzero_unrolling_step")) + 1;
    if (game.observations.size() > recent) {
        std::string& old = game.observations[game.observations.size() - recent];
        old.clear();
        old.shrink_to_fit();
    }
}

bool Worker::atariTerminal(const Game& game) const { return static_cast<int>(game.moves.size()) * 4 >= 108000 || game.emu.gameOver(); }

bool Worker::loadModel(const std::string& path)
{
    std::string err;
    if (!loadNetwork(path, engines_[0], err)) {
        std::cerr << "Failed to load model \"" << path << "\": " << err << std::endl;
        return false;
    }
    const int n = static_cast<int>(engines_.size());
    if (n > 1) {
        for (int i = 1; i < n; ++i) {
            if (!configureEmpty(net_, engines_[i], err)) {
                std::cerr << "network allocation failed: " << err << std::endl;
                return false;
            }
        }
        ncclComm_t* comms = static_cast<ncclComm_t*>(nccl_comms_);
        ncclGroupStart();
        for (int i = 0; i < n; ++i) {
            void* blob = nullptr;
            int64_t bytes = 0;
            mz_net_blob(engines_[i], &blob, &bytes);
            cudaSetDevice(i);
            ncclBroadcast(blob, blob, static_cast<size_t>(bytes), ncclUint8, 0, comms[i], nullptr);
        }
        ncclGroupEnd();
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(i);
            cudaDeviceSynchronize();
        }
    }
    cfg_.set("nn_file_name", path);
    header_.model_file = path;
    return true;
}

// ---- commands (actor_group.cpp:189-252) --------------------------------------------------------------------------
void Worker::handleIO()
{
    std::string command;
    while (std::getline(std::cin, command)) {
        std::lock_guard<std::mutex> lock(mutex_);
        commands_.push_back(command);
    }
    std::lock_guard<std::mutex> lock(mutex_);
    commands_.push_back("quit"); // stdin closed: the server is gone
}

void Worker::handleCommands()
{
    std::deque<std::string> todo;
    {
        std::lock_guard<std::mutex> lock(mutex_);
        todo.swap(commands_);
    }
    const std::string ignored = cfg_.getString("zero_actor_ignored_command");
    for (const std::string& command : todo) {
        const std::string prefix = (command.find(" ") == std::string::npos ? command : command.substr(0, command.find(" ")));
        {
            std::istringstream iss(ignored);
            std::string word;
            bool skip = false;
            while (iss >> word) { skip |= (word == prefix); }
            if (skip) {
                std::cerr << "[ignored command] " << command << std::endl;
                continue;
            }
        }
        if (prefix == "reset_actors") {
            std::cerr << "[command] " << command << std::endl;
            for (int g = 0; g < num_games_; ++g) { resetGameHost(g); }
            for (mz_engine* e : engines_) { mz_reset_game(e, -1); }
            {   // every actor's next beforeNNEvaluation draws the rotation of cycle 0 afresh (zero_actor.cpp:56), in actor order. (The reference draws
                // the resign switches above on its main-thread generator and these on the slave threads'; here both come from the calling thread's.)
                const bool random_rotation = cfg_.getBool("actor_use_random_rotation_features") && !muzero_;
                const int ne = static_cast<int>(engine_games_.size());
                for (int g = 0; g < num_games_; ++g) { rotations_[g % ne][g / ne] = (random_rotation ? static_cast<uint8_t>(rng().randInt() % 8) : 0); }
            }
        } else if (prefix == "load_model") {
            std::cerr << "[command] " << command << std::endl;
            if (command.find(" ") != std::string::npos && !loadModel(command.substr(command.find(" ") + 1))) { std::exit(0); }
        } else if (prefix == "update_config") {
            std::cerr << "[command] " << command << std::endl;
            // settings the engines captured when they were created: the reference reads most of them on every use, here they need a restart
            static const char* const fixed[] = {"actor_num_simulation", "zero_num_parallel_games", "actor_mcts_puct_base", "actor_mcts_puct_init",
                                                "actor_mcts_reward_discount", "actor_dirichlet_noise_epsilon", "actor_use_gumbel", "actor_use_gumbel_noise",
                                                "actor_gumbel_sample_size", "actor_gumbel_sigma_visit_c", "actor_gumbel_sigma_scale_c", "env_board_size",
                                                "env_go_komi", "env_go_ko_rule", "env_gomoku_rule", "env_gomoku_exactly_five_stones", "env_hex_use_swap_rule"};
            std::vector<std::string> before;
            for (const char* k : fixed) { before.push_back(cfg_.getString(k)); }
            // the engines were told at start-up whether root noise goes to the logits (Gumbel) or to the priors (Dirichlet): with Gumbel noise
            // configured, the Dirichlet switch decides that and is fixed too
            const bool noise_kind_fixed = cfg_.getBool("actor_use_gumbel_noise");
            const std::string dirichlet_before = cfg_.getString("actor_use_dirichlet_noise");
            if (command.find(" ") == std::string::npos || !cfg_.loadFromString(command.substr(command.find(" ") + 1))) {
                std::cerr << "Failed to load configuration string." << std::endl;
                std::exit(0);
            }
            if (noise_kind_fixed && cfg_.getString("actor_use_dirichlet_noise") != dirichlet_before) {
                std::cerr << "[warning] actor_use_dirichlet_noise is fixed when the worker starts with actor_use_gumbel_noise; the running engines keep " << dirichlet_before << std::endl;
                cfg_.set("actor_use_dirichlet_noise", dirichlet_before);
            }
            for (size_t i = 0; i < before.size(); ++i) {
                if (cfg_.getString(fixed[i]) != before[i]) {
                    std::cerr << "[warning] " << fixed[i] << " is fixed when the worker starts; the running engines keep " << before[i] << std::endl;
                    cfg_.set(fixed[i], before[i]); // keep the host's view consistent with the engines
                }
            }
        } else if (prefix == "start") {
            std::cerr << "[command] " << command << std::endl;
            running_ = true;
        } else if (prefix == "stop") {
            std::cerr << "[command] " << command << std::endl;
            running_ = false;
        } else if (prefix == "quit") {
            std::cerr << "[command] " << command << std::endl;
            quit_ = true;
        } // anything else (keep_alive ...) falls through silently, as in the reference
    }
}

// ---- move decision (zero_actor.cpp:178-192, mcts.cpp:84-124) -----------------------------------------------------
// MCTSNode::getNormalizedMean (mcts.cpp:40-53) of a root child, or of the root itself (child < 0; its reward is 0): reward + discount *
// mean, min-max rescaled with the tree's value bounds under actor_mcts_value_rescale, negated for White's nodes; no virtual loss
float Worker::normalizedMean(const RootView& r, int child, int player) const
{
    const float discount = cfg_.getFloat("actor_mcts_reward_discount");
    const float mean = (child < 0 ? r.root_mean : r.mean[child]), count = (child < 0 ? static_cast<float>(sims_ + 1) : r.cnt[child]);
    float value = (child >= 0 && r.reward ? r.reward[child] : 0.0f) + discount * mean;
    if (cfg_.getBool("actor_mcts_value_rescale")) {
        if (r.bound_size < 2) { return 1.0f; }
        value = (value - r.bound_lo) / (r.bound_hi - r.bound_lo);
        value = fmin(1, fmax(-1, 2 * value - 1));
    }
    value = (player == 2 ? -value : value);
    value = (value * count - 0.0f) / (count + 0.0f);
    return value;
}

int Worker::decideAction(int g, const RootView& r, bool& resign, int& child_index)
{
    const Game& game = games_[g];
    const int* actions = r.acts;
    const float* counts = r.cnt;
    const int num_children = r.num_children;
    const int child_player = game.turn; // children of the root carry the side to move (zero_actor.cpp:33,217)
    // selectChildByMaxCount (mcts.cpp:91-104)
    int best = -1;
    float max_count = 0.0f;
    for (int i = 0; i < num_children; ++i) {
        if (counts[i] <= max_count) { continue; }
        max_count = counts[i];
        best = i;
    }
    int selected = best;
    if (!cfg_.getBool("actor_select_action_by_count") && cfg_.getBool("actor_select_action_by_softmax_count")) {
        // selectChildBySoftmaxCount (mcts.cpp:106-124)
        const float temperature = cfg_.getFloat("actor_select_action_softmax_temperature"), value_threshold = 0.1f;
        const float best_mean = normalizedMean(r, best, child_player);
        float sum = 0.0f;
        selected = -1;
        for (int i = 0; i < num_children; ++i) {
            float count = std::pow(counts[i], 1 / temperature);
            float mean = (counts[i] == 0 ? 0.0f / 0.0f : normalizedMean(r, i, child_player));
            if (count == 0 || (mean < best_mean - value_threshold)) { continue; }
            sum += count;
            float rand = rng().randReal(sum);
            if (selected == -1 || rand < count) { selected = i; }
        }
    }
    child_index = selected;
    // isResign (zero_actor.h:42, mcts.cpp:84-89); the root's action player is the previous player (zero_actor.cpp:33; player 1 in a one-player game)
    const float root_win_rate = normalizedMean(r, -1, atari_ ? 1 : 3 - game.turn);
    const float action_win_rate = normalizedMean(r, selected, child_player);
    const float threshold = cfg_.getFloat("actor_resign_threshold");
    resign = game.enable_resign && (-root_win_rate < threshold && action_win_rate < threshold);
    return actions[selected];
}

bool Worker::hostTerminal(const Game& game) const
{
    const int n = static_cast<int>(game.moves.size());
    if (game_type_ == MZ_GAME_GO) {
        const int pass = board_ * board_;
        if (n >= 2 && game.moves[n - 1].action == pass && game.moves[n - 2].action == pass) { return true; } // go.cpp:249-251
        return n > 2 * board_ * board_;                                                                        // go.cpp:254
    }
    if (game_type_ == MZ_GAME_KILLALLGO) { // killallgo.cpp:34-40, then GoEnv::isTerminal
        if (mz_ka_terminal(game.ka[0], game.ka[1])) { return true; }
        const int pass = board_ * board_;
        if (n >= 2 && game.moves[n - 1].action == pass && game.moves[n - 2].action == pass) { return true; }
        return n > 2 * board_ * board_;
    }
    if (game_type_ == MZ_GAME_NOGO) { return !nogoHasLegalMove(game); } // nogo.h:61-68
    if (game_type_ == MZ_GAME_HEX) { // hex.cpp:96-99: a player connects its two edges (Black columns 0 / N-1, White rows 0 / N-1; hex.cpp:47-58,305-347)
        const int N = board_;
        for (int player = 1; player <= 2; ++player) {
            std::vector<int> stack;
            std::vector<uint8_t> seen(N * N, 0);
            for (int i = 0; i < N; ++i) {
                const int p = (player == 1 ? i * N : i);
                if (game.stones[p] == player) { seen[p] = 1, stack.push_back(p); }
            }
            static const int dx[6] = {-1, 0, -1, 1, 0, 1}, dy[6] = {-1, -1, 0, 0, 1, 1};
            while (!stack.empty()) {
                const int p = stack.back(), x = p % N, y = p / N;
                stack.pop_back();
                if (player == 1 ? (x == N - 1) : (y == N - 1)) { return true; }
                for (int k = 0; k < 6; ++k) {
                    const int qx = x + dx[k], qy = y + dy[k];
                    if (qx < 0 || qx >= N || qy < 0 || qy >= N) { continue; }
                    const int q = qy * N + qx;
                    if (game.stones[q] == player && !seen[q]) { seen[q] = 1, stack.push_back(q); }
                }
            }
        }
        return false;
    }
    if (game_type_ == MZ_GAME_GOMOKU) { // gomoku.cpp:60-63,140-164: five in a row through the last move, or a full board
        if (n == 0) { return false; }
        const int pos = game.moves[n - 1].action, who = game.stones[pos], N = board_;
        static const int dirs[4][2] = {{1, 0}, {0, 1}, {1, 1}, {1, -1}};
        for (const auto& d : dirs) {
            int count = 1;
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                int x = pos % N + sgn * d[0], y = pos / N + sgn * d[1];
                while (x >= 0 && x < N && y >= 0 && y < N && game.stones[y * N + x] == who) { ++count, x += sgn * d[0], y += sgn * d[1]; }
            }
            if (cfg_.getBool("env_gomoku_exactly_five_stones") ? (count == 5) : (count >= 5)) { return true; }
        }
        return n == N * N;
    }
    if (game_type_ == MZ_GAME_OTHELLO) { // othello.cpp:201-207
        const int pass = board_ * board_;
        return n >= 2 && game.moves[n - 1].action == pass && game.moves[n - 2].action == pass;
    }
    static const int lines[8][3] = {{0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {0, 3, 6}, {1, 4, 7}, {2, 5, 8}, {0, 4, 8}, {2, 4, 6}};
    for (const auto& l : lines) {
        if (game.ttt[l[0]] != 0 && game.ttt[l[0]] == game.ttt[l[1]] && game.ttt[l[1]] == game.ttt[l[2]]) { return true; }
    }
    return n == 9; // tictactoe.cpp:51-55
}

// NoGoEnv::isLegalAction for the side to move (nogo.h:27-59): an empty point with an empty neighbour or an own neighbouring block
// with more than one liberty, and no opposing neighbouring block in atari. Stones never leave the board in NoGo.
bool Worker::nogoHasLegalMove(const Game& game) const
{
    const int n = board_, nn = n * n, me = game.turn, opp = 3 - me;
    const std::vector<uint8_t>& b = game.stones;
    auto liberties = [&](int start) {
        std::vector<int> stack{start};
        std::vector<uint8_t> seen(nn, 0), lib(nn, 0);
        seen[start] = 1;
        int libs = 0;
        while (!stack.empty()) {
            const int p = stack.back();
            stack.pop_back();
            const int x = p % n, y = p / n;
            const int nb[4] = {y + 1 < n ? p + n : -1, x + 1 < n ? p + 1 : -1, y > 0 ? p - n : -1, x > 0 ? p - 1 : -1};
            for (int q : nb) {
                if (q < 0) { continue; }
                if (b[q] == 0) {
                    if (!lib[q]) { lib[q] = 1, ++libs; }
                } else if (b[q] == b[start] && !seen[q]) {
                    seen[q] = 1;
                    stack.push_back(q);
                }
            }
        }
        return libs;
    };
    for (int pos = 0; pos < nn; ++pos) {
        if (b[pos] != 0) { continue; }
        const int x = pos % n, y = pos / n;
        const int nb[4] = {y + 1 < n ? pos + n : -1, x + 1 < n ? pos + 1 : -1, y > 0 ? pos - n : -1, x > 0 ? pos - 1 : -1};
        bool legal = false, captures = false;
        for (int q : nb) {
            if (q < 0) { continue; }
            if (b[q] == 0) {
                legal = true;
            } else if (b[q] == me) {
                if (liberties(q) > 1) { legal = true; }
            } else if (b[q] == opp && liberties(q) == 1) {
                captures = true;
            }
        }
        if (legal && !captures) { return true; }
    }
    return false;
}

void Worker::emitGame(int g, bool terminal, float eval_score)
{
    AtariRecord at;
    if (atari_) { at.observations = &games_[g].observations, at.lives_history = &games_[g].lives_history, at.seed = games_[g].seed, at.total_reward = games_[g].total_reward; }
    const float ka_score = (game_type_ == MZ_GAME_KILLALLGO ? (mz_ka_winner(games_[g].ka[0], games_[g].ka[1]) == 1 ? 1.0f : -1.0f) : 0.0f);
    const std::string line = selfPlayLine(header_, games_[g].moves, terminal, eval_score, games_[g].turn, sequenceConfig(), atari_ ? &at : nullptr,
                                          game_type_ == MZ_GAME_KILLALLGO ? &ka_score : nullptr);
    const std::string out = line + "\n"; // the only thing this process ever writes to the server (zero_server.cpp:111-139)
    std::lock_guard<std::mutex> lock(emit_mutex_); // engine threads share the wire
    size_t done = 0;
    while (done < out.size()) {
        const ssize_t n = write(wire_fd_, out.data() + done, out.size() - done);
        if (n <= 0) {
            std::cerr << "server connection closed" << std::endl;
            std::exit(0);
        }
        done += static_cast<size_t>(n);
    }
    ++games_finished_;
}

// (1) randomness of cycles 1 .. S of one search in the reference's per-cycle, per-actor order: in cycle 1
//     every actor first receives its root noise (afterNNEvaluation of the root) and then draws the rotation of its next leaf
void Worker::drawSearchRandomness(int e0, int e1)
{
    const int ne = static_cast<int>(engine_games_.size()), S1 = sims_ + 1, A = actions_;
    const bool use_dirichlet = cfg_.getBool("actor_use_dirichlet_noise"), use_gumbel_noise = (!use_dirichlet && cfg_.getBool("actor_use_gumbel_noise"));
    const bool use_noise = use_dirichlet || use_gumbel_noise, random_rotation = cfg_.getBool("actor_use_random_rotation_features") && !muzero_;
    const float alpha = cfg_.getFloat("actor_dirichlet_noise_alpha");
    for (int c = 1; c < S1; ++c) {
        for (int g = 0; g < num_games_; ++g) {
            const int e = g % ne, slot = g / ne;
            if (e < e0 || e >= e1) { continue; }
            if (c == 1 && use_noise) {
                std::vector<float> dir = (use_dirichlet ? rng().randDirichlet(alpha, games_[g].num_legal) : rng().randGumbel(games_[g].num_legal)); // zero_actor.cpp:194-213
                float* dst = noise_[e].data() + static_cast<size_t>(slot) * A;
                std::fill(dst, dst + A, 0.0f);
                std::copy(dir.begin(), dir.end(), dst);
            }
            rotations_[e][static_cast<size_t>(c) * engine_games_[e] + slot] = (random_rotation ? static_cast<uint8_t>(rng().randInt() % 8) : 0);
        }
    }
}

// (4) for one actor, in the reference's order: decide the move (softmax-count draws), act or resign, detect the end of the game, draw the
//     resign switch of the next game if it ended, draw the first rotation of the next search. Returns the action to play (-1: resigned).
int Worker::advanceGame(int g, const RootView& r, bool& resign, bool& end)
{
    const int ne = static_cast<int>(engine_games_.size()), e = g % ne, slot = g / ne;
    const bool random_rotation = cfg_.getBool("actor_use_random_rotation_features") && !muzero_;
    Game& game = games_[g];
    resign = false;
    int child = -1;
    int action = decideAction(g, r, resign, child);
    if (gumbel_ && cfg_.getBool("actor_select_action_by_count")) { // GumbelZero::decideActionNode: best-scoring candidate (gumbel_zero.cpp:61-66)
        action = r.gumbel_best;
        for (int i = 0; i < r.num_children; ++i) {
            if (r.acts[i] == action) { child = i; }
        }
        const float threshold = cfg_.getFloat("actor_resign_threshold");
        const float root_win_rate = normalizedMean(r, -1, atari_ ? 1 : 3 - game.turn);
        const float action_win_rate = normalizedMean(r, child, game.turn);
        resign = game.enable_resign && (-root_win_rate < threshold && action_win_rate < threshold); // mcts.cpp:84-89
    }
    end = resign;
    int play = -1;
    if (!resign) { // BaseActor::act + getActionInfo (base_actor.cpp:22-30,59-66)
        MoveRecord m;
        m.action = action, m.player = game.turn;
        float value_pi = r.root_value; // gumbel_zero.cpp:20-31: the root's own value goes through the same min-max rescaling
        if (cfg_.getBool("actor_mcts_value_rescale")) {
            if (r.bound_size < 2) {
                value_pi = 1.0f;
            } else {
                value_pi = (value_pi - r.bound_lo) / (r.bound_hi - r.bound_lo);
                value_pi = fmin(1, fmax(-1, 2 * value_pi - 1));
            }
        }
        m.policy = (gumbel_ ? gumbelPolicy(r.acts, r.cnt, r.policy, r.logit, r.noise, r.num_children, value_pi, game.turn, sims_, cfg_.getFloat("actor_gumbel_sigma_visit_c"),
                                           cfg_.getFloat("actor_gumbel_sigma_scale_c"), [&](int i) { return normalizedMean(r, i, game.turn); })
                            : searchDistribution(r.acts, r.cnt, r.num_children)); // zero_actor.h:48
        m.value = std::to_string(r.root_mean); // zero_actor.h:49
        m.reward = "0";                        // operator<< of Environment::getReward() == 0.0f (go.h:50, tictactoe.h:25)
        if (atari_) { // the emulator answers: reward of the four frames (atari.cpp:66-70), streamed with operator<< (zero_actor.cpp:122-127)
            atariAct(game, action);
            std::ostringstream oss;
            oss << game.reward;
            m.reward = oss.str();
        }
        game.moves.push_back(m);
        if (game_type_ == MZ_GAME_TICTACTOE && action >= 0 && action < 9) { game.ttt[action] = static_cast<uint8_t>(game.turn); }
        if ((game_type_ == MZ_GAME_NOGO || game_type_ == MZ_GAME_GOMOKU) && action >= 0 && action < board_ * board_) {
            game.stones[action] = static_cast<uint8_t>(game.turn);
        }
        if (game_type_ == MZ_GAME_KILLALLGO && action >= 0 && action < board_ * board_) { // GoEnv::act with captures (go.cpp:150-178)
            mz_ka_place(game.ka[game.turn - 1], game.ka[2 - game.turn], (action / board_) * 8 + action % board_);
        }
        if (game_type_ == MZ_GAME_HEX && action >= 0 && action < board_ * board_) { // hex.cpp:21-66 (game.moves already holds this move)
            int id = action;
            if (cfg_.getBool("env_hex_use_swap_rule") && game.moves.size() == 2 && action == game.moves[0].action) { // swap: mirrored, first stone removed
                id = (board_ - 1 - action % board_) * board_ + (board_ - 1 - action / board_);
                game.stones[action] = 0;
            }
            game.stones[id] = static_cast<uint8_t>(game.turn);
        }
        if (!atari_) { game.turn = 3 - game.turn; } // a one-player game stays with player 1 (atari.h:18)
        play = action;
        end = (atari_ ? atariTerminal(game) : hostTerminal(game));
    }
    if (g == 0 && !cfg_.getBool("program_quiet")) {
        std::cerr << "[actor 0] move " << game.moves.size() << " action " << action << (resign ? " (resign)" : "") << " root mean " << r.root_mean << std::endl;
    }
    // actor->reset() draws the resign switch of the next game (zero_actor.cpp:26); the deferred state reset must not consume
    // randomness, so only the draw happens here, in order
    if (end) {
        if (atari_) { next_seed_[g] = rng().randInt(); } // AtariEnv::reset() of the next game (atari.h:54), before the resign switch
        game.enable_resign = (rng().randReal() < cfg_.getFloat("zero_disable_resign_ratio") ? false : true);
    }
    rotations_[e][slot] = (random_rotation ? static_cast<uint8_t>(rng().randInt() % 8) : 0); // cycle 0 of the next search
    return play;
}

// BaseActor::reset of a finished game on the host (the resign switch of the next game was already drawn, in order)
void Worker::restartGameHost(int g)
{
    Game& game = games_[g];
    game.moves.clear();
    game.turn = 1;
    game.num_legal = initialNumLegal();
    std::fill(game.ttt, game.ttt + 9, 0);
    std::fill(game.stones.begin(), game.stones.end(), 0);
    game.ka[0] = game.ka[1] = 0;
    if (atari_) { atariReset(game, next_seed_[g]); } // the seed was drawn in order when the game ended
}

// ---- one move for every game ------------------------------------------------------------------------------------------
// engines e0 .. e1-1 and their games: the whole worker from the main thread (zero_num_threads = 1: the reference's single-thread draw order, seed-exact),
// or one engine from its own host thread (run(): the engines then never wait for each other or for another engine's host work)
bool Worker::playOneMove(int e0, int e1)
{
    const int ne = static_cast<int>(engines_.size()), A = actions_;
    const bool use_dirichlet = cfg_.getBool("actor_use_dirichlet_noise"), use_gumbel_noise = (!use_dirichlet && cfg_.getBool("actor_use_gumbel_noise"));
    const bool use_noise = use_dirichlet || use_gumbel_noise, random_rotation = cfg_.getBool("actor_use_random_rotation_features") && !muzero_;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    nvtxRangePushA("move"); // phases of a move on the host timeline: draw, search, decide, play, emit
    struct PopAtExit { ~PopAtExit() { nvtxRangePop(); } } pop_move;
    nvtxRangePushA("draw");
    drawSearchRandomness(e0, e1); // (1)
    nvtxRangePop();
    const double t1 = now();
    // (2) the whole search on the devices, all engines in flight together
    for (int e = e0; e < e1; ++e) {
        if (mz_search_set_inputs(engines_[e], random_rotation ? rotations_[e].data() : nullptr, use_noise ? noise_[e].data() : nullptr) != MZ_OK ||
            mz_search_run(engines_[e], 0, nullptr) != MZ_OK) {
            std::cerr << "search failed: " << mz_last_error() << std::endl;
            return false;
        }
    }
    // (3) root tables
    struct Roots {
        std::vector<mz_root_info> info;
        std::vector<int32_t> action;
        std::vector<float> count, mean, policy, logit, noise, reward, bound_lo, bound_hi;
        std::vector<int32_t> gumbel_best, bound_size;
    };
    const bool rescale = cfg_.getBool("actor_mcts_value_rescale");
    std::vector<Roots> roots(ne);
    for (int e = e0; e < e1; ++e) {
        const size_t n = engine_games_[e];
        roots[e].info.resize(n), roots[e].action.resize(n * A), roots[e].count.resize(n * A), roots[e].mean.resize(n * A);
        if (gumbel_) { roots[e].policy.resize(n * A), roots[e].logit.resize(n * A), roots[e].noise.resize(n * A), roots[e].gumbel_best.resize(n); }
        if (mz_get_roots(engines_[e], roots[e].info.data(), roots[e].action.data(), roots[e].count.data(), roots[e].mean.data(), gumbel_ ? roots[e].policy.data() : nullptr,
                         gumbel_ ? roots[e].logit.data() : nullptr, gumbel_ ? roots[e].noise.data() : nullptr, nullptr) != MZ_OK) {
            std::cerr << "mz_get_roots failed: " << mz_last_error() << std::endl;
            return false;
        }
        if (atari_ || rescale) { // rewards of the root children and the value bounds: what getNormalizedMean reads beside count / mean (mcts.cpp:40-53)
            roots[e].reward.resize(n * A), roots[e].bound_size.resize(n), roots[e].bound_lo.resize(n), roots[e].bound_hi.resize(n);
            if (mz_get_root_rewards(engines_[e], roots[e].reward.data(), roots[e].bound_size.data(), roots[e].bound_lo.data(), roots[e].bound_hi.data()) != MZ_OK) {
                std::cerr << "mz_get_root_rewards failed: " << mz_last_error() << std::endl;
                return false;
            }
        }
        if (gumbel_ && mz_gumbel_best_actions(engines_[e], roots[e].gumbel_best.data()) != MZ_OK) {
            std::cerr << "mz_gumbel_best_actions failed: " << mz_last_error() << std::endl;
            return false;
        }
    }
    const double t2 = now();
    nvtxMarkA("decide + records");
    // (4) per actor, in order: decide, act or resign, restart finished games, draw the first rotation of the next search
    std::vector<std::vector<int32_t>> play(ne);
    for (int e = e0; e < e1; ++e) { play[e].assign(engine_games_[e], -1); }
    std::vector<int> ended; // games to emit after the devices have applied the moves (the final score comes from there)
    std::vector<char> ended_by_resign(num_games_, 0);
    for (int g = 0; g < num_games_; ++g) {
        const int e = g % ne, slot = g / ne;
        if (e < e0 || e >= e1) { continue; }
        const mz_root_info& ri = roots[e].info[slot];
        RootView r;
        r.num_children = ri.num_children, r.root_mean = ri.mean, r.root_value = ri.value;
        r.acts = roots[e].action.data() + static_cast<size_t>(slot) * A;
        r.cnt = roots[e].count.data() + static_cast<size_t>(slot) * A, r.mean = roots[e].mean.data() + static_cast<size_t>(slot) * A;
        if (gumbel_) {
            r.policy = roots[e].policy.data() + static_cast<size_t>(slot) * A, r.logit = roots[e].logit.data() + static_cast<size_t>(slot) * A;
            r.noise = roots[e].noise.data() + static_cast<size_t>(slot) * A, r.gumbel_best = roots[e].gumbel_best[slot];
        }
        if (!roots[e].reward.empty()) {
            r.reward = roots[e].reward.data() + static_cast<size_t>(slot) * A;
            r.bound_size = roots[e].bound_size[slot], r.bound_lo = roots[e].bound_lo[slot], r.bound_hi = roots[e].bound_hi[slot];
        }
        bool resign = false, end = false;
        play[e][slot] = advanceGame(g, r, resign, end);
        if (end) {
            ended.push_back(g);
            ended_by_resign[g] = resign ? 1 : 0;
        }
    }
    const double t3 = now();
    // (5) apply the moves on the devices; finished games are emitted with the device's score and restarted
    std::vector<std::vector<mz_play_result>> res(ne);
    for (int e = e0; e < e1; ++e) {
        res[e].resize(engine_games_[e]);
        if (mz_play(engines_[e], play[e].data(), res[e].data()) != MZ_OK) {
            std::cerr << "mz_play failed: " << mz_last_error() << std::endl;
            return false;
        }
    }
    for (int g = 0; g < num_games_; ++g) {
        const int e = g % ne, slot = g / ne;
        if (e < e0 || e >= e1) { continue; }
        if (play[e][slot] >= 0) {
            if (!res[e][slot].applied) {
                std::cerr << "device rejected action " << play[e][slot] << " of game " << g << std::endl;
                return false;
            }
            games_[g].num_legal = res[e][slot].num_legal;
        }
    }
    if (atari_) { // the emulator's answers join the observation histories on the devices (AtariEnv::act's history update)
        for (int e = e0; e < e1; ++e) {
            std::vector<int32_t> acts(engine_games_[e], -2);
            std::vector<uint8_t> frames(static_cast<size_t>(engine_games_[e]) * 3 * 96 * 96);
            for (int slot = 0; slot < engine_games_[e]; ++slot) {
                const int g = slot * ne + e;
                if (play[e][slot] < 0 || std::find(ended.begin(), ended.end(), g) != ended.end()) { continue; }
                acts[slot] = play[e][slot];
                const std::string& obs = games_[g].observations.back();
                std::copy(obs.begin(), obs.end(), frames.begin() + static_cast<size_t>(slot) * obs.size());
            }
            if (mz_atari_observe(engines_[e], acts.data(), frames.data()) != MZ_OK) {
                std::cerr << "mz_atari_observe failed: " << mz_last_error() << std::endl;
                return false;
            }
        }
    }
    const double t4 = now();
    // games that go on: an intermediate sequence may be due (actor_group.cpp:126-132)
    {
        const SequenceConfig seq = sequenceConfig();
        for (int g = 0; g < num_games_ && seq.sequence_length > 0; ++g) {
            const int e = g % ne, slot = g / ne;
            if (e < e0 || e >= e1) { continue; }
            if (play[e][slot] < 0 || std::find(ended.begin(), ended.end(), g) != ended.end()) { continue; }
            if (!intermediateSequenceDue(static_cast<int>(games_[g].moves.size()), seq)) { continue; }
            emitGame(g, false, 0.0f);
            {
                std::lock_guard<std::mutex> lock(emit_mutex_);
                --games_finished_;
            }
            clearSentActionInfo(games_[g].moves, false, seq);
        }
    }
    for (int g : ended) {
        const int e = g % ne, slot = g / ne;
        const bool terminal = !ended_by_resign[g];
        if (terminal && !atari_ && !res[e][slot].terminal) { // (the end of an Atari episode is the emulator's word alone)
            std::cerr << "host / device disagree on the end of game " << g << std::endl;
            return false;
        }
        const bool keep_resign = games_[g].enable_resign; // already drawn for the next game
        emitGame(g, terminal, terminal ? res[e][slot].eval_score : 0.0f);
        mz_reset_game(engines_[e], slot);
        restartGameHost(g);
        games_[g].enable_resign = keep_resign;
    }
    if (atari_ && !ended.empty()) { // first screens of the restarted episodes
        for (int e = e0; e < e1; ++e) {
            std::vector<int32_t> acts(engine_games_[e], -2);
            std::vector<uint8_t> frames(static_cast<size_t>(engine_games_[e]) * 3 * 96 * 96);
            bool any = false;
            for (int g : ended) {
                if (g % ne != e) { continue; }
                const int slot = g / ne;
                acts[slot] = -1, any = true;
                const std::string& obs = games_[g].observations.back();
                std::copy(obs.begin(), obs.end(), frames.begin() + static_cast<size_t>(slot) * obs.size());
            }
            if (any && mz_atari_observe(engines_[e], acts.data(), frames.data()) != MZ_OK) {
                std::cerr << "mz_atari_observe failed: " << mz_last_error() << std::endl;
                return false;
            }
        }
    }
    {
        std::lock_guard<std::mutex> lock(stat_mutex_);
        moves_played_ += e1 - e0; // engine moves: one search of one engine's games
        t_draw_ += t1 - t0, t_search_ += t2 - t1, t_decide_ += t3 - t2, t_play_ += t4 - t3, t_emit_ += now() - t4;
    }
    return true;
}

// `-mode rng_test` (CPU only): the host's draw sequence of a run, driven by root tables read from stdin instead of the devices, so that
// it can be compared with what the reference drew for the same seed (rotations, Dirichlet / Gumbel noise, moves, resignations).
// stdin: "setup <game_type> <board> <actions>", then per move and game "root <g> <k> <root_mean_hex> <root_value_hex> a:count_hex:mean_hex ..."
// (B lines per move). stdout: "rot <cycle> <g> <r>", "noise <g> <hex ...>", "act <g> <action> <resign> <ended>".
int Worker::rngTest(std::istream& in)
{
    auto hexf = [](const std::string& t) {
        const uint32_t bits = static_cast<uint32_t>(std::stoul(t, nullptr, 16));
        float f;
        std::memcpy(&f, &bits, 4);
        return f;
    };
    std::string word;
    in >> word >> game_type_ >> board_ >> actions_;
    sims_ = cfg_.getInt("actor_num_simulation"), num_games_ = cfg_.getInt("zero_num_parallel_games");
    muzero_ = (cfg_.getString("nn_type_name") == "muzero"), gumbel_ = cfg_.getBool("actor_use_gumbel");
    atari_ = (game_type_ == MZ_GAME_ATARI);
    if (atari_) { header_.game_name = "atari_" + cfg_.getString("env_atari_name"), header_.model_file = cfg_.getString("nn_file_name"); }
    engine_games_.assign(1, num_games_);
    rotations_.assign(1, std::vector<uint8_t>(static_cast<size_t>(sims_ + 1) * num_games_, 0));
    noise_.assign(1, std::vector<float>(static_cast<size_t>(num_games_) * actions_, 0.0f));
    rng().seed(cfg_.getInt("program_seed")); // main thread generator (console/mode_handler.cpp:62)
    startGames();
    const int A = actions_;
    long cycle = 0;
    std::vector<std::vector<int32_t>> acts(num_games_);
    std::vector<std::vector<float>> cnt(num_games_), mean(num_games_), pol(num_games_), lgt(num_games_), nse(num_games_), rwd(num_games_);
    std::vector<RootView> views(num_games_);
    for (;;) {
        bool ok = true;
        for (int i = 0; i < num_games_ && ok; ++i) {
            int g, k;
            std::string mean_hex, value_hex;
            if (!(in >> word)) {
                ok = false;
                break;
            }
            int bsize = 0;
            std::string blo = "0", bhi = "0";
            if (word == "bounds") { in >> bsize >> blo >> bhi >> word; } // "bounds <size> <lo_hex> <hi_hex>" precedes the root line of a rescaled search
            if (!(in >> g >> k >> mean_hex >> value_hex)) {
                ok = false;
                break;
            }
            acts[g].assign(A, -1), cnt[g].assign(A, 0.0f), mean[g].assign(A, 0.0f), pol[g].assign(A, 0.0f), lgt[g].assign(A, 0.0f), nse[g].assign(A, 0.0f), rwd[g].assign(A, 0.0f);
            for (int j = 0; j < k; ++j) {
                std::string tok, part;
                in >> tok;
                std::istringstream ts(tok);
                std::getline(ts, part, ':');
                acts[g][j] = std::stoi(part);
                std::getline(ts, part, ':');
                cnt[g][j] = hexf(part);
                std::getline(ts, part, ':');
                mean[g][j] = hexf(part);
                if (std::getline(ts, part, ':')) { pol[g][j] = hexf(part); } // optional: policy, logit, noise (Gumbel records)
                if (std::getline(ts, part, ':')) { lgt[g][j] = hexf(part); }
                if (std::getline(ts, part, ':')) { nse[g][j] = hexf(part); }
                if (std::getline(ts, part, ':')) { rwd[g][j] = hexf(part); }
            }
            RootView& r = views[g];
            r.num_children = k, r.root_mean = hexf(mean_hex), r.root_value = hexf(value_hex);
            r.acts = acts[g].data(), r.cnt = cnt[g].data(), r.mean = mean[g].data();
            r.policy = pol[g].data(), r.logit = lgt[g].data(), r.noise = nse[g].data();
            r.reward = rwd[g].data(), r.bound_size = bsize, r.bound_lo = hexf(blo), r.bound_hi = hexf(bhi);
            games_[g].num_legal = k; // the root's children are its legal actions (Dirichlet / Gumbel draws, zero_actor.cpp:197,206)
        }
        if (!ok) { break; }
        for (int g = 0; g < num_games_; ++g) { std::cout << "rot " << cycle << " " << g << " " << static_cast<int>(rotations_[0][g]) << "\n"; }
        drawSearchRandomness(0, static_cast<int>(engine_games_.size()));
        for (int c = 1; c <= sims_; ++c) {
            for (int g = 0; g < num_games_; ++g) { std::cout << "rot " << cycle + c << " " << g << " " << static_cast<int>(rotations_[0][static_cast<size_t>(c) * num_games_ + g]) << "\n"; }
        }
        cycle += sims_ + 1;
        for (int g = 0; g < num_games_; ++g) {
            std::cout << "noise " << g;
            for (int j = 0; j < views[g].num_children; ++j) {
                uint32_t bits;
                std::memcpy(&bits, &noise_[0][static_cast<size_t>(g) * A + j], 4);
                std::cout << " " << std::hex << bits << std::dec;
            }
            std::cout << "\n";
        }
        for (int g = 0; g < num_games_; ++g) {
            bool resign = false, end = false;
            const int action = advanceGame(g, views[g], resign, end);
            std::cout << "act " << g << " " << action << " " << (resign ? 1 : 0) << " " << (end ? 1 : 0) << "\n";
            if (atari_) { // the emulator lives on the host: whole records can be checked without a device
                AtariRecord at;
                at.observations = &games_[g].observations, at.lives_history = &games_[g].lives_history, at.seed = games_[g].seed, at.total_reward = games_[g].total_reward;
                const SequenceConfig seq = sequenceConfig();
                if (end) {
                    std::cout << "line " << selfPlayLine(header_, games_[g].moves, !resign, 0.0f, 1, seq, &at) << "\n";
                } else if (intermediateSequenceDue(static_cast<int>(games_[g].moves.size()), seq)) {
                    std::cout << "line " << selfPlayLine(header_, games_[g].moves, false, 0.0f, 1, seq, &at) << "\n";
                    clearSentActionInfo(games_[g].moves, false, seq);
                }
            }
            if (end) {
                const bool keep = games_[g].enable_resign;
                restartGameHost(g);
                games_[g].enable_resign = keep;
            }
        }
    }
    std::cout.flush();
    return 0;
}

int Worker::run()
{
    // main thread generator: program_seed (console/mode_handler.cpp:62)
    rng().seed(cfg_.getInt("program_seed"));
    if (!initialize()) { return -1; }
    const int ne = static_cast<int>(engines_.size());
    if (ne > 1 && cfg_.getInt("zero_num_threads") > 1) { return runThreaded(); }
    while (!quit_) {
        handleCommands();
        if (quit_) { break; }
        if (!running_) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
            continue;
        }
        if (!playOneMove(0, ne)) { return -1; }
    }
    reportTiming();
    return 0;
}

// One host thread per engine (GPU), like the reference's slave threads (actor_group.cpp:66-70,81-134): each draws from its own generator seeded
// with program_seed + thread id, searches, decides and records its own games, and shares nothing with the others but the wire. The reference hands
// actors to its threads through a shared counter, so its draw order is not reproducible with more than one thread either; seed-exact runs use
// zero_num_threads = 1 (the loop in run()). Commands are executed by the main thread while every engine thread is parked between two moves.
int Worker::runThreaded()
{
    const int ne = static_cast<int>(engines_.size());
    engine_rng_.assign(ne, Random());
    for (int e = 0; e < ne; ++e) { engine_rng_[e].seed(cfg_.getBool("program_auto_seed") ? static_cast<int>(std::random_device()()) : cfg_.getInt("program_seed") + e); }
    std::mutex ctl;
    std::condition_variable cv;
    int parked = 0;
    bool pause = true, failed = false;
    std::vector<std::thread> threads;
    for (int e = 0; e < ne; ++e) {
        threads.emplace_back([&, e] {
            tl_rng_ = &engine_rng_[e];
            for (;;) {
                {
                    std::unique_lock<std::mutex> lock(ctl);
                    ++parked;
                    cv.notify_all();
                    cv.wait(lock, [&] { return quit_ || failed || (running_ && !pause); });
                    --parked;
                    if (quit_ || failed) { return; }
                }
                for (;;) {
                    {
                        std::lock_guard<std::mutex> lock(ctl);
                        if (quit_ || failed || pause || !running_) { break; }
                    }
                    if (!playOneMove(e, e + 1)) {
                        std::lock_guard<std::mutex> lock(ctl);
                        failed = true;
                        cv.notify_all();
                        break;
                    }
                }
            }
        });
    }
    for (;;) {
        bool pending;
        {
            std::lock_guard<std::mutex> lock(mutex_);
            pending = !commands_.empty();
        }
        {
            std::unique_lock<std::mutex> lock(ctl);
            if (failed) { break; }
            if (pending) {
                pause = true;
                cv.wait(lock, [&] { return parked == ne || failed; }); // every engine thread is between two moves
                lock.unlock();
                handleCommands(); // running_ / quit_ change here, under no lock the threads could be reading them through: they are parked
                lock.lock();
                pause = false;
                cv.notify_all();
                if (quit_) { break; }
            }
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    {
        std::lock_guard<std::mutex> lock(ctl);
        quit_ = true;
        cv.notify_all();
    }
    for (std::thread& t : threads) { t.join(); }
    reportTiming();
    return failed ? -1 : 0;
}

void Worker::reportTiming()
{
    if (moves_played_ > 0) { // a "move" is one engine's move (one search of its games); the milliseconds are host wall time of the thread that drove it
        const double k = 1e3 / static_cast<double>(moves_played_);
        std::cerr << "[timing] " << moves_played_ << " moves, " << games_finished_ << " games; ms per move: draw " << t_draw_ * k << ", search + root tables " << t_search_ * k
                  << ", decide + records " << t_decide_ * k << ", play " << t_play_ * k << ", emit + restart " << t_emit_ * k << std::endl;
    }
}

} // namespace mzhost
