/*
 * mz_b200 — C ABI of the B200-native MiniZero self-play engine (libmzb200.so).
 *
 * The reference (rlglab/minizero) has no FFI layer; its replaceable seams are the C++ classes
 * ActorGroup / BaseActor / Network and the `-mode sp` process. This header is the thin
 * boundary between a host that keeps those seams (minizero_b200/host: the ActorGroup-compatible worker;
 * the minizero_b200 Python package: the ctypes mirror used by tests and bench) and the CUDA library. Each entry point
 * names the reference interface it replaces (paths relative to /root/reference/minizero).
 *
 * Conventions: plain pointers and sizes, caller-owned host buffers, `int` status returns (0 = ok, <0 =
 * error, text via mz_last_error()), no exceptions and no stdout/stderr writes across the boundary, one
 * engine per device context, entry points of one engine are not re-entrant. There is NO CPU path: every
 * call fails with MZ_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MZ_B200_H
#define MZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZ_OK 0
#define MZ_ERR_ARG (-1)
#define MZ_ERR_CUDA (-2)
#define MZ_ERR_STATE (-3)

#define MZ_GAME_TICTACTOE 0 /* environment/tictactoe */
#define MZ_GAME_GO 1        /* environment/go        */
#define MZ_GAME_OTHELLO 2   /* environment/othello   */
#define MZ_GAME_NOGO 3      /* environment/nogo (GoEnv with its own legality / end / result; 9x9 in the reference) */
#define MZ_GAME_GOMOKU 4    /* environment/gomoku (N x N, no pass, five in a row through the last move) */
#define MZ_GAME_HEX 5       /* environment/hex (N x N, no pass, swap rule, connect the two own edges; features and policy are never rotated) */

#define MZ_GAME_KILLALLGO 7 /* environment/killallgo: GoEnv on 7 x 7; Black opens with two stones, wins when the whole board is unconditionally his (Benson), White wins
                             * with any unconditionally alive group or by surviving (killallgo.cpp:27-48; env_killallgo_use_seki = false, the default). AlphaZero only */
#define MZ_GAME_ATARI 6     /* environment/atari: one player, 18 actions, 32 x 96 x 96 planes; the emulator stays with the host (mz_atari_observe); MuZero only */

typedef struct mz_engine mz_engine;

/* Search / environment configuration: the config keys the actor path reads (config/configuration.cpp:13-28,80-81). */
typedef struct {
    int32_t device;             /* CUDA device ordinal */
    int32_t game;               /* MZ_GAME_* (compile-time GAME_TYPE in the reference, environment/environment.h:5-110) */
    int32_t board_size;         /* env_board_size (ignored for tictactoe) */
    int32_t num_games;          /* zero_num_parallel_games handled by this engine */
    int32_t num_simulation;     /* actor_num_simulation */
    float puct_base;            /* actor_mcts_puct_base */
    float puct_init;            /* actor_mcts_puct_init */
    float reward_discount;      /* actor_mcts_reward_discount */
    float komi;                 /* env_go_komi */
    int32_t ko_situational;     /* env_go_ko_rule == "situational" */
    float dirichlet_epsilon;    /* actor_dirichlet_noise_epsilon */
    int32_t muzero;             /* nn_type_name == "muzero" (actor/zero_actor.cpp:59-67,86-90): hidden-state search, no environment below the root */
    int32_t use_gumbel;         /* actor_use_gumbel (actor/gumbel_zero.cpp) */
    int32_t gumbel_noise;       /* actor_use_gumbel_noise && !actor_use_dirichlet_noise: supplied root noise is added to the logits (zero_actor.cpp:205-211) */
    int32_t gumbel_sample_size; /* actor_gumbel_sample_size */
    float gumbel_sigma_visit_c; /* actor_gumbel_sigma_visit_c */
    float gumbel_sigma_scale_c; /* actor_gumbel_sigma_scale_c */
    int32_t gomoku_exactly_five; /* env_gomoku_exactly_five_stones (reference default: true) */
    int32_t gomoku_outer_open;   /* env_gomoku_rule == "outer_open" */
    int32_t hex_swap_rule;       /* env_hex_use_swap_rule (reference default: true) */
    int32_t value_rescale;       /* actor_mcts_value_rescale: min-max normalised Q from the tree's value bounds (actor/mcts.cpp:43-49,219-228) */
    uint32_t atari_legal_mask;   /* MZ_GAME_ATARI: the game's minimal action set as a bit mask over the 18 actions (AtariEnv::isLegalAction, atari.h:57) */
    int32_t think_batch_size;    /* actor_mcts_think_batch_size (config/configuration.cpp:106; 0 or 1: off). K > 1 is the console search, ZeroActor::think /
                                  * ZeroActor::step (actor/zero_actor.cpp:36-49,129-157): every step selects K leaves of ONE tree one after the other under
                                  * virtual loss (actor/mcts.h:33-34, mcts.cpp:51-61,185), evaluates the distinct ones together and applies them in selection
                                  * order. num_games is then the number of trees; every per-game array of this API has num_games * K entries, lane-major
                                  * (index lane * num_games + tree): rotations, features, path lengths and network outputs per lane, while roots / moves /
                                  * noise use the first num_games entries (one per tree). PUCT selection on the board games, AlphaZero and MuZero networks (a MuZero root's
                                  * initial inference is a batch of one lane, zero_actor.cpp:134-135); Gumbel, Atari and value rescaling are refused. */
} mz_config;

/* Hyper-parameters the reference reads from the TorchScript module (network/network.cpp:30-41). */
typedef struct {
    int32_t num_input_channels, input_height, input_width;
    int32_t num_hidden_channels, num_blocks, action_size, num_value_hidden_channels, discrete_value_size;
    int32_t num_action_feature_channels; /* MuZero only (network/muzero_network.h:51); board games: 1 */
    int32_t is_muzero;                   /* get_type_name() == "muzero"; must match mz_config.muzero */
} mz_net_dims;

typedef struct {
    int32_t applied;   /* 1 = the action was legal and has been played (Environment::act return value) */
    int32_t terminal;  /* Environment::isTerminal() of the new position */
    int32_t num_legal; /* number of legal actions of the new position (= root children of the next search) */
    int32_t turn;      /* side to move: 1 = Black / first player, 2 = White */
    float eval_score;  /* Environment::getEvalScore(false) when terminal, else 0 */
} mz_play_result;

typedef struct {
    int32_t num_children;
    float count, mean, value; /* MCTSNode::getCount / getMean / getValue of the root */
} mz_root_info;

/* ---- lifetime -------------------------------------------------------------------------------------- */
/* replaces ActorGroup::initialize / createActors (actor/actor_group.cpp:150-187): allocates the node pools
 * (1 + (S+1)*A nodes per game, actor_group.cpp:183) and environments in HBM and resets every game. */
int mz_create(const mz_config* cfg, mz_engine** out);
void mz_destroy(mz_engine* e);
const char* mz_last_error(void);
int mz_action_size(const mz_engine* e);
int mz_num_features(const mz_engine* e); /* C * H * W of one position */

/* ---- network (replaces Network::loadModel + AlphaZeroNetwork, network/network.cpp:14-42,
 *      network/alphazero_network.h:48-104). The host reads the .pt with libtorch / torch.jit and hands over
 *      the hyper-parameters and every state_dict tensor by name (fp32, contiguous). ---------------------- */
int mz_net_configure(mz_engine* e, const mz_net_dims* dims);
int mz_net_set_tensor(mz_engine* e, const char* state_dict_name, const float* data, int64_t numel);
/* folds BatchNorm (eval mode, eps 1e-5) into the convolutions, converts to the kernels' fp16 layout and
 * uploads one packed blob. */
int mz_net_finalize(mz_engine* e);
/* the packed device blob, for an NCCL broadcast from the rank that read the .pt */
int mz_net_blob(mz_engine* e, void** device_ptr, int64_t* bytes);
/* ranks that did not read the .pt: allocate the blob from the dims alone, then receive it */
int mz_net_finalize_empty(mz_engine* e);
/* AlphaZeroNetwork::pushBack x n + forward(): features [n][C][H][W] fp32 (n <= num_games) ->
 * policy [n][A], policy_logit [n][A], value [n] */
int mz_eval_batch(mz_engine* e, const float* features, int32_t n, float* policy, float* logits, float* value);
/* MuZeroNetwork::pushBackInitialData x n + initialInference() (network/muzero_network.h:64-76,95-103; module
 * network/py/muzero_network.py:136-142): features [n][C][H][W] -> policy, policy_logit [n][A], value [n],
 * hidden_state [n][Ch][H][W] fp32 (scaled to [0, 1]); any output may be NULL */
int mz_eval_initial(mz_engine* e, const float* features, int32_t n, float* policy, float* logits, float* value, float* hidden_out);
/* MuZeroNetwork::pushBackRecurrentData x n + recurrentInference() (network/muzero_network.h:78-93,105-116; module
 * muzero_network.py:144-150): hidden [n][Ch][H][W] fp32 and the action ids whose planes Environment::getActionFeatures
 * would produce (one-hot cell, all zero for a pass; environment/othello/othello.cpp:257-262) */
int mz_eval_recurrent(mz_engine* e, const float* hidden, const int32_t* actions, int32_t n, float* policy, float* logits, float* value, float* hidden_out);

/* reward head output of the last mz_eval_recurrent / mz_eval_initial (muzero_atari networks; after the expectation over the
 * bins and utils::invertValue, network/muzero_network.h:165-171): reward [n] */
int mz_eval_rewards(mz_engine* e, int32_t n, float* reward);

/* ---- games ----------------------------------------------------------------------------------------- */
/* BaseActor::reset (actor/base_actor.cpp:8-13); g < 0 resets every game */
int mz_reset_game(mz_engine* e, int32_t g);
/* BaseActor::act + resetSearch for every game with actions[g] >= 0 (actor/base_actor.cpp:22-30,
 * actor/actor_group.cpp:116-134); results [num_games] */
int mz_play(mz_engine* e, const int32_t* actions, mz_play_result* results);
/* actor_select_action_by_count=true decided on the device: MCTS::selectChildByMaxCount at every root
 * (actor/mcts.cpp:91-104) followed by mz_play of that action; auto_reset != 0 also restarts finished games in place.
 * actions_out / results [num_games] may both be NULL, in which case the call is asynchronous (no host read-back). */
int mz_play_max_count(mz_engine* e, int32_t auto_reset, int32_t* actions_out, mz_play_result* results);
/* MZ_GAME_ATARI: what the host's emulator answered. For every game, actions[g] >= 0: AtariEnv::act's history update (the screen
 * joins the 8-entry observation history with the action that produced it, environment/atari/atari.cpp:82-85; call after mz_play);
 * -1: the initial screen after a reset (atari.cpp:52-56); -2: nothing for this game. frames [num_games][3][96][96]: the screen
 * resized to 96 x 96, RGB bytes, channel-major (AtariEnv::getObservation, atari.cpp:136-160) */
int mz_atari_observe(mz_engine* e, const int32_t* actions, const uint8_t* frames);
/* beside mz_get_roots: MCTSNode::getReward of the root children [B][A] and the tree's value bounds (number of distinct values,
 * smallest, largest: MCTS::getTreeValueBound, actor/mcts.h:106) [B]; what the move choice and the resign test of a rescaled search
 * read (actor/mcts.cpp:40-53,85-124). Any pointer may be NULL */
int mz_get_root_rewards(mz_engine* e, float* reward, int32_t* bound_size, float* bound_lo, float* bound_hi);
/* root child tables, children in stored (policy-sorted) order; any array may be NULL.
 * info [B]; the others [B][A] (MCTSNode getters, actor/mcts.h:44-52) */
int mz_get_roots(mz_engine* e, mz_root_info* info, int32_t* action, float* count, float* mean, float* policy, float* logit, float* noise,
                 float* value);

/* ---- learner data path -------------------------------------------------------------- */
/* BaseEnvLoader::getFeatures (environment/base/base_env.h:235-241; called per training sample by
 * learner/data_loader.cpp:134-200): for n <= num_games samples, replay the first positions[i] actions of record i (actions
 * [n][max_len], -1 padded, players alternate from the first player as Environment::act applies them) from the initial position and
 * return the feature planes of the position reached under rotations[i] (NULL = none): features_out [n][C*H*W]. Board games only:
 * Atari records carry their observations. The engine's search state is not touched (its network input rows are). */
int mz_replay_features(mz_engine* e, const int32_t* actions, int32_t max_len, const int32_t* positions, const uint8_t* rotations, int32_t n, float* features_out);

/* ---- search, one phase at a time (parity hooks; NN outputs supplied by the caller) ------------------ */
/* ZeroActor::beforeNNEvaluation for every game (actor/zero_actor.cpp:51-58). rotations [B] or NULL;
 * features_out [B][C*H*W] or NULL; path_len_out [B] or NULL */
int mz_search_select(mz_engine* e, const uint8_t* rotations, float* features_out, int32_t* path_len_out);
/* think mode: path_len_out per lane is > 0 for a leaf to evaluate, < 0 (-length) for a leaf an earlier lane of this step already selected
 * (zero_actor.cpp:140-142: it is not evaluated again, its planes are not produced), 0 for a lane beyond the simulations left (:133-135) */
/* ZeroActor::afterNNEvaluation for every game (actor/zero_actor.cpp:74-98). policy/logits [B][A], value [B],
 * noise [B][A] by root child index or NULL (Dirichlet values drawn by the host, utils/random.h:15-24) */
int mz_search_apply(mz_engine* e, const float* policy, const float* logits, const float* value, const float* noise);
/* the same with the reward head's outputs [B] (MuZeroNetworkOutput::reward_, zero_actor.cpp:88); reward may be NULL (= 0) */
int mz_search_apply_reward(mz_engine* e, const float* policy, const float* logits, const float* value, const float* reward, const float* noise);

/* MuZero / Gumbel parity hooks after mz_search_select: for every game the evaluation slot of the leaf's parent (-1 at the root;
 * the hidden state the recurrent inference reads, zero_actor.cpp:62-66), the leaf's action id (-1 at the root) and the action ids
 * along the selected path ([B][S + 2], -1 padded; path_actions may be NULL) */
int mz_search_leaf(mz_engine* e, int32_t* parent_slot, int32_t* leaf_action, int32_t* path_actions);
/* GumbelZero::decideActionNode with actor_select_action_by_count (gumbel_zero.cpp:61-66) for every game: action ids [B] */
int mz_gumbel_best_actions(mz_engine* e, int32_t* actions_out);

/* ---- search, whole move on the device --------------------------------------------------------------- */
/* host-drawn randomness of one search: rotations [(S+1)][B] (cycle-major) or NULL, root noise [B][A] or NULL */
int mz_search_set_inputs(mz_engine* e, const uint8_t* rotations, const float* noise);
/* runs num_evals (<= S+1; 0 = S+1) cycles of select -> features -> network -> expand/backup for all games as
 * one CUDA graph (the ActorGroup::run loop, actor/actor_group.cpp:136-148, without its host round trips).
 * device_ms receives the CUDA-event time of the graph on the engine's stream (the call then waits for it);
 * with device_ms == NULL the graph is only enqueued. */
int mz_search_run(mz_engine* e, int32_t num_evals, float* device_ms);

/* device_ms == NULL makes mz_search_run asynchronous; these bracket any sequence of calls with CUDA events on the
 * engine's stream (mz_timer_end and mz_sync wait for the stream). */
int mz_sync(mz_engine* e);
int mz_timer_begin(mz_engine* e);
int mz_timer_end(mz_engine* e, float* device_ms);

/* ---- measurement hooks ------------------------------------------------------------------------------ */
/* average device time (CUDA events on the engine's stream) of `iters` back-to-back launches of: the tower
 * conv kernel (one launch = mz_conv_layers_per_launch layers), the tree step kernel, the heads kernel; any pointer may be NULL */
int mz_profile_kernels(mz_engine* e, int32_t iters, float* conv_ms, float* tree_ms, float* heads_ms);
/* per-game cycle counters accumulated by every tree step since the last call, when the engine was created with the
 * environment variable MZ_DEBUG_TREE=1 (profiling only): out [num_games][16] = cycles in {selection, transition, leaf
 * analysis, feature planes, expand + backup}, number of steps, longest path, sum of path lengths, then selection detail:
 * cycles in {guess checking, hint chasing, serial finish}, check rounds, levels checked, levels finished serially, 2 unused */
int mz_debug_tree_timing(mz_engine* e, uint64_t* out);
/* per-CTA cycle counters of one launch of the fused tower kernel (engine created with MZ_DEBUG_TOWER=1): out [max_ctas][8] =
 * {producer total, producer waiting for the previous layer's groups, producer waiting for a free weight stage,
 *  MMA total, MMA waiting for the input block, for a free accumulator, for weights, epilogue busy}; returns the CTAs written */
int mz_debug_tower_timing(mz_engine* e, uint64_t* out, int32_t max_ctas);
/* how many 3x3 conv layers one launch of the conv kernel covers: all of them for the fused tower (then the conv time of
 * mz_profile_kernels is the whole tower), else 1 */
int mz_conv_layers_per_launch(const mz_engine* e);
/* kernels launched by this engine so far */
int64_t mz_launch_count(const mz_engine* e);
/* think mode: batched steps (network forwards) the last mz_search_run took to bring every tree to S + 1 simulations (zero_actor.cpp:39-44) */
int mz_think_steps(const mz_engine* e);
/* The fused conv tower's CTAs wait for each other through completion counters, so its whole grid must be resident. By default it is an ordinary
 * cluster launch: mz_net_finalize checks (cudaOccupancyMaxActiveClusters) that the device / context can hold the grid, the engine's kernels are
 * stream-ordered, and a wait that never ends traps after 10 s instead of hanging. Where OTHER work may compete for the SMs (a second engine on the
 * same device, another process under MPS) switch the cooperative launch on: the driver then guarantees co-residency or refuses the launch. It is
 * not the default because Nsight Compute fails every cooperative cluster launch of this kernel (LaunchFailed under ncu 2025.2 on B200, also for a
 * 2-CTA grid), which would make the library unprofilable. Results do not depend on the mode; captured search graphs are rebuilt.
 * mz_tower_is_cooperative: 1 / 0. mz_set_tower_cooperative(e, 1) probes with one launch and fails (leaving the mode off) if the device refuses. */
int mz_tower_is_cooperative(const mz_engine* e);
/* 1 when this engine's network runs through conv_tower_wide_kernel (two row tiles per CTA: chosen when a layer holds at least two such units per CTA pair,
 * e.g. 19x19 x 128 boards x 256 channels), 0 for conv_tower_kernel; a position's outputs are bit-identical either way */
int mz_tower_is_wide(const mz_engine* e);
int mz_set_tower_cooperative(mz_engine* e, int32_t on);

#ifdef __cplusplus
}
#endif
#endif
