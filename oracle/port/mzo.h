/*
 * mzo — CPU restatement ("port") of MiniZero's batched-MCTS self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This is the parity oracle for the CUDA path in minizero_b200/csrc.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product
 * path never links, imports or executes anything under oracle/.
 *
 * Every function cites the reference file:line (relative to /root/reference/minizero) whose
 * behaviour it restates. Pinned against outputs of the compiled reference itself
 * (oracle/_ref/ref_stepper_*, fixtures under tests/golden/, generator oracle/gen_golden.py).
 */
#ifndef MZO_H
#define MZO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZO_GAME_TICTACTOE 0
#define MZO_GAME_GO 1
#define MZO_GAME_OTHELLO 2
#define MZO_GAME_NOGO 3
#define MZO_GAME_GOMOKU 4
#define MZO_GAME_HEX 5
#define MZO_GAME_ATARI 6 /* environment/atari: one player, 18 actions, host-side emulator; MuZero only (6 x 6 hidden state) */
#define MZO_GAME_KILLALLGO 7 /* environment/killallgo: GoEnv on 7 x 7 with its own legality in the opening, terminal test (Benson) and result */
#define MZO_ATARI_RES 96
#define MZO_ATARI_HIST 8
#define MZO_HEX_SWAP_RULE 4       /* env_hex_use_swap_rule (default true); shares the flags word with the Gomoku options */
#define MZO_GOMOKU_EXACTLY_FIVE 1 /* env_gomoku_exactly_five_stones (default true) */
#define MZO_GOMOKU_OUTER_OPEN 2   /* env_gomoku_rule == "outer_open" */

#define MZO_MAX_N 19
#define MZO_MAX_CELLS (MZO_MAX_N * MZO_MAX_N)
#define MZO_MAX_ACTIONS (MZO_MAX_CELLS + 1)
#define MZO_MAX_MOVES (2 * MZO_MAX_CELLS + 2)
#define MZO_HIST 8

typedef struct {
    int32_t game;           /* MZO_GAME_* */
    int32_t board_size;     /* 3 for tictactoe, 2..19 for go */
    int32_t num_games;      /* B */
    int32_t num_simulation; /* actor_num_simulation (S); a search is S+1 evaluations */
    float puct_base;        /* actor_mcts_puct_base */
    float puct_init;        /* actor_mcts_puct_init */
    float reward_discount;  /* actor_mcts_reward_discount */
    float komi;             /* env_go_komi */
    int32_t ko_situational; /* env_go_ko_rule == "situational" */
    int32_t value_rescale;  /* actor_mcts_value_rescale */
    float dirichlet_epsilon; /* actor_dirichlet_noise_epsilon (used when noise is supplied) */
    int32_t muzero;          /* nn_type_name == "muzero": no environment below the root (zero_actor.cpp:59-67,86-90) */
    int32_t use_gumbel;      /* actor_use_gumbel */
    int32_t gumbel_noise;    /* actor_use_gumbel_noise: supplied noise is added to the root children's logits (zero_actor.cpp:205-211) */
    int32_t gumbel_sample_size; /* actor_gumbel_sample_size */
    float gumbel_sigma_visit_c; /* actor_gumbel_sigma_visit_c */
    float gumbel_sigma_scale_c; /* actor_gumbel_sigma_scale_c */
    int32_t gomoku_flags;       /* MZO_GOMOKU_* */
    uint32_t atari_legal_mask;  /* Atari: ALE minimal action set of the game as a bit mask over the 18 actions (atari.h:57) */
} mzo_config;

/* ---- environment (environment/go/go.cpp, environment/tictactoe/tictactoe.cpp) ---- */
typedef struct {
    int32_t game, n, turn, num_moves;
    int32_t flags; /* MZO_GOMOKU_* */
    float komi;
    uint64_t turn_key, hash;
    uint8_t board[MZO_MAX_CELLS];
    uint8_t hist[MZO_HIST][MZO_MAX_CELLS]; /* ring: position after move i lives in hist[i % 8] */
    int16_t actions[MZO_MAX_MOVES];
    uint64_t hashes[MZO_MAX_MOVES];
} mzo_env;

void mzo_env_init(mzo_env* e, int game, int n, float komi, int ko_situational);
void mzo_env_set_flags(mzo_env* e, int flags);
int mzo_env_num_actions(const mzo_env* e);
int mzo_env_input_channels(const mzo_env* e);
int mzo_env_is_legal(const mzo_env* e, int action, int player);
int mzo_env_act(mzo_env* e, int action, int player);
int mzo_env_is_terminal(const mzo_env* e);
float mzo_env_eval_score(const mzo_env* e, int is_resign);
void mzo_env_features(const mzo_env* e, int rotation, float* out);
void mzo_env_action_features(const mzo_env* e, int action, float* out);
int mzo_rotate_position(int rotation, int pos, int n);
int mzo_reversed_rotation(int rotation);
uint64_t mzo_go_key(int pos, int player);
uint64_t mzo_mt19937_64_nth(uint64_t seed, int nth);

/* ---- batched search (actor/mcts.cpp, actor/zero_actor.cpp) ---- */
typedef struct mzo_batch mzo_batch;

typedef struct {
    int32_t num_children;
    float count, mean, value;
    int32_t action[MZO_MAX_ACTIONS];
    float c_count[MZO_MAX_ACTIONS];
    float c_mean[MZO_MAX_ACTIONS];
    float c_policy[MZO_MAX_ACTIONS];
    float c_logit[MZO_MAX_ACTIONS];
    float c_noise[MZO_MAX_ACTIONS];
    float c_value[MZO_MAX_ACTIONS];
} mzo_root_out;

mzo_batch* mzo_create(const mzo_config* cfg);
void mzo_destroy(mzo_batch* b);
void mzo_reset_game(mzo_batch* b, int g);
/* ZeroActor::resetSearch for every game */
void mzo_reset_search(mzo_batch* b, int g);
/* ZeroActor::beforeNNEvaluation for every game: features [B][C*H*W], rotations [B] (NULL = none) */
void mzo_select(mzo_batch* b, const uint8_t* rotations, float* features);
/* ZeroActor::afterNNEvaluation for every game; noise [B][A] by child index (NULL = no noise) */
void mzo_apply(mzo_batch* b, const float* policy, const float* logits, const float* value, const float* noise);
int mzo_num_simulation_done(const mzo_batch* b, int g);
int mzo_path_len(const mzo_batch* b, int g);
void mzo_root(const mzo_batch* b, int g, mzo_root_out* out);
const mzo_env* mzo_root_env(const mzo_batch* b, int g);
/* BaseActor::act on the root environment; returns 1 if the move was legal and applied */
int mzo_play(mzo_batch* b, int g, int action);
int mzo_select_by_max_count(const mzo_batch* b, int g);
/* MuZero: action id of the leaf chosen by the last mzo_select (-1 for the root) and of its parent's evaluation slot
 * (the simulation index whose hidden state feeds the dynamics network, zero_actor.cpp:62-66) */
int mzo_leaf_action(const mzo_batch* b, int g);
int mzo_leaf_parent_slot(const mzo_batch* b, int g);
/* FNV-1a over the action ids of the selected path (same formula as oracle/drivers/ref_stepper.cpp) */
int mzo_path_hash(const mzo_batch* b, int g);
/* GumbelZero::decideActionNode with actor_select_action_by_count (gumbel_zero.cpp:61-66): action id */
int mzo_gumbel_best_action(mzo_batch* b, int g);
/* GumbelZero::getMCTSPolicy (gumbel_zero.cpp:9-59): fills action ids / probabilities of the entries the reference prints
 * (completed-Q softmax, entries below -38 dropped) in ascending action id; returns how many */
int mzo_gumbel_policy(const mzo_batch* b, int g, int32_t* actions, float* probs);

/* ---- Atari MuZero (BASELINE configs[4]) ----
 * mzo_apply with the reward head's output (muzero_network.h:165-171; zero_actor.cpp:88); reward may be NULL (= 0) */
void mzo_apply_mz(mzo_batch* b, const float* policy, const float* logits, const float* value, const float* reward, const float* noise);
/* rewards of the root children and the value bounds (mcts.h:106): what the move choice / resign test read beside mzo_root */
void mzo_root_extra(const mzo_batch* b, int g, float* c_reward, int32_t* bound_size, float* bound_lo, float* bound_hi);
/* MCTSNode::getNormalizedMean of root child i (i < 0: the root itself) with the tree's value bounds (mcts.cpp:40-53) */
float mzo_root_normalized_mean(const mzo_batch* b, int g, int i);
/* AtariEnv: reset() (action < 0: the initial screen) or act() (atari.cpp:47-93): frame = resized screen, RGB bytes [3][96][96] */
void mzo_atari_observe(mzo_batch* b, int g, int action, const uint8_t* frame_chw, int terminal);

/* ---- console think() with actor_mcts_think_batch_size = K > 1 (zero_actor.cpp:36-49,129-157; AlphaZero networks) ----
 * One ZeroActor::step per call pair. mzo_think_select: min(K, simulations left) selections per tree under virtual loss; lane-major arrays:
 * rotations [K][B], features [K][B][C*H*W] (every selection's position, used or not), path_len [K][B] (> 0: leaf to evaluate, < 0: duplicate
 * of an earlier lane, 0: lane unused). mzo_think_apply: the evaluated lanes in selection order; policy / logits [K][B][A], value [K][B]. */
void mzo_think_select(mzo_batch* b, int K, const uint8_t* rotations, float* features, int32_t* path_len);
void mzo_think_apply(mzo_batch* b, const float* policy, const float* logits, const float* value, const float* noise);
void mzo_think_leaf(const mzo_batch* b, int k, int g, int32_t* parent_slot, int32_t* action);

/* std::sort(candidates, policy descending) exactly as libstdc++ orders them, ties included (mzo_sort.c) */
void mzo_std_sort_candidates(int n, int32_t* action, float* policy, float* logit);

/* ---- network forward, fp32 (network/py/alphazero_network.py, network_unit.py) ---- */
typedef struct mzo_net mzo_net;
mzo_net* mzo_net_create(int c, int h, int w, int hidden, int blocks, int actions, int value_hidden);
void mzo_net_destroy(mzo_net* n);
/* name = state_dict key; returns 0 on success */
int mzo_net_set(mzo_net* n, const char* name, const float* data, int64_t numel);
void mzo_net_forward(const mzo_net* n, const float* features, int batch, float* policy, float* logits, float* value);

#ifdef __cplusplus
}
#endif
#endif
