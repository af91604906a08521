// Search core: PUCT / Gumbel selection, leaf environment transition (Go, NoGo, Othello, Gomoku, TicTacToe rules on
// row bitboards), feature packing, expansion and backup over a flat node pool; AlphaZero and MuZero branches.
//
// One thread block owns one game (tree + environment): warp-collective routines for one level / one flood fill, block-wide
// ones for the leaf analysis and the level-parallel re-evaluation of a path. All functions are written against a tiny warp
// abstraction (MZ_W lanes, ballot / any / butterfly reductions, lane-strided loops, flood fills
// iterated to their unique fixpoint) so that the same source also compiles as plain C++ with
// MZ_W == 1. That second build (tests/hostsim) exists ONLY so the CPU test-suite can check this
// file's logic against the recordings of the reference before GPU time is spent; it is never
// linked into libmzb200.so, whose entry points fail loudly without a CUDA device.
//
// Reference behaviour restated here (paths relative to /root/reference/minizero):
//   actor/mcts.cpp:20-28,40-61,139-217   node update, normalised mean, PUCT score, selection, init-Q
//   actor/mcts.cpp:151-179               expand, backup
//   actor/zero_actor.cpp:51-98           beforeNNEvaluation / afterNNEvaluation (AlphaZero branch)
//   actor/zero_actor.cpp:194-229,247-252 root noise mix, legal-filtered sorted candidates, env transition
//   environment/go/go.cpp:132-308,690-723  act, isLegalAction (superko), isTerminal, Tromp-Taylor, features
//   environment/tictactoe/tictactoe.cpp:19-146
//   environment/othello/othello.cpp:13-262 act (flips), legal boards, pass, terminal, score, features
//   actor/zero_actor.cpp:59-67,86-90,231-245  MuZero branch (no environment below the root, all actions expanded)
//   actor/gumbel_zero.cpp:61-137           Gumbel: candidate selection, sequential halving, score order
//   utils/rotation.h:22-93
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(MZ_HOSTSIM)
#define MZ_DEV __device__ __forceinline__
#define MZ_W 32
#define MZ_FULL 0xffffffffu
MZ_DEV long long mz_clock() { return clock64(); }
MZ_DEV void mz_block_sync() { __syncthreads(); }
MZ_DEV void mz_atomic_min(int* p, int v) { atomicMin(p, v); }
MZ_DEV void mz_atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
MZ_DEV void mz_atomic_inc(int* p) { atomicAdd(p, 1); }
MZ_DEV void mz_atomic_xor64(uint64_t* p, uint64_t v) { atomicXor(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v)); }
MZ_DEV unsigned mz_ballot(int p) { return __ballot_sync(MZ_FULL, p); }
MZ_DEV int mz_any(int p) { return __any_sync(MZ_FULL, p); }
MZ_DEV void mz_sync() { __syncwarp(); }
MZ_DEV float mz_fmul(float a, float b) { return __fmul_rn(a, b); }
MZ_DEV float mz_fadd(float a, float b) { return __fadd_rn(a, b); }
MZ_DEV float mz_fsub(float a, float b) { return __fsub_rn(a, b); }
MZ_DEV float mz_fdiv(float a, float b) { return __fdiv_rn(a, b); }
MZ_DEV double mz_dmul(double a, double b) { return __dmul_rn(a, b); }
MZ_DEV double mz_ddiv(double a, double b) { return __ddiv_rn(a, b); }
MZ_DEV double mz_dsqrt(double a) { return __dsqrt_rn(a); }
MZ_DEV int mz_popc(uint32_t x) { return __popc(x); }
MZ_DEV int mz_ffs0(uint32_t x) { return __ffs(x) - 1; }
MZ_DEV int mz_reduce_add(int v)
{
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(MZ_FULL, v, o); }
    return v;
}
MZ_DEV uint32_t mz_reduce_or(uint32_t v)
{
    for (int o = 16; o > 0; o >>= 1) { v |= __shfl_xor_sync(MZ_FULL, v, o); }
    return v;
}
MZ_DEV int mz_reduce_min(int v)
{
    for (int o = 16; o > 0; o >>= 1) { v = min(v, __shfl_xor_sync(MZ_FULL, v, o)); }
    return v;
}
MZ_DEV uint64_t mz_reduce_xor64(uint64_t v)
{
    for (int o = 16; o > 0; o >>= 1) { v ^= __shfl_xor_sync(MZ_FULL, v, o); }
    return v;
}
MZ_DEV uint32_t mz_redux_max(uint32_t v) { return __reduce_max_sync(MZ_FULL, v); }
MZ_DEV uint32_t mz_redux_min(uint32_t v) { return __reduce_min_sync(MZ_FULL, v); }
MZ_DEV void mz_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
MZ_DEV uint32_t mz_float_bits(float f) { return __float_as_uint(f); }
// lexicographic arg-max over (score desc, policy desc, index asc): the order-independent form of
// the serial scan in mcts.cpp:187-194 (replace when score > best, or score == best and policy > best)
MZ_DEV void mz_reduce_best(float& s, float& p, int& i)
{
    for (int o = 16; o > 0; o >>= 1) {
        float s2 = __shfl_xor_sync(MZ_FULL, s, o), p2 = __shfl_xor_sync(MZ_FULL, p, o);
        int i2 = __shfl_xor_sync(MZ_FULL, i, o);
        bool take = (i2 >= 0) && (i < 0 || s2 > s || (s2 == s && (p2 > p || (p2 == p && i2 < i))));
        if (take) { s = s2, p = p2, i = i2; }
    }
}
struct __align__(16) mz_hot {
    float count, mean, policy;
    uint32_t link;
};
MZ_DEV mz_hot mz_load_hot(const mz_hot* p)
{
    float4 v = *reinterpret_cast<const float4*>(p);
    mz_hot h;
    h.count = v.x, h.mean = v.y, h.policy = v.z, h.link = __float_as_uint(v.w);
    return h;
}
MZ_DEV void mz_store_hot(mz_hot* p, float count, float mean, float policy, uint32_t link)
{
    *reinterpret_cast<float4*>(p) = make_float4(count, mean, policy, __uint_as_float(link));
}
#else
#include <math.h>
#define MZ_DEV static inline
#define MZ_W 1
static inline long long mz_clock() { return 0; }
static inline void mz_block_sync() {}
static inline void mz_atomic_min(int* p, int v) { *p = (v < *p ? v : *p); }
static inline void mz_atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
static inline void mz_atomic_inc(int* p) { ++*p; }
static inline void mz_atomic_xor64(uint64_t* p, uint64_t v) { *p ^= v; }
static inline unsigned mz_ballot(int p) { return p ? 1u : 0u; }
static inline int mz_any(int p) { return p; }
static inline void mz_sync() {}
static inline float mz_fmul(float a, float b) { return a * b; }
static inline float mz_fadd(float a, float b) { return a + b; }
static inline float mz_fsub(float a, float b) { return a - b; }
static inline float mz_fdiv(float a, float b) { return a / b; }
static inline double mz_dmul(double a, double b) { return a * b; }
static inline double mz_ddiv(double a, double b) { return a / b; }
static inline double mz_dsqrt(double a) { return sqrt(a); }
static inline int mz_popc(uint32_t x) { return __builtin_popcount(x); }
static inline int mz_ffs0(uint32_t x) { return __builtin_ffs((int)x) - 1; }
static inline int mz_reduce_add(int v) { return v; }
static inline uint32_t mz_reduce_or(uint32_t v) { return v; }
static inline int mz_reduce_min(int v) { return v; }
static inline uint64_t mz_reduce_xor64(uint64_t v) { return v; }
static inline uint32_t mz_redux_max(uint32_t v) { return v; }
static inline uint32_t mz_redux_min(uint32_t v) { return v; }
static inline void mz_prefetch(const void*) {}
static inline uint32_t mz_float_bits(float f)
{
    uint32_t u;
    __builtin_memcpy(&u, &f, 4);
    return u;
}
static inline void mz_reduce_best(float&, float&, int&) {}
struct alignas(16) mz_hot {
    float count, mean, policy;
    uint32_t link;
};
static inline mz_hot mz_load_hot(const mz_hot* p) { return *p; }
static inline void mz_store_hot(mz_hot* p, float count, float mean, float policy, uint32_t link)
{
    p->count = count, p->mean = mean, p->policy = policy, p->link = link;
}
#endif

// visited children of a node, in child order: what selection below the root has to look at (the unvisited children are
// represented by the first of them, see mz_select_level). n > MZ_VIS_MAX: list overflowed, scan the children instead.
#define MZ_VIS_MAX 6
struct alignas(16) mz_vis {
    uint16_t idx[MZ_VIS_MAX]; // child indices (relative to first_child), ascending
    uint16_t n;               // number of visited children
    uint16_t fu;              // index of the first unvisited child (== num_children when all are visited)
};
#if defined(__CUDACC__) && !defined(MZ_HOSTSIM)
MZ_DEV mz_vis mz_load_vis(const mz_vis* p)
{
    union {
        uint4 u;
        mz_vis v;
    } x;
    x.u = *reinterpret_cast<const uint4*>(p);
    return x.v;
}
MZ_DEV void mz_store_vis(mz_vis* p, const mz_vis& v)
{
    union {
        uint4 u;
        mz_vis v;
    } x;
    x.v = v;
    *reinterpret_cast<uint4*>(p) = x.u;
}
#else
static inline mz_vis mz_load_vis(const mz_vis* p) { return *p; }
static inline void mz_store_vis(mz_vis* p, const mz_vis& v) { *p = v; }
#endif

#include "killallgo_rules.h"

#define MZ_GAME_TICTACTOE 0
#define MZ_GAME_GO 1
#define MZ_GAME_OTHELLO 2
#define MZ_GAME_NOGO 3       // environment/nogo/nogo.h: GoEnv with its own legality, terminal test and result
#define MZ_GAME_GOMOKU 4     // environment/gomoku: N x N, no pass, five in a row through the last move
#define MZ_GAME_HEX 5        // environment/hex: N x N, no pass, swap rule, connect the two own edges; never rotated
#define MZ_GAME_ATARI 6      // environment/atari: one player, 18 actions, the emulator is the host's; MuZero only (hidden state 6 x 6)
#define MZ_GAME_KILLALLGO 7  // environment/killallgo: GoEnv on 7 x 7 (mz_dims.game is MZ_GAME_GO with mz_dims.killall set: every Go rule applies) + its own opening
                             // legality, terminal test (Benson's unconditional life) and result
#define MZ_ATARI_RES 96      // kAtariResolution, atari.h:24
#define MZ_ATARI_FRAME (3 * MZ_ATARI_RES * MZ_ATARI_RES)
#define MZ_GO_FAMILY(game) ((game) == MZ_GAME_GO || (game) == MZ_GAME_NOGO)
#define MZ_GUMBEL_LEVELS 12  // halvings of actor_gumbel_sample_size that can ever happen (m <= 362)
#define MZ_MAXN 19
#define MZ_MAXA (MZ_MAXN * MZ_MAXN + 1)
#define MZ_ROWS 32           // row slots per bitboard (N <= 19 used)
#define MZ_HIST 8            // positions kept for the feature planes (go.cpp:291-299)
#define MZ_LEGAL_WORDS 12    // ceil(362 / 32)
#define MZ_LINK_SHIFT 20     // link = first_child | num_children << 20
#define MZ_NN_CPAD 64        // input channels of the first conv, padded to one K block
#define MZ_HALF_ONE 0x3C00   // fp16 1.0
#define MZ_SEL_AHEAD 3       // chunks of 32 children fetched together by a selection level (3 covers the 82 actions of 9x9 Go)

struct mz_dims {
    int game, N, A, C, S, NP, B;
    int slots;      // NN rows per board: (N + 1) * (N + 1)   (one shared zero column / row, see nn_conv.cu)
    int max_hashes; // per game: 2 * N * N + 4
    float puct_init, puct_base, discount, komi, eps;
    uint64_t turn_key; // go.cpp:45-49 (0 unless situational superko)
    // MuZero (nn_type_name == "muzero"): no environment below the root; every evaluated node keeps its hidden state
    int muzero;
    int hid_c;   // channels per cell of a stored hidden state (the network's padded hidden width)
    int dyn_c;   // channels per row of the dynamics network's input: hidden state, then the action planes, zero padded
    int act_col; // column of the (single) action plane in a dynamics input row = num_hidden_channels
    // Gumbel (actor_use_gumbel, gumbel_zero.cpp)
    int hex_swap_rule; // env_hex_use_swap_rule
    int gomoku_exactly_five, gomoku_outer_open; // env_gomoku_exactly_five_stones, env_gomoku_rule == "outer_open"
    int gumbel, gumbel_noise, gumbel_m;
    float sigma_visit_c, sigma_scale_c;
    int gumbel_budget0;                   // max(1, floor(S / (log2(m) * m))), gumbel_zero.cpp:99 (host-computed in double)
    int gumbel_next[MZ_GUMBEL_LEVELS];
    // Atari MuZero (BASELINE configs[4])
    int num_players;    // 1 for Atari: the value is never flipped, every node is player 1's (atari.h:18, base_env.h getNextPlayer)
    int value_rescale;  // actor_mcts_value_rescale: Q is min-max normalised with the tree's value bounds (mcts.cpp:43-49)
    int atari_init_q;   // the `#if ATARI` branch of MCTS::calculateInitQValue (mcts.cpp:211-213)
    int has_reward;     // the network has a reward head (muzero_atari): nodes carry rewards, backup discounts through them
    int act_planes;     // action planes of the dynamics input: 1 (one-hot cell) or 18 (one-hot channel block, atari.cpp:124-130)
    uint32_t legal_mask; // Atari: minimal action set of the game (the root's legal actions, atari.h:57)
    int vb_cap;         // capacity of the value-bound table per game    // floor(S / (log2(m) * (m >> level) / 2)) in double, gumbel_zero.cpp:109
    int killall;        // KillAllGo on top of the Go rules (game == MZ_GAME_GO, N == 7): killallgo.cpp:27-48 with env_killallgo_use_seki = false
    int think_k;        // actor_mcts_think_batch_size when > 1 (console think(), zero_actor.cpp:129-157): leaves selected per tree and network forward
};

struct mz_state {
    // node pool, per game NP entries
    mz_hot* hot;
    int16_t* action;
    float* logit;
    float* value;
    float* root_noise; // [B][A] policy_noise_ of the root children
    int32_t* cursor;   // [B]
    int16_t* node_slot; // [B][NP] index of the cached environment of an evaluated node (-1: none)
    int32_t* last_child; // [B][NP] child chosen the last time selection passed through the node (-1: never): speculation hint only
    mz_vis* vis;         // [B][NP] visited-children list of every expanded node (selection accelerator; null = always scan)
    // environment of every evaluated node of the current search, slot = simulation index (0 .. S)
    uint32_t* slot_st;   // [B][S + 1][2][N] stone rows
    uint64_t* slot_hash; // [B][S + 1]
    int32_t* slot_meta;  // [B][S + 1][4] turn, num_moves, last action, action before last
    // root environment
    uint32_t* root_st;   // [B][2][MZ_ROWS]
    uint32_t* root_hist; // [B][MZ_HIST][2][MZ_ROWS]
    uint64_t* root_hash; // [B]
    int32_t* root_meta;  // [B][4] turn, num_moves, last action, action before last
    uint64_t* hashes;    // [B][max_hashes] position hashes of the game so far (superko history of the root)
    // leaf of the current simulation
    int32_t* path;        // [B][S + 2]
    int32_t* path_len;    // [B]
    int32_t* spec_len;    // [B] length of the previous simulation's path while it is still valid for speculation (0: none)
    uint32_t* leaf_legal; // [B][MZ_LEGAL_WORDS]
    int32_t* leaf_meta;   // [B][4] terminal, turn, rotation, num_legal
    float* leaf_score;    // [B]
    // network io
    uint16_t* nn_in; // [B * slots][MZ_NN_CPAD] fp16 bits, NHWC rows; only 0 / 1.0 are ever written
    float* policy;   // [B][A]
    float* logits;   // [B][A]
    float* nn_value; // [B]
    // per-search inputs
    const uint8_t* rotations; // [B] for this cycle (may be null = identity)
    const float* noise_in;    // [B][A] by root child index (may be null)
    const float* puct_bias;   // [S + 2] host-computed: (float)(init + log((1 + n + base) / base)), mcts.cpp:57
    const double* sqrt_table; // [S + 2] sqrt((double)n): IEEE-exact, identical to the host's sqrt (mcts.cpp:58)
    const uint64_t* keys;     // [2][361] Zobrist stone keys, go.cpp:19-32
    unsigned long long* dbg;  // optional [B][16] per-phase cycle counters (profiling only)
    // MuZero: hidden states of the evaluated nodes of the current search, slot = simulation index (TreeData<HiddenStateData>, tree.h)
    uint16_t* hid;       // [B][S + 1][N * N][hid_c] fp16, written by the hidden-state scaling kernel (null when not MuZero / host build)
    uint16_t* dyn_in;    // [B * slots][dyn_c] fp16 rows of the dynamics network's input (parent hidden state + action plane)
    int32_t* eval_slot;  // [B] slot the hidden state of this cycle's evaluation goes to
    int32_t* leaf_parent; // [B][2] parent's slot (-1 for the root), leaf action: what the recurrent inference consumes (parity hook)
    // Gumbel: GumbelZero::candidates_ / sample_size_ / simulation_budget_ (gumbel_zero.h:20-23)
    int32_t* gum_cand;   // [B][A] node indices
    int32_t* gum_meta;   // [B][4] number of candidates, sample size, budget, halving level
    // rewards and value bounds (null unless has_reward / value_rescale)
    float* reward;       // [B][NP] MCTSNode::reward_
    float* nn_reward;    // [B] reward head output of this cycle's evaluation (after expectation + invertValue)
    float* vb_key;       // [B][vb_cap] MCTS::tree_value_bound_ (std::map<float, int>, mcts.h:117): distinct keys, unordered
    int32_t* vb_cnt;     // [B][vb_cap] their multiplicities
    int32_t* vb_n;       // [B] number of keys
    // Atari root environment: the last 8 screens and the actions that led to them (atari.cpp:47-93); ring, oldest entry at at_head
    uint8_t* at_frames;  // [B][8][3][96][96]
    int32_t* at_meta;    // [B][16]: [0] head, [1..8] action id per ring slot (-1: zero plane), [9] bit mask of slots holding a screen
    // console think() with actor_mcts_think_batch_size = K > 1 (zero_actor.cpp:129-157): K leaves of ONE tree are selected one after the other, each
    // selection leaving a virtual loss on its path; the K positions are evaluated together; the results are applied in selection order. A step runs as
    // K "before" passes and K "after" passes over a VIEW of this struct per lane: the per-leaf arrays (path, path_len, leaf_*, nn_in, policy, logits,
    // nn_value, rotations) are offset to the lane's section, everything belonging to the tree is shared.
    float* vloss;           // [B][NP] MCTSNode::virtual_loss_ (null unless think_k > 1)
    int32_t* think_pending; // [B] leaves of the current step that will be evaluated so far: their slots follow the finished simulations
    int think_lane;         // lane of this view, 0 .. K-1
};

// view of the state for lane k of a batched think() step over `trees` trees: the per-leaf arrays point at the lane's section
// (lane-major: index k * trees + tree), the tree's own arrays are shared by all lanes
static inline mz_state mz_lane_view(const mz_dims& d, const mz_state& s, int k, int trees)
{
    mz_state v = s;
    const size_t o = (size_t)k * trees;
    v.path += o * (d.S + 2), v.path_len += o, v.leaf_legal += o * MZ_LEGAL_WORDS, v.leaf_meta += o * 4, v.leaf_score += o;
    v.nn_in += o * d.slots * MZ_NN_CPAD, v.policy += o * d.A, v.logits += o * d.A, v.nn_value += o;
    if (v.dyn_in) { v.dyn_in += o * d.slots * d.dyn_c; } // MuZero: the lane's dynamics input rows, evaluation slot and parity-hook record;
    if (v.eval_slot) { v.eval_slot += o; }               // the hidden states themselves (hid) belong to the tree
    if (v.leaf_parent) { v.leaf_parent += o * 2; }
    if (v.rotations) { v.rotations += o; }
    v.think_lane = k;
    return v;
}

// value bounds of the tree being searched + the game's reward column: what MCTSNode::getNormalizedMean reads beside the node
struct mz_qb {
    const float* reward; // game-local rewards (null: all zero)
    float lo, hi;
    int n;               // number of distinct keys (< 2: Q is 1, mcts.cpp:44)
};

// per-warp scratch (shared memory on the device)
struct mz_scratch {
    uint32_t st[2][MZ_ROWS];
    uint32_t hist[MZ_HIST][2][MZ_ROWS];
    uint32_t fill[MZ_ROWS], checked[MZ_ROWS], tmp[MZ_ROWS], legal_rows[MZ_ROWS];
    uint64_t cap_hash[MZ_MAXA];
    float pol[MZ_MAXA], lg[MZ_MAXA], q[MZ_MAXA];
    uint32_t legal[MZ_LEGAL_WORDS];
    uint64_t hash;
    int turn, num_moves, last, last2;
    uint64_t* path_hashes; // [S + 2] position hashes of the nodes on the current path (shared memory on the device)
    int32_t* sel;          // [S + 2] child chosen at every level of the previous path by the speculative re-evaluation
    mz_hot* lvl_h;         // [S + 2] hot record of the node at every level of the guessed path
    mz_vis* lvl_v;         // [S + 2] its visited-children list
    float* q_warp;         // [num_warps][A] per-warp Q scratch of the level evaluation
    int mismatch;          // first level whose re-evaluated choice differs from the previous path
    // block-wide leaf analysis (mz_env_legal_block)
    int label[MZ_MAXN * MZ_MAXN];  // block id of a stone = smallest cell index of its block
    int libcnt[MZ_MAXN * MZ_MAXN]; // liberties per block id
    uint32_t bloom[64];            // 2048-bit filter over the superko history
    int flag, shared_len, shared_count;
    mz_qb qb;                      // value bounds during selection (uniform for the block)
};

MZ_DEV uint32_t mz_rowmask(int N) { return (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)); }

// utils/rotation.h:51-93 in doubled integer coordinates
MZ_DEV int mz_rotate(int rotation, int pos, int N)
{
    if (pos == N * N) { return pos; }
    int x = 2 * (pos % N) - (N - 1), y = 2 * (pos / N) - (N - 1), rx = x, ry = y;
    switch (rotation) {
        case 1: rx = y, ry = -x; break;
        case 2: rx = -x, ry = -y; break;
        case 3: rx = -y, ry = x; break;
        case 4: rx = x, ry = -y; break;
        case 5: rx = -y, ry = -x; break;
        case 6: rx = -x, ry = y; break;
        case 7: rx = y, ry = x; break;
        default: break;
    }
    return ((ry + (N - 1)) / 2) * N + (rx + (N - 1)) / 2;
}
MZ_DEV int mz_reversed_rotation(int r) { return (r == 1 ? 3 : (r == 3 ? 1 : r)); } // rotation.h:22-31

// ---------------------------------------------------------------------------------------------
// row-bitboard helpers: rows[r] bit x = cell (x, y = r); every routine is a warp collective
// ---------------------------------------------------------------------------------------------

// grow `f` through `mask` along the row until it stops changing
MZ_DEV uint32_t mz_hfill(uint32_t f, uint32_t mask)
{
    uint32_t prev;
    do {
        prev = f;
        f |= ((f << 1) | (f >> 1)) & mask;
    } while (f != prev);
    return f;
}

// flood `fill` (seeded by the caller, seeds inside `mask`) to the connected component(s) within mask
MZ_DEV void mz_flood(uint32_t* fill, const uint32_t* mask, int N, int lane)
{
    int changed;
    do {
        changed = 0;
        uint32_t g[(MZ_ROWS + MZ_W - 1) / MZ_W];
        int k = 0;
        for (int r = lane; r < N; r += MZ_W, ++k) {
            uint32_t f = fill[r];
            uint32_t v = f | (r > 0 ? fill[r - 1] : 0u) | (r + 1 < N ? fill[r + 1] : 0u);
            v = mz_hfill(v & mask[r], mask[r]);
            g[k] = v;
            changed |= (v != f);
        }
        mz_sync();
        k = 0;
        for (int r = lane; r < N; r += MZ_W, ++k) { fill[r] = g[k]; }
        mz_sync();
        changed = mz_any(changed);
    } while (changed);
}

// 4-neighbourhood dilation of rows `f` (including f itself), clipped to the board
MZ_DEV uint32_t mz_dilate_row(const uint32_t* f, int r, int N)
{
    uint32_t v = f[r] | (f[r] << 1) | (f[r] >> 1) | (r > 0 ? f[r - 1] : 0u) | (r + 1 < N ? f[r + 1] : 0u);
    return v & mz_rowmask(N);
}

MZ_DEV uint64_t mz_block_hash(const uint32_t* fill, int colour_idx, const uint64_t* keys, int N, int lane)
{
    uint64_t h = 0;
    for (int r = lane; r < N; r += MZ_W) {
        uint32_t m = fill[r];
        while (m) {
            int x = mz_ffs0(m);
            m &= m - 1;
            h ^= keys[colour_idx * 361 + r * N + x];
        }
    }
    return mz_reduce_xor64(h);
}

// ---------------------------------------------------------------------------------------------
// environment
// ---------------------------------------------------------------------------------------------

MZ_DEV void mz_env_load_root(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int lane)
{
    const uint32_t* st = s.root_st + (size_t)g * 2 * MZ_ROWS;
    const uint32_t* hist = s.root_hist + (size_t)g * MZ_HIST * 2 * MZ_ROWS;
    for (int i = lane; i < 2 * MZ_ROWS; i += MZ_W) { (&w->st[0][0])[i] = st[i]; }
    for (int i = lane; i < MZ_HIST * 2 * MZ_ROWS; i += MZ_W) { (&w->hist[0][0][0])[i] = hist[i]; }
    if (lane == 0) {
        w->hash = s.root_hash[g];
        w->turn = s.root_meta[g * 4 + 0];
        w->num_moves = s.root_meta[g * 4 + 1];
        w->last = s.root_meta[g * 4 + 2];
        w->last2 = s.root_meta[g * 4 + 3];
    }
    mz_sync();
}

MZ_DEV void mz_env_store_root(const mz_dims& d, const mz_state& s, int g, const mz_scratch* w, int lane)
{
    uint32_t* st = s.root_st + (size_t)g * 2 * MZ_ROWS;
    uint32_t* hist = s.root_hist + (size_t)g * MZ_HIST * 2 * MZ_ROWS;
    for (int i = lane; i < 2 * MZ_ROWS; i += MZ_W) { st[i] = (&w->st[0][0])[i]; }
    for (int i = lane; i < MZ_HIST * 2 * MZ_ROWS; i += MZ_W) { hist[i] = (&w->hist[0][0][0])[i]; }
    if (lane == 0) {
        s.root_hash[g] = w->hash;
        s.root_meta[g * 4 + 0] = w->turn;
        s.root_meta[g * 4 + 1] = w->num_moves;
        s.root_meta[g * 4 + 2] = w->last;
        s.root_meta[g * 4 + 3] = w->last2;
    }
}

MZ_DEV void mz_env_reset(const mz_dims& d, mz_scratch* w, int lane)
{
    for (int i = lane; i < 2 * MZ_ROWS; i += MZ_W) { (&w->st[0][0])[i] = 0u; }
    for (int i = lane; i < MZ_HIST * 2 * MZ_ROWS; i += MZ_W) { (&w->hist[0][0][0])[i] = 0u; }
    mz_sync();
    if (lane == 0 && d.game == MZ_GAME_OTHELLO) { // othello.cpp:22-27: Black (player 1) on init_place and its diagonal
        const int bs = d.N, ip = bs * (bs / 2 - (1 - bs % 2)) + (bs / 2 - 1);
        w->st[1][(ip + 1) / bs] |= 1u << ((ip + 1) % bs), w->st[1][(ip + bs) / bs] |= 1u << ((ip + bs) % bs);
        w->st[0][ip / bs] |= 1u << (ip % bs), w->st[0][(ip + bs + 1) / bs] |= 1u << ((ip + bs + 1) % bs);
    }
    if (lane == 0) {
        w->hash = 0; // go.cpp:106
        w->turn = 1; // go.cpp:105, tictactoe.cpp:13, othello.cpp:15
        w->num_moves = 0;
        w->last = -1;
        w->last2 = -1;
    }
    mz_sync();
}


// Othello: stones `player` (1 / 2) would flip by playing the EMPTY cell (x0, y0), as row masks OR-ed into flips[] when it
// is not null; returns their number (OthelloEnv::getFlipPoint over the 8 directions, othello.cpp:63-82,119-121). The
// reference's shift-and-mask sweeps implement the standard rule: a run of opposing stones closed by an own stone.
MZ_DEV int mz_othello_flips(const mz_scratch* w, int N, int x0, int y0, int player, uint32_t* flips)
{
    const int me = player - 1, opp = 1 - me;
    int total = 0;
    for (int dir = 0; dir < 8; ++dir) {
        const int dx = (dir == 2 || dir == 4 || dir == 7) ? -1 : ((dir == 3 || dir == 5 || dir == 6) ? 1 : 0);
        const int dy = (dir == 0 || dir == 4 || dir == 5) ? 1 : ((dir == 1 || dir == 6 || dir == 7) ? -1 : 0);
        int x = x0 + dx, y = y0 + dy, run = 0;
        while (x >= 0 && x < N && y >= 0 && y < N && ((w->st[opp][y] >> x) & 1u)) { x += dx, y += dy, ++run; }
        if (run == 0 || x < 0 || x >= N || y < 0 || y >= N || !((w->st[me][y] >> x) & 1u)) { continue; }
        total += run;
        if (flips) {
            for (int k = 1; k <= run; ++k) { flips[y0 + k * dy] |= 1u << (x0 + k * dx); }
        }
    }
    return total;
}

MZ_DEV int mz_othello_has_move(const mz_scratch* w, int N, int player)
{
    for (int c = 0; c < N * N; ++c) {
        const int x = c % N, y = c / N;
        if (((w->st[0][y] | w->st[1][y]) >> x) & 1u) { continue; }
        if (mz_othello_flips(w, N, x, y, player, nullptr) > 0) { return 1; }
    }
    return 0;
}

// GoEnv::act (go.cpp:132-190) / TicTacToeEnv::act (tictactoe.cpp:19-26) / OthelloEnv::act (othello.cpp:102-139) for a move already known to
// be legal. The caller appends the new w->hash to the superko history (go.cpp:145-147,180-182).
MZ_DEV void mz_env_act(const mz_dims& d, const mz_state& s, mz_scratch* w, int a, int player, int lane)
{
    const int N = d.N, me = player - 1, opp = 1 - me;
    uint64_t hash = w->hash ^ d.turn_key; // go.cpp:141
    const int num_moves = w->num_moves;
    if (MZ_GO_FAMILY(d.game)) { // NoGoEnv inherits GoEnv::act; its legality rules out every capture
        if (a != N * N) {
            const int r = a / N, x = a % N;
            if (lane == 0) { w->st[me][r] |= (1u << x); }
            hash ^= s.keys[me * 361 + a]; // go.cpp:154
            for (int i = lane; i < N; i += MZ_W) { w->checked[i] = 0u; }
            mz_sync();
            // neighbours in the order of go_grid.h:43-54 (up, right, down, left); the final position
            // does not depend on the order
            for (int k = 0; k < 4; ++k) {
                int nr = r + (k == 0) - (k == 2), nx = x + (k == 1) - (k == 3);
                if (nr < 0 || nr >= N || nx < 0 || nx >= N) { continue; }
                const uint32_t bit = 1u << nx;
                if (!(w->st[opp][nr] & bit) || (w->checked[nr] & bit)) { continue; } // warp-uniform
                for (int i = lane; i < N; i += MZ_W) { w->fill[i] = (i == nr ? bit : 0u); }
                mz_sync();
                mz_flood(w->fill, w->st[opp], N, lane);
                int has_lib = 0;
                for (int i = lane; i < N; i += MZ_W) {
                    uint32_t empty = ~(w->st[0][i] | w->st[1][i]) & mz_rowmask(N);
                    has_lib |= ((mz_dilate_row(w->fill, i, N) & empty) != 0u);
                }
                has_lib = mz_any(has_lib);
                if (!has_lib) { // removeBlockFromBoard, go.cpp:388-433
                    hash ^= mz_block_hash(w->fill, opp, s.keys, N, lane);
                    for (int i = lane; i < N; i += MZ_W) { w->st[opp][i] &= ~w->fill[i]; }
                } else {
                    for (int i = lane; i < N; i += MZ_W) { w->checked[i] |= w->fill[i]; }
                }
                mz_sync();
            }
        }
    } else if (d.game == MZ_GAME_HEX) { // hex.cpp:21-66
        if (lane == 0) {
            int id = a;
            if (d.hex_swap_rule && num_moves == 1 && a == w->last) { // swap: the first stone changes sides, mirrored over the other diagonal
                const int row = a / N, col = a % N;
                id = (N - 1 - col) * N + (N - 1 - row);
                w->st[0][row] &= ~(1u << col);
                w->st[1][row] &= ~(1u << col);
            }
            w->st[me][id / N] |= (1u << (id % N));
        }
        mz_sync();
    } else if (d.game == MZ_GAME_OTHELLO) {
        if (lane == 0 && a != N * N) { // a pass only hands the turn over (othello.cpp:111)
            for (int i = 0; i < N; ++i) { w->tmp[i] = 0u; }
            mz_othello_flips(w, N, a % N, a / N, player, w->tmp);
            w->st[me][a / N] |= (1u << (a % N));
            for (int i = 0; i < N; ++i) { w->st[me][i] |= w->tmp[i], w->st[opp][i] &= ~w->tmp[i]; }
        }
        mz_sync();
    } else {
        if (lane == 0) { w->st[me][a / N] |= (1u << (a % N)); }
        mz_sync();
    }
    const int slot = num_moves % MZ_HIST;
    for (int i = lane; i < N; i += MZ_W) {
        w->hist[slot][0][i] = w->st[0][i];
        w->hist[slot][1][i] = w->st[1][i];
    }
    mz_sync();
    if (lane == 0) {
        w->hash = hash;
        w->turn = 3 - player; // go.cpp:140
        w->num_moves = num_moves + 1;
        w->last2 = w->last;
        w->last = a;
    }
    mz_sync();
}

// TicTacToeEnv::eval (tictactoe.cpp:124-146): 1 / 2 if that player owns a line, else 0
MZ_DEV int mz_ttt_eval(const mz_scratch* w)
{
    for (int p = 0; p < 2; ++p) {
        const uint32_t r0 = w->st[p][0], r1 = w->st[p][1], r2 = w->st[p][2];
        bool win = (r0 == 7u) || (r1 == 7u) || (r2 == 7u) || ((r0 & r1 & r2) != 0u) ||
                   ((r0 & 1u) && (r1 & 2u) && (r2 & 4u)) || ((r0 & 4u) && (r1 & 2u) && (r2 & 1u));
        if (win) { return p + 1; }
    }
    return 0;
}

// Hex: does `player` connect its two edges (Black: columns 0 and N-1, White: rows 0 and N-1; hex.cpp:47-58)? The reference merges
// per-cell edge flags through the six neighbours of every new stone (hex.cpp:305-347); the same relation as a flood fill from
// edge 1 through the player's stones, dilated over the six hex neighbours (x-1,y-1) (x,y-1) (x-1,y) (x+1,y) (x,y+1) (x+1,y+1).
// Single thread; rows of at most 19 bits.
MZ_DEV int mz_hex_connected(const mz_scratch* w, int N, int player)
{
    const uint32_t* st = w->st[player - 1];
    uint32_t f[MZ_MAXN];
    for (int y = 0; y < N; ++y) { f[y] = (player == 1 ? (st[y] & 1u) : (y == 0 ? st[y] : 0u)); }
    bool changed = true;
    while (changed) {
        changed = false;
        for (int y = 0; y < N; ++y) {
            uint32_t v = f[y] | (f[y] << 1) | (f[y] >> 1);
            if (y > 0) { v |= f[y - 1] | (f[y - 1] << 1); }
            if (y + 1 < N) { v |= f[y + 1] | (f[y + 1] >> 1); }
            v &= st[y];
            if (v != f[y]) {
                f[y] = v;
                changed = true;
            }
        }
    }
    if (player == 1) {
        for (int y = 0; y < N; ++y) {
            if ((f[y] >> (N - 1)) & 1u) { return 1; }
        }
        return 0;
    }
    return f[N - 1] != 0u;
}
MZ_DEV int mz_hex_winner(const mz_dims& d, const mz_scratch* w) { return mz_hex_connected(w, d.N, 1) ? 1 : (mz_hex_connected(w, d.N, 2) ? 2 : 0); }

// GomokuEnv::updateWinner for the last move (gomoku.cpp:140-164): winner_ is a function of the board and the last action
MZ_DEV int mz_gomoku_winner(const mz_dims& d, const mz_scratch* w)
{
    const int N = d.N;
    if (w->num_moves == 0 || w->last < 0) { return 0; }
    const int x0 = w->last % N, y0 = w->last / N;
    const int p = ((w->st[0][y0] >> x0) & 1u) ? 0 : 1; // the mover's colour index
    for (int dir = 0; dir < 4; ++dir) {
        const int dx = (dir == 1 ? 0 : 1), dy = (dir == 0 ? 0 : (dir == 3 ? -1 : 1));
        int c = 1;
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            int x = x0 + sgn * dx, y = y0 + sgn * dy;
            while (x >= 0 && x < N && y >= 0 && y < N && ((w->st[p][y] >> x) & 1u)) { ++c, x += sgn * dx, y += sgn * dy; }
        }
        if (d.gomoku_exactly_five ? (c == 5) : (c >= 5)) { return p + 1; } // gomoku.h:46
    }
    return 0;
}

// ---- KillAllGo (environment/killallgo/killallgo.cpp:27-48) on top of the Go rules; Benson's unconditional life lives in killallgo_rules.h ----
MZ_DEV uint64_t mz_ka_board(const mz_scratch* w, int colour) // row bitboards -> one 64-bit board (bit y * 8 + x)
{
    uint64_t b = 0;
    for (int y = 0; y < 7; ++y) { b |= (uint64_t)(w->st[colour][y] & 0x7fu) << (8 * y); }
    return b;
}

MZ_DEV int mz_env_is_terminal(const mz_dims& d, const mz_scratch* w)
{
    const int N = d.N;
    if (d.game == MZ_GAME_HEX) { return mz_hex_winner(d, w) != 0; } // hex.cpp:96-99
    if (d.game == MZ_GAME_GOMOKU) { // gomoku.cpp:60-63
        if (mz_gomoku_winner(d, w) != 0) { return 1; }
        for (int r = 0; r < N; ++r) {
            if (((w->st[0][r] | w->st[1][r]) & mz_rowmask(N)) != mz_rowmask(N)) { return 0; }
        }
        return 1;
    }
    if (d.game == MZ_GAME_GO) {
        if (d.killall && mz_ka_terminal(mz_ka_board(w, 0), mz_ka_board(w, 1))) { return 1; } // killallgo.cpp:34-40
        if (w->num_moves >= 2 && w->last == N * N && w->last2 == N * N) { return 1; } // go.cpp:249-251
        return w->num_moves > 2 * N * N;                                              // go.cpp:254
    }
    if (d.game == MZ_GAME_OTHELLO) { return w->num_moves >= 2 && w->last == N * N && w->last2 == N * N; } // othello.cpp:201-207
    if (d.game == MZ_GAME_NOGO) { return 0; } // "no legal move left" (nogo.h:61-68): decided by the callers from the legal set
    if (mz_ttt_eval(w) != 0) { return 1; } // tictactoe.cpp:51-55
    uint32_t occ = (w->st[0][0] | w->st[1][0]) & (w->st[0][1] | w->st[1][1]) & (w->st[0][2] | w->st[1][2]);
    return occ == 7u;
}

// getEvalScore(false): Tromp-Taylor area (go.cpp:259-278,703-723) / line owner (tictactoe.cpp:57-65).
// An empty region counts for Black unless it touches a White stone, else for White unless it touches a
// Black stone — which two multi-seed floods through the empty cells decide for all regions at once.
MZ_DEV float mz_env_eval_score(const mz_dims& d, mz_scratch* w, int lane)
{
    const int N = d.N;
    int winner;
    if (d.game == MZ_GAME_NOGO) { // the side to move has lost (nogo.h:70-78)
        winner = 3 - w->turn;
    } else if (d.game == MZ_GAME_GOMOKU) { // gomoku.cpp:65-73
        winner = mz_gomoku_winner(d, w);
    } else if (d.game == MZ_GAME_HEX) { // hex.cpp:101-111
        winner = mz_hex_winner(d, w);
    } else if (d.game == MZ_GAME_GO && d.killall) { // killallgo.cpp:42-48: Black wins when no White stone is left or the whole board is unconditionally Black's
        winner = mz_ka_winner(mz_ka_board(w, 0), mz_ka_board(w, 1));
    } else if (d.game == MZ_GAME_GO) {
        int cnt_b = 0, cnt_w = 0;
        for (int i = lane; i < N; i += MZ_W) {
            uint32_t empty = ~(w->st[0][i] | w->st[1][i]) & mz_rowmask(N);
            w->tmp[i] = empty;
            cnt_b += mz_popc(w->st[0][i]);
            cnt_w += mz_popc(w->st[1][i]);
        }
        mz_sync();
        for (int c = 0; c < 2; ++c) { // c = 0: cells reaching Black -> checked[], c = 1: reaching White -> fill[]
            uint32_t* out = (c == 0 ? w->checked : w->fill);
            for (int i = lane; i < N; i += MZ_W) { out[i] = mz_dilate_row(w->st[c], i, N) & w->tmp[i]; }
            mz_sync();
            mz_flood(out, w->tmp, N, lane);
        }
        for (int i = lane; i < N; i += MZ_W) {
            cnt_b += mz_popc(w->tmp[i] & ~w->fill[i]);
            cnt_w += mz_popc(w->tmp[i] & w->fill[i] & ~w->checked[i]);
        }
        cnt_b = mz_reduce_add(cnt_b);
        cnt_w = mz_reduce_add(cnt_w);
        mz_sync();
        const float tb = (float)cnt_b, tw = mz_fadd((float)cnt_w, d.komi);
        winner = (tb > tw ? 1 : (tb < tw ? 2 : 0));
    } else if (d.game == MZ_GAME_OTHELLO) { // OthelloEnv::eval, othello.cpp:219-236
        winner = 0;
        if (!mz_othello_has_move(w, N, 1) && !mz_othello_has_move(w, N, 2)) {
            int c1 = 0, c2 = 0;
            for (int i = 0; i < N; ++i) { c1 += mz_popc(w->st[0][i]), c2 += mz_popc(w->st[1][i]); }
            winner = (c1 > c2 ? 1 : (c1 < c2 ? 2 : 0));
        }
    } else {
        winner = mz_ttt_eval(w);
    }
    return winner == 1 ? 1.0f : (winner == 2 ? -1.0f : 0.0f);
}

MZ_DEV int mz_cell_colour(const mz_scratch* w, int c, int N)
{
    const int r = c / N, x = c % N;
    return ((w->st[0][r] >> x) & 1u) ? 1 : (((w->st[1][r] >> x) & 1u) ? 2 : 0);
}

// Same result as mz_env_legal, computed by ALL threads of the block (tid / nthreads), one cell per thread:
// blocks are labelled by min-propagation with pointer jumping (label = smallest cell index of the block, the unique
// fixpoint whatever the update order), liberties are counted per label with shared-memory atomics, block hashes are
// XOR-folded per label, and every empty cell then tests its <= 4 neighbouring blocks (go.cpp:208-244). The superko
// lookup goes through a 2048-bit filter of the history first; only filter hits scan the list.
MZ_DEV int mz_env_legal_block(const mz_dims& d, const mz_state& s, mz_scratch* w, const uint64_t* root_list, int root_n, const uint64_t* path_list, int path_n,
                              int tid, int nthreads)
{
    const int N = d.N, NN = N * N, me = w->turn - 1;
    for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { w->legal[i] = 0u; }
    if (d.game == MZ_GAME_OTHELLO) { // legal_board_ of the side to move (othello.cpp:125-134), pass iff it is empty (:135-136,194-196)
        mz_block_sync();
        for (int c = tid; c < NN; c += nthreads) {
            const int x = c % N, y = c / N;
            if (((w->st[0][y] | w->st[1][y]) >> x) & 1u) { continue; }
            if (mz_othello_flips(w, N, x, y, w->turn, nullptr) > 0) { mz_atomic_or(&w->legal[c >> 5], 1u << (c & 31)); }
        }
        mz_block_sync();
        int n = 0;
        for (int i = 0; i < MZ_LEGAL_WORDS; ++i) { n += mz_popc(w->legal[i]); }
        mz_block_sync();
        if (n == 0) {
            if (tid == 0) { w->legal[NN >> 5] |= (1u << (NN & 31)); }
            mz_block_sync();
            n = 1;
        }
        return n;
    }
    if (d.game == MZ_GAME_GOMOKU || d.game == MZ_GAME_HEX) {
        // gomoku.cpp:49-58: empty points; "outer_open": Black's first stone within two lines of an edge
        // hex.cpp:83-94: empty points; with the swap rule every point on the second move (playing the first stone's point swaps)
        mz_block_sync();
        for (int c = tid; c < NN; c += nthreads) {
            const int x = c % N, y = c / N;
            bool ok = !(((w->st[0][y] | w->st[1][y]) >> x) & 1u);
            if (d.game == MZ_GAME_GOMOKU && w->num_moves == 0 && d.gomoku_outer_open) { ok = (y < 2 || y >= N - 2) || (x < 2 || x >= N - 2); }
            if (d.game == MZ_GAME_HEX && w->num_moves == 1 && d.hex_swap_rule) { ok = true; }
            if (ok) { mz_atomic_or(&w->legal[c >> 5], 1u << (c & 31)); }
        }
        mz_block_sync();
        int n = 0;
        for (int i = 0; i < MZ_LEGAL_WORDS; ++i) { n += mz_popc(w->legal[i]); }
        mz_block_sync();
        return n;
    }
    if (!MZ_GO_FAMILY(d.game)) {
        mz_block_sync();
        if (tid == 0) {
            uint32_t bits = 0;
            for (int r = 0; r < N; ++r) { bits |= (~(w->st[0][r] | w->st[1][r]) & mz_rowmask(N)) << (r * N); }
            w->legal[0] = bits; // tictactoe.cpp:44-49
        }
        mz_block_sync();
        return mz_popc(w->legal[0]);
    }
    uint64_t* bhash = w->cap_hash; // per-label block hash
    uint8_t* cinfo = reinterpret_cast<uint8_t*>(w->pol); // per cell: colour | up << 2 | right << 3 | down << 4 | left << 5 (pol[] is free here)
    for (int c = tid; c < NN; c += nthreads) {
        const int r = c / N, x = c % N, col = mz_cell_colour(w, c, N);
        cinfo[c] = (uint8_t)(col | ((r + 1 < N) << 2) | ((x + 1 < N) << 3) | ((r > 0) << 4) | ((x > 0) << 5));
        w->label[c] = (col != 0 ? c : -1);
        w->libcnt[c] = 0;
        bhash[c] = 0;
    }
    for (int i = tid; i < 64; i += nthreads) { w->bloom[i] = 0u; }
    mz_block_sync();
    // superko filter
    for (int i = tid; i < root_n + path_n; i += nthreads) {
        const uint64_t h = (i < root_n ? root_list[i] : path_list[i - root_n]);
        mz_atomic_or(&w->bloom[(h >> 5) & 63], 1u << (h & 31));
    }
    // block labels
    for (;;) {
        if (tid == 0) { w->flag = 0; }
        mz_block_sync();
        int changed = 0;
        for (int c = tid; c < NN; c += nthreads) {
            const int l = w->label[c];
            if (l < 0) { continue; }
            const int info = cinfo[c], col = info & 3;
            int m = l;
            if ((info & 4) && (cinfo[c + N] & 3) == col) { m = (w->label[c + N] < m ? w->label[c + N] : m); }
            if ((info & 8) && (cinfo[c + 1] & 3) == col) { m = (w->label[c + 1] < m ? w->label[c + 1] : m); }
            if ((info & 16) && (cinfo[c - N] & 3) == col) { m = (w->label[c - N] < m ? w->label[c - N] : m); }
            if ((info & 32) && (cinfo[c - 1] & 3) == col) { m = (w->label[c - 1] < m ? w->label[c - 1] : m); }
            const int mm = w->label[m]; // pointer jumping: the label of my label's cell
            m = (mm < m ? mm : m);
            if (m < l) {
                w->label[c] = m;
                changed = 1;
            }
        }
        if (changed) { w->flag = 1; }
        mz_block_sync();
        if (!w->flag) { break; }
        mz_block_sync();
    }
    // liberties and hashes per label
    for (int c = tid; c < NN; c += nthreads) {
        const int info = cinfo[c], col = info & 3;
        if (col != 0) {
            mz_atomic_xor64(&bhash[w->label[c]], s.keys[(col - 1) * 361 + c]);
            continue;
        }
        int nb[4], k = 0;
        if ((info & 4) && w->label[c + N] >= 0) { nb[k++] = w->label[c + N]; }
        if ((info & 8) && w->label[c + 1] >= 0) { nb[k++] = w->label[c + 1]; }
        if ((info & 16) && w->label[c - N] >= 0) { nb[k++] = w->label[c - N]; }
        if ((info & 32) && w->label[c - 1] >= 0) { nb[k++] = w->label[c - 1]; }
        for (int i = 0; i < k; ++i) {
            bool dup = false;
            for (int j = 0; j < i; ++j) { dup |= (nb[j] == nb[i]); }
            if (!dup) { mz_atomic_inc(&w->libcnt[nb[i]]); }
        }
    }
    mz_block_sync();
    // legality of every empty cell (go.cpp:208-244)
    const uint64_t base = w->hash ^ d.turn_key;
    for (int c = tid; c < NN; c += nthreads) {
        const int info = cinfo[c];
        if ((info & 3) != 0) { continue; }
        int nbc[4], k = 0;
        if (info & 4) { nbc[k++] = c + N; }
        if (info & 8) { nbc[k++] = c + 1; }
        if (info & 16) { nbc[k++] = c - N; }
        if (info & 32) { nbc[k++] = c - 1; }
        bool legal = false, forbidden = false;
        uint64_t nh = base ^ s.keys[me * 361 + c];
        int seen_lab[4], ns = 0;
        for (int i = 0; i < k; ++i) {
            const int l = w->label[nbc[i]];
            if (l < 0) { // empty neighbour (go.cpp:225-226)
                legal = true;
                continue;
            }
            bool dup = false;
            for (int j = 0; j < ns; ++j) { dup |= (seen_lab[j] == l); }
            if (dup) { continue; } // block already examined (go.cpp:229)
            seen_lab[ns++] = l;
            const int lc = w->libcnt[l];
            if ((cinfo[l] & 3) - 1 == me) {
                if (lc > 1) { legal = true; } // go.cpp:232-233
            } else if (lc == 1) {             // capture (go.cpp:235-238)
                nh ^= bhash[l];
                legal = true;
                forbidden = (d.game == MZ_GAME_NOGO); // "illegal when suicide or capture opponent's stones", nogo.h:40-56
            }
        }
        if (!legal || forbidden) { continue; }
        bool seen = false;
        if (d.game == MZ_GAME_GO && ((w->bloom[(nh >> 5) & 63] >> (nh & 31)) & 1u)) { // NoGo has no repetition rule
            for (int i = 0; i < root_n; ++i) { seen |= (root_list[i] == nh); }
            for (int i = 0; i < path_n; ++i) { seen |= (path_list[i] == nh); }
        }
        if (!seen) { mz_atomic_or(&w->legal[c >> 5], 1u << (c & 31)); }
    }
    mz_block_sync();
    if (tid == 0 && d.game == MZ_GAME_GO) { w->legal[NN >> 5] |= (1u << (NN & 31)); } // pass, go.cpp:213 (never legal in NoGo, nogo.h:32)
    mz_block_sync();
    if (d.killall && w->num_moves < 3) { // KillAllGoEnv::isLegalAction, killallgo.cpp:27-32: Black opens with two stones, White's first move is the pass
        if (tid == 0) {
            if (w->num_moves == 1) {
                for (int i = 0; i < MZ_LEGAL_WORDS; ++i) { w->legal[i] = 0u; }
                w->legal[NN >> 5] = (1u << (NN & 31));
            } else {
                w->legal[NN >> 5] &= ~(1u << (NN & 31));
            }
        }
        mz_block_sync();
    }
    int n = 0;
    for (int i = 0; i < MZ_LEGAL_WORDS; ++i) { n += mz_popc(w->legal[i]); }
    return n;
}

// getFeatures (go.cpp:280-308, tictactoe.cpp:67-90) written as fp16 NHWC rows of the first conv's input:
// board cell (x, y) of game g lives at row g * slots + (y + 1) * (N + 1) + x, channel c at column c.
MZ_DEV void mz_env_features(const mz_dims& d, const mz_state& s, int g, const mz_scratch* w, int rotation, int lane, int stride = MZ_W)
{
    const int N = d.N, rev = (d.game == MZ_GAME_HEX ? 0 : mz_reversed_rotation(rotation)); // HexEnv::getFeatures ignores it (hex.cpp:123-124)
    const int turn = w->turn, me = turn - 1, opp = 1 - me;
    uint16_t* base = s.nn_in + (size_t)g * d.slots * MZ_NN_CPAD;
    for (int pos = lane; pos < N * N; pos += stride) {
        const int rp = mz_rotate(rev, pos, N), rr = rp / N, rx = rp % N;
        uint16_t* out = base + (size_t)((pos / N + 1) * (N + 1) + pos % N) * MZ_NN_CPAD;
        if (MZ_GO_FAMILY(d.game)) {
            for (int c = 0; c < 16; ++c) {
                const int idx = w->num_moves - 1 - c / 2;
                uint16_t v = 0;
                if (idx >= 0) { v = ((w->hist[idx % MZ_HIST][(c & 1) ? opp : me][rr] >> rx) & 1u) ? MZ_HALF_ONE : 0; }
                out[c] = v;
            }
            out[16] = (turn == 1 ? MZ_HALF_ONE : 0);
            out[17] = (turn == 2 ? MZ_HALF_ONE : 0);
        } else {
            out[0] = ((w->st[me][rr] >> rx) & 1u) ? MZ_HALF_ONE : 0;
            out[1] = ((w->st[opp][rr] >> rx) & 1u) ? MZ_HALF_ONE : 0;
            out[2] = (turn == 1 ? MZ_HALF_ONE : 0);
            out[3] = (turn == 2 ? MZ_HALF_ONE : 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tree
// ---------------------------------------------------------------------------------------------

// MCTSNode::getNormalizedMean (mcts.cpp:40-53), no virtual loss: reward + discount * mean, min-max rescaled by the tree's value
// bounds when actor_mcts_value_rescale (Atari), negated for White's nodes
MZ_DEV float mz_normalized_mean(const mz_dims& d, const mz_qb& qb, int node, float mean, float count, int player, float vloss = 0.0f)
{
    float v = mz_fadd(qb.reward ? qb.reward[node] : 0.0f, mz_fmul(d.discount, mean));
    if (d.value_rescale) {
        if (qb.n < 2) { return 1.0f; }
        v = mz_fdiv(mz_fsub(v, qb.lo), mz_fsub(qb.hi, qb.lo));
        v = mz_fsub(mz_fmul(2.0f, v), 1.0f);
        v = (v < -1.0f ? -1.0f : v), v = (v > 1.0f ? 1.0f : v); // fmin(1, fmax(-1, x)) on a non-NaN x
    }
    if (player == 2) { v = -v; } // actor_mcts_value_flipping_player == 'W'
    return mz_fdiv(mz_fsub(mz_fmul(v, count), vloss), mz_fadd(count, vloss)); // value with virtual loss (0 outside think(), mcts.cpp:51)
}

// MCTS::calculateInitQValue (mcts.cpp:200-217) from the ordered sum over the visited children
MZ_DEV float mz_init_q(const mz_dims& d, float sum_win, float sum_n)
{
    if (d.atari_init_q) { return sum_n > 0.0f ? mz_fdiv(sum_win, sum_n) : 1.0f; } // #if ATARI
    return mz_fdiv(mz_fsub(sum_win, 1.0f), mz_fadd(sum_n, 1.0f));
}

// player of the CHILDREN of a node at depth `level` below a root where `root_turn` is to move (a one-player game has only player 1)
MZ_DEV int mz_child_player(const mz_dims& d, int root_turn, int level) { return d.num_players == 1 ? 1 : ((level & 1) ? 3 - root_turn : root_turn); }

// order-preserving map float -> uint32 (for REDUX arg-max); +0.0 and -0.0 are made equal first
MZ_DEV uint32_t mz_sortable(float f)
{
    uint32_t u = mz_float_bits(f == 0.0f ? 0.0f : f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One level of MCTS::select: MCTS::selectChildByPUCTScore (mcts.cpp:181-198) for the node whose hot record is `h`.
// Returns the index (0 .. nc-1) of the chosen child and its hot record in `out`. Warp collective; `q` is this warp's
// scratch of MZ_MAXA floats.
//
// One coalesced read of the children's hot records, the ordered f32 sum of the visited children's Q for init-Q
// (mcts.cpp:200-217), then the arg-max of the PUCT score (mcts.cpp:55-61). Below the root the children are stored in
// non-increasing prior order and all unvisited children share the same Q (init-Q), so the best unvisited child is the
// FIRST unvisited one (score is monotone in the prior; ties go to the higher prior, then to the lower index —
// mcts.cpp:191): only the visited children and that one candidate are scored. The root's priors are mixed with noise
// after sorting (zero_actor.cpp:194-204), so every root child is scored.
MZ_DEV int mz_select_level(const mz_dims& d, const mz_state& s, const mz_qb& qb, const mz_hot* hot, const mz_hot& h, bool score_all, int child_player, float* q, int lane,
                           mz_hot& out)
{
    const int nc = (int)(h.link >> MZ_LINK_SHIFT), fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
    const int total = (int)mz_fsub(h.count, 1.0f); // mcts.cpp:185
    const float bias = s.puct_bias[total];
    const double sqrt_n = s.sqrt_table[total];
    float sum_win = 0.0f, sum_n = 0.0f;
    int first_unvisited = nc;
    // candidates of this lane among the VISITED children (their score does not depend on init-Q)
    float best_s = 0.0f, best_p = 0.0f;
    int best_i = -1;
    mz_hot best_h = h, h_fu = h;
    // the children's records are fetched MZ_SEL_AHEAD chunks at a time, so that a node with up to MZ_SEL_AHEAD * 32 children costs
    // one memory round trip instead of one per chunk (the ordered init-Q sum below serialises the chunks)
    for (int base0 = 0; base0 < nc; base0 += MZ_SEL_AHEAD * MZ_W) {
        mz_hot cbuf[MZ_SEL_AHEAD];
#pragma unroll
        for (int u = 0; u < MZ_SEL_AHEAD; ++u) {
            const int i = base0 + u * MZ_W + lane;
            cbuf[u] = h;
            if (i < nc) { cbuf[u] = mz_load_hot(hot + fc + i); }
        }
#pragma unroll
        for (int u = 0; u < MZ_SEL_AHEAD; ++u) {
            const int base = base0 + u * MZ_W;
            if (base >= nc) { break; } // warp-uniform
            const int i = base + lane;
            int visited = 0;
            const mz_hot c = cbuf[u];
            if (i < nc) {
                visited = (c.count != 0.0f);
                if (visited) {
                    const float qv = mz_normalized_mean(d, qb, fc + i, c.mean, c.count, child_player);
                    q[i] = qv;
                    const double num = mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n);
                    const float score = mz_fadd((float)mz_ddiv(num, (double)mz_fadd(1.0f, c.count)), qv);
                    if (best_i < 0 || score > best_s || (score == best_s && c.policy > best_p)) { best_s = score, best_p = c.policy, best_i = i, best_h = c; }
                }
            }
            unsigned m = mz_ballot(visited);
            const unsigned valid = (nc - base >= MZ_W ? ~0u >> (32 - MZ_W) : ((1u << (nc - base)) - 1u));
            const unsigned unv = ~m & valid;
            if (first_unvisited == nc && unv) {
                const int fl = mz_ffs0(unv);
                first_unvisited = base + fl;
#if MZ_W > 1
                h_fu.count = __shfl_sync(MZ_FULL, c.count, fl), h_fu.mean = __shfl_sync(MZ_FULL, c.mean, fl);
                h_fu.policy = __shfl_sync(MZ_FULL, c.policy, fl), h_fu.link = __shfl_sync(MZ_FULL, c.link, fl);
#else
                h_fu = c;
#endif
            }
            mz_sync();
            while (m) { // ordered f32 sum of the visited children's Q (mcts.cpp:200-217)
                const int b = mz_ffs0(m);
                m &= m - 1;
                sum_win = mz_fadd(sum_win, q[base + b]);
                sum_n = mz_fadd(sum_n, 1.0f);
            }
        }
    }
    const float init_q = mz_init_q(d, sum_win, sum_n);
    if (score_all) { // root: the unvisited children's priors are not sorted (noise), score every one of them
        for (int i = lane; i < nc; i += MZ_W) {
            const mz_hot c = mz_load_hot(hot + fc + i); // L1
            if (c.count != 0.0f) { continue; }
            const float score = mz_fadd((float)mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n), init_q); // n = 0: division by 1.0
            if (best_i < 0 || score > best_s || (score == best_s && (c.policy > best_p || (c.policy == best_p && i < best_i)))) {
                best_s = score, best_p = c.policy, best_i = i, best_h = c;
            }
        }
    } else if (first_unvisited < nc && (lane == first_unvisited % MZ_W)) { // the one unvisited candidate, held by its own lane
        const float score = mz_fadd((float)mz_dmul((double)mz_fmul(bias, h_fu.policy), sqrt_n), init_q);
        const int i = first_unvisited;
        if (best_i < 0 || score > best_s || (score == best_s && (h_fu.policy > best_p || (h_fu.policy == best_p && i < best_i)))) {
            best_s = score, best_p = h_fu.policy, best_i = i, best_h = h_fu;
        }
    }
    // lexicographic arg-max (score desc, prior desc, index asc) over the lanes' candidates (mcts.cpp:187-194)
#if MZ_W > 1
    {
        const int mine = best_i;
        const uint32_t ks = (mine >= 0 ? mz_sortable(best_s) : 0u);
        const uint32_t top_s = mz_redux_max(ks);
        const bool in_s = (mine >= 0 && ks == top_s);
        const uint32_t kp = (in_s ? mz_sortable(best_p) : 0u);
        const uint32_t top_p = mz_redux_max(kp);
        const bool in_p = (in_s && kp == top_p);
        best_i = (int)mz_redux_min(in_p ? (uint32_t)mine : 0xffffffffu);
        const int owner = mz_ffs0(mz_ballot(in_p && mine == best_i));
        out.count = __shfl_sync(MZ_FULL, best_h.count, owner);
        out.mean = __shfl_sync(MZ_FULL, best_h.mean, owner);
        out.policy = __shfl_sync(MZ_FULL, best_h.policy, owner);
        out.link = __shfl_sync(MZ_FULL, best_h.link, owner);
    }
#else
    out = best_h;
#endif
    mz_sync();
    return best_i;
}

// The same choice made by ONE thread scanning the children in order, exactly like the loops of mcts.cpp:181-217
// (ordered f32 sum for init-Q, first-best arg-max), with the unvisited children below the root reduced to the first
// one as explained above. Used by the level-parallel re-evaluation of deep paths: one thread per level.
MZ_DEV int mz_select_level_serial(const mz_dims& d, const mz_state& s, const mz_qb& qb, const mz_hot* hot, const mz_hot& h, int child_player)
{
    const int nc = (int)(h.link >> MZ_LINK_SHIFT), fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
    const int total = (int)mz_fsub(h.count, 1.0f);
    const float bias = s.puct_bias[total];
    const double sqrt_n = s.sqrt_table[total];
    float sum_win = 0.0f, sum_n = 0.0f, best_s = 0.0f, best_p = 0.0f, p_fu = 0.0f;
    int first_unvisited = nc, best_i = -1;
#pragma unroll 4
    for (int i = 0; i < nc; ++i) {
        const mz_hot c = mz_load_hot(hot + fc + i);
        if (c.count != 0.0f) {
            const float qv = mz_normalized_mean(d, qb, fc + i, c.mean, c.count, child_player);
            sum_win = mz_fadd(sum_win, qv);
            sum_n = mz_fadd(sum_n, 1.0f);
            const double num = mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n);
            const float score = mz_fadd((float)mz_ddiv(num, (double)mz_fadd(1.0f, c.count)), qv);
            if (best_i < 0 || score > best_s || (score == best_s && c.policy > best_p)) { best_s = score, best_p = c.policy, best_i = i; }
        } else if (first_unvisited == nc) {
            first_unvisited = i;
            p_fu = c.policy;
        }
    }
    if (first_unvisited < nc) {
        const float init_q = mz_init_q(d, sum_win, sum_n);
        const float score = mz_fadd((float)mz_dmul((double)mz_fmul(bias, p_fu), sqrt_n), init_q);
        if (best_i < 0 || score > best_s || (score == best_s && (p_fu > best_p || (p_fu == best_p && first_unvisited < best_i)))) { best_i = first_unvisited; }
    }
    return fc + best_i;
}


// mz_select_level for a node below the root whose visited children are listed in `v` (v.n <= MZ_VIS_MAX): the same choice
// from the same arithmetic in the same child order — the listed children and the first unvisited one are the only candidates
// (see mz_select_level) — but one gather of <= 7 records instead of a scan of all children. Warp collective.
MZ_DEV int mz_select_level_vis(const mz_dims& d, const mz_state& s, const mz_qb& qb, const mz_hot* hot, const mz_hot& h, const mz_vis& v, int child_player, int lane)
{
    const int nc = (int)(h.link >> MZ_LINK_SHIFT), fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
    const int total = (int)mz_fsub(h.count, 1.0f); // mcts.cpp:185
    const float bias = s.puct_bias[total];
    const double sqrt_n = s.sqrt_table[total];
    const int n = v.n;
#if MZ_W > 1
    const int my = (lane < n ? (int)v.idx[lane] : (lane == n && v.fu < nc ? (int)v.fu : -1));
    mz_hot c = h;
    if (my >= 0) { c = mz_load_hot(hot + fc + my); }
    float qv = 0.0f, score = 0.0f;
    if (lane < n) {
        qv = mz_normalized_mean(d, qb, fc + my, c.mean, c.count, child_player);
        const double num = mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n);
        score = mz_fadd((float)mz_ddiv(num, (double)mz_fadd(1.0f, c.count)), qv);
    }
    float sum_win = 0.0f, sum_n = 0.0f;
    for (int i = 0; i < n; ++i) { // ordered f32 sum of the visited children's Q (mcts.cpp:200-217)
        sum_win = mz_fadd(sum_win, __shfl_sync(MZ_FULL, qv, i));
        sum_n = mz_fadd(sum_n, 1.0f);
    }
    if (lane == n && my >= 0) {
        const float init_q = mz_init_q(d, sum_win, sum_n);
        score = mz_fadd((float)mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n), init_q);
    }
    const uint32_t ks = (my >= 0 ? mz_sortable(score) : 0u);
    const uint32_t top_s = mz_redux_max(ks);
    const bool in_s = (my >= 0 && ks == top_s);
    const uint32_t kp = (in_s ? mz_sortable(c.policy) : 0u);
    const uint32_t top_p = mz_redux_max(kp);
    const bool in_p = (in_s && kp == top_p);
    return fc + (int)mz_redux_min(in_p ? (uint32_t)my : 0xffffffffu);
#else
    float sum_win = 0.0f, sum_n = 0.0f, best_s = 0.0f, best_p = 0.0f;
    int best_i = -1;
    for (int k = 0; k < n; ++k) {
        const int i = v.idx[k];
        const mz_hot c = mz_load_hot(hot + fc + i);
        const float qv = mz_normalized_mean(d, qb, fc + i, c.mean, c.count, child_player);
        sum_win = mz_fadd(sum_win, qv);
        sum_n = mz_fadd(sum_n, 1.0f);
        const double num = mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n);
        const float score = mz_fadd((float)mz_ddiv(num, (double)mz_fadd(1.0f, c.count)), qv);
        if (best_i < 0 || score > best_s || (score == best_s && c.policy > best_p)) { best_s = score, best_p = c.policy, best_i = i; }
    }
    if (v.fu < nc) {
        const mz_hot c = mz_load_hot(hot + fc + v.fu);
        const float init_q = mz_init_q(d, sum_win, sum_n);
        const float score = mz_fadd((float)mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n), init_q);
        if (best_i < 0 || score > best_s || (score == best_s && (c.policy > best_p || (c.policy == best_p && (int)v.fu < best_i)))) { best_i = v.fu; }
    }
    return fc + best_i;
#endif
}

// the same by ONE thread (level-parallel re-evaluation of deep paths): all records are requested before the first is used
MZ_DEV int mz_select_level_vis_serial(const mz_dims& d, const mz_state& s, const mz_qb& qb, const mz_hot* hot, const mz_hot& h, const mz_vis& v, int child_player)
{
    const int nc = (int)(h.link >> MZ_LINK_SHIFT), fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
    const int total = (int)mz_fsub(h.count, 1.0f);
    const float bias = s.puct_bias[total];
    const double sqrt_n = s.sqrt_table[total];
    const int n = v.n;
    mz_hot c[MZ_VIS_MAX + 1];
#pragma unroll
    for (int k = 0; k < MZ_VIS_MAX; ++k) {
        c[k] = h;
        if (k < n) { c[k] = mz_load_hot(hot + fc + v.idx[k]); }
    }
    c[MZ_VIS_MAX] = h;
    if (v.fu < nc) { c[MZ_VIS_MAX] = mz_load_hot(hot + fc + v.fu); }
    float sum_win = 0.0f, sum_n = 0.0f, best_s = 0.0f, best_p = 0.0f;
    int best_i = -1;
#pragma unroll
    for (int k = 0; k < MZ_VIS_MAX; ++k) {
        if (k < n) {
            const float qv = mz_normalized_mean(d, qb, fc + v.idx[k], c[k].mean, c[k].count, child_player);
            sum_win = mz_fadd(sum_win, qv);
            sum_n = mz_fadd(sum_n, 1.0f);
            const double num = mz_dmul((double)mz_fmul(bias, c[k].policy), sqrt_n);
            const float score = mz_fadd((float)mz_ddiv(num, (double)mz_fadd(1.0f, c[k].count)), qv);
            if (best_i < 0 || score > best_s || (score == best_s && c[k].policy > best_p)) { best_s = score, best_p = c[k].policy, best_i = v.idx[k]; }
        }
    }
    if (v.fu < nc) {
        const float p_fu = c[MZ_VIS_MAX].policy;
        const float init_q = mz_init_q(d, sum_win, sum_n);
        const float score = mz_fadd((float)mz_dmul((double)mz_fmul(bias, p_fu), sqrt_n), init_q);
        if (best_i < 0 || score > best_s || (score == best_s && (p_fu > best_p || (p_fu == best_p && (int)v.fu < best_i)))) { best_i = v.fu; }
    }
    return fc + best_i;
}

// a child of `parent` received its first visit: keep the parent's list in child order, move first-unvisited on
MZ_DEV void mz_vis_insert(const mz_state& s, size_t base, int parent, int child_index, int num_children)
{
    mz_vis v = mz_load_vis(s.vis + base + parent);
    if (v.n < MZ_VIS_MAX) {
        int k = v.n;
        while (k > 0 && v.idx[k - 1] > child_index) {
            v.idx[k] = v.idx[k - 1];
            --k;
        }
        v.idx[k] = (uint16_t)child_index;
    }
    if (v.n <= MZ_VIS_MAX) { ++v.n; } // MZ_VIS_MAX + 1 = overflowed: selection scans this node's children from now on
    if (v.n <= MZ_VIS_MAX) {
        int fu = v.fu;
        for (int k = 0; k < v.n; ++k) { // idx[] ascending: one pass finds the first index not in the list at or after fu
            if (v.idx[k] == fu) { ++fu; }
        }
        v.fu = (uint16_t)fu;
    }
    (void)num_children;
    mz_store_vis(s.vis + base + parent, v);
}

// ---- MCTS::tree_value_bound_ (mcts.h:117, mcts.cpp:219-228): a std::map<float, int> used as a multiset of the Q values of the
// evaluated nodes. Kept per game as unordered (key, multiplicity) arrays; the map's only observable properties are its size,
// its smallest / largest key and whether a key is present. Keys compare with == (what !(a < b) && !(b < a) is for non-NaN floats).

// smallest / largest key and the number of keys of game g -> qb (warp collective; every lane returns the same)
MZ_DEV mz_qb mz_vb_bounds(const mz_dims& d, const mz_state& s, int g, int lane)
{
    mz_qb qb;
    qb.reward = (s.reward ? s.reward + (size_t)g * d.NP : nullptr);
    qb.lo = 0.0f, qb.hi = 0.0f, qb.n = 0;
    if (!d.value_rescale) { return qb; }
    const float* key = s.vb_key + (size_t)g * d.vb_cap;
    const int n = s.vb_n[g];
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
    for (int i = lane; i < n; i += MZ_W) {
        const float k = key[i];
        lo = (k < lo ? k : lo), hi = (k > hi ? k : hi);
    }
#if MZ_W > 1
    for (int o = 16; o > 0; o >>= 1) {
        const float l2 = __shfl_xor_sync(MZ_FULL, lo, o), h2 = __shfl_xor_sync(MZ_FULL, hi, o);
        lo = (l2 < lo ? l2 : lo), hi = (h2 > hi ? h2 : hi);
    }
#endif
    qb.lo = lo, qb.hi = hi, qb.n = n;
    return qb;
}

// index of `x` among the n keys, or -1 (warp collective)
MZ_DEV int mz_vb_find(const float* key, int n, float x, int lane)
{
    for (int base = 0; base < n; base += MZ_W) {
        const int i = base + lane;
        const unsigned m = mz_ballot(i < n && key[i] == x);
        if (m) { return base + mz_ffs0(m); }
    }
    return -1;
}

// MCTS::updateTreeValueBound (mcts.cpp:219-228): one less of the old value when it is present (erased at zero), one more of the new
MZ_DEV void mz_vb_update(const mz_dims& d, const mz_state& s, int g, float old_value, float new_value, int lane)
{
    float* key = s.vb_key + (size_t)g * d.vb_cap;
    int32_t* cnt = s.vb_cnt + (size_t)g * d.vb_cap;
    int n = s.vb_n[g];
    int pos = mz_vb_find(key, n, old_value, lane);
    if (pos >= 0) {
        const int left = cnt[pos] - 1;
        mz_sync();
        if (left == 0) {
            if (lane == 0) { key[pos] = key[n - 1], cnt[pos] = cnt[n - 1]; }
            --n;
        } else if (lane == 0) {
            cnt[pos] = left;
        }
        mz_sync();
    }
    pos = mz_vb_find(key, n, new_value, lane);
    if (pos >= 0) {
        if (lane == 0) { cnt[pos] += 1; }
    } else if (n < d.vb_cap) { // cannot overflow: every evaluated node holds at most one key (S + 1 nodes)
        if (lane == 0) { key[n] = new_value, cnt[n] = 1; }
        ++n;
    }
    if (lane == 0) { s.vb_n[g] = n; }
    mz_sync();
}

// MCTS::select (mcts.cpp:139-148): returns the path length (valid in warp 0); path[] holds node indices from the root.
//
// Every level's choice depends only on that level's node, so a GUESSED path can be checked level-parallel: the block
// re-evaluates all levels of the guess at once (the root by a whole warp, the deeper levels warp-per-level when the
// path is short and thread-per-level when it is long), finds the first level whose choice differs, takes the
// re-evaluated child there and extends the guess below it along the `last_child` hints (the child chosen the last time
// selection passed through a node). The loop ends when a guess is confirmed down to a node without hint, from where
// warp 0 finishes serially (normally the one freshly expanded leaf). The first guess is the previous simulation's
// path. Every choice is made by mz_select_level / _serial on the current statistics: the path is exactly the serial one.
MZ_DEV int mz_select(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int root_turn, int lane, int wid, int nw)
{
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    int32_t* path = s.path + (size_t)g * (d.S + 2);
    int32_t* last_child = s.last_child + (size_t)g * d.NP;
    float* q = w->q_warp + (size_t)wid * d.A;
    const int tid = wid * MZ_W + lane;
    if (wid == 0) { // value bounds are fixed during a selection: read once (uniform for the block after the first sync below)
        const mz_qb qb = mz_vb_bounds(d, s, g, lane);
        if (lane == 0) { w->qb = qb; }
    }
    int glen = s.spec_len[g]; // length of the guessed path (uniform across the block)
    if (glen == 0) {
        if (tid == 0) { path[0] = 0; }
        glen = 1;
    }
    int start = 0; // levels below `start` are confirmed
    long long t_verify = 0, t_chase = 0, n_rounds = 0, n_checked = 0;
    for (;;) {
        if (tid == 0) { w->mismatch = glen - 1; }
        mz_block_sync();
        const long long tv0 = mz_clock();
        ++n_rounds;
        n_checked += (glen - 1 - start > 0 ? glen - 1 - start : 0);
        // ---- check levels start .. glen-2 of the guess
        const int nlev = glen - 1 - start;
        // all node records of the guess at once, and their children blocks on their way into L2, so that the per-level
        // evaluations below do not each pay two dependent misses
        for (int j = start + tid; j < glen - 1; j += nw * MZ_W) {
            const mz_hot h = mz_load_hot(hot + path[j]);
            w->lvl_h[j] = h;
            const int cnc = (int)(h.link >> MZ_LINK_SHIFT), cfc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
            bool listed = false;
            if (s.vis && j > 0) {
                const mz_vis v = mz_load_vis(s.vis + (size_t)g * d.NP + path[j]);
                w->lvl_v[j] = v;
                listed = (v.n <= MZ_VIS_MAX);
                if (listed) { // only these records will be read
                    for (int l = 0; l < v.n; ++l) { mz_prefetch(hot + cfc + v.idx[l]); }
                    if (v.fu < cnc) { mz_prefetch(hot + cfc + v.fu); }
                }
            }
            if (!listed) {
                for (int l = 0; l < cnc; l += 8) { mz_prefetch(hot + cfc + l); }
            }
        }
        mz_block_sync();
        if (nlev > 0) {
            const bool by_thread = (nw > 1 && nlev > 3 * nw);
            if (by_thread) {
                const int root_warp = nw - 1;
                if (start == 0 && wid == root_warp) {
                    const mz_hot h = w->lvl_h[0];
                    mz_hot c;
                    const int chosen = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u)) + mz_select_level(d, s, w->qb, hot, h, true, mz_child_player(d, root_turn, 0), q, lane, c);
                    if (lane == 0) {
                        w->sel[0] = chosen;
                        last_child[path[0]] = chosen;
                        if (chosen != path[1]) { mz_atomic_min(&w->mismatch, 0); }
                    }
                }
                if (wid != root_warp) {
                    for (int j = (start == 0 ? 1 : start) + tid; j < glen - 1; j += (nw - 1) * MZ_W) {
                        const int node = path[j];
                        const int cp = mz_child_player(d, root_turn, j);
                        const int chosen = (s.vis && w->lvl_v[j].n <= MZ_VIS_MAX ? mz_select_level_vis_serial(d, s, w->qb, hot, w->lvl_h[j], w->lvl_v[j], cp)
                                                                                  : mz_select_level_serial(d, s, w->qb, hot, w->lvl_h[j], cp));
                        w->sel[j] = chosen;
                        last_child[node] = chosen;
                        if (chosen != path[j + 1]) { mz_atomic_min(&w->mismatch, j); }
                    }
                }
            } else {
                for (int j = start + wid; j < glen - 1; j += nw) {
                    const int node = path[j];
                    const mz_hot h = w->lvl_h[j];
                    mz_hot c;
                    const int cp = mz_child_player(d, root_turn, j);
                    const int chosen = (s.vis && j > 0 && w->lvl_v[j].n <= MZ_VIS_MAX)
                                           ? mz_select_level_vis(d, s, w->qb, hot, h, w->lvl_v[j], cp, lane)
                                           : (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u)) + mz_select_level(d, s, w->qb, hot, h, j == 0, cp, q, lane, c);
                    if (lane == 0) {
                        w->sel[j] = chosen;
                        last_child[node] = chosen;
                        if (chosen != path[j + 1]) { mz_atomic_min(&w->mismatch, j); }
                    }
                }
            }
        }
        mz_block_sync();
        const long long tv1 = mz_clock();
        t_verify += tv1 - tv0;
        // ---- take the re-evaluated child at the first changed level, extend the guess along the hints (thread 0)
        if (tid == 0) {
            int level = w->mismatch;
            if (level < glen - 1) {
                ++level;
                path[level] = w->sel[level - 1];
            }
            const int confirmed = level; // path[0 .. level] is now the true prefix
            int node = path[level];
            for (;;) {
                const int lc = last_child[node];
                if (lc < 0) { break; }
                node = lc;
                path[++level] = node;
            }
            w->shared_len = level + 1;
            w->flag = confirmed;
        }
        mz_block_sync();
        const int confirmed = w->flag, new_len = w->shared_len;
        mz_block_sync();
        t_chase += mz_clock() - tv1;
        if (new_len - 1 == confirmed) { // nothing left to check: finish serially from path[confirmed]
            glen = new_len;
            break;
        }
        start = confirmed;
        glen = new_len;
    }
    int len = glen;
    const long long ts0 = mz_clock();
    if (wid == 0) {
        int level = glen - 1;
        mz_hot h = mz_load_hot(hot + path[level]);
        while ((h.link >> MZ_LINK_SHIFT) != 0) {
            const int fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
            mz_hot c;
            const int best = mz_select_level(d, s, w->qb, hot, h, level == 0, mz_child_player(d, root_turn, level), q, lane, c);
            if (lane == 0) {
                last_child[path[level]] = fc + best;
                path[level + 1] = fc + best;
            }
            h = c;
            ++level;
            mz_sync();
        }
        len = level + 1;
        if (s.dbg && lane == 0) {
            unsigned long long* o = s.dbg + (size_t)g * 16;
            o[8] += (unsigned long long)t_verify, o[9] += (unsigned long long)t_chase, o[10] += (unsigned long long)(mz_clock() - ts0);
            o[11] += (unsigned long long)n_rounds, o[12] += (unsigned long long)n_checked, o[13] += (unsigned long long)(len - glen);
        }
    }
    return len;
}

// ---- console think() (zero_actor.cpp:129-157): selection under virtual loss ----
// MCTS::selectChildByPUCTScore with virtual losses (mcts.cpp:181-217, 40-61): a child counts as visited when count + virtual loss != 0, its Q is
// (q * count - vloss) / (count + vloss), the exploration term divides by 1 + count + vloss and the parent's total is count + vloss - 1. Every child
// is scored (a virtual loss breaks the "unvisited children share one score" shortcut of mz_select_level); ordered f32 sum for init-Q as there.
#define MZ_Q_NONE 3.0e38f
MZ_DEV int mz_select_level_think(const mz_dims& d, const mz_state& s, const mz_qb& qb, const mz_hot* hot, const float* vl, int node, const mz_hot& h, int child_player,
                                 float* q, int lane)
{
    const int nc = (int)(h.link >> MZ_LINK_SHIFT), fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
    const int total = (int)mz_fsub(mz_fadd(h.count, vl[node]), 1.0f); // getCountWithVirtualLoss() - 1, mcts.cpp:185
    const float bias = s.puct_bias[total];
    const double sqrt_n = s.sqrt_table[total];
    for (int i = lane; i < nc; i += MZ_W) {
        const mz_hot c = mz_load_hot(hot + fc + i);
        const float v = vl[fc + i];
        q[i] = (mz_fadd(c.count, v) != 0.0f ? mz_normalized_mean(d, qb, fc + i, c.mean, c.count, child_player, v) : MZ_Q_NONE);
    }
    mz_sync();
    float sum_win = 0.0f, sum_n = 0.0f;
    for (int i = 0; i < nc; ++i) { // every lane folds the same ordered sum
        const float x = q[i];
        if (x != MZ_Q_NONE) { sum_win = mz_fadd(sum_win, x), sum_n = mz_fadd(sum_n, 1.0f); }
    }
    const float init_q = mz_init_q(d, sum_win, sum_n);
    float best_s = 0.0f, best_p = 0.0f;
    int best_i = -1;
    for (int i = lane; i < nc; i += MZ_W) {
        const mz_hot c = mz_load_hot(hot + fc + i);
        const float cv = mz_fadd(c.count, vl[fc + i]);
        const float u = (float)mz_ddiv(mz_dmul((double)mz_fmul(bias, c.policy), sqrt_n), (double)mz_fadd(1.0f, cv));
        const float score = mz_fadd(u, (cv == 0.0f ? init_q : q[i]));
        if (best_i < 0 || score > best_s || (score == best_s && c.policy > best_p)) { best_s = score, best_p = c.policy, best_i = i; }
    }
#if MZ_W > 1
    {   // lexicographic arg-max (score desc, prior desc, index asc) over the lanes' candidates (mcts.cpp:187-194)
        const int mine = best_i;
        const uint32_t ks = (mine >= 0 ? mz_sortable(best_s) : 0u);
        const uint32_t top_s = mz_redux_max(ks);
        const bool in_s = (mine >= 0 && ks == top_s);
        const uint32_t kp = (in_s ? mz_sortable(best_p) : 0u);
        const uint32_t top_p = mz_redux_max(kp);
        const bool in_p = (in_s && kp == top_p);
        best_i = (int)mz_redux_min(in_p ? (uint32_t)mine : 0xffffffffu);
    }
#endif
    mz_sync();
    return best_i;
}

// MCTS::select (mcts.cpp:139-148) under virtual loss, one warp, level by level; returns the path length, path[] in s.path
MZ_DEV int mz_select_think(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int root_turn, int lane)
{
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    const float* vl = s.vloss + (size_t)g * d.NP;
    int32_t* path = s.path + (size_t)g * (d.S + 2);
    {
        const mz_qb qb = mz_vb_bounds(d, s, g, lane);
        if (lane == 0) { w->qb = qb, path[0] = 0; }
    }
    mz_sync();
    int level = 0, node = 0;
    mz_hot h = mz_load_hot(hot);
    while ((h.link >> MZ_LINK_SHIFT) != 0) {
        const int fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
        node = fc + mz_select_level_think(d, s, w->qb, hot, vl, node, h, mz_child_player(d, root_turn, level), w->q_warp, lane);
        h = mz_load_hot(hot + node);
        ++level;
        if (lane == 0) { path[level] = node; }
        mz_sync();
    }
    return level + 1;
}

MZ_DEV void mz_slot_store(const mz_dims& d, const mz_state& s, int g, int slot, const mz_scratch* w, int lane)
{
    const size_t e = (size_t)g * (d.S + 1) + slot;
    uint32_t* st = s.slot_st + e * 2 * d.N;
    for (int i = lane; i < 2 * d.N; i += MZ_W) { st[i] = w->st[i / d.N][i % d.N]; }
    if (lane == 0) {
        s.slot_hash[e] = w->hash;
        s.slot_meta[e * 4 + 0] = w->turn, s.slot_meta[e * 4 + 1] = w->num_moves, s.slot_meta[e * 4 + 2] = w->last, s.slot_meta[e * 4 + 3] = w->last2;
    }
}

MZ_DEV void mz_slot_load(const mz_dims& d, const mz_state& s, int g, int slot, mz_scratch* w, int lane)
{
    const size_t e = (size_t)g * (d.S + 1) + slot;
    const uint32_t* st = s.slot_st + e * 2 * d.N;
    for (int i = lane; i < 2 * d.N; i += MZ_W) { w->st[i / d.N][i % d.N] = st[i]; }
    if (lane == 0) {
        w->hash = s.slot_hash[e];
        w->turn = s.slot_meta[e * 4 + 0], w->num_moves = s.slot_meta[e * 4 + 1], w->last = s.slot_meta[e * 4 + 2], w->last2 = s.slot_meta[e * 4 + 3];
    }
    mz_sync();
}


// GumbelZero::sortCandidatesByScore (gumbel_zero.cpp:120-137): candidates by descending
// logit + (c_visit + max child count) * c_scale * q, unvisited ones last. One thread; the lists are short
// (actor_gumbel_sample_size) and this runs only at a halving and at the move decision.
MZ_DEV void mz_gumbel_sort_by_score(const mz_dims& d, const mz_state& s, int g, int root_turn)
{
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    int32_t* cand = s.gum_cand + (size_t)g * d.A;
    const int n = s.gum_meta[g * 4 + 0];
    const mz_hot root = mz_load_hot(hot);
    const int nc = (int)(root.link >> MZ_LINK_SHIFT), fc = (int)(root.link & ((1u << MZ_LINK_SHIFT) - 1u));
    float max_count = 0.0f;
    for (int i = 0; i < nc; ++i) {
        const float c = mz_load_hot(hot + fc + i).count;
        max_count = (c > max_count ? c : max_count);
    }
    const float scale = mz_fmul(mz_fadd(d.sigma_visit_c, max_count), d.sigma_scale_c);
    mz_qb qb; // one thread: the bounds are read serially (getNormalizedMean(tree_value_bound), gumbel_zero.cpp:128)
    qb.reward = (s.reward ? s.reward + (size_t)g * d.NP : nullptr), qb.lo = 0.0f, qb.hi = 0.0f, qb.n = 0;
    if (d.value_rescale) {
        const float* key = s.vb_key + (size_t)g * d.vb_cap;
        qb.n = s.vb_n[g];
        for (int i = 0; i < qb.n; ++i) { qb.lo = (i == 0 || key[i] < qb.lo ? key[i] : qb.lo), qb.hi = (i == 0 || key[i] > qb.hi ? key[i] : qb.hi); }
    }
    root_turn = mz_child_player(d, root_turn, 0);
    for (int i = 1; i < n; ++i) { // stable insertion sort, descending score
        const int c = cand[i];
        const mz_hot hc = mz_load_hot(hot + c);
        const float sc = (hc.count > 0.0f ? mz_fadd(s.logit[(size_t)g * d.NP + c], mz_fmul(scale, mz_normalized_mean(d, qb, c, hc.mean, hc.count, root_turn))) : -3.402823466e+38f);
        int j = i;
        while (j > 0) {
            const int p = cand[j - 1];
            const mz_hot hp = mz_load_hot(hot + p);
            const float sp = (hp.count > 0.0f ? mz_fadd(s.logit[(size_t)g * d.NP + p], mz_fmul(scale, mz_normalized_mean(d, qb, p, hp.mean, hp.count, root_turn))) : -3.402823466e+38f);
            if (!(sp < sc)) { break; }
            cand[j] = p;
            --j;
        }
        cand[j] = c;
    }
}

// GumbelZero::sequentialHalving (gumbel_zero.cpp:87-118), after every backup. Block collective.
MZ_DEV void mz_gumbel_halving(const mz_dims& d, const mz_state& s, int g, int root_turn, int tid, int nthreads)
{
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    int32_t* cand = s.gum_cand + (size_t)g * d.A;
    int32_t* meta = s.gum_meta + g * 4;
    const mz_hot root = mz_load_hot(hot);
    const int nc = (int)(root.link >> MZ_LINK_SHIFT), fc = (int)(root.link & ((1u << MZ_LINK_SHIFT) - 1u));
    if ((int)root.count == 1) { // first call of a search: the m root children with the largest (noisy) logits
        const float* lg = s.logit + (size_t)g * d.NP + fc;
        for (int i = tid; i < nc; i += nthreads) {
            const float l = lg[i];
            int rank = 0;
            for (int j = 0; j < nc; ++j) { rank += (lg[j] > l || (lg[j] == l && j < i)) ? 1 : 0; }
            if (rank < d.gumbel_m) { cand[rank] = fc + i; }
        }
        if (tid == 0) { meta[0] = (nc < d.gumbel_m ? nc : d.gumbel_m), meta[1] = d.gumbel_m, meta[2] = d.gumbel_budget0, meta[3] = 0; }
        mz_block_sync();
        return;
    }
    if (tid == 0) {
        const int n = meta[0];
        bool all = true;
        for (int i = 0; i < n && all; ++i) { all = (mz_load_hot(hot + cand[i]).count >= (float)meta[2]); }
        if (all) {
            const int level = meta[3];
            const int next_budget = (level < MZ_GUMBEL_LEVELS ? d.gumbel_next[level] : 0);
            if (next_budget > 0 && meta[1] > 2) {
                meta[1] /= 2;
                meta[3] = level + 1;
                mz_gumbel_sort_by_score(d, s, g, root_turn);
                if (n > meta[1]) { meta[0] = meta[1]; }
                meta[2] = (int)mz_fadd(mz_load_hot(hot + cand[0]).count, (float)next_budget);
            }
        }
    }
    mz_block_sync();
}

// GumbelZero::selection's root-level choice (gumbel_zero.cpp:76-81): least visited candidate, ties to the larger logit
MZ_DEV int mz_gumbel_pick(const mz_dims& d, const mz_state& s, int g)
{
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    const int32_t* cand = s.gum_cand + (size_t)g * d.A;
    const int n = s.gum_meta[g * 4 + 0];
    int best = cand[0];
    float bc = mz_load_hot(hot + best).count, bl = s.logit[(size_t)g * d.NP + best];
    for (int i = 1; i < n; ++i) {
        const int c = cand[i];
        const float cc = mz_load_hot(hot + c).count, cl = s.logit[(size_t)g * d.NP + c];
        if (cc < bc || (cc == bc && cl > bl)) { best = c, bc = cc, bl = cl; }
    }
    return best;
}

// ZeroActor::beforeNNEvaluation, MuZero branch (zero_actor.cpp:51-72): selection (PUCT, or Gumbel at the root level), then
// either the root position's feature planes (initial inference) or the parent's hidden state + the leaf's action plane
// (recurrent inference) are laid out as the network's input rows. No environment exists below the root.
MZ_DEV void mz_before_nn_muzero(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int lane, int wid, int nw)
{
    const int N = d.N, tid = wid * MZ_W + lane, nthreads = nw * MZ_W;
    const int root_turn = s.root_meta[g * 4 + 0];
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    int32_t* path = s.path + (size_t)g * (d.S + 2);
    const int sims = (int)mz_load_hot(hot).count; // MCTS::getNumSimulation, mcts.h:100
    int slot_of_eval = sims; // hidden states are stored in evaluation order (zero_actor.cpp:90)
    if (s.vloss) { // one lane of a batched think() step (zero_actor.cpp:129-145); the root's initial inference is a batch of one (:134-135)
        if (s.think_lane >= (sims == 0 ? 1 : d.S + 1 - sims)) {
            if (tid == 0) { s.path_len[g] = 0; }
            return;
        }
        if (wid == 0) {
            const int len0 = mz_select_think(d, s, g, w, root_turn, lane);
            float* vl = s.vloss + (size_t)g * d.NP;
            const int dup = (vl[path[len0 - 1]] != 0.0f); // selected earlier in this step: evaluated once (:140-142)
            mz_sync();
            for (int i = lane; i < len0; i += MZ_W) { vl[path[i]] = mz_fadd(vl[path[i]], 1.0f); }
            if (lane == 0) {
                w->shared_len = (dup ? -len0 : len0);
                w->shared_count = sims + s.think_pending[g];
                if (!dup) { s.think_pending[g] += 1; }
            }
        }
        mz_block_sync();
        if (w->shared_len < 0) {
            if (tid == 0) { s.path_len[g] = w->shared_len; }
            return;
        }
        slot_of_eval = w->shared_count;
    } else if (d.gumbel && sims > 0) {
        if (wid == 0) {
            int level = 1;
            const mz_qb qb = mz_vb_bounds(d, s, g, lane);
            if (lane == 0) {
                w->qb = qb;
                path[0] = 0;
                path[1] = mz_gumbel_pick(d, s, g);
            }
            mz_sync();
            mz_hot h = mz_load_hot(hot + path[1]);
            while ((h.link >> MZ_LINK_SHIFT) != 0) { // MCTS::selectFromNode below the candidate, mcts.cpp:139-148
                const int fc = (int)(h.link & ((1u << MZ_LINK_SHIFT) - 1u));
                mz_hot c;
                const int best = mz_select_level(d, s, w->qb, hot, h, false, mz_child_player(d, root_turn, level), w->q_warp, lane, c);
                if (lane == 0) { path[level + 1] = fc + best; }
                h = c;
                ++level;
                mz_sync();
            }
            if (lane == 0) { w->shared_len = level + 1; }
        }
    } else {
        const int len0 = mz_select(d, s, g, w, root_turn, lane, wid, nw);
        if (wid == 0 && lane == 0) { w->shared_len = len0; }
    }
    mz_block_sync();
    const int len = w->shared_len, L = len - 1, leaf = path[L];
    int num_legal = d.A;
    if (L == 0 && d.game == MZ_GAME_ATARI) {
        // the planes come from the screen ring (packed by its own kernel, engine.cu); legal actions = the game's minimal action set
        for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { w->legal[i] = (i == 0 ? d.legal_mask : 0u); }
        num_legal = mz_popc(d.legal_mask);
    } else if (L == 0) { // initial inference: env_.getFeatures() and the root's legal set (zero_actor.cpp:60,238)
        if (wid == 0) { mz_env_load_root(d, s, g, w, lane); }
        mz_block_sync();
        const uint64_t* root_list = s.hashes + (size_t)g * d.max_hashes;
        num_legal = mz_env_legal_block(d, s, w, root_list, s.root_meta[g * 4 + 1], root_list, 0, tid, nthreads);
        mz_env_features(d, s, g, w, 0, tid, nthreads);
    } else {
        for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { // every action is expanded below the root (zero_actor.cpp:238)
            const int lo = i * 32;
            w->legal[i] = (d.A >= lo + 32 ? 0xffffffffu : (d.A > lo ? ((1u << (d.A - lo)) - 1u) : 0u));
        }
        const int pslot = s.node_slot[(size_t)g * d.NP + path[L - 1]];
        const int a = s.action[(size_t)g * d.NP + leaf];
#if MZ_W > 1
        if (s.hid) { // recurrent inference input: parent's hidden state (zero_actor.cpp:65) + getActionFeatures(leaf action) (:66)
            const uint4* src = reinterpret_cast<const uint4*>(s.hid + ((size_t)g * (d.S + 1) + pslot) * N * N * d.hid_c);
            uint16_t* dst = s.dyn_in + (size_t)g * d.slots * d.dyn_c;
            const int per_cell = d.hid_c / 8; // 16-byte chunks
            for (int i = tid; i < N * N * per_cell; i += nthreads) {
                const int cell = i / per_cell, k = i - cell * per_cell;
                *reinterpret_cast<uint4*>(dst + (size_t)((cell / N + 1) * (N + 1) + cell % N) * d.dyn_c + k * 8) = src[i];
            }
            mz_block_sync(); // the action planes may overwrite padded (zero) hidden columns copied above
            if (d.act_planes > 1) { // Atari: plane `a` of the 18 action planes is all ones (atari.cpp:124-130)
                for (int i = tid; i < N * N * d.act_planes; i += nthreads) {
                    const int cell = i / d.act_planes, k = i - cell * d.act_planes;
                    dst[(size_t)((cell / N + 1) * (N + 1) + cell % N) * d.dyn_c + d.act_col + k] = (k == a ? MZ_HALF_ONE : 0);
                }
            } else {
                for (int cell = tid; cell < N * N; cell += nthreads) { // one-hot plane; all zero for a pass (othello.cpp:257-262)
                    dst[(size_t)((cell / N + 1) * (N + 1) + cell % N) * d.dyn_c + d.act_col] = (cell == a ? MZ_HALF_ONE : 0);
                }
            }
        }
#endif
        if (tid == 0 && s.leaf_parent) { s.leaf_parent[g * 2 + 0] = pslot, s.leaf_parent[g * 2 + 1] = a; }
    }
    mz_block_sync();
    for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { s.leaf_legal[g * MZ_LEGAL_WORDS + i] = w->legal[i]; }
    if (tid == 0) {
        s.node_slot[(size_t)g * d.NP + leaf] = (int16_t)slot_of_eval;
        if (s.eval_slot) { s.eval_slot[g] = slot_of_eval; }
        if (L == 0 && s.leaf_parent) { s.leaf_parent[g * 2 + 0] = -1, s.leaf_parent[g * 2 + 1] = -1; }
        s.path_len[g] = len;
        s.spec_len[g] = (d.gumbel ? 0 : len);
        s.leaf_meta[g * 4 + 0] = 0;
        s.leaf_meta[g * 4 + 1] = (d.num_players == 1 ? 1 : ((L & 1) ? 3 - root_turn : root_turn));
        s.leaf_meta[g * 4 + 2] = 0;
        s.leaf_meta[g * 4 + 3] = num_legal;
        s.leaf_score[g] = 0.0f;
    }
}

// One "before NN evaluation" step of game g (zero_actor.cpp:51-58): select, transition, analyse the leaf
// (terminal / score / legal set) and emit its feature planes.
//
// The reference rebuilds the leaf position by copying the root environment and replaying the whole path
// (zero_actor.cpp:247-252). Here every evaluated node keeps its position (stone rows, hash, last two actions) in a
// slot indexed by the simulation that evaluated it, so the transition is "parent's slot + one move"; the 8-position
// feature history and the superko hash list are gathered from the slots of the nodes on the path (and from the root
// environment for positions older than the root). Same positions, same results, cost independent of the depth.
MZ_DEV void mz_before_nn(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int lane, int wid, int nw)
{
    if (d.muzero) {
        mz_before_nn_muzero(d, s, g, w, lane, wid, nw);
        return;
    }
    const int N = d.N, tid = wid * MZ_W + lane, nthreads = nw * MZ_W;
    const int root_turn = s.root_meta[g * 4 + 0], root_moves = s.root_meta[g * 4 + 1];
    const long long t0 = mz_clock();
    if (s.vloss) { // one lane of a batched think() step (zero_actor.cpp:129-145)
        const int done = (int)mz_load_hot(s.hot + (size_t)g * d.NP).count;
        if (s.think_lane >= d.S + 1 - done) { // batch_size = min(K, simulations left), :133-135
            if (tid == 0) { s.path_len[g] = 0; }
            return;
        }
        if (wid == 0) {
            const int len0 = mz_select_think(d, s, g, w, root_turn, lane);
            float* vl = s.vloss + (size_t)g * d.NP;
            const int32_t* sel = s.path + (size_t)g * (d.S + 2);
            // a leaf that already carries a virtual loss was selected earlier in this step: it is evaluated once (:140-142) ...
            const int dup = (vl[sel[len0 - 1]] != 0.0f);
            mz_sync();
            for (int i = lane; i < len0; i += MZ_W) { vl[sel[i]] = mz_fadd(vl[sel[i]], 1.0f); } // ... but every selection leaves its virtual loss (:143)
            if (lane == 0) {
                w->shared_len = (dup ? -len0 : len0);
                w->shared_count = done + s.think_pending[g]; // slot of this evaluation: finished simulations + leaves queued before it in this step
                if (!dup) { s.think_pending[g] += 1; }
            }
        }
        mz_block_sync();
        if (w->shared_len < 0) { // duplicate: nothing to evaluate, the lane reports -length
            if (tid == 0) { s.path_len[g] = w->shared_len; }
            return;
        }
    } else {
        const int len0 = mz_select(d, s, g, w, root_turn, lane, wid, nw);
        if (wid == 0 && lane == 0) {
            w->shared_len = len0;
            w->shared_count = (int)mz_load_hot(s.hot + (size_t)g * d.NP).count; // simulations finished so far
        }
    }
    mz_block_sync();
    const long long t1 = mz_clock();
    const int len = w->shared_len, slot = w->shared_count;
    const int32_t* path = s.path + (size_t)g * (d.S + 2);
    const int16_t* node_slot = s.node_slot + (size_t)g * d.NP;
    const uint64_t* root_list = s.hashes + (size_t)g * d.max_hashes;
    const int L = len - 1; // depth of the leaf
    const int leaf = path[L];
    // ---- gather (all threads): hashes of the path nodes 1 .. L-1, the ancestors' positions for the feature history,
    //      and the parent's position
    for (int j = 1 + tid; j < L; j += nthreads) { w->path_hashes[j - 1] = s.slot_hash[(size_t)g * (d.S + 1) + node_slot[path[j]]]; }
    if (L == 0) {
        if (wid == 0) { mz_env_load_root(d, s, g, w, lane); }
    } else {
        for (int i = tid; i < (MZ_HIST - 1) * 2 * N; i += nthreads) {
            const int k = 1 + i / (2 * N), e = i % (2 * N);
            const int pos = root_moves + L - 1 - k; // index of the position k moves before the leaf's
            if (pos < 0) { continue; }
            const int ring = pos % MZ_HIST;
            uint32_t v;
            if (pos >= root_moves) {
                v = s.slot_st[((size_t)g * (d.S + 1) + node_slot[path[L - k]]) * 2 * N + e];
            } else {
                v = s.root_hist[((size_t)g * MZ_HIST + ring) * 2 * MZ_ROWS + (e / N) * MZ_ROWS + e % N];
            }
            w->hist[ring][e / N][e % N] = v;
        }
        if (L == 1) {
            const uint32_t* st = s.root_st + (size_t)g * 2 * MZ_ROWS;
            for (int i = tid; i < 2 * N; i += nthreads) { w->st[i / N][i % N] = st[(i / N) * MZ_ROWS + i % N]; }
            if (tid == 0) {
                w->hash = s.root_hash[g];
                w->turn = root_turn, w->num_moves = root_moves, w->last = s.root_meta[g * 4 + 2], w->last2 = s.root_meta[g * 4 + 3];
            }
        } else {
            const size_t e = (size_t)g * (d.S + 1) + node_slot[path[L - 1]];
            for (int i = tid; i < 2 * N; i += nthreads) { w->st[i / N][i % N] = s.slot_st[e * 2 * N + i]; }
            if (tid == 0) {
                w->hash = s.slot_hash[e];
                w->turn = s.slot_meta[e * 4 + 0], w->num_moves = s.slot_meta[e * 4 + 1], w->last = s.slot_meta[e * 4 + 2], w->last2 = s.slot_meta[e * 4 + 3];
            }
        }
    }
    mz_block_sync();
    // ---- transition (warp 0): getEnvironmentTransition's last step, zero_actor.cpp:250
    if (wid == 0 && L > 0) {
        const int a = s.action[(size_t)g * d.NP + leaf];
        mz_env_act(d, s, w, a, w->turn, lane);
        if (lane == 0) { w->path_hashes[L - 1] = w->hash; }
    }
    mz_block_sync();
    {
        const size_t e = (size_t)g * (d.S + 1) + slot;
        for (int i = tid; i < 2 * N; i += nthreads) { s.slot_st[e * 2 * N + i] = w->st[i / N][i % N]; }
        if (tid == 0) {
            s.slot_hash[e] = w->hash;
            s.slot_meta[e * 4 + 0] = w->turn, s.slot_meta[e * 4 + 1] = w->num_moves, s.slot_meta[e * 4 + 2] = w->last, s.slot_meta[e * 4 + 3] = w->last2;
            s.node_slot[(size_t)g * d.NP + leaf] = (int16_t)slot;
        }
    }
    const long long t2 = mz_clock();
    // ---- leaf analysis and feature planes (all threads)
    const int rotation = (s.rotations ? s.rotations[g] : 0);
    int terminal = mz_env_is_terminal(d, w);
    float score = 0.0f;
    int num_legal = 0;
    if (terminal) {
        if (wid == 0) { score = mz_env_eval_score(d, w, lane); }
        for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { w->legal[i] = 0u; }
        mz_block_sync();
    } else {
        num_legal = mz_env_legal_block(d, s, w, root_list, root_moves, w->path_hashes, L, tid, nthreads);
        if (d.game == MZ_GAME_NOGO && num_legal == 0) { // nogo.h:61-68: the game is over when the side to move has no legal move
            terminal = 1;
            score = mz_env_eval_score(d, w, lane);
        }
    }
    const long long t3 = mz_clock();
    mz_env_features(d, s, g, w, rotation, tid, nthreads);
    for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { s.leaf_legal[g * MZ_LEGAL_WORDS + i] = w->legal[i]; }
    if (tid == 0) {
        s.path_len[g] = len;
        s.spec_len[g] = len;
        s.leaf_meta[g * 4 + 0] = terminal;
        s.leaf_meta[g * 4 + 1] = w->turn;
        s.leaf_meta[g * 4 + 2] = rotation;
        s.leaf_meta[g * 4 + 3] = num_legal;
        s.leaf_score[g] = score;
        if (s.dbg) {
            unsigned long long* o = s.dbg + (size_t)g * 16;
            o[0] += (unsigned long long)(t1 - t0), o[1] += (unsigned long long)(t2 - t1), o[2] += (unsigned long long)(t3 - t2);
            o[3] += (unsigned long long)(mz_clock() - t3), o[5] += 1ull, o[6] = ((unsigned long long)len > o[6] ? (unsigned long long)len : o[6]);
            o[7] += (unsigned long long)len;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// std::sort(candidates, lhs.policy_ > rhs.policy_) (zero_actor.cpp:225-227,241-243) EXACTLY as libstdc++ orders them.
// std::sort is unstable: where candidates with equal priors end up is decided by the library's algorithm, and the child
// order is observable (PUCT ties, the record's P[...] tag). libstdc++'s std::__sort (bits/stl_algo.h, unchanged since
// GCC 4.9): introsort — median-of-three quicksort down to ranges of 16 with a depth limit of 2 * floor(log2 n) and heapsort
// beyond it — then one insertion-sort pass. Restated here for ONE thread over elements packed as
// (float bits of the prior) << 32 | action id, given in ascending action id (the order the reference pushes them). Only
// used when the parallel rank sort found an exact tie; without ties every correct sort agrees.
// ---------------------------------------------------------------------------------------------
MZ_DEV float mz_key_policy(uint64_t e)
{
#if MZ_W > 1
    return __uint_as_float((uint32_t)(e >> 32));
#else
    uint32_t u = (uint32_t)(e >> 32);
    float f;
    __builtin_memcpy(&f, &u, 4);
    return f;
#endif
}
MZ_DEV bool mz_cand_gt(uint64_t x, uint64_t y) { return mz_key_policy(x) > mz_key_policy(y); }
MZ_DEV void mz_cand_swap(uint64_t* e, int i, int j)
{
    const uint64_t t = e[i];
    e[i] = e[j];
    e[j] = t;
}
MZ_DEV void mz_cand_adjust_heap(uint64_t* f, int hole, int len, uint64_t value) // stl_heap.h:224-249 + __push_heap :135-148
{
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (mz_cand_gt(f[second], f[second - 1])) { second--; }
        f[hole] = f[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        f[hole] = f[second - 1];
        hole = second - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && mz_cand_gt(f[parent], value)) {
        f[hole] = f[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    f[hole] = value;
}
MZ_DEV void mz_cand_unguarded_linear_insert(uint64_t* e, int last) // stl_algo.h:1792-1807
{
    const uint64_t val = e[last];
    int next = last - 1;
    while (mz_cand_gt(val, e[next])) {
        e[last] = e[next];
        last = next;
        --next;
    }
    e[last] = val;
}
MZ_DEV void mz_cand_insertion_sort(uint64_t* e, int first, int last) // stl_algo.h:1812-1831
{
    if (first == last) { return; }
    for (int i = first + 1; i != last; ++i) {
        if (mz_cand_gt(e[i], e[first])) {
            const uint64_t val = e[i];
            for (int j = i; j != first; --j) { e[j] = e[j - 1]; }
            e[first] = val;
        } else {
            mz_cand_unguarded_linear_insert(e, i);
        }
    }
}
MZ_DEV void mz_std_sort_candidates(uint64_t* e, int n)
{
    if (n <= 0) { return; }
    int lg = 0;
    for (int m = n; m > 1; m >>= 1) { ++lg; }
    // __introsort_loop (stl_algo.h:1918-1935); the recursion on the right part becomes a stack (the two parts are disjoint,
    // so the order in which they are finished does not matter)
    int st_first[40], st_last[40], st_depth[40], sp = 0;
    st_first[0] = 0, st_last[0] = n, st_depth[0] = lg * 2, sp = 1;
    while (sp > 0) {
        --sp;
        const int first = st_first[sp];
        int last = st_last[sp], depth = st_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { // __partial_sort(first, last, last): make_heap + sort_heap (stl_heap.h:340-362,415-427)
                uint64_t* f = e + first;
                const int len = last - first;
                for (int parent = (len - 2) / 2;; --parent) {
                    mz_cand_adjust_heap(f, parent, len, f[parent]);
                    if (parent == 0) { break; }
                }
                for (int l = len - 1; l >= 1; --l) {
                    const uint64_t value = f[l];
                    f[l] = f[0];
                    mz_cand_adjust_heap(f, 0, l, value);
                }
                break;
            }
            --depth;
            // __unguarded_partition_pivot (stl_algo.h:1893-1900): median of (first + 1, mid, last - 1) to first
            const int mid = first + (last - first) / 2, a = first + 1, b = mid, c = last - 1;
            if (mz_cand_gt(e[a], e[b])) {
                if (mz_cand_gt(e[b], e[c])) {
                    mz_cand_swap(e, first, b);
                } else if (mz_cand_gt(e[a], e[c])) {
                    mz_cand_swap(e, first, c);
                } else {
                    mz_cand_swap(e, first, a);
                }
            } else if (mz_cand_gt(e[a], e[c])) {
                mz_cand_swap(e, first, a);
            } else if (mz_cand_gt(e[b], e[c])) {
                mz_cand_swap(e, first, c);
            } else {
                mz_cand_swap(e, first, b);
            }
            int lo = first + 1, hi = last; // __unguarded_partition (stl_algo.h:1871-1888)
            for (;;) {
                while (mz_cand_gt(e[lo], e[first])) { ++lo; }
                --hi;
                while (mz_cand_gt(e[first], e[hi])) { --hi; }
                if (!(lo < hi)) { break; }
                mz_cand_swap(e, lo, hi);
                ++lo;
            }
            st_first[sp] = lo, st_last[sp] = last, st_depth[sp] = depth, ++sp; // __introsort_loop(cut, last, depth)
            last = lo;
        }
    }
    if (n > 16) { // __final_insertion_sort (stl_algo.h:1854-1865)
        mz_cand_insertion_sort(e, 0, 16);
        for (int i = 16; i != n; ++i) { mz_cand_unguarded_linear_insert(e, i); }
    } else {
        mz_cand_insertion_sort(e, 0, n);
    }
}

// One "after NN evaluation" step of game g (zero_actor.cpp:74-98): expand the leaf with the legal actions in
// descending policy order (zero_actor.cpp:215-229, mcts.cpp:151-164), back the value up (mcts.cpp:166-179)
// and mix the root noise in (zero_actor.cpp:194-204).
MZ_DEV void mz_after_nn(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int tid, int nthreads)
{
    const int len = s.path_len[g];
    if (len <= 0) { return; } // uniform for the block
    const int A = d.A, N = d.N;
    mz_hot* hot = s.hot + (size_t)g * d.NP;
    const int32_t* path = s.path + (size_t)g * (d.S + 2);
    const int leaf = path[len - 1];
    const int terminal = s.leaf_meta[g * 4 + 0], rotation = s.leaf_meta[g * 4 + 2];
    float v;
    if (!terminal) {
        for (int i = tid; i < MZ_LEGAL_WORDS; i += nthreads) { w->legal[i] = s.leaf_legal[g * MZ_LEGAL_WORDS + i]; }
        for (int a = tid; a < A; a += nthreads) {
            const int ra = (d.game == MZ_GAME_HEX ? a : mz_rotate(rotation, a, N)); // getRotateAction, zero_actor.cpp:222 (identity for Hex, hex.h:65)
            w->pol[a] = s.policy[(size_t)g * A + ra];
            w->lg[a] = s.logits[(size_t)g * A + ra];
        }
        const int first = s.cursor[g];
        if (tid == 0) { w->flag = 0; }
        mz_block_sync();
        int k = 0;
        for (int i = 0; i < MZ_LEGAL_WORDS; ++i) { k += mz_popc(w->legal[i]); }
        // rank sort, one candidate per thread: descending policy. Exact ties are ranked by ascending action id here and
        // re-ordered below the way std::sort leaves them (mz_std_sort_candidates)
        for (int a = tid; a < A; a += nthreads) {
            if (!((w->legal[a >> 5] >> (a & 31)) & 1u)) { continue; }
            const float p = w->pol[a];
            int rank = 0, tie = 0;
            for (int b = 0; b < A; ++b) {
                const float pb = w->pol[b];
                const bool legal_b = ((w->legal[b >> 5] >> (b & 31)) & 1u) != 0u;
                rank += (legal_b && ((pb > p) || (pb == p && b < a))) ? 1 : 0;
                tie |= (legal_b && pb == p && b != a) ? 1 : 0;
            }
            if (tie) { w->flag = 1; }
            const int c = first + rank;
            mz_store_hot(hot + c, 0.0f, 0.0f, p, 0u);
            s.action[(size_t)g * d.NP + c] = (int16_t)a;
            s.logit[(size_t)g * d.NP + c] = w->lg[a];
            s.value[(size_t)g * d.NP + c] = 0.0f;
            if (s.reward) { s.reward[(size_t)g * d.NP + c] = 0.0f; }
            s.node_slot[(size_t)g * d.NP + c] = -1;
            s.last_child[(size_t)g * d.NP + c] = -1;
        }
        mz_block_sync();
        if (w->flag) { // rare: candidates with exactly equal priors
            if (tid == 0) {
                uint64_t* e = w->cap_hash;
                int n = 0;
                for (int a = 0; a < A; ++a) {
                    if ((w->legal[a >> 5] >> (a & 31)) & 1u) { e[n++] = ((uint64_t)mz_float_bits(w->pol[a]) << 32) | (uint32_t)a; }
                }
                mz_std_sort_candidates(e, n);
                for (int i = 0; i < n; ++i) {
                    const int a = (int)(e[i] & 0xffffffffu);
                    mz_store_hot(hot + first + i, 0.0f, 0.0f, w->pol[a], 0u);
                    s.action[(size_t)g * d.NP + first + i] = (int16_t)a;
                    s.logit[(size_t)g * d.NP + first + i] = w->lg[a];
                }
            }
            mz_block_sync();
        }
        if (tid == 0) {
            s.cursor[g] = first + k;
            const mz_hot h = mz_load_hot(hot + leaf);
            mz_store_hot(hot + leaf, h.count, h.mean, h.policy, (uint32_t)first | ((uint32_t)k << MZ_LINK_SHIFT));
            if (s.vis) {
                mz_vis v;
                for (int i = 0; i < MZ_VIS_MAX; ++i) { v.idx[i] = 0xffffu; }
                v.n = 0, v.fu = 0;
                mz_store_vis(s.vis + (size_t)g * d.NP + leaf, v);
            }
        }
        v = s.nn_value[g];
        if (leaf == 0) {
            const float eps = d.eps, one_minus = mz_fsub(1.0f, eps);
            for (int i = tid; i < A; i += nthreads) {
                float nz = 0.0f;
                if (s.noise_in && i < k) {
                    nz = s.noise_in[(size_t)g * A + i];
                    if (d.gumbel_noise) { // Gumbel noise goes to the logit (zero_actor.cpp:205-211)
                        s.logit[(size_t)g * d.NP + first + i] = mz_fadd(s.logit[(size_t)g * d.NP + first + i], nz);
                    } else {
                        const mz_hot h = mz_load_hot(hot + first + i);
                        mz_store_hot(hot + first + i, h.count, h.mean, mz_fadd(mz_fmul(one_minus, h.policy), mz_fmul(eps, nz)), h.link);
                    }
                }
                s.root_noise[(size_t)g * A + i] = nz;
            }
        }
    } else {
        v = s.leaf_score[g];
    }
    mz_block_sync();
    // backup (mcts.cpp:166-179): node i of the path receives the leaf value carried up through the rewards and the discount of the
    // nodes below it; every thread folds the chain for its own node (paths are short where rewards exist)
    float* rew = (s.reward ? s.reward + (size_t)g * d.NP : nullptr);
    if (tid == 0) {
        s.value[(size_t)g * d.NP + leaf] = v;
        if (rew) { rew[leaf] = (d.has_reward && s.nn_reward ? s.nn_reward[g] : 0.0f); } // muzero_output->reward_, zero_actor.cpp:88
    }
    if (rew) { mz_block_sync(); }
    for (int i = tid; i < len; i += nthreads) {
        float x = v;
        if (rew) {
            for (int j = len - 1; j > i; --j) { x = mz_fadd(rew[path[j]], mz_fmul(d.discount, x)); }
        } else if (d.discount == 1.0f) {
            if (i < len - 1 && x == 0.0f) { x = 0.0f; } // 0 + 1 * x: only -0 changes (to +0)
        } else {
            for (int j = len - 1; j > i; --j) { x = mz_fadd(0.0f, mz_fmul(d.discount, x)); }
        }
        const int n = path[i];
        const mz_hot h = mz_load_hot(hot + n);
        const float cnt = mz_fadd(h.count, 1.0f);
        const float mean = mz_fadd(h.mean, mz_fdiv(mz_fmul(1.0f, mz_fsub(x, h.mean)), cnt));
        mz_store_hot(hot + n, cnt, mean, h.policy, h.link);
        if (d.value_rescale) { // Q of the node before / after this visit, for the value bounds below (path_hashes is free here)
            const float r = (rew ? rew[n] : 0.0f);
            float* qq = reinterpret_cast<float*>(w->path_hashes + i);
            qq[0] = mz_fadd(r, mz_fmul(d.discount, h.mean)), qq[1] = mz_fadd(r, mz_fmul(d.discount, mean));
        }
    }
    mz_block_sync();
    if (d.value_rescale && tid < MZ_W) { // updateTreeValueBound leaf -> root, in that order (the map's state is order dependent)
        for (int i = len - 1; i >= 0; --i) {
            const float* qq = reinterpret_cast<const float*>(w->path_hashes + i);
            mz_vb_update(d, s, g, qq[0], qq[1], tid);
        }
    }
    if (d.value_rescale) { mz_block_sync(); }
    // a simulation gives exactly one node its first visit: the leaf, unless it is a terminal node seen before (count now > 1)
    if (tid == 0 && s.vis && len >= 2 && mz_load_hot(hot + leaf).count == 1.0f) {
        const int parent = path[len - 2];
        const mz_hot ph = mz_load_hot(hot + parent);
        mz_vis_insert(s, (size_t)g * d.NP, parent, leaf - (int)(ph.link & ((1u << MZ_LINK_SHIFT) - 1u)), (int)(ph.link >> MZ_LINK_SHIFT));
    }
    if (d.gumbel) { mz_gumbel_halving(d, s, g, s.root_meta[g * 4 + 0], tid, nthreads); } // zero_actor.cpp:97
    if (s.vloss) { // think(): the leaf's virtual loss — how often this step selected it — comes off every node of its path (zero_actor.cpp:153-154)
        float* vl = s.vloss + (size_t)g * d.NP;
        mz_block_sync();
        const float x = vl[leaf];
        mz_block_sync();
        for (int i = tid; i < len; i += nthreads) { vl[path[i]] = mz_fsub(vl[path[i]], x); }
    }
    if (tid == 0) { s.path_len[g] = 0; }
}

// Tree::reset + ZeroActor::resetSearch (tree.h:64-69, zero_actor.cpp:29-34)
MZ_DEV void mz_tree_reset(const mz_dims& d, const mz_state& s, int g, int lane)
{
    if (lane == 0) {
        mz_store_hot(s.hot + (size_t)g * d.NP, 0.0f, 0.0f, 0.0f, 0u);
        s.value[(size_t)g * d.NP] = 0.0f;
        s.logit[(size_t)g * d.NP] = 0.0f;
        s.action[(size_t)g * d.NP] = -1;
        s.node_slot[(size_t)g * d.NP] = -1;
        s.last_child[(size_t)g * d.NP] = -1;
        s.cursor[g] = 1;
        s.path_len[g] = 0;
        s.spec_len[g] = 0;
        if (s.reward) { s.reward[(size_t)g * d.NP] = 0.0f; }
        if (s.vb_n) { s.vb_n[g] = 0; } // MCTS::reset clears tree_value_bound_ (mcts.cpp:78-83)
        if (s.gum_meta) { s.gum_meta[g * 4 + 0] = 0, s.gum_meta[g * 4 + 1] = 0, s.gum_meta[g * 4 + 2] = 0, s.gum_meta[g * 4 + 3] = 0; }
    }
}

// BaseActor::reset (base_actor.cpp:8-13)
MZ_DEV void mz_game_reset(const mz_dims& d, const mz_state& s, int g, mz_scratch* w, int lane)
{
    mz_env_reset(d, w, lane);
    mz_env_store_root(d, s, g, w, lane);
    mz_tree_reset(d, s, g, lane);
    if (d.game == MZ_GAME_ATARI && s.at_meta && lane == 0) { // AtariEnv::reset, atari.cpp:52-58: empty screen and action histories
        int32_t* m = s.at_meta + (size_t)g * 16;
        m[0] = 0, m[9] = 0;
        for (int i = 0; i < MZ_HIST; ++i) { m[1 + i] = -1; }
    }
}

// AtariEnv::reset's first screen (action < 0) or AtariEnv::act (atari.cpp:82-85): the screen the emulator answered with joins the
// 8-entry history together with the action that produced it; the oldest entry leaves. Block collective (any number of threads);
// the frame bytes are copied by the caller's threads, the bookkeeping by thread 0. Returns the ring slot the frame belongs in.
MZ_DEV int mz_atari_push(const mz_state& s, int g, int action, int tid)
{
    int32_t* m = s.at_meta + (size_t)g * 16;
    const int slot = m[0];
    mz_block_sync();
    if (tid == 0) {
        m[0] = (slot + 1) % MZ_HIST;
        m[1 + slot] = action; // -1 after a reset: the all-zero action plane (atari.cpp:57-58)
        m[9] |= 1 << slot;
    }
    return slot;
}

// BaseActor::act on the root environment + resetSearch (base_actor.cpp:22-30, actor_group.cpp:116-134).
// out[0] = 1 if the action was legal and applied, out[1] = terminal, out[2] = number of legal actions of the
// new position (what the host needs to draw the next root's Dirichlet noise), out[3] = side to move;
// *score = getEvalScore(false) of the new position when terminal.
MZ_DEV void mz_play(const mz_dims& d, const mz_state& s, int g, int action, mz_scratch* w, int32_t* out, float* score, int lane)
{
    if (d.game == MZ_GAME_ATARI) { // the emulator is the host's: legality against the minimal action set, move count, new search
        const int ok = (action >= 0 && action < d.A && ((d.legal_mask >> action) & 1u)) ? 1 : 0;
        if (ok) {
            if (lane == 0) { s.root_meta[g * 4 + 1] += 1, s.root_meta[g * 4 + 3] = s.root_meta[g * 4 + 2], s.root_meta[g * 4 + 2] = action; }
            mz_tree_reset(d, s, g, lane);
        }
        if (lane == 0) {
            out[0] = ok, out[1] = 0, out[2] = mz_popc(d.legal_mask), out[3] = 1;
            *score = 0.0f;
        }
        return;
    }
    mz_env_load_root(d, s, g, w, lane);
    uint64_t* hash_list = s.hashes + (size_t)g * d.max_hashes;
    mz_env_legal_block(d, s, w, hash_list, w->num_moves, hash_list, 0, lane, MZ_W);
    const int ok = (action >= 0 && action < d.A && ((w->legal[action >> 5] >> (action & 31)) & 1u)) ? 1 : 0;
    int terminal = 0, num_legal = 0;
    float sc = 0.0f;
    if (ok) {
        mz_env_act(d, s, w, action, w->turn, lane);
        if (lane == 0 && w->num_moves - 1 < d.max_hashes) { hash_list[w->num_moves - 1] = w->hash; } // hashkey_history_ / hash_table_, go.cpp:145-147,180-182
        mz_sync();
        mz_env_store_root(d, s, g, w, lane);
        mz_tree_reset(d, s, g, lane);
    }
    terminal = mz_env_is_terminal(d, w);
    if (terminal) {
        sc = mz_env_eval_score(d, w, lane);
    } else {
        num_legal = mz_env_legal_block(d, s, w, hash_list, w->num_moves, hash_list, 0, lane, MZ_W);
        if (d.game == MZ_GAME_NOGO && num_legal == 0) {
            terminal = 1;
            sc = mz_env_eval_score(d, w, lane);
        }
    }
    if (lane == 0) {
        out[0] = ok, out[1] = terminal, out[2] = num_legal, out[3] = w->turn;
        *score = sc;
    }
}

// MCTS::selectChildByMaxCount at the root (mcts.cpp:91-104): first child with the strictly largest count.
// Returns its action id (-1 if the root has no visited child).
MZ_DEV int mz_root_max_count_action(const mz_dims& d, const mz_state& s, int g, int lane)
{
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    const mz_hot root = mz_load_hot(hot);
    const int nc = (int)(root.link >> MZ_LINK_SHIFT), fc = (int)(root.link & ((1u << MZ_LINK_SHIFT) - 1u));
    float best_c = 0.0f, zero = 0.0f;
    int best_i = -1;
    for (int i = lane; i < nc; i += MZ_W) {
        const mz_hot c = mz_load_hot(hot + fc + i);
        if (c.count > best_c) { best_c = c.count, best_i = i; }
    }
    mz_reduce_best(best_c, zero, best_i); // (count desc, index asc)
    return best_i < 0 ? -1 : (int)s.action[(size_t)g * d.NP + fc + best_i];
}

// GumbelZero::decideActionNode with actor_select_action_by_count (gumbel_zero.cpp:61-66): the best-scoring candidate.
// Returns its action id (-1 without candidates). Result valid in lane 0.
MZ_DEV int mz_root_gumbel_action(const mz_dims& d, const mz_state& s, int g, int lane)
{
    int a = -1;
    if (lane == 0 && s.gum_meta[g * 4 + 0] > 0) {
        mz_gumbel_sort_by_score(d, s, g, s.root_meta[g * 4 + 0]);
        a = (int)s.action[(size_t)g * d.NP + s.gum_cand[(size_t)g * d.A]];
    }
    return a;
}
