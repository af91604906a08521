// TEST-ONLY build of minizero_b200/csrc/search_core.cuh as plain C++ (MZ_W == 1): lets the CPU
// test-suite replay the reference recordings through the exact source the CUDA kernels are
// compiled from. Never linked into the product library (libmzb200.so has no CPU path).
#define MZ_HOSTSIM 1
#include "../../minizero_b200/csrc/search_core.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

struct sim {
    mz_dims d;
    mz_state s;
    std::vector<float> bias;
    std::vector<double> sqrt_table;
    std::vector<uint64_t> keys;
    std::vector<uint8_t> rot;
    std::vector<float> noise;
    mz_scratch w;
};

template <class T>
static T* zalloc(size_t n) { return (T*)calloc(n, sizeof(T)); }

extern "C" {

// search options beyond the AlphaZero defaults: muzero, use_gumbel, gumbel_noise, gumbel_sample_size, sigma_visit_c, sigma_scale_c
static int g_opt_i[4] = {0, 0, 0, 16};
static float g_opt_f[2] = {50.0f, 1.0f};
static int g_opt_atari[2] = {0, 0}; // value_rescale, legal mask (Atari MuZero)
static int g_think_k = 0; // actor_mcts_think_batch_size of the next hs_create (console think(), zero_actor.cpp:129-157); 0 / 1: off
void hs_set_think(int k) { g_think_k = k; }
void hs_set_atari_options(int value_rescale, int legal_mask) { g_opt_atari[0] = value_rescale, g_opt_atari[1] = legal_mask; }
void hs_set_options(int muzero, int use_gumbel, int gumbel_noise, int m, float visit_c, float scale_c)
{
    g_opt_i[0] = muzero, g_opt_i[1] = use_gumbel, g_opt_i[2] = gumbel_noise, g_opt_i[3] = m;
    g_opt_f[0] = visit_c, g_opt_f[1] = scale_c;
}

sim* hs_create(int game, int N, int B, int S, float puct_base, float puct_init, float discount, float komi, float eps)
{
    sim* h = new sim();
    mz_dims& d = h->d;
    memset(&d, 0, sizeof(d));
    memset(&h->s, 0, sizeof(h->s));
    d.game = game, d.N = N, d.A = (game == MZ_GAME_TICTACTOE ? 9 : ((game == MZ_GAME_GOMOKU || game == MZ_GAME_HEX) ? N * N : N * N + 1)), d.C = (MZ_GO_FAMILY(game) ? 18 : 4), d.S = S, d.B = B;
    if (game == MZ_GAME_KILLALLGO) { d.game = MZ_GAME_GO, d.killall = 1, d.C = 18; } // every Go rule + killallgo.cpp:27-48 (engine.cu maps it the same way)
    d.gomoku_exactly_five = 1, d.gomoku_outer_open = 0, d.hex_swap_rule = 1; // reference defaults (configuration.cpp:82-85)
    d.num_players = 2, d.act_planes = 1;
    if (game == MZ_GAME_ATARI) { // atari.h:18-26
        d.A = 18, d.C = 32, d.num_players = 1, d.atari_init_q = 1, d.has_reward = 1, d.act_planes = 18;
        d.value_rescale = g_opt_atari[0], d.legal_mask = (uint32_t)g_opt_atari[1];
    }
    d.vb_cap = S + 2;
    d.muzero = g_opt_i[0], d.gumbel = g_opt_i[1], d.gumbel_noise = g_opt_i[2], d.gumbel_m = g_opt_i[3];
    d.sigma_visit_c = g_opt_f[0], d.sigma_scale_c = g_opt_f[1];
    if (d.gumbel) { // gumbel_zero.cpp:99,109 in the reference's double arithmetic
        const double lg = std::log2((double)d.gumbel_m);
        d.gumbel_budget0 = (int)std::max(1.0, std::floor(S / (lg * d.gumbel_m)));
        for (int l = 0; l < MZ_GUMBEL_LEVELS; ++l) {
            const int size = (d.gumbel_m >> l);
            d.gumbel_next[l] = (size > 0 ? (int)std::floor(S / (lg * size / 2)) : 0);
        }
    }
    d.NP = 1 + (S + 1) * d.A;
    d.think_k = (g_think_k > 1 ? g_think_k : 0);
    const int K = (d.think_k ? d.think_k : 1); // lanes: sections of the per-leaf arrays
    d.slots = (N + 1) * (N + 1);
    d.max_hashes = 2 * N * N + 4;
    d.puct_init = puct_init, d.puct_base = puct_base, d.discount = discount, d.komi = komi, d.eps = eps, d.turn_key = 0;
    mz_state& s = h->s;
    size_t np = (size_t)B * d.NP;
    s.hot = zalloc<mz_hot>(np), s.action = zalloc<int16_t>(np), s.logit = zalloc<float>(np), s.value = zalloc<float>(np);
    s.root_noise = zalloc<float>((size_t)B * d.A), s.cursor = zalloc<int32_t>(B);
    s.node_slot = zalloc<int16_t>(np);
    s.last_child = zalloc<int32_t>(np);
    s.slot_st = zalloc<uint32_t>((size_t)B * (S + 1) * 2 * N), s.slot_hash = zalloc<uint64_t>((size_t)B * (S + 1)), s.slot_meta = zalloc<int32_t>((size_t)B * (S + 1) * 4);
    h->w.path_hashes = zalloc<uint64_t>(S + 2);
    h->w.sel = zalloc<int32_t>(S + 2);
    h->w.lvl_h = zalloc<mz_hot>(S + 2);
    h->w.lvl_v = zalloc<mz_vis>(S + 2);
    if (!getenv("HS_NO_VIS") && !d.think_k) { s.vis = zalloc<mz_vis>(np); }
    if (d.think_k) { s.vloss = zalloc<float>(np), s.think_pending = zalloc<int32_t>(B); }
    h->w.q_warp = zalloc<float>(MZ_MAXA);
    s.spec_len = zalloc<int32_t>(B);
    s.gum_cand = zalloc<int32_t>((size_t)B * d.A), s.gum_meta = zalloc<int32_t>((size_t)B * 4);
    s.leaf_parent = zalloc<int32_t>((size_t)K * B * 2); // per-leaf array: one section per think() lane
    if (d.has_reward) { s.reward = zalloc<float>(np), s.nn_reward = zalloc<float>(B); }
    if (d.value_rescale) { s.vb_key = zalloc<float>((size_t)B * d.vb_cap), s.vb_cnt = zalloc<int32_t>((size_t)B * d.vb_cap), s.vb_n = zalloc<int32_t>(B); }
    if (game == MZ_GAME_ATARI) { s.at_meta = zalloc<int32_t>((size_t)B * 16); }
    h->sqrt_table.resize(S + 2 + K);
    for (int n = 0; n < S + 2 + K; ++n) { h->sqrt_table[n] = sqrt((double)n); }
    s.sqrt_table = h->sqrt_table.data();
    s.root_st = zalloc<uint32_t>((size_t)B * 2 * MZ_ROWS), s.root_hist = zalloc<uint32_t>((size_t)B * MZ_HIST * 2 * MZ_ROWS);
    s.root_hash = zalloc<uint64_t>(B), s.root_meta = zalloc<int32_t>((size_t)B * 4), s.hashes = zalloc<uint64_t>((size_t)B * d.max_hashes);
    const size_t KB = (size_t)K * B;
    s.path = zalloc<int32_t>(KB * (S + 2)), s.path_len = zalloc<int32_t>(KB), s.leaf_legal = zalloc<uint32_t>(KB * MZ_LEGAL_WORDS);
    s.leaf_meta = zalloc<int32_t>(KB * 4), s.leaf_score = zalloc<float>(KB);
    s.nn_in = zalloc<uint16_t>(KB * d.slots * MZ_NN_CPAD);
    s.policy = zalloc<float>(KB * d.A), s.logits = zalloc<float>(KB * d.A), s.nn_value = zalloc<float>(KB);
    h->bias.resize(S + 2 + K); // a node's total under virtual loss reaches S + K
    for (int n = 0; n < S + 2 + K; ++n) {
        float t = (float)(1 + n) + puct_base;
        t = t / puct_base;
        h->bias[n] = (float)((double)puct_init + log((double)t));
    }
    std::mt19937_64 gen(0);
    h->keys.resize(2 * 361);
    (void)gen(); // turn key
    for (int pos = 0; pos < 361; ++pos) {
        (void)gen();
        h->keys[0 * 361 + pos] = gen();
        h->keys[1 * 361 + pos] = gen();
    }
    s.puct_bias = h->bias.data(), s.keys = h->keys.data();
    h->rot.assign(KB, 0), h->noise.assign((size_t)B * d.A, 0.0f);
    s.rotations = h->rot.data(), s.noise_in = nullptr;
    for (int g = 0; g < B; ++g) { mz_game_reset(d, s, g, &h->w, 0); }
    return h;
}

void hs_select(sim* h, const uint8_t* rotations, float* features)
{
    const mz_dims& d = h->d;
    for (int g = 0; g < d.B; ++g) { h->rot[g] = (rotations ? rotations[g] : 0); }
    for (int g = 0; g < d.B; ++g) { mz_before_nn(d, h->s, g, &h->w, 0, 0, 1); }
    if (features && d.game != MZ_GAME_ATARI) { // Atari planes are produced by the device's screen-ring kernel, not by search_core.cuh
        const int N = d.N;
        for (int g = 0; g < d.B; ++g) {
            for (int c = 0; c < d.C; ++c) {
                for (int pos = 0; pos < N * N; ++pos) {
                    uint16_t v = h->s.nn_in[((size_t)g * d.slots + (pos / N + 1) * (N + 1) + pos % N) * MZ_NN_CPAD + c];
                    features[((size_t)g * d.C + c) * N * N + pos] = (v == MZ_HALF_ONE ? 1.0f : 0.0f);
                }
            }
        }
    }
}

void hs_apply_mz(sim* h, const float* policy, const float* logits, const float* value, const float* reward, const float* noise);
void hs_apply(sim* h, const float* policy, const float* logits, const float* value, const float* noise) { hs_apply_mz(h, policy, logits, value, nullptr, noise); }

void hs_apply_mz(sim* h, const float* policy, const float* logits, const float* value, const float* reward, const float* noise)
{
    const mz_dims& d = h->d;
    if (h->s.nn_reward) {
        for (int g = 0; g < d.B; ++g) { h->s.nn_reward[g] = (reward ? reward[g] : 0.0f); }
    }
    memcpy(h->s.policy, policy, sizeof(float) * (size_t)d.B * d.A);
    memcpy(h->s.logits, logits, sizeof(float) * (size_t)d.B * d.A);
    memcpy(h->s.nn_value, value, sizeof(float) * (size_t)d.B);
    if (noise) { memcpy(h->noise.data(), noise, sizeof(float) * (size_t)d.B * d.A); }
    h->s.noise_in = (noise ? h->noise.data() : nullptr);
    for (int g = 0; g < d.B; ++g) { mz_after_nn(d, h->s, g, &h->w, 0, 1); }
}

// one batched think() step (zero_actor.cpp:129-157), "before" half: K selections per tree, one after the other. rotations / features / path_len
// are lane-major [K][B]; path_len: > 0 leaf to evaluate, < 0 duplicate of an earlier lane (-length), 0 lane beyond the simulations left
void hs_think_select(sim* h, const uint8_t* rotations, float* features, int32_t* path_len)
{
    const mz_dims& d = h->d;
    const int K = d.think_k, N = d.N;
    for (int i = 0; i < K * d.B; ++i) { h->rot[i] = (rotations ? rotations[i] : 0); }
    for (int g = 0; g < d.B; ++g) { h->s.think_pending[g] = 0; }
    for (int k = 0; k < K; ++k) {
        const mz_state v = mz_lane_view(d, h->s, k, d.B);
        for (int g = 0; g < d.B; ++g) { mz_before_nn(d, v, g, &h->w, 0, 0, 1); }
        for (int g = 0; g < d.B; ++g) {
            const size_t l = (size_t)k * d.B + g;
            path_len[l] = v.path_len[g];
            if (!features || v.path_len[g] <= 0) { continue; }
            for (int c = 0; c < d.C; ++c) {
                for (int pos = 0; pos < N * N; ++pos) {
                    uint16_t x = v.nn_in[((size_t)g * d.slots + (pos / N + 1) * (N + 1) + pos % N) * MZ_NN_CPAD + c];
                    features[(l * d.C + c) * N * N + pos] = (x == MZ_HALF_ONE ? 1.0f : 0.0f);
                }
            }
        }
    }
}

// "after" half: the evaluated lanes' results are applied in selection order; policy / logits [K][B][A], value [K][B], noise [B][A] by root child index
void hs_think_apply(sim* h, const float* policy, const float* logits, const float* value, const float* noise)
{
    const mz_dims& d = h->d;
    const size_t KB = (size_t)d.think_k * d.B;
    memcpy(h->s.policy, policy, sizeof(float) * KB * d.A);
    memcpy(h->s.logits, logits, sizeof(float) * KB * d.A);
    memcpy(h->s.nn_value, value, sizeof(float) * KB);
    if (noise) { memcpy(h->noise.data(), noise, sizeof(float) * (size_t)d.B * d.A); }
    h->s.noise_in = (noise ? h->noise.data() : nullptr);
    for (int k = 0; k < d.think_k; ++k) {
        const mz_state v = mz_lane_view(d, h->s, k, d.B);
        for (int g = 0; g < d.B; ++g) { mz_after_nn(d, v, g, &h->w, 0, 1); }
    }
}

int hs_path_len(sim* h, int g) { return h->s.path_len[g]; }
int hs_leaf_action(sim* h, int g) { return h->s.leaf_parent[g * 2 + 1]; }
int hs_leaf_parent_slot(sim* h, int g) { return h->s.leaf_parent[g * 2 + 0]; }
int hs_path_hash(sim* h, int g)
{
    const int32_t* path = h->s.path + (size_t)g * (h->d.S + 2);
    uint32_t x = 2166136261u;
    for (int k = 1; k < h->s.path_len[g]; ++k) { x = (x ^ (uint32_t)(int32_t)h->s.action[(size_t)g * h->d.NP + path[k]]) * 16777619u; }
    return (int)(x & 0x7fffffffu);
}
int hs_gumbel_best_action(sim* h, int g) { return mz_root_gumbel_action(h->d, h->s, g, 0); }
int hs_sims_done(sim* h, int g) { return (int)h->s.hot[(size_t)g * h->d.NP].count; }

// out_i: [1 + A] num_children, actions; out_f: [3 + 6 * A] root count/mean/value then count, mean, policy, logit, noise, value per child
void hs_root(sim* h, int g, int32_t* out_i, float* out_f)
{
    const mz_dims& d = h->d;
    const mz_hot* hot = h->s.hot + (size_t)g * d.NP;
    const int nc = (int)(hot[0].link >> MZ_LINK_SHIFT), fc = (int)(hot[0].link & ((1u << MZ_LINK_SHIFT) - 1u));
    out_i[0] = nc;
    out_f[0] = hot[0].count, out_f[1] = hot[0].mean, out_f[2] = h->s.value[(size_t)g * d.NP];
    for (int i = 0; i < nc; ++i) {
        out_i[1 + i] = h->s.action[(size_t)g * d.NP + fc + i];
        out_f[3 + 0 * d.A + i] = hot[fc + i].count;
        out_f[3 + 1 * d.A + i] = hot[fc + i].mean;
        out_f[3 + 2 * d.A + i] = hot[fc + i].policy;
        out_f[3 + 3 * d.A + i] = h->s.logit[(size_t)g * d.NP + fc + i];
        out_f[3 + 4 * d.A + i] = h->s.root_noise[(size_t)g * d.A + i];
        out_f[3 + 5 * d.A + i] = h->s.value[(size_t)g * d.NP + fc + i];
    }
}

// returns ok | terminal << 1 ; num_legal and score through pointers
int hs_play(sim* h, int g, int action, int* num_legal, float* score)
{
    int32_t out[4];
    mz_play(h->d, h->s, g, action, &h->w, out, score, 0);
    *num_legal = out[2];
    return out[0] | (out[1] << 1);
}

void hs_reset_game(sim* h, int g) { mz_game_reset(h->d, h->s, g, &h->w, 0); }

// Atari: the history bookkeeping of a new screen (the bytes themselves only matter to the device's plane kernel)
void hs_atari_observe(sim* h, int g, int action) { mz_atari_push(h->s, g, action, 0); }

// rewards of the root children, number of value-bound keys, smallest / largest key
void hs_root_extra(sim* h, int g, float* c_reward, int32_t* bound)
{
    const mz_dims& d = h->d;
    const mz_hot* hot = h->s.hot + (size_t)g * d.NP;
    const int nc = (int)(hot[0].link >> MZ_LINK_SHIFT), fc = (int)(hot[0].link & ((1u << MZ_LINK_SHIFT) - 1u));
    for (int i = 0; i < nc; ++i) { c_reward[i] = (h->s.reward ? h->s.reward[(size_t)g * d.NP + fc + i] : 0.0f); }
    const mz_qb qb = mz_vb_bounds(d, h->s, g, 0);
    bound[0] = qb.n;
    memcpy(&bound[1], &qb.lo, 4), memcpy(&bound[2], &qb.hi, 4);
}

// property-test hook: policies[n] (candidates in ascending action id) -> action order left by (a) the restatement in
// search_core.cuh and (b) the real std::sort with the reference's comparator (zero_actor.cpp:225-227); returns 1 if equal
int hs_sort_matches_std(int n, const float* policy, int32_t* order_out)
{
    struct Cand {
        int a;
        float p;
    };
    std::vector<Cand> ref(n);
    std::vector<uint64_t> e(n);
    for (int i = 0; i < n; ++i) {
        ref[i] = {i, policy[i]};
        e[i] = ((uint64_t)mz_float_bits(policy[i]) << 32) | (uint32_t)i;
    }
    std::sort(ref.begin(), ref.end(), [](const Cand& lhs, const Cand& rhs) { return lhs.p > rhs.p; });
    mz_std_sort_candidates(e.data(), n);
    int same = 1;
    for (int i = 0; i < n; ++i) {
        order_out[i] = (int)(e[i] & 0xffffffffu);
        same &= (order_out[i] == ref[i].a);
    }
    return same;
}

// differential check of the four per-level selection routines on every expanded node below the root of game g's current
// tree: the full scans (warp-collective and single-thread forms) and the visited-list forms must choose the same child.
// Returns the number of nodes compared, or -(node index) - 1 of the first disagreement.
int hs_check_level_variants(sim* h, int g)
{
    const mz_dims& d = h->d;
    const mz_state& s = h->s;
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    if (!s.vis) { return 0; }
    int compared = 0;
    const int used = s.cursor[g];
    for (int node = 1; node < used; ++node) {
        const mz_hot hn = hot[node];
        if ((hn.link >> MZ_LINK_SHIFT) == 0 || hn.count < 1.0f) { continue; }
        const mz_vis v = s.vis[(size_t)g * d.NP + node];
        for (int player = 1; player <= d.num_players; ++player) {
            mz_hot c;
            const int fc = (int)(hn.link & ((1u << MZ_LINK_SHIFT) - 1u));
            const mz_qb qb = mz_vb_bounds(d, s, g, 0);
            const int a = fc + mz_select_level(d, s, qb, hot, hn, false, player, h->w.q_warp, 0, c);
            const int b = mz_select_level_serial(d, s, qb, hot, hn, player);
            if (a != b) { return -node - 1; }
            if (v.n <= MZ_VIS_MAX) {
                if (mz_select_level_vis(d, s, qb, hot, hn, v, player, 0) != a || mz_select_level_vis_serial(d, s, qb, hot, hn, v, player) != a) { return -node - 1; }
                ++compared;
            }
        }
    }
    return compared;
}

void hs_destroy(sim* h) { delete h; } // test helper: buffers are reclaimed at process exit
}
