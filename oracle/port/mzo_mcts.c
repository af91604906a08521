/*
 * Oracle port — batched AlphaZero MCTS (PUCT select, expand, backup). TEST INFRASTRUCTURE ONLY
 * (see mzo.h). Plain serial C, one game after another; arithmetic restated operation by
 * operation, in the precision the compiled reference uses:
 *
 *   actor/mcts.cpp:20-28     MCTSNode::add                 -> node_add
 *   actor/mcts.cpp:40-53     MCTSNode::getNormalizedMean   -> normalized_mean
 *   actor/mcts.cpp:55-61     getNormalizedPUCTScore        -> puct_score
 *   actor/mcts.cpp:91-104    selectChildByMaxCount         -> mzo_select_by_max_count
 *   actor/mcts.cpp:139-148   selectFromNode                -> select_path
 *   actor/mcts.cpp:151-164   expand                        -> expand
 *   actor/mcts.cpp:166-179   backup                        -> backup
 *   actor/mcts.cpp:181-198   selectChildByPUCTScore        -> select_child
 *   actor/mcts.cpp:200-217   calculateInitQValue           -> init_q
 *   actor/tree.h:64-77       Tree::reset / allocateNodes   -> tree_reset / cursor
 *   actor/zero_actor.cpp:29-34   resetSearch (root action = (-1, previous player))
 *   actor/zero_actor.cpp:51-72   beforeNNEvaluation        -> mzo_select
 *   actor/zero_actor.cpp:74-98   afterNNEvaluation         -> mzo_apply
 *   actor/zero_actor.cpp:194-204 addNoiseToNodeChildren (Dirichlet mix) -> in mzo_apply
 *   actor/zero_actor.cpp:215-229 calculateAlphaZeroActionPolicy -> candidates
 *   actor/zero_actor.cpp:247-252 getEnvironmentTransition  -> transition
 *
 * Build with -ffp-contract=off (oracle/Makefile): the reference is compiled for baseline
 * x86-64 and therefore never fuses a multiply with an add.
 *
 * Candidate order: the reference sorts with std::sort (introsort, unstable); mzo_sort.c restates
 * libstdc++'s algorithm so that candidates with exactly equal priors land where the reference puts them.
 */
#include "mzo.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t cursor; /* Tree::current_node_size_ */
    int32_t* first_child;
    int32_t* num_children;
    int16_t* action;
    uint8_t* player;
    float *count, *mean, *policy, *logit, *noise, *value, *reward;
    float* vloss; /* MCTSNode::virtual_loss_ (mcts.h:59): non-zero only inside a batched think() step */
    int16_t* slot; /* MuZero: simulation index that evaluated the node = index of its hidden state (tree.h hidden_state_data_index_) */
    /* GumbelZero state (gumbel_zero.h:20-23) */
    int32_t cand[MZO_MAX_ACTIONS];
    int32_t num_cand, sample_size, budget;
    /* MCTS::tree_value_bound_ (mcts.h:117): std::map<float, int>, kept as arrays sorted by key */
    float* vb_key;
    int32_t* vb_cnt;
    int32_t vb_n;
} tree;

/* Atari root environment (atari.cpp:47-93): the last 8 screens and the actions that led to them; the emulator is the host's */
typedef struct {
    uint8_t frame[MZO_ATARI_HIST][3 * MZO_ATARI_RES * MZO_ATARI_RES];
    int8_t has_frame[MZO_ATARI_HIST]; /* 0: the all-zero planes the history starts with (atari.cpp:52-54) */
    int32_t action[MZO_ATARI_HIST];   /* action id whose plane is id / 18; -1: all-zero plane (atari.cpp:57-58) */
    int32_t terminal;
} atari_env;

struct mzo_batch {
    mzo_config cfg;
    int A, F, NP;
    mzo_env* root_env;
    mzo_env* leaf_env;
    tree* trees;
    int32_t* path;     /* [B][S+2] */
    int32_t* path_len; /* [B] */
    uint8_t* rotation; /* [B] */
    atari_env* atari;  /* [B] when cfg.game == MZO_GAME_ATARI */
    /* batched think() step (zero_actor.cpp:129-157): the lanes of mzo_think_select, lane-major [K][B] */
    int think_k;
    int32_t* t_path;     /* [K][B][S+2] */
    int32_t* t_len;      /* [K][B]: > 0 evaluate, < 0 duplicate leaf (-length), 0 lane unused */
    uint8_t* t_rot;      /* [K][B] */
    mzo_env* t_env;      /* [K][B] leaf environments */
};

static int other(int p) { return p == 1 ? 2 : 1; }

static void node_reset(tree* t, int i)
{
    t->num_children[i] = 0;
    t->first_child[i] = -1;
    t->mean[i] = t->count[i] = t->policy[i] = t->logit[i] = t->noise[i] = t->value[i] = t->reward[i] = t->vloss[i] = 0.0f;
    t->slot[i] = -1;
}

static void tree_reset(tree* t)
{
    t->cursor = 1;
    t->vb_n = 0; /* MCTS::reset clears tree_value_bound_, mcts.cpp:78-83 */
    node_reset(t, 0);
}

mzo_batch* mzo_create(const mzo_config* cfg)
{
    mzo_batch* b = (mzo_batch*)calloc(1, sizeof(mzo_batch));
    b->cfg = *cfg;
    mzo_env tmp;
    mzo_env_init(&tmp, cfg->game, cfg->board_size, cfg->komi, cfg->ko_situational);
    mzo_env_set_flags(&tmp, cfg->gomoku_flags);
    b->A = mzo_env_num_actions(&tmp);
    b->F = mzo_env_input_channels(&tmp) * tmp.n * tmp.n;
    if (cfg->game == MZO_GAME_ATARI) { /* atari.h:68-74: 8 x (action plane + RGB) at 96 x 96 */
        b->F = MZO_ATARI_HIST * 4 * MZO_ATARI_RES * MZO_ATARI_RES;
        b->atari = (atari_env*)calloc((size_t)cfg->num_games, sizeof(atari_env));
    }
    b->NP = 1 + (cfg->num_simulation + 1) * b->A; /* actor_group.cpp:183, tree.h:66 */
    int B = cfg->num_games;
    b->root_env = (mzo_env*)calloc((size_t)B, sizeof(mzo_env));
    b->leaf_env = (mzo_env*)calloc((size_t)B, sizeof(mzo_env));
    b->trees = (tree*)calloc((size_t)B, sizeof(tree));
    b->path = (int32_t*)calloc((size_t)B * (size_t)(cfg->num_simulation + 2), sizeof(int32_t));
    b->path_len = (int32_t*)calloc((size_t)B, sizeof(int32_t));
    b->rotation = (uint8_t*)calloc((size_t)B, 1);
    for (int g = 0; g < B; ++g) {
        tree* t = &b->trees[g];
        size_t n = (size_t)b->NP;
        t->first_child = (int32_t*)malloc(n * 4);
        t->num_children = (int32_t*)malloc(n * 4);
        t->action = (int16_t*)malloc(n * 2);
        t->player = (uint8_t*)malloc(n);
        t->count = (float*)malloc(n * 4);
        t->mean = (float*)malloc(n * 4);
        t->policy = (float*)malloc(n * 4);
        t->logit = (float*)malloc(n * 4);
        t->noise = (float*)malloc(n * 4);
        t->value = (float*)malloc(n * 4);
        t->reward = (float*)malloc(n * 4);
        t->vloss = (float*)malloc(n * 4);
        t->slot = (int16_t*)malloc(n * 2);
        t->vb_key = (float*)malloc(sizeof(float) * (size_t)(2 * cfg->num_simulation + 8));
        t->vb_cnt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * cfg->num_simulation + 8));
        mzo_reset_game(b, g);
    }
    return b;
}

void mzo_destroy(mzo_batch* b)
{
    if (!b) { return; }
    for (int g = 0; g < b->cfg.num_games; ++g) {
        tree* t = &b->trees[g];
        free(t->first_child), free(t->num_children), free(t->action), free(t->player);
        free(t->count), free(t->mean), free(t->policy), free(t->logit), free(t->noise), free(t->value), free(t->reward), free(t->slot), free(t->vloss);
        free(t->vb_key), free(t->vb_cnt);
    }
    free(b->atari);
    free(b->t_path), free(b->t_len), free(b->t_rot), free(b->t_env);
    free(b->root_env), free(b->leaf_env), free(b->trees), free(b->path), free(b->path_len), free(b->rotation);
    free(b);
}

/* ZeroActor::resetSearch, zero_actor.cpp:29-34 */
void mzo_reset_search(mzo_batch* b, int g)
{
    tree* t = &b->trees[g];
    tree_reset(t);
    t->action[0] = -1;
    t->player[0] = (uint8_t)(b->cfg.game == MZO_GAME_ATARI ? 1 : other(b->root_env[g].turn)); /* env::getPreviousPlayer: the other player of a 2-player game, player 1 of a 1-player game */
    b->path_len[g] = 0;
}

/* BaseActor::reset, base_actor.cpp:8-13 */
void mzo_reset_game(mzo_batch* b, int g)
{
    mzo_env_init(&b->root_env[g], b->cfg.game, b->cfg.board_size, b->cfg.komi, b->cfg.ko_situational);
    mzo_env_set_flags(&b->root_env[g], b->cfg.gomoku_flags);
    if (b->atari) {
        memset(&b->atari[g], 0, sizeof(atari_env));
        for (int i = 0; i < MZO_ATARI_HIST; ++i) { b->atari[g].action[i] = -1; }
    }
    mzo_reset_search(b, g);
}

/* mcts.cpp:40-53; virtual_loss_ is 0 outside a batched think() step and stays in the formula as an explicit term */
static float normalized_mean(const mzo_batch* b, const tree* t, int i)
{
    float value = t->reward[i] + b->cfg.reward_discount * t->mean[i];
    if (b->cfg.value_rescale) { /* mcts.cpp:43-49 */
        if (t->vb_n < 2) { return 1.0f; }
        const float lower = t->vb_key[0], upper = t->vb_key[t->vb_n - 1];
        value = (value - lower) / (upper - lower);
        value = (float)fmin(1, fmax(-1, 2 * value - 1)); /* the double overloads of fmin / fmax, as compiled (SURVEY a-3) */
    }
    if (t->player[i] == 2) { value = -value; } /* actor_mcts_value_flipping_player == 'W' */
    const float vloss = t->vloss[i];
    value = (value * t->count[i] - vloss) / (t->count[i] + vloss);
    return value;
}

/* MCTS::updateTreeValueBound, mcts.cpp:219-228: decrement (and erase at zero) the old key when present, then count the new one */
static void update_value_bound(const mzo_batch* b, tree* t, float old_value, float new_value)
{
    if (!b->cfg.value_rescale) { return; }
    int i = 0;
    while (i < t->vb_n && t->vb_key[i] < old_value) { ++i; }
    if (i < t->vb_n && !(old_value < t->vb_key[i])) { /* map::count(old_value): neither key orders before the other */
        if (--t->vb_cnt[i] == 0) {
            memmove(t->vb_key + i, t->vb_key + i + 1, sizeof(float) * (size_t)(t->vb_n - i - 1));
            memmove(t->vb_cnt + i, t->vb_cnt + i + 1, sizeof(int32_t) * (size_t)(t->vb_n - i - 1));
            --t->vb_n;
        }
    }
    i = 0;
    while (i < t->vb_n && t->vb_key[i] < new_value) { ++i; }
    if (i < t->vb_n && !(new_value < t->vb_key[i])) {
        ++t->vb_cnt[i];
    } else {
        memmove(t->vb_key + i + 1, t->vb_key + i, sizeof(float) * (size_t)(t->vb_n - i));
        memmove(t->vb_cnt + i + 1, t->vb_cnt + i, sizeof(int32_t) * (size_t)(t->vb_n - i));
        t->vb_key[i] = new_value, t->vb_cnt[i] = 1;
        ++t->vb_n;
    }
}

/* mcts.cpp:55-61 */
static float puct_score(const mzo_batch* b, const tree* t, int i, int total_simulation, float init_q_value)
{
    float tt = (float)(1 + total_simulation) + b->cfg.puct_base;
    tt = tt / b->cfg.puct_base;
    float puct_bias = (float)((double)b->cfg.puct_init + log((double)tt));
    float bp = puct_bias * t->policy[i];
    const float count_vl = t->count[i] + t->vloss[i]; /* getCountWithVirtualLoss, mcts.h:47 */
    float value_u = (float)(((double)bp * sqrt((double)total_simulation)) / (double)(1.0f + count_vl));
    float value_q = (count_vl == 0.0f ? init_q_value : normalized_mean(b, t, i));
    return value_u + value_q;
}

/* mcts.cpp:200-217 (board-game branch) */
static float init_q(const mzo_batch* b, const tree* t, int node)
{
    float sum_of_win = 0.0f, sum = 0.0f;
    for (int k = 0; k < t->num_children[node]; ++k) {
        int c = t->first_child[node] + k;
        if (t->count[c] + t->vloss[c] == 0.0f) { continue; }
        sum_of_win += normalized_mean(b, t, c);
        sum += 1;
    }
    if (b->cfg.game == MZO_GAME_ATARI) { return (sum > 0 ? sum_of_win / sum : 1.0f); } /* the #if ATARI branch, mcts.cpp:211-213 */
    return (sum_of_win - 1) / (sum + 1);
}

/* mcts.cpp:181-198 */
static int select_child(const mzo_batch* b, const tree* t, int node)
{
    int total_simulation = (int)(t->count[node] + t->vloss[node] - 1);
    float iq = init_q(b, t, node);
    float best_score = -3.402823466e+38f, best_policy = -3.402823466e+38f;
    int selected = -1;
    for (int k = 0; k < t->num_children[node]; ++k) {
        int c = t->first_child[node] + k;
        float score = puct_score(b, t, c, total_simulation, iq);
        if (score < best_score || (score == best_score && t->policy[c] <= best_policy)) { continue; }
        best_score = score;
        best_policy = t->policy[c];
        selected = c;
    }
    return selected;
}


/* ---- Gumbel (actor/gumbel_zero.cpp) ---- */
static float gumbel_max_child_count(const tree* t)
{
    float m = 0;
    for (int i = 0; i < t->num_children[0]; ++i) { m = (float)fmax(m, t->count[t->first_child[0] + i]); }
    return m;
}

/* gumbel_zero.cpp:120-137: candidates by descending logit + (c_visit + max_count) * c_scale * q; unvisited ones last */
static void sort_candidates_by_score(const mzo_batch* b, tree* t)
{
    const float max_child_count = gumbel_max_child_count(t);
    float score[MZO_MAX_ACTIONS];
    for (int i = 0; i < t->num_cand; ++i) {
        int c = t->cand[i];
        float v = normalized_mean(b, t, c);
        float s = t->logit[c] + (b->cfg.gumbel_sigma_visit_c + max_child_count) * b->cfg.gumbel_sigma_scale_c * v;
        score[i] = (t->count[c] > 0 ? s : -3.402823466e+38f);
    }
    for (int i = 1; i < t->num_cand; ++i) { /* stable insertion sort, descending */
        int c = t->cand[i];
        float s = score[i];
        int j = i;
        while (j > 0 && score[j - 1] < s) {
            t->cand[j] = t->cand[j - 1], score[j] = score[j - 1];
            --j;
        }
        t->cand[j] = c, score[j] = s;
    }
}

/* gumbel_zero.cpp:87-118 */
static void sequential_halving(const mzo_batch* b, tree* t)
{
    const int S = b->cfg.num_simulation, m = b->cfg.gumbel_sample_size;
    if ((int)t->count[0] == 1) {
        t->num_cand = 0;
        for (int i = 0; i < t->num_children[0]; ++i) { /* descending logit (noise already added), stable */
            int c = t->first_child[0] + i, j = t->num_cand++;
            while (j > 0 && t->logit[t->cand[j - 1]] < t->logit[c]) {
                t->cand[j] = t->cand[j - 1];
                --j;
            }
            t->cand[j] = c;
        }
        if (t->num_cand > m) { t->num_cand = m; }
        t->sample_size = m;
        t->budget = (int)fmax(1.0, floor(S / (log2((double)m) * t->sample_size)));
    } else {
        for (int i = 0; i < t->num_cand; ++i) {
            if (!(t->count[t->cand[i]] >= (float)t->budget)) { return; }
        }
        int next_budget = (int)floor(S / (log2((double)m) * t->sample_size / 2)); /* (double * int) / 2: odd sample sizes are not truncated */
        if (next_budget > 0 && t->sample_size > 2) {
            t->sample_size /= 2;
            sort_candidates_by_score(b, t);
            if (t->num_cand > t->sample_size) { t->num_cand = t->sample_size; }
            t->budget = (int)(t->count[t->cand[0]] + next_budget);
        }
    }
}

int mzo_gumbel_best_action(mzo_batch* b, int g)
{
    tree* t = &b->trees[g];
    if (t->num_cand <= 0) { return -1; }
    sort_candidates_by_score(b, t);
    return t->action[t->cand[0]];
}

/* gumbel_zero.cpp:9-59 */
int mzo_gumbel_policy(const mzo_batch* b, int g, int32_t* actions, float* probs)
{
    const tree* t = &b->trees[g];
    const int nc = t->num_children[0], fc = t->first_child[0];
    float pi_sum = 0.0f, q_sum = 0.0f;
    for (int i = 0; i < nc; ++i) {
        int c = fc + i;
        if (t->count[c] == 0) { continue; }
        float value = normalized_mean(b, t, c);
        pi_sum += t->policy[c];
        q_sum += t->policy[c] * value;
    }
    float value_pi = t->value[0];
    if (b->cfg.value_rescale) { /* gumbel_zero.cpp:21-30 */
        if (t->vb_n < 2) {
            value_pi = 1.0f;
        } else {
            value_pi = (value_pi - t->vb_key[0]) / (t->vb_key[t->vb_n - 1] - t->vb_key[0]);
            value_pi = (float)fmin(1, fmax(-1, 2 * value_pi - 1));
        }
    }
    value_pi = (t->player[fc] == 2 ? -value_pi : value_pi);
    const int S = b->cfg.num_simulation;
    float non_visited = (float)(1.0 / (1 + S) * (value_pi + (S / pi_sum) * q_sum));
    float max_logit = -3.402823466e+38f, max_child_count = gumbel_max_child_count(t);
    float score[MZO_MAX_ACTIONS];
    for (int i = 0; i < nc; ++i) {
        int c = fc + i;
        float value = (t->count[c] == 0 ? non_visited : normalized_mean(b, t, c));
        float lw = t->logit[c] - t->noise[c];
        score[i] = lw + (b->cfg.gumbel_sigma_visit_c + max_child_count) * b->cfg.gumbel_sigma_scale_c * value;
        max_logit = (float)fmax(max_logit, score[i]);
    }
    int n = 0;
    for (int a = 0; a < b->A; ++a) { /* ascending action id (the reference prints in unordered_map order) */
        for (int i = 0; i < nc; ++i) {
            if (t->action[fc + i] != a) { continue; }
            float l = score[i] - max_logit;
            if (l < -38) { continue; }
            actions[n] = a;
            probs[n++] = (float)exp(l);
        }
    }
    return n;
}

/* AtariEnv::getFeatures, atari.cpp:106-116: for each of the 8 history entries the action plane (id / 18) then R, G, B (byte / 255) */
static void atari_features(const atari_env* e, float* out)
{
    const int hw = MZO_ATARI_RES * MZO_ATARI_RES;
    for (int i = 0; i < MZO_ATARI_HIST; ++i) {
        const float a = (e->action[i] < 0 ? 0.0f : e->action[i] * 1.0f / 18); /* atari.cpp:82 */
        for (int k = 0; k < hw; ++k) { out[(size_t)(4 * i) * hw + k] = a; }
        for (int k = 0; k < 3 * hw; ++k) {
            float v = (float)e->frame[i][k];
            out[(size_t)(4 * i + 1) * hw + k] = (e->has_frame[i] ? v / 255.0f : 0.0f); /* atari.cpp:152-156 */
        }
    }
}

void mzo_atari_observe(mzo_batch* b, int g, int action, const uint8_t* frame_chw, int terminal)
{
    atari_env* e = &b->atari[g];
    if (action < 0) { /* AtariEnv::reset, atari.cpp:47-59: seven zero screens + the initial one, eight zero action planes */
        memset(e, 0, sizeof(*e));
        for (int i = 0; i < MZO_ATARI_HIST; ++i) { e->action[i] = -1; }
    } else { /* AtariEnv::act, atari.cpp:82-85: the action plane joins the history together with the screen it produced */
        memmove(e->action, e->action + 1, sizeof(int32_t) * (MZO_ATARI_HIST - 1));
        e->action[MZO_ATARI_HIST - 1] = action;
    }
    memmove(e->frame[0], e->frame[1], sizeof(e->frame[0]) * (MZO_ATARI_HIST - 1));
    memmove(e->has_frame, e->has_frame + 1, MZO_ATARI_HIST - 1);
    memcpy(e->frame[MZO_ATARI_HIST - 1], frame_chw, sizeof(e->frame[0]));
    e->has_frame[MZO_ATARI_HIST - 1] = 1;
    e->terminal = terminal;
}

/* ZeroActor::beforeNNEvaluation, zero_actor.cpp:51-58 */
void mzo_select(mzo_batch* b, const uint8_t* rotations, float* features)
{
    int S2 = b->cfg.num_simulation + 2;
    for (int g = 0; g < b->cfg.num_games; ++g) {
        tree* t = &b->trees[g];
        int32_t* path = b->path + (size_t)g * S2;
        int len = 0, node = 0;
        path[len++] = 0;
        if (b->cfg.use_gumbel && t->count[0] != 0.0f) {
            /* GumbelZero::selection, gumbel_zero.cpp:70-85: the least visited candidate (ties: larger logit), PUCT below it */
            int best = 0;
            for (int i = 1; i < t->num_cand; ++i) {
                int c = t->cand[i], d = t->cand[best];
                if (t->count[c] < t->count[d] || (t->count[c] == t->count[d] && t->logit[c] > t->logit[d])) { best = i; }
            }
            node = t->cand[best];
            path[len++] = node;
        }
        while (t->num_children[node] > 0) { /* mcts.cpp:139-148 */
            node = select_child(b, t, node);
            path[len++] = node;
        }
        b->path_len[g] = len;
        if (b->cfg.muzero) { /* zero_actor.cpp:59-67: root features for the initial inference, nothing else touches the environment */
            b->rotation[g] = 0;
            if (features) {
                if (len == 1 && b->atari) {
                    atari_features(&b->atari[g], features + (size_t)g * b->F);
                } else if (len == 1) {
                    mzo_env_features(&b->root_env[g], 0, features + (size_t)g * b->F);
                } else {
                    memset(features + (size_t)g * b->F, 0, sizeof(float) * (size_t)b->F);
                }
            }
            continue;
        }
        /* getEnvironmentTransition, zero_actor.cpp:247-252 */
        mzo_env* e = &b->leaf_env[g];
        *e = b->root_env[g];
        for (int i = 1; i < len; ++i) { mzo_env_act(e, t->action[path[i]], t->player[path[i]]); }
        b->rotation[g] = (rotations ? rotations[g] : 0);
        if (features) { mzo_env_features(e, b->rotation[g], features + (size_t)g * b->F); }
    }
}

/* mcts.cpp:20-28 with weight 1 */
static void node_add(tree* t, int i, float value)
{
    t->count[i] += 1.0f;
    t->mean[i] += 1.0f * (value - t->mean[i]) / t->count[i];
}

void mzo_apply(mzo_batch* b, const float* policy, const float* logits, const float* value, const float* noise)
{
    mzo_apply_mz(b, policy, logits, value, NULL, noise);
}

static void apply_game(mzo_batch* b, int g, const float* policy, const float* logits, const float* value, const float* reward, const float* noise);

/* ZeroActor::afterNNEvaluation for every game */
void mzo_apply_mz(mzo_batch* b, const float* policy, const float* logits, const float* value, const float* reward, const float* noise)
{
    for (int g = 0; g < b->cfg.num_games; ++g) { apply_game(b, g, policy, logits, value, reward, noise); }
}

/* ZeroActor::afterNNEvaluation, zero_actor.cpp:74-98, for game g: the network outputs are indexed [g] like the batch they came from */
static void apply_game(mzo_batch* b, int g, const float* policy, const float* logits, const float* value, const float* reward, const float* noise)
{
    int S2 = b->cfg.num_simulation + 2, A = b->A;
    {
        tree* t = &b->trees[g];
        const int32_t* path = b->path + (size_t)g * S2;
        int len = b->path_len[g];
        if (len == 0) { return; }
        int leaf = path[len - 1];
        const mzo_env* e = &b->leaf_env[g];
        float v;
        if (b->cfg.muzero) {
            /* calculateMuZeroActionPolicy, zero_actor.cpp:231-245: every action below the root, the legal ones at the root */
            const mzo_env* re = &b->root_env[g];
            const int turn = (b->cfg.game == MZO_GAME_ATARI ? 1 : (t->player[leaf] == 1 ? 2 : 1)); /* leaf_node->getAction().nextPlayer() */
            int32_t cand_a[MZO_MAX_ACTIONS];
            int k = 0;
            float cand_p[MZO_MAX_ACTIONS], cand_l[MZO_MAX_ACTIONS];
            for (int a = 0; a < A; ++a) {
                if (leaf == 0 && !(b->cfg.game == MZO_GAME_ATARI ? (int)((b->cfg.atari_legal_mask >> a) & 1u) : mzo_env_is_legal(re, a, turn))) { continue; }
                cand_a[k] = a, cand_p[k] = policy[(size_t)g * A + a], cand_l[k] = logits[(size_t)g * A + a];
                ++k;
            }
            mzo_std_sort_candidates(k, cand_a, cand_p, cand_l); /* zero_actor.cpp:241-243 */
            t->first_child[leaf] = t->cursor;
            t->num_children[leaf] = k;
            for (int i = 0; i < k; ++i) {
                int c = t->cursor + i;
                node_reset(t, c);
                t->action[c] = (int16_t)cand_a[i];
                t->player[c] = (uint8_t)turn;
                t->policy[c] = cand_p[i];
                t->logit[c] = cand_l[i];
            }
            t->cursor += k;
            v = value[g];
            t->slot[leaf] = (int16_t)t->count[0]; /* hidden states are stored in evaluation order (zero_actor.cpp:90) */
        } else if (!mzo_env_is_terminal(e)) {
            /* calculateAlphaZeroActionPolicy, zero_actor.cpp:215-229 */
            int32_t cand_a[MZO_MAX_ACTIONS];
            int k = 0;
            float cand_p[MZO_MAX_ACTIONS], cand_l[MZO_MAX_ACTIONS];
            for (int a = 0; a < A; ++a) {
                if (!mzo_env_is_legal(e, a, e->turn)) { continue; }
                int ra = (e->game == MZO_GAME_HEX ? a : mzo_rotate_position(b->rotation[g], a, e->n)); /* getRotateAction, zero_actor.cpp:222 (identity for Hex, hex.h:65); the pass does not rotate (rotation.h:54) */
                cand_a[k] = a, cand_p[k] = policy[(size_t)g * A + ra], cand_l[k] = logits[(size_t)g * A + ra];
                ++k;
            }
            mzo_std_sort_candidates(k, cand_a, cand_p, cand_l); /* zero_actor.cpp:225-227 */
            /* expand, mcts.cpp:151-164 */
            t->first_child[leaf] = t->cursor;
            t->num_children[leaf] = k;
            for (int i = 0; i < k; ++i) {
                int c = t->cursor + i;
                node_reset(t, c);
                t->action[c] = (int16_t)cand_a[i];
                t->player[c] = (uint8_t)e->turn;
                t->policy[c] = cand_p[i];
                t->logit[c] = cand_l[i];
            }
            t->cursor += k;
            v = value[g];
        } else {
            v = mzo_env_eval_score(e, 0);
        }
        /* backup, mcts.cpp:166-179. AlphaZero: env reward, 0 for the board games (go.h:50, tictactoe.h:25); MuZero: the reward head's
         * output, 0 unless the network is muzero_atari (muzero_network.h:25,165-171) */
        float updated = v;
        t->value[leaf] = v;
        t->reward[leaf] = (b->cfg.muzero && reward ? reward[g] : 0.0f);
        for (int i = len - 1; i >= 0; --i) {
            int n = path[i];
            float old_mean = t->reward[n] + b->cfg.reward_discount * t->mean[n];
            node_add(t, n, updated);
            update_value_bound(b, t, old_mean, t->reward[n] + b->cfg.reward_discount * t->mean[n]);
            updated = t->reward[n] + b->cfg.reward_discount * updated;
        }
        /* addNoiseToNodeChildren, zero_actor.cpp:194-204: only when the evaluated leaf is the root */
        if (leaf == 0 && noise && t->num_children[0] > 0) {
            const float eps = b->cfg.dirichlet_epsilon;
            for (int i = 0; i < t->num_children[0]; ++i) {
                int c = t->first_child[0] + i;
                float nz = noise[(size_t)g * A + i];
                t->noise[c] = nz;
                if (b->cfg.gumbel_noise) { /* zero_actor.cpp:205-211 */
                    t->logit[c] = t->logit[c] + nz;
                } else {
                    t->policy[c] = (1 - eps) * t->policy[c] + eps * nz;
                }
            }
        }
        if (b->cfg.use_gumbel) { sequential_halving(b, t); } /* zero_actor.cpp:97 */
        b->path_len[g] = 0;
    }
}

/* ---- console think() with actor_mcts_think_batch_size = K (zero_actor.cpp:36-49,129-157), AlphaZero networks ----
 * ZeroActor::step's first loop for every game: batch_size = min(K, simulations left) selections, one after the other; a leaf whose virtual
 * loss is still 0 joins the batch (:140-142), every selection adds a virtual loss to its path (:143). Arrays are lane-major [K][B]. */
void mzo_think_select(mzo_batch* b, int K, const uint8_t* rotations, float* features, int32_t* path_len)
{
    const int B = b->cfg.num_games, S2 = b->cfg.num_simulation + 2;
    if (b->think_k != K) {
        free(b->t_path), free(b->t_len), free(b->t_rot), free(b->t_env);
        b->think_k = K;
        b->t_path = (int32_t*)calloc((size_t)K * B * S2, sizeof(int32_t));
        b->t_len = (int32_t*)calloc((size_t)K * B, sizeof(int32_t));
        b->t_rot = (uint8_t*)calloc((size_t)K * B, 1);
        b->t_env = (mzo_env*)calloc((size_t)K * B, sizeof(mzo_env));
    }
    for (int g = 0; g < B; ++g) {
        tree* t = &b->trees[g];
        const int left = b->cfg.num_simulation + 1 - (int)t->count[0];
        /* zero_actor.cpp:133-135: an AlphaZero network also batches the root's first evaluation, a MuZero network's initial inference is a batch of one */
        const int batch = ((b->cfg.muzero && t->count[0] == 0.0f) ? 1 : (K < left ? K : left));
        for (int k = 0; k < K; ++k) {
            const size_t l = (size_t)k * B + g;
            b->t_len[l] = 0;
            if (path_len) { path_len[l] = 0; }
            if (k >= batch) { continue; }
            int32_t* path = b->t_path + l * S2;
            int len = 0, node = 0;
            path[len++] = 0;
            while (t->num_children[node] > 0) { /* mcts.cpp:139-148 */
                node = select_child(b, t, node);
                path[len++] = node;
            }
            const int fresh = (t->vloss[node] == 0.0f);
            if (b->cfg.muzero) { /* zero_actor.cpp:59-67: root planes for the initial inference; below the root the network consumes (hidden state, action) */
                b->t_rot[l] = 0;
                if (features) {
                    if (len == 1) {
                        mzo_env_features(&b->root_env[g], 0, features + l * (size_t)b->F);
                    } else {
                        memset(features + l * (size_t)b->F, 0, sizeof(float) * (size_t)b->F);
                    }
                }
            } else {
                /* beforeNNEvaluation pushes the position whether or not it will be used (zero_actor.cpp:54-57) */
                mzo_env* e = &b->t_env[l];
                *e = b->root_env[g];
                for (int i = 1; i < len; ++i) { mzo_env_act(e, t->action[path[i]], t->player[path[i]]); }
                b->t_rot[l] = (rotations ? rotations[l] : 0);
                if (features) { mzo_env_features(e, b->t_rot[l], features + l * (size_t)b->F); }
            }
            for (int i = 0; i < len; ++i) { t->vloss[path[i]] += 1.0f; }
            b->t_len[l] = (fresh ? len : -len);
            if (path_len) { path_len[l] = b->t_len[l]; }
        }
    }
}

/* ZeroActor::step's second loop (zero_actor.cpp:147-156): afterNNEvaluation for the queried lanes in selection order, then the leaf's virtual loss
 * comes off every node of the path. policy / logits [K][B][A], value [K][B]; noise [B][A] by root child index (NULL = none) */
void mzo_think_apply(mzo_batch* b, const float* policy, const float* logits, const float* value, const float* noise)
{
    const int B = b->cfg.num_games, S2 = b->cfg.num_simulation + 2, K = b->think_k;
    for (int g = 0; g < B; ++g) {
        tree* t = &b->trees[g];
        for (int k = 0; k < K; ++k) {
            const size_t l = (size_t)k * B + g;
            const int len = b->t_len[l];
            if (len <= 0) { continue; }
            const int32_t* path = b->t_path + l * S2;
            memcpy(b->path + (size_t)g * S2, path, sizeof(int32_t) * (size_t)len);
            b->path_len[g] = len, b->rotation[g] = b->t_rot[l];
            if (!b->cfg.muzero) { b->leaf_env[g] = b->t_env[l]; }
            /* apply_game reads the outputs at index g of the arrays it is given: hand it the lane's section */
            apply_game(b, g, policy + (size_t)k * B * b->A, logits + (size_t)k * B * b->A, value + (size_t)k * B, NULL, noise);
            const float vl = t->vloss[path[len - 1]];
            for (int i = 0; i < len; ++i) { t->vloss[path[i]] -= vl; }
        }
    }
}

/* MuZero think(): what lane k of tree g hands to the recurrent inference — the evaluation slot of the leaf's parent and the leaf's action (-1, -1: the root) */
void mzo_think_leaf(const mzo_batch* b, int k, int g, int32_t* parent_slot, int32_t* action)
{
    const size_t l = (size_t)k * b->cfg.num_games + g;
    const int len = b->t_len[l] < 0 ? -b->t_len[l] : b->t_len[l];
    const int32_t* path = b->t_path + l * (size_t)(b->cfg.num_simulation + 2);
    *parent_slot = (len > 1 ? b->trees[g].slot[path[len - 2]] : -1);
    *action = (len > 1 ? b->trees[g].action[path[len - 1]] : -1);
}

int mzo_num_simulation_done(const mzo_batch* b, int g) { return (int)b->trees[g].count[0]; }
int mzo_path_len(const mzo_batch* b, int g) { return b->path_len[g]; }
int mzo_leaf_action(const mzo_batch* b, int g)
{
    const int32_t* path = b->path + (size_t)g * (b->cfg.num_simulation + 2);
    return b->path_len[g] > 0 ? b->trees[g].action[path[b->path_len[g] - 1]] : -1;
}
int mzo_leaf_parent_slot(const mzo_batch* b, int g)
{
    const int32_t* path = b->path + (size_t)g * (b->cfg.num_simulation + 2);
    return b->path_len[g] > 1 ? b->trees[g].slot[path[b->path_len[g] - 2]] : -1;
}
int mzo_path_hash(const mzo_batch* b, int g)
{
    const int32_t* path = b->path + (size_t)g * (b->cfg.num_simulation + 2);
    uint32_t h = 2166136261u;
    for (int k = 1; k < b->path_len[g]; ++k) { h = (h ^ (uint32_t)(int32_t)b->trees[g].action[path[k]]) * 16777619u; }
    return (int)(h & 0x7fffffffu);
}
const mzo_env* mzo_root_env(const mzo_batch* b, int g) { return &b->root_env[g]; }

void mzo_root(const mzo_batch* b, int g, mzo_root_out* out)
{
    const tree* t = &b->trees[g];
    memset(out, 0, sizeof(*out));
    out->num_children = t->num_children[0];
    out->count = t->count[0], out->mean = t->mean[0], out->value = t->value[0];
    for (int i = 0; i < b->A; ++i) { out->action[i] = -1; }
    for (int i = 0; i < t->num_children[0]; ++i) {
        int c = t->first_child[0] + i;
        out->action[i] = t->action[c];
        out->c_count[i] = t->count[c], out->c_mean[i] = t->mean[c], out->c_policy[i] = t->policy[c];
        out->c_logit[i] = t->logit[c], out->c_noise[i] = t->noise[c], out->c_value[i] = t->value[c];
    }
}

void mzo_root_extra(const mzo_batch* b, int g, float* c_reward, int32_t* bound_size, float* bound_lo, float* bound_hi)
{
    const tree* t = &b->trees[g];
    for (int i = 0; i < t->num_children[0]; ++i) { c_reward[i] = t->reward[t->first_child[0] + i]; }
    *bound_size = t->vb_n;
    *bound_lo = (t->vb_n ? t->vb_key[0] : 0.0f), *bound_hi = (t->vb_n ? t->vb_key[t->vb_n - 1] : 0.0f);
}

float mzo_root_normalized_mean(const mzo_batch* b, int g, int i)
{
    const tree* t = &b->trees[g];
    return normalized_mean(b, t, i < 0 ? 0 : t->first_child[0] + i);
}

/* mcts.cpp:91-104: first child with the strictly largest count */
int mzo_select_by_max_count(const mzo_batch* b, int g)
{
    const tree* t = &b->trees[g];
    float max_count = 0.0f;
    int selected = -1;
    for (int i = 0; i < t->num_children[0]; ++i) {
        int c = t->first_child[0] + i;
        if (t->count[c] <= max_count) { continue; }
        max_count = t->count[c];
        selected = t->action[c];
    }
    return selected;
}

/* BaseActor::act (base_actor.cpp:22-30) + SlaveThread::handleSearchDone's resetSearch (actor_group.cpp:116-134) */
int mzo_play(mzo_batch* b, int g, int action)
{
    mzo_env* e = &b->root_env[g];
    if (b->cfg.game == MZO_GAME_ATARI) { /* the emulator is the host's (mzo_atari_observe): only legality and the move count live here */
        if (action < 0 || action >= 18 || !((b->cfg.atari_legal_mask >> action) & 1u)) { return 0; }
        ++e->num_moves;
        mzo_reset_search(b, g);
        return 1;
    }
    int ok = mzo_env_act(e, action, e->turn);
    if (ok) { mzo_reset_search(b, g); }
    return ok;
}
