/*
 * Oracle port — environments (TicTacToe, Go). TEST INFRASTRUCTURE ONLY (see mzo.h).
 *
 * Restates, with plain arrays and flood fill instead of the reference's incremental
 * block/area/Benson bookkeeping (which is not observable through legality, terminal test,
 * score or features in plain Go):
 *   environment/go/go.cpp:19-43     Zobrist keys from std::mt19937_64(0)
 *   environment/go/go.cpp:132-190   act
 *   environment/go/go.cpp:208-244   isLegalAction (suicide + positional/situational superko)
 *   environment/go/go.cpp:246-257   isTerminal
 *   environment/go/go.cpp:259-278,703-723  getEvalScore / Tromp-Taylor territory
 *   environment/go/go.cpp:280-308   getFeatures (18 planes)
 *   environment/tictactoe/tictactoe.cpp:11-146
 *   utils/rotation.h:22-93          rotation tables
 */
#include "mzo.h"
#include <string.h>

/* ---- std::mt19937_64 (ISO C++ [rand.predef]; parameters of the 64-bit Mersenne twister) ---- */
typedef struct {
    uint64_t mt[312];
    int idx;
} mt64;

static void mt64_seed(mt64* g, uint64_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 312; ++i) { g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i; }
    g->idx = 312;
}

static uint64_t mt64_next(mt64* g)
{
    if (g->idx >= 312) {
        for (int i = 0; i < 312; ++i) {
            uint64_t x = (g->mt[i] & 0xFFFFFFFF80000000ULL) | (g->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
            uint64_t xa = x >> 1;
            if (x & 1ULL) { xa ^= 0xB5026F5AA96619E9ULL; }
            g->mt[i] = g->mt[(i + 156) % 312] ^ xa;
        }
        g->idx = 0;
    }
    uint64_t x = g->mt[g->idx++];
    x ^= (x >> 29) & 0x5555555555555555ULL;
    x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
    x ^= (x << 37) & 0xFFF7EEE000000000ULL;
    x ^= (x >> 43);
    return x;
}

uint64_t mzo_mt19937_64_nth(uint64_t seed, int nth)
{
    mt64 g;
    mt64_seed(&g, seed);
    uint64_t v = 0;
    for (int i = 0; i <= nth; ++i) { v = mt64_next(&g); }
    return v;
}

/* go.cpp:19-32: turn key, then for each of the 361 positions: empty, black, white */
static uint64_t g_turn_key;
static uint64_t g_grid_key[MZO_MAX_CELLS][3];
static int g_keys_ready = 0;

static void go_init_keys(void)
{
    if (g_keys_ready) { return; }
    mt64 g;
    mt64_seed(&g, 0);
    g_turn_key = mt64_next(&g);
    for (int pos = 0; pos < MZO_MAX_CELLS; ++pos) {
        g_grid_key[pos][0] = mt64_next(&g);
        g_grid_key[pos][1] = mt64_next(&g);
        g_grid_key[pos][2] = mt64_next(&g);
    }
    g_keys_ready = 1;
}

uint64_t mzo_go_key(int pos, int player)
{
    go_init_keys();
    return (pos < 0 ? g_turn_key : g_grid_key[pos][player]);
}

/* ---- rotation (utils/rotation.h:22-31 reversed table, :51-93 getPositionByRotating) ---- */
int mzo_reversed_rotation(int r)
{
    static const int rev[8] = {0, 3, 2, 1, 4, 5, 6, 7};
    return rev[r];
}

int mzo_rotate_position(int rotation, int pos, int n)
{
    if (pos == n * n) { return pos; }
    /* doubled coordinates relative to the centre keep everything in integers */
    int x = 2 * (pos % n) - (n - 1), y = 2 * (pos / n) - (n - 1), rx = x, ry = y;
    switch (rotation) {
        case 0: rx = x, ry = y; break;
        case 1: rx = y, ry = -x; break;
        case 2: rx = -x, ry = -y; break;
        case 3: rx = -y, ry = x; break;
        case 4: rx = x, ry = -y; break;
        case 5: rx = -y, ry = -x; break;
        case 6: rx = -x, ry = y; break;
        case 7: rx = y, ry = x; break;
    }
    return ((ry + (n - 1)) / 2) * n + (rx + (n - 1)) / 2;
}

/* ---- common ---- */
static int other(int p) { return p == 1 ? 2 : 1; }

void mzo_env_init(mzo_env* e, int game, int n, float komi, int ko_situational)
{
    go_init_keys();
    memset(e, 0, sizeof(*e));
    e->game = game;
    e->n = (game == MZO_GAME_TICTACTOE ? 3 : (game == MZO_GAME_ATARI ? 6 : n)); /* Atari: n = side of the hidden state (atari.h:25-26) */
    e->turn = 1; /* go.cpp:105, tictactoe.cpp:13 */
    e->komi = komi;
    e->turn_key = (ko_situational ? g_turn_key : 0); /* go.cpp:45-49 */
    if (game == MZO_GAME_OTHELLO) { /* othello.cpp:13-31: initial four stones; Black (kPlayer1) to move */
        const int bs = e->n, init_place = bs * (bs / 2 - (1 - bs % 2)) + (bs / 2 - 1);
        e->board[init_place + 1] = 2, e->board[init_place + bs] = 2;
        e->board[init_place] = 1, e->board[init_place + bs + 1] = 1;
    }
}

void mzo_env_set_flags(mzo_env* e, int flags) { e->flags = flags; }
int mzo_env_num_actions(const mzo_env* e)
{
    if (e->game == MZO_GAME_ATARI) { return 18; } /* atari.h:20 */
    if (e->game == MZO_GAME_TICTACTOE) { return 9; }
    return (e->game == MZO_GAME_GOMOKU || e->game == MZO_GAME_HEX) ? e->n * e->n : e->n * e->n + 1; /* gomoku.h:32, hex.h:56: no pass */
}
int mzo_env_input_channels(const mzo_env* e) { return (e->game == MZO_GAME_GO || e->game == MZO_GAME_NOGO || e->game == MZO_GAME_KILLALLGO) ? 18 : 4; }

/* neighbour order of go_grid.h:43-54: up(+n), right(+1), down(-n), left(-1) */
static int neighbours(int n, int pos, int* out)
{
    int x = pos % n, y = pos / n, k = 0;
    if (y + 1 < n) { out[k++] = pos + n; }
    if (x + 1 < n) { out[k++] = pos + 1; }
    if (y - 1 >= 0) { out[k++] = pos - n; }
    if (x - 1 >= 0) { out[k++] = pos - 1; }
    return k;
}

/* flood the block containing `start`; mark cells with `tag`; return #liberties and block hash */
static int go_block(const mzo_env* e, int start, int* mark, int tag, int* cells, int* ncells, uint64_t* hash)
{
    int n = e->n, colour = e->board[start], top = 0, libs = 0, count = 0;
    int stack[MZO_MAX_CELLS];
    uint8_t libmark[MZO_MAX_CELLS];
    memset(libmark, 0, (size_t)(n * n));
    uint64_t h = 0;
    stack[top++] = start;
    mark[start] = tag;
    while (top > 0) {
        int p = stack[--top];
        cells[count++] = p;
        h ^= g_grid_key[p][colour];
        int nb[4], k = neighbours(n, p, nb);
        for (int i = 0; i < k; ++i) {
            int q = nb[i];
            if (e->board[q] == 0) {
                if (!libmark[q]) {
                    libmark[q] = 1;
                    ++libs;
                }
            } else if (e->board[q] == colour && mark[q] != tag) {
                mark[q] = tag;
                stack[top++] = q;
            }
        }
    }
    *ncells = count;
    *hash = h;
    return libs;
}

static int hash_seen(const mzo_env* e, uint64_t h)
{
    for (int i = 0; i < e->num_moves; ++i) {
        if (e->hashes[i] == h) { return 1; }
    }
    return 0;
}

static int go_is_legal(const mzo_env* e, int action, int player)
{
    int n = e->n;
    if (action == n * n) { return 1; }                  /* go.cpp:213 */
    if (action < 0 || action > n * n) { return 0; }
    if (e->board[action] != 0) { return 0; }            /* go.cpp:218 */
    int legal = 0;
    uint64_t new_hash = e->hash ^ e->turn_key ^ g_grid_key[action][player]; /* go.cpp:222 */
    int mark[MZO_MAX_CELLS];
    memset(mark, 0, sizeof(int) * (size_t)(n * n));
    int nb[4], k = neighbours(n, action, nb), cells[MZO_MAX_CELLS], nc;
    for (int i = 0; i < k; ++i) {
        int q = nb[i];
        if (e->board[q] == 0) {
            legal = 1; /* go.cpp:225-226 */
            continue;
        }
        if (mark[q]) { continue; } /* block already examined, go.cpp:229 */
        uint64_t bh;
        int libs = go_block(e, q, mark, i + 1, cells, &nc, &bh);
        if (e->board[q] == player) {
            if (libs > 1) { legal = 1; } /* go.cpp:232-233 */
        } else if (libs == 1) {          /* capture, go.cpp:235-238 */
            new_hash ^= bh;
            legal = 1;
        }
    }
    return legal && !hash_seen(e, new_hash); /* go.cpp:243 */
}

/* NoGoEnv::isLegalAction (nogo.h:27-59): no pass, no suicide, no capture, no repetition rule */
static int nogo_is_legal(const mzo_env* e, int action, int player)
{
    int n = e->n;
    if (action < 0 || action >= n * n) { return 0; } /* the pass is never legal, nogo.h:32 */
    if (e->board[action] != 0) { return 0; }
    int legal = 0;
    int mark[MZO_MAX_CELLS];
    memset(mark, 0, sizeof(int) * (size_t)(n * n));
    int nb[4], k = neighbours(n, action, nb), cells[MZO_MAX_CELLS], nc;
    for (int i = 0; i < k; ++i) {
        int q = nb[i];
        if (e->board[q] == 0) {
            legal = 1;
            continue;
        }
        if (mark[q]) { continue; }
        uint64_t bh;
        int libs = go_block(e, q, mark, i + 1, cells, &nc, &bh);
        if (e->board[q] == player) {
            if (libs > 1) { legal = 1; }
        } else if (libs == 1) {
            return 0; /* would capture */
        }
    }
    return legal;
}

static void push_history(mzo_env* e)
{
    memcpy(e->hist[e->num_moves % MZO_HIST], e->board, (size_t)(e->n * e->n));
    e->hashes[e->num_moves] = e->hash;
}

static int killallgo_is_legal(const mzo_env* e, int action, int player);

static int go_act(mzo_env* e, int action, int player)
{
    if (!(e->game == MZO_GAME_NOGO ? nogo_is_legal(e, action, player)
                                   : (e->game == MZO_GAME_KILLALLGO ? killallgo_is_legal(e, action, player) : go_is_legal(e, action, player)))) { return 0; } /* go.cpp:134 (virtual call) */
    int n = e->n;
    e->turn = other(player);   /* go.cpp:140 */
    e->hash ^= e->turn_key;    /* go.cpp:141 */
    e->actions[e->num_moves] = (int16_t)action;
    if (action != n * n) {
        e->board[action] = (uint8_t)player;
        e->hash ^= g_grid_key[action][player]; /* go.cpp:154 */
        int nb[4], k = neighbours(n, action, nb), cells[MZO_MAX_CELLS], nc;
        int mark[MZO_MAX_CELLS];
        memset(mark, 0, sizeof(int) * (size_t)(n * n));
        for (int i = 0; i < k; ++i) {
            int q = nb[i];
            if (e->board[q] != other(player) || mark[q]) { continue; }
            uint64_t bh;
            int libs = go_block(e, q, mark, i + 1, cells, &nc, &bh);
            if (libs == 0) { /* go.cpp:174, removeBlockFromBoard go.cpp:388-433 */
                for (int c = 0; c < nc; ++c) { e->board[cells[c]] = 0; }
                e->hash ^= bh;
                /* cells are now empty: marks of removed stones must not suppress later blocks */
            }
        }
    }
    push_history(e); /* go.cpp:145-147,180-182 */
    e->num_moves++;
    return 1;
}

static int go_is_terminal(const mzo_env* e)
{
    int n2 = e->n * e->n, m = e->num_moves;
    if (m >= 2 && e->actions[m - 1] == n2 && e->actions[m - 2] == n2) { return 1; } /* go.cpp:249-251 */
    return m > 2 * n2;                                                                /* go.cpp:254 */
}

/* ---- KillAllGo (environment/killallgo/killallgo.cpp:27-48; env_killallgo_use_seki = false, the default) ----
 * Unconditional life (Benson) of `player`'s stones, GoEnv::findBensonBitboard (go.cpp:614-676) evaluated on the whole position: the reference keeps
 * its area / Benson bitboards incrementally (go.cpp:464-612) and re-evaluates them only where a move can change them; the restatement recomputes
 * them from the position after every move (pinned against playouts of the reference's own environment, tests/golden/env_killallgo7*).
 * Blocks = connected groups of the player's stones; areas = connected regions of everything else (empty points and opposing stones, go.cpp:573-589);
 * an area is vital to a neighbouring block when every EMPTY point of it is a liberty of that block (go.cpp:627-633). Returns the number of points
 * of the Benson bitboard (surviving blocks and their surviving areas). */
static int benson_count(const mzo_env* e, int player)
{
    const int n = e->n, n2 = n * n;
    int block_of[MZO_MAX_CELLS], area_of[MZO_MAX_CELLS], stack[MZO_MAX_CELLS];
    int nblocks = 0, nareas = 0;
    for (int i = 0; i < n2; ++i) { block_of[i] = area_of[i] = -1; }
    for (int s0 = 0; s0 < n2; ++s0) {
        const int own = (e->board[s0] == player);
        int* lab = own ? block_of : area_of;
        if (lab[s0] >= 0) { continue; }
        const int id = own ? nblocks++ : nareas++;
        int top = 0;
        stack[top++] = s0, lab[s0] = id;
        while (top > 0) {
            int p = stack[--top], nb[4], k = neighbours(n, p, nb);
            for (int j = 0; j < k; ++j) {
                int q = nb[j];
                if ((e->board[q] == player) == own && lab[q] < 0) { lab[q] = id, stack[top++] = q; }
            }
        }
    }
    if (nblocks == 0) { return 0; }
    /* adjacency and vitality */
    static uint8_t adj[MZO_MAX_CELLS][MZO_MAX_CELLS], vital[MZO_MAX_CELLS][MZO_MAX_CELLS]; /* [block][area] */
    for (int b = 0; b < nblocks; ++b) {
        memset(adj[b], 0, (size_t)nareas), memset(vital[b], 0, (size_t)nareas);
    }
    for (int p = 0; p < n2; ++p) {
        if (block_of[p] < 0) { continue; }
        int nb[4], k = neighbours(n, p, nb);
        for (int j = 0; j < k; ++j) {
            if (area_of[nb[j]] >= 0) { adj[block_of[p]][area_of[nb[j]]] = 1; }
        }
    }
    for (int b = 0; b < nblocks; ++b) {
        for (int a = 0; a < nareas; ++a) {
            if (!adj[b][a]) { continue; }
            int ok = 1;
            for (int p = 0; p < n2 && ok; ++p) {
                if (area_of[p] != a || e->board[p] != 0) { continue; }
                int nb[4], k = neighbours(n, p, nb), lib = 0;
                for (int j = 0; j < k; ++j) { lib |= (block_of[nb[j]] == b); }
                ok = lib;
            }
            vital[b][a] = (uint8_t)ok;
        }
    }
    uint8_t in_b[MZO_MAX_CELLS], in_a[MZO_MAX_CELLS];
    memset(in_b, 0, sizeof(in_b)), memset(in_a, 0, sizeof(in_a));
    for (int b = 0; b < nblocks; ++b) {
        for (int a = 0; a < nareas; ++a) {
            if (vital[b][a]) { in_b[b] = 1, in_a[a] = 1; } /* go.cpp:630-632 */
        }
    }
    for (int changed = 1; changed;) { /* go.cpp:638-671 */
        changed = 0;
        for (int b = 0; b < nblocks; ++b) {
            if (!in_b[b]) { continue; }
            int cnt = 0;
            for (int a = 0; a < nareas; ++a) { cnt += (vital[b][a] && in_a[a]); }
            if (cnt < 2) { in_b[b] = 0, changed = 1; }
        }
        for (int a = 0; a < nareas; ++a) {
            if (!in_a[a]) { continue; }
            for (int b = 0; b < nblocks; ++b) {
                if (adj[b][a] && !in_b[b]) { in_a[a] = 0, changed = 1; }
            }
        }
    }
    int count = 0;
    for (int p = 0; p < n2; ++p) { count += (block_of[p] >= 0 ? in_b[block_of[p]] : in_a[area_of[p]]); }
    return count;
}

/* KillAllGoEnv::isLegalAction, killallgo.cpp:27-32: Black opens with two stones (move 0 and 2 must be stones, move 1 must be the pass) */
static int killallgo_is_legal(const mzo_env* e, int action, int player)
{
    const int pass = e->n * e->n;
    if (e->num_moves == 1) { return action == pass; }
    if (e->num_moves < 3) { return action != pass && go_is_legal(e, action, player); }
    return go_is_legal(e, action, player);
}

/* go.cpp:703-723 */
static float go_eval_score(const mzo_env* e, int is_resign)
{
    int n = e->n, n2 = n * n, winner;
    if (is_resign) {
        winner = other(e->turn); /* go.cpp:262-263 */
    } else {
        float terr_b = 0.0f, terr_w = 0.0f;
        int cb = 0, cw = 0;
        for (int p = 0; p < n2; ++p) {
            cb += (e->board[p] == 1);
            cw += (e->board[p] == 2);
        }
        terr_b = (float)cb;
        terr_w = (float)cw + e->komi; /* GamePair<float>(count, count + komi_) */
        uint8_t seen[MZO_MAX_CELLS];
        memset(seen, 0, (size_t)n2);
        for (int s = 0; s < n2; ++s) {
            if (e->board[s] != 0 || seen[s]) { continue; }
            int stack[MZO_MAX_CELLS], top = 0, size = 0, touch_b = 0, touch_w = 0;
            stack[top++] = s;
            seen[s] = 1;
            while (top > 0) {
                int p = stack[--top];
                ++size;
                int nb[4], k = neighbours(n, p, nb);
                for (int i = 0; i < k; ++i) {
                    int q = nb[i];
                    if (e->board[q] == 1) {
                        touch_b = 1;
                    } else if (e->board[q] == 2) {
                        touch_w = 1;
                    } else if (!seen[q]) {
                        seen[q] = 1;
                        stack[top++] = q;
                    }
                }
            }
            /* go.cpp:713-717: "surrounded only by black" is tested first and is also true for
             * a region with no surrounding stones at all (empty board) */
            if (!touch_w) {
                terr_b += (float)size;
            } else if (!touch_b) {
                terr_w += (float)size;
            }
        }
        winner = (terr_b > terr_w ? 1 : (terr_b < terr_w ? 2 : 0)); /* go.cpp:266-270 */
    }
    return winner == 1 ? 1.0f : (winner == 2 ? -1.0f : 0.0f);
}

/* go.cpp:280-308 */
static void go_features(const mzo_env* e, int rotation, float* out)
{
    int n = e->n, n2 = n * n, rev = mzo_reversed_rotation(rotation);
    for (int c = 0; c < 18; ++c) {
        for (int pos = 0; pos < n2; ++pos) {
            int rp = mzo_rotate_position(rev, pos, n);
            float v = 0.0f;
            if (c < 16) {
                int idx = e->num_moves - 1 - c / 2;
                if (idx >= 0) {
                    int player = (c % 2 == 0 ? e->turn : other(e->turn));
                    v = (e->hist[idx % MZO_HIST][rp] == player ? 1.0f : 0.0f);
                }
            } else if (c == 16) {
                v = (e->turn == 1 ? 1.0f : 0.0f);
            } else {
                v = (e->turn == 2 ? 1.0f : 0.0f);
            }
            out[c * n2 + pos] = v;
        }
    }
}


/* ---- othello (environment/othello/othello.cpp) ----
 * The reference keeps per-player "legal boards" refreshed by every non-pass act() with shift-and-mask
 * bitboard sweeps (othello.cpp:63-99,125-137). Those sweeps compute exactly the standard rule — a point is
 * playable when, in at least one of the 8 directions, a run of one or more opposing stones is followed by
 * an own stone — so legality here is that rule evaluated on the current board. legal_pass_ is "the legal
 * board is empty" (othello.cpp:135-136); it starts false (othello.cpp:17) and the initial position has
 * moves for both sides, so it too is a function of the board alone. */
static const int oth_dx[8] = {0, 0, -1, 1, -1, 1, 1, -1};
static const int oth_dy[8] = {1, -1, 0, 0, 1, 1, -1, -1};

/* stones flipped by `player` playing the empty point pos (othello.cpp:63-82); flips may be NULL */
static int othello_flips(const mzo_env* e, int pos, int player, int* flips)
{
    const int n = e->n, x0 = pos % n, y0 = pos / n, opp = other(player);
    int total = 0;
    for (int d = 0; d < 8; ++d) {
        int x = x0 + oth_dx[d], y = y0 + oth_dy[d], run = 0;
        while (x >= 0 && x < n && y >= 0 && y < n && e->board[y * n + x] == opp) { x += oth_dx[d], y += oth_dy[d], ++run; }
        if (run == 0 || x < 0 || x >= n || y < 0 || y >= n || e->board[y * n + x] != player) { continue; }
        for (int k = 1; k <= run; ++k) {
            if (flips) { flips[total] = (y0 + k * oth_dy[d]) * n + x0 + k * oth_dx[d]; }
            ++total;
        }
    }
    return total;
}

static int othello_can_put(const mzo_env* e, int pos, int player) { return e->board[pos] == 0 && othello_flips(e, pos, player, 0) > 0; }

static int othello_has_move(const mzo_env* e, int player)
{
    for (int pos = 0; pos < e->n * e->n; ++pos) {
        if (othello_can_put(e, pos, player)) { return 1; }
    }
    return 0;
}

/* othello.cpp:192-199 */
static int othello_is_legal(const mzo_env* e, int action, int player)
{
    if (action < 0 || action > e->n * e->n) { return 0; }
    if (action == e->n * e->n) { return !othello_has_move(e, player); }
    return othello_can_put(e, action, player);
}

/* othello.cpp:102-139 */
static int othello_act(mzo_env* e, int action, int player)
{
    if (!othello_is_legal(e, action, player)) { return 0; }
    e->actions[e->num_moves++] = (int16_t)action;
    e->turn = other(player);
    if (action == e->n * e->n) { return 1; }
    int flips[MZO_MAX_CELLS];
    const int k = othello_flips(e, action, player, flips);
    e->board[action] = (uint8_t)player;
    for (int i = 0; i < k; ++i) { e->board[flips[i]] = (uint8_t)player; }
    return 1;
}

/* othello.cpp:201-207 */
static int othello_is_terminal(const mzo_env* e)
{
    const int pass = e->n * e->n;
    return e->num_moves >= 2 && e->actions[e->num_moves - 1] == pass && e->actions[e->num_moves - 2] == pass;
}

/* othello.cpp:209-236 */
static float othello_eval_score(const mzo_env* e, int is_resign)
{
    int r = 0;
    if (is_resign) {
        r = other(e->turn);
    } else if (!othello_has_move(e, 1) && !othello_has_move(e, 2)) {
        int c1 = 0, c2 = 0;
        for (int i = 0; i < e->n * e->n; ++i) { c1 += (e->board[i] == 1), c2 += (e->board[i] == 2); }
        r = (c1 > c2 ? 1 : (c1 < c2 ? 2 : 0));
    }
    return r == 1 ? 1.0f : (r == 2 ? -1.0f : 0.0f);
}

/* getActionFeatures (othello.cpp:257-262): one-hot plane of the action, all zero for a pass */
void mzo_env_action_features(const mzo_env* e, int action, float* out)
{
    const int cells = e->n * e->n;
    for (int i = 0; i < cells; ++i) { out[i] = 0.0f; }
    if (action >= 0 && action < cells) { out[action] = 1.0f; }
}

/* ---- gomoku (environment/gomoku/gomoku.cpp) ---- */
/* calculateNumberOfConnection, gomoku.cpp:150-164 */
static int gomoku_run(const mzo_env* e, int start, int dx, int dy)
{
    int n = e->n, x = start % n, y = start / n, count = 0, who = e->board[start];
    while (x >= 0 && x < n && y >= 0 && y < n && e->board[y * n + x] == who) { ++count, x += dx, y += dy; }
    return count;
}

/* updateWinner for the last move (gomoku.cpp:140-148); winner_ is a function of the board and the last action */
static int gomoku_winner(const mzo_env* e)
{
    if (e->num_moves == 0) { return 0; }
    const int pos = e->actions[e->num_moves - 1];
    static const int dirs[4][2] = {{1, 0}, {0, 1}, {1, 1}, {1, -1}};
    for (int d = 0; d < 4; ++d) {
        int c = gomoku_run(e, pos, dirs[d][0], dirs[d][1]) + gomoku_run(e, pos, -dirs[d][0], -dirs[d][1]) - 1;
        if ((e->flags & MZO_GOMOKU_EXACTLY_FIVE) ? (c == 5) : (c >= 5)) { return e->board[pos]; } /* gomoku.h:46 */
    }
    return 0;
}

/* gomoku.cpp:49-58 */
static int gomoku_is_legal(const mzo_env* e, int action, int player)
{
    (void)player;
    int n = e->n;
    if (action < 0 || action >= n * n) { return 0; }
    if (e->num_moves == 0 && (e->flags & MZO_GOMOKU_OUTER_OPEN)) {
        int i = action / n, j = action % n;
        return (i < 2 || i >= n - 2) || (j < 2 || j >= n - 2);
    }
    return e->board[action] == 0;
}

/* ---- hex (environment/hex/hex.cpp) ---- */
/* The reference keeps per-cell edge-connection flags, merged through the six neighbours of every new stone (hex.cpp:305-347);
 * a group's cells all carry the union of the edges the group touches, so "winner_" is: the group of the stone just placed
 * touches both of its owner's edges — Black (player 1) the columns x = 0 and x = n-1, White the rows y = 0 and y = n-1
 * (hex.cpp:47-58). Restated as a flood fill over the same neighbourhood. */
static int hex_connected(const mzo_env* e, int player)
{
    int n = e->n, top = 0, stack[MZO_MAX_CELLS];
    uint8_t seen[MZO_MAX_CELLS];
    memset(seen, 0, (size_t)(n * n));
    for (int i = 0; i < n; ++i) {
        int p = (player == 1 ? i * n : i); /* edge 1: x == 0 for Black, y == 0 for White */
        if (e->board[p] == player) {
            seen[p] = 1;
            stack[top++] = p;
        }
    }
    static const int dx[6] = {-1, 0, -1, 1, 0, 1}, dy[6] = {-1, -1, 0, 0, 1, 1}; /* hex.cpp:313-316 */
    while (top > 0) {
        int p = stack[--top], x = p % n, y = p / n;
        if (player == 1 ? (x == n - 1) : (y == n - 1)) { return 1; }
        for (int k = 0; k < 6; ++k) {
            int qx = x + dx[k], qy = y + dy[k];
            if (qx < 0 || qx >= n || qy < 0 || qy >= n) { continue; }
            int q = qy * n + qx;
            if (e->board[q] == player && !seen[q]) {
                seen[q] = 1;
                stack[top++] = q;
            }
        }
    }
    return 0;
}

static int hex_winner(const mzo_env* e) { return hex_connected(e, 1) ? 1 : (hex_connected(e, 2) ? 2 : 0); }

/* hex.cpp:83-94 */
static int hex_is_legal(const mzo_env* e, int action, int player)
{
    if (action < 0 || action >= e->n * e->n) { return 0; }
    return player == e->turn && (((e->flags & MZO_HEX_SWAP_RULE) && e->num_moves == 1) || e->board[action] == 0);
}

/* hex.cpp:21-66 */
static int hex_act(mzo_env* e, int action, int player)
{
    if (!hex_is_legal(e, action, player)) { return 0; }
    int n = e->n, id = action;
    if ((e->flags & MZO_HEX_SWAP_RULE) && e->num_moves == 1 && action == e->actions[0]) { /* swap: the first stone changes sides, mirrored */
        int row = e->actions[0] / n, col = e->actions[0] % n;
        id = (n - 1 - col) * n + (n - 1 - row);
        e->board[e->actions[0]] = 0;
    }
    e->board[id] = (uint8_t)player;
    e->actions[e->num_moves++] = (int16_t)action;
    e->turn = other(player);
    return 1;
}

/* ---- tictactoe (tictactoe.cpp:124-146 eval) ---- */
static int ttt_eval(const mzo_env* e)
{
    const uint8_t* b = e->board;
    int c;
    for (int i = 0; i < 3; ++i) {
        c = 3;
        for (int j = 0; j < 3; ++j) { c &= b[i * 3 + j]; }
        if (c) { return c; }
        c = 3;
        for (int j = 0; j < 3; ++j) { c &= b[j * 3 + i]; }
        if (c) { return c; }
    }
    c = 3;
    for (int i = 0; i < 3; ++i) { c &= b[i * 3 + i]; }
    if (c) { return c; }
    c = 3;
    for (int i = 0; i < 3; ++i) { c &= b[i * 3 + (2 - i)]; }
    return c;
}

int mzo_env_is_legal(const mzo_env* e, int action, int player)
{
    if (e->game == MZO_GAME_GO) { return go_is_legal(e, action, player); }
    if (e->game == MZO_GAME_KILLALLGO) { return killallgo_is_legal(e, action, player); }
    if (e->game == MZO_GAME_NOGO) { return nogo_is_legal(e, action, player); }
    if (e->game == MZO_GAME_GOMOKU) { return gomoku_is_legal(e, action, player); }
    if (e->game == MZO_GAME_HEX) { return hex_is_legal(e, action, player); }
    if (e->game == MZO_GAME_OTHELLO) { return othello_is_legal(e, action, player); }
    return action >= 0 && action < 9 && e->board[action] == 0; /* tictactoe.cpp:44-49 */
}

int mzo_env_act(mzo_env* e, int action, int player)
{
    if (e->game == MZO_GAME_GO || e->game == MZO_GAME_NOGO || e->game == MZO_GAME_KILLALLGO) { return go_act(e, action, player); }
    if (e->game == MZO_GAME_OTHELLO) { return othello_act(e, action, player); }
    if (e->game == MZO_GAME_HEX) { return hex_act(e, action, player); }
    if (e->game == MZO_GAME_GOMOKU) { /* gomoku.cpp:23-31 */
        if (!gomoku_is_legal(e, action, player)) { return 0; }
        e->actions[e->num_moves++] = (int16_t)action;
        e->board[action] = (uint8_t)player;
        e->turn = other(player);
        return 1;
    }
    if (!mzo_env_is_legal(e, action, player)) { return 0; } /* tictactoe.cpp:19-26 */
    e->actions[e->num_moves++] = (int16_t)action;
    e->board[action] = (uint8_t)player;
    e->turn = other(player);
    return 1;
}

int mzo_env_is_terminal(const mzo_env* e)
{
    if (e->game == MZO_GAME_GO) { return go_is_terminal(e); }
    if (e->game == MZO_GAME_KILLALLGO) { /* killallgo.cpp:34-40: all of the board unconditionally Black's, or any unconditionally alive White group */
        if (benson_count(e, 1) == e->n * e->n || benson_count(e, 2) > 0) { return 1; }
        return go_is_terminal(e);
    }
    if (e->game == MZO_GAME_NOGO) { /* nogo.h:61-68 */
        for (int pos = 0; pos < e->n * e->n; ++pos) {
            if (nogo_is_legal(e, pos, e->turn)) { return 0; }
        }
        return 1;
    }
    if (e->game == MZO_GAME_OTHELLO) { return othello_is_terminal(e); }
    if (e->game == MZO_GAME_HEX) { return hex_winner(e) != 0; } /* hex.cpp:96-99 */
    if (e->game == MZO_GAME_GOMOKU) { /* gomoku.cpp:60-63 */
        if (gomoku_winner(e) != 0) { return 1; }
        for (int i = 0; i < e->n * e->n; ++i) {
            if (e->board[i] == 0) { return 0; }
        }
        return 1;
    }
    if (ttt_eval(e) != 0) { return 1; } /* tictactoe.cpp:51-55 */
    for (int i = 0; i < 9; ++i) {
        if (e->board[i] == 0) { return 0; }
    }
    return 1;
}

float mzo_env_eval_score(const mzo_env* e, int is_resign)
{
    if (e->game == MZO_GAME_GO) { return go_eval_score(e, is_resign); }
    if (e->game == MZO_GAME_KILLALLGO) { /* killallgo.cpp:42-48 (is_resign is not consulted) */
        int white = 0;
        for (int p = 0; p < e->n * e->n; ++p) { white += (e->board[p] == 2); }
        return (white == 0 || benson_count(e, 1) == e->n * e->n) ? 1.0f : -1.0f;
    }
    if (e->game == MZO_GAME_NOGO) { return other(e->turn) == 1 ? 1.0f : -1.0f; } /* nogo.h:70-78: whoever is to move has lost */
    if (e->game == MZO_GAME_OTHELLO) { return othello_eval_score(e, is_resign); }
    if (e->game == MZO_GAME_HEX) { /* hex.cpp:101-111 */
        int w = (is_resign ? other(e->turn) : hex_winner(e));
        return w == 1 ? 1.0f : (w == 2 ? -1.0f : 0.0f);
    }
    if (e->game == MZO_GAME_GOMOKU) { /* gomoku.cpp:65-73 */
        int w = (is_resign ? other(e->turn) : gomoku_winner(e));
        return w == 1 ? 1.0f : (w == 2 ? -1.0f : 0.0f);
    }
    int r = (is_resign ? other(e->turn) : ttt_eval(e)); /* tictactoe.cpp:57-65 */
    return r == 1 ? 1.0f : (r == 2 ? -1.0f : 0.0f);
}

void mzo_env_features(const mzo_env* e, int rotation, float* out)
{
    if (e->game == MZO_GAME_GO || e->game == MZO_GAME_NOGO || e->game == MZO_GAME_KILLALLGO) {
        go_features(e, rotation, out);
        return;
    }
    int rev = mzo_reversed_rotation(rotation); /* tictactoe.cpp:67-90, othello.cpp:237-255: same four planes */
    if (e->game == MZO_GAME_HEX) { rev = 0; }  /* HexEnv::getFeatures ignores the rotation (hex.cpp:123-124) */
    const int cells = e->n * e->n;
    for (int c = 0; c < 4; ++c) {
        for (int pos = 0; pos < cells; ++pos) {
            int rp = mzo_rotate_position(rev, pos, e->n);
            float v;
            if (c == 0) {
                v = (e->board[rp] == e->turn);
            } else if (c == 1) {
                v = (e->board[rp] == other(e->turn));
            } else if (c == 2) {
                v = (e->turn == 1);
            } else {
                v = (e->turn == 2);
            }
            out[c * cells + pos] = v;
        }
    }
}
