// KillAllGo's own rules on 64-bit boards (7 x 7, bit y * 8 + x), shared by the device kernels (search_core.cuh) and the host worker
// (host/worker.cpp keeps the position of every game to detect the end of a game in the reference's draw order): one source for both.
//
// Unconditional life of one colour, GoEnv::findBensonBitboard (go.cpp:614-676): blocks = groups of the colour's stones, areas = connected regions
// of everything else (go.cpp:573-589); an area is vital to a neighbouring block when all its EMPTY points are liberties of the block; blocks with
// fewer than two vital areas and areas touching a removed block are dropped until nothing changes. The reference maintains these boards
// incrementally (go.cpp:464-612); here they are recomputed from the position — pinned against 143 k positions of random playouts of the
// reference's environment, legal sets, terminal flags and results agreeing on every one (oracle/gen_env_golden.py, tests/golden/env_killallgo7).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MZ_KA_FN __host__ __device__ inline
#define MZ_KA_COLD __host__ __device__ __noinline__ // Benson's algorithm keeps ~700 bytes of block / area tables: a real call, so that the tree-step kernel's
                                                    // own frame and register allocation do not carry them for the games that never use it
#else
#define MZ_KA_FN static inline
#define MZ_KA_COLD static
#endif

#define MZ_KA_BOARD 0x007f7f7f7f7f7f7full

MZ_KA_FN int mz_ka_popc(uint32_t x)
{
    x = x - ((x >> 1) & 0x55555555u);
    x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
    return (int)((((x + (x >> 4)) & 0x0f0f0f0fu) * 0x01010101u) >> 24);
}
MZ_KA_FN uint64_t mz_ka_dilate(uint64_t x) { return (x | (x << 1) | (x >> 1) | (x << 8) | (x >> 8)) & MZ_KA_BOARD; }
MZ_KA_FN uint64_t mz_ka_flood(uint64_t seed, uint64_t mask)
{
    uint64_t f = seed & mask;
    for (;;) {
        const uint64_t g = mz_ka_dilate(f) & mask;
        if (g == f) { return f; }
        f = g;
    }
}
MZ_KA_COLD uint64_t mz_ka_benson(uint64_t own, uint64_t opp)
{
    uint64_t blk[28], area[28];
    uint32_t adj[28], vital[28];
    int nb = 0, na = 0;
    for (uint64_t o = own; o;) {
        const uint64_t b = mz_ka_flood(o & (0 - o), own);
        blk[nb++] = b, o &= ~b;
    }
    if (nb == 0) { return 0; }
    const uint64_t rest = MZ_KA_BOARD & ~own, empty = rest & ~opp;
    for (uint64_t o = rest; o;) {
        const uint64_t a = mz_ka_flood(o & (0 - o), rest);
        area[na++] = a, o &= ~a;
    }
    uint32_t in_b = 0, in_a = 0;
    for (int b = 0; b < nb; ++b) {
        const uint64_t around = mz_ka_dilate(blk[b]) & ~blk[b];
        adj[b] = vital[b] = 0;
        for (int a = 0; a < na; ++a) {
            if (!(around & area[a])) { continue; }
            adj[b] |= 1u << a;
            if (!(area[a] & empty & ~around)) { vital[b] |= 1u << a; } // every empty point of the area is a liberty of the block (go.cpp:627)
        }
        if (vital[b]) { in_b |= 1u << b, in_a |= vital[b]; } // go.cpp:630-632
    }
    for (bool changed = true; changed;) { // go.cpp:638-671
        changed = false;
        for (int b = 0; b < nb; ++b) {
            if (((in_b >> b) & 1u) && mz_ka_popc(vital[b] & in_a) < 2) { in_b &= ~(1u << b), changed = true; }
        }
        for (int b = 0; b < nb; ++b) {
            if (!((in_b >> b) & 1u) && (adj[b] & in_a)) { in_a &= ~adj[b], changed = true; } // areas with a surrounding block that is not alive
        }
    }
    uint64_t out = 0;
    for (int b = 0; b < nb; ++b) {
        if ((in_b >> b) & 1u) { out |= blk[b]; }
    }
    for (int a = 0; a < na; ++a) {
        if ((in_a >> a) & 1u) { out |= area[a]; }
    }
    return out;
}
// all of the board unconditionally Black's, or any unconditionally alive White group (killallgo.cpp:37-38)
MZ_KA_FN int mz_ka_terminal(uint64_t black, uint64_t white) { return mz_ka_benson(black, white) == MZ_KA_BOARD || mz_ka_benson(white, black) != 0; }
// KillAllGoEnv::getEvalScore (killallgo.cpp:42-48): 1 = Black wins, 2 = White wins
MZ_KA_FN int mz_ka_winner(uint64_t black, uint64_t white) { return (white == 0 || mz_ka_benson(black, white) == MZ_KA_BOARD) ? 1 : 2; }
// GoEnv::act for a stone of `own` at bit `pos` (go.cpp:150-178): opposing blocks left without a liberty are removed (suicide is illegal, so the
// stone's own block keeps one). Host side only: the device applies moves with its row-bitboard rules (mz_env_act).
MZ_KA_FN void mz_ka_place(uint64_t& own, uint64_t& opp, int pos)
{
    const uint64_t stone = 1ull << pos;
    own |= stone;
    uint64_t nb = mz_ka_dilate(stone) & opp;
    while (nb) {
        const uint64_t grp = mz_ka_flood(nb & (0 - nb), opp);
        nb &= ~grp;
        if (!(mz_ka_dilate(grp) & MZ_KA_BOARD & ~(own | opp))) { opp &= ~grp; }
    }
}
