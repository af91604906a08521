// Network kernels for sm_100a: BN-folded 3x3 convolution as an implicit GEMM on tcgen05 tensor cores
// (TMA-staged operands, fp32 accumulators in TMEM, fused bias + residual + ReLU epilogue) and the fused
// policy / value heads.
//
// Math restated (paths relative to /root/reference/minizero):
//   network/py/alphazero_network.py:90-113  stem conv-BN-ReLU, residual tower, heads, softmax
//   network/py/network_unit.py:6-23         ResidualBlock: conv3x3-BN-ReLU-conv3x3-BN-(+x)-ReLU
//   network/py/network_unit.py:26-42        PolicyNetwork: conv1x1-BN-ReLU-fc
//   network/py/network_unit.py:45-65        ValueNetwork: conv1x1-BN-ReLU-fc1-ReLU-fc2-tanh
//
// Activation layout ("shared-halo rows"): fp16 [rows][C], one row per board slot. A board of N x N cells owns
// (N+1) x (N+1) consecutive rows: slot (yy, xx) = yy * (N+1) + xx holds cell (x = xx, y = yy - 1); slots with
// yy == 0 or xx == N are zero and serve as the halo of the cells next to them — and, because boards are stored
// back to back, as the halo of the neighbouring board too. A 3x3 tap (ky, kx) of EVERY output row is then the
// input row at the constant offset (ky-1)*(N+1) + (kx-1), so the convolution is 9 accumulated GEMMs whose A
// tiles are plain 2-D TMA boxes of the same matrix shifted by a row offset (rows outside the matrix are
// zero-filled by TMA). Cost: the halo rows are computed and then zeroed, (N+1)^2 / N^2 of the useful FLOPs.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mznn {

constexpr int BM = 128;      // rows (board slots) per tile = UMMA M
constexpr int BK = 64;       // K per pipeline stage: 64 fp16 = one 128-byte swizzle row
constexpr float MZ_HALF_MAX = 65504.0f; // activations are stored as fp16: the epilogues saturate instead of producing inf
constexpr int UMMA_K = 16;   // K per tcgen05.mma for 16-bit inputs
constexpr int CONV_THREADS = 192; // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout [61,64) with SWIZZLE_128B = 2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// InstrDescriptor (same header): D = F32 [4,6) = 1, A/B = F16 (0), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n)
{
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 -> two fp16 activations: round to nearest even, saturate at +-65504 (activations are stored as fp16: the epilogues saturate
// instead of producing inf), ReLU if asked for — one F2FP.SATFINITE(.RELU).F16.F32.PACK_AB. x0 goes to the low half.
__device__ __forceinline__ uint32_t pack_act_f16x2(float x0, float x1, int relu)
{
    uint32_t r;
    if (relu) {
        asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    } else {
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    }
    return r;
}

struct ConvParams {
    __half* out;            // [rows_alloc][cout]
    const __half* residual; // [rows_alloc][cout] or null
    const float* bias;      // [cout] BN-folded
    int rows_valid;         // B * slots
    int n1;                 // N + 1: slots per board row
    int slots;              // (N + 1)^2
    int cin;                // multiple of BK
    int cout;               // multiple of BN
    int relu;
    int krot;               // rotate each CTA's K-loop start so that CTAs do not stream the same weight tile at the same time
};

template <int BN, int STAGES>
struct ConvSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024; // + alignment slack
};

// out[r][n0 + j] = act( bias[n0 + j] + sum_{tap, ci} in[r + off(tap)][ci] * w[tap][n0 + j][ci] (+ residual[r][n0 + j]) ),
// halo rows forced to zero. grid = (row tiles, cout / BN).
template <int BN, int STAGES>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_tcgen05_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w, const ConvParams p)
{
    using L = ConvSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kb_per_tap = p.cin / BK, num_k = 9 * kb_per_tap;
    const int rot = (p.krot ? static_cast<int>((blockIdx.x * 7u + blockIdx.y * 3u) % static_cast<unsigned>(num_k)) : 0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_in)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) { // TMEM accumulator: BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) { // ===== TMA producer =====
            for (int k = 0; k < num_k; ++k) {
                const int s = k % STAGES;
                if (k >= STAGES) { mbar_wait(&empty_bar[s], ((k / STAGES) - 1) & 1); }
                const int kr = (k + rot >= num_k ? k + rot - num_k : k + rot);
                const int tap = kr / kb_per_tap, kb = kr - tap * kb_per_tap;
                const int off = (tap / 3 - 1) * p.n1 + (tap % 3 - 1);
                uint8_t* a_dst = smem + s * L::STAGE_BYTES;
                uint8_t* b_dst = a_dst + L::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
                tma_load_2d(a_dst, &map_in, &full_bar[s], kb * BK, m0 + off);
                tma_load_2d(b_dst, &map_w, &full_bar[s], kb * BK, tap * p.cout + n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { // ===== MMA issuer =====
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
            for (int k = 0; k < num_k; ++k) {
                const int s = k % STAGES;
                mbar_wait(&full_bar[s], (k / STAGES) & 1);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES), b_addr = a_addr + L::A_BYTES;
#pragma unroll
                for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                    umma_f16(tmem_base, umma_desc_sw128(a_addr + kk * UMMA_K * 2), umma_desc_sw128(b_addr + kk * UMMA_K * 2), idesc, (k | kk) != 0);
                }
                tcgen05_commit(&empty_bar[s]); // frees the stage once these MMAs have read it
            }
            tcgen05_commit(accum_bar); // accumulator complete
        }
    } else { // ===== epilogue: TMEM -> registers -> bias (+ residual) (ReLU) -> fp16 rows =====
        const int quarter = warp & 3; // TMEM lanes [32 * (warp % 4), +32) are the only ones this warp may read
        const int r = m0 + quarter * 32 + lane;
        const int rr = r % p.slots;
        const bool live = (r < p.rows_valid) && (rr / p.n1 != 0) && (rr % p.n1 != p.n1 - 1);
        mbar_wait(accum_bar, 0);
        tcgen05_fence_after();
        __half* out_row = p.out + static_cast<size_t>(r) * p.cout + n0;
        const __half* res_row = (p.residual ? p.residual + static_cast<size_t>(r) * p.cout + n0 : nullptr);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c, v);
            tmem_ld_wait();
            uint4 packed[4];
            uint32_t* pk = reinterpret_cast<uint32_t*>(packed);
            uint4 res[4];
            if (res_row && live) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { res[q] = *reinterpret_cast<const uint4*>(res_row + c + q * 8); }
            }
            const __half2* rh = reinterpret_cast<const __half2*>(res);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x0 = __uint_as_float(v[2 * j]) + __ldg(p.bias + n0 + c + 2 * j);
                float x1 = __uint_as_float(v[2 * j + 1]) + __ldg(p.bias + n0 + c + 2 * j + 1);
                if (res_row && live) {
                    const float2 rf = __half22float2(rh[j]);
                    x0 += rf.x, x1 += rf.y;
                }
                if (!live) { x0 = 0.0f, x1 = 0.0f; }
                pk[j] = pack_act_f16x2(x0, x1, p.relu); // ReLU + saturation at the fp16 range + rounding in one F2FP
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { *reinterpret_cast<uint4*>(out_row + c + q * 8) = packed[q]; }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Resident-A variant (the default): the 9 taps of one row tile read the SAME input rows shifted by at most
// +-(N+2), so the tile's input block (128 + 2*(N+2) rows, all input channels) is brought into shared memory ONCE
// and every tap's A operand is a descriptor into that block at a different start row; only the weight tiles
// stream through the TMA ring. L2 -> SM traffic per tile drops from 9 x (A + B) to A + 9 x B, and a persistent
// CTA that handles both output-channel halves of a row tile reuses the block again. Accumulators are double
// buffered in TMEM so the epilogue of one unit overlaps the MMAs of the next.
// ---------------------------------------------------------------------------------------------
struct ConvResParams {
    ConvParams c;
    int rows_ext;   // rows of the resident block: 128 + 2 * halo, rounded up to 8
    int halo;       // N + 2
    int num_mtiles; // row tiles
    int base_off_mode; // 1: descriptor base-offset field = (start address >> 7) & 7 for row-shifted starts
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// spin on an mbarrier phase with the minimum of instructions (labels are local to the PTX block)
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar_addr, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}" ::"r"(bar_addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xFFFFFFFF;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tcgen05_commit_u32(uint32_t bar_addr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// The producer and the MMA issuer are single threads: every instruction in their per-K-block loops is on the
// critical path (measured: ~650 cycles of address arithmetic per K block starve a 256-cycle MMA), so the loops keep
// stage / phase / descriptor words incrementally and contain no division.
// CL > 1: thread-block clusters of CL CTAs work on CL different row tiles with the SAME sequence of weight tiles; each CTA
// fetches 1/CL of every weight tile and TMA-multicasts it into the shared memory of all CTAs of the cluster, so the
// L2 -> SM weight traffic (the dominant stream once the input block is resident) drops by CL. A stage may be overwritten
// only after every CTA of the cluster has consumed it: the MMA warps commit their "stage free" arrival to all CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc_u32(uint32_t bar_addr, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr), "h"(mask) : "memory");
}

template <int BN, int STAGES, int CL>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_resident_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w, const ConvResParams rp)
{
    const ConvParams& p = rp.c;
    constexpr int B_BYTES = BN * BK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int a_kb = p.cin / BK;
    const int a_kb_bytes = rp.rows_ext * 128;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + a_kb * a_kb_bytes;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_BYTES);
    uint64_t* b_empty = b_full + STAGES;
    uint64_t* a_full = b_empty + STAGES;
    uint64_t* a_empty = a_full + 1;
    uint64_t* acc_full = a_empty + 1;   // [2]
    uint64_t* acc_empty = acc_full + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = p.cout / BN;
    // work units: (group of CL consecutive row tiles, output-channel half); contiguous ranges of units per cluster;
    // CTA `crank` of the cluster takes row tile group * CL + crank
    const int crank = (CL > 1 ? static_cast<int>(cluster_ctarank()) : 0);
    const int cid = blockIdx.x / CL, num_clusters = gridDim.x / CL;
    const int units = ((rp.num_mtiles + CL - 1) / CL) * nh;
    const int u_begin = static_cast<int>((static_cast<long long>(cid) * units) / num_clusters);
    const int u_end = static_cast<int>((static_cast<long long>(cid + 1) * units) / num_clusters);
    constexpr uint16_t mc_mask = static_cast<uint16_t>((1u << CL) - 1u);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_in)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], CL); // one "stage consumed" arrival from the MMA warp of every CTA of the cluster
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4); // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CL > 1) { cluster_sync_all(); } // every CTA's barriers exist before any peer signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t full0 = smem_u32(b_full), empty0 = smem_u32(b_empty);

    if (warp == 0) {
        { // ===== TMA producer: the whole warp runs the loop (uniform control flow), one elected lane issues =====
            int s = 0, grp = u_begin / nh, half = u_begin - grp * nh, cur_mt = -1;
            int mt = grp * CL + crank;
            uint32_t ph = 1, a_ph = 1; // "empty" barriers: the first pass over the ring must not block
            const uint32_t b_dst0 = smem_u32(smem_b), a_dst0 = smem_u32(smem_a);
            const uint64_t map_w_ptr = reinterpret_cast<uint64_t>(&map_w), map_in_ptr = reinterpret_cast<uint64_t>(&map_in);
            for (int u = u_begin; u < u_end; ++u) {
                if (mt != cur_mt) {
                    mbar_wait_u32(smem_u32(a_empty), a_ph); // every MMA on the previous input block has completed
                    a_ph ^= 1;
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(a_full, a_kb * a_kb_bytes);
                        for (int kb = 0; kb < a_kb; ++kb) {
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(a_dst0 + kb * a_kb_bytes),
                                         "l"(map_in_ptr), "r"(smem_u32(a_full)), "r"(kb * BK), "r"(mt * BM - rp.halo)
                                         : "memory");
                        }
                    }
                    __syncwarp();
                    cur_mt = mt;
                }
                int wrow = half * BN; // row of the weight matrix: tap * cout + half * BN
                for (int tap = 0; tap < 9; ++tap, wrow += p.cout) {
                    for (int kc = 0; kc < p.cin; kc += BK) {
                        const uint32_t full = full0 + s * 8;
                        mbar_wait_u32(empty0 + s * 8, ph);
                        if (elect_one_sync()) {
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(B_BYTES) : "memory");
                            if constexpr (CL == 1) {
                                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(b_dst0 + s * B_BYTES),
                                             "l"(map_w_ptr), "r"(full), "r"(kc), "r"(wrow)
                                             : "memory");
                            } else { // this CTA's 1/CL slice of the tile, delivered to every CTA of the cluster
                                asm volatile(
                                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                                        b_dst0 + s * B_BYTES + crank * (B_BYTES / CL)),
                                    "l"(map_w_ptr), "r"(full), "r"(kc), "r"(wrow + crank * (BN / CL)), "h"(mc_mask)
                                    : "memory");
                            }
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0, ph ^= 1; }
                    }
                }
                if (++half == nh) { half = 0, mt += CL; }
            }
        }
    } else if (warp == 1) {
        { // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
            // descriptor words (cute/arch/mma_sm100_desc.hpp): lo = start>>4 | LBO(1)<<16 ; hi = SBO(1024>>4) | version 1<<14 | SWIZZLE_128B 2<<29
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t a_kb_step = static_cast<uint32_t>(a_kb_bytes) >> 4;
            int s = 0, grp = u_begin / nh, half = u_begin - grp * nh, cur_mt = -1, buf = 0;
            int mt = grp * CL + crank;
            uint32_t ph = 0, a_ph = 0, acc_ph0 = 1, acc_ph1 = 1;
            for (int u = u_begin; u < u_end; ++u) {
                if (mt != cur_mt) {
                    mbar_wait_u32(smem_u32(a_full), a_ph);
                    a_ph ^= 1;
                    cur_mt = mt;
                }
                if (buf == 0) {
                    mbar_wait_u32(smem_u32(&acc_empty[0]), acc_ph0);
                    acc_ph0 ^= 1;
                } else {
                    mbar_wait_u32(smem_u32(&acc_empty[1]), acc_ph1);
                    acc_ph1 ^= 1;
                }
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                uint32_t accumulate = 0;
                int row0 = rp.halo - p.n1 - 1; // first tap: (ky, kx) = (0, 0) -> offset -(N+1) - 1
                for (int ty = 0; ty < 3; ++ty, row0 += p.n1 - 3) {
                    for (int tx = 0; tx < 3; ++tx, ++row0) {
                        uint32_t a_lo = a_lo0 + static_cast<uint32_t>(row0) * 8u; // + row0 * 128 bytes
                        for (int kc = 0; kc < p.cin; kc += BK, a_lo += a_kb_step) {
                            mbar_wait_u32(full0 + s * 8, ph);
                            tcgen05_fence_after();
                            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(s) * (B_BYTES >> 4);
                            if (elect_one_sync()) {
                                umma_f16_lohi(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate);
                                umma_f16_lohi(tmem_d, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
                                umma_f16_lohi(tmem_d, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
                                umma_f16_lohi(tmem_d, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
                                if constexpr (CL == 1) {
                                    tcgen05_commit_u32(empty0 + s * 8);
                                } else {
                                    tcgen05_commit_mc_u32(empty0 + s * 8, mc_mask);
                                }
                            }
                            __syncwarp();
                            accumulate = 1u;
                            if (++s == STAGES) { s = 0, ph ^= 1; }
                        }
                    }
                }
                if (++half == nh) { half = 0, mt += CL; }
                if (elect_one_sync()) {
                    tcgen05_commit_u32(smem_u32(&acc_full[buf]));
                    if (mt != cur_mt || u + 1 == u_end) { tcgen05_commit_u32(smem_u32(a_empty)); } // input block no longer needed
                }
                __syncwarp();
                buf ^= 1;
            }
        }
    } else { // ===== epilogue =====
        const int quarter = warp & 3;
        int ucount = 0;
        for (int u = u_begin; u < u_end; ++u, ++ucount) {
            const int grp = u / nh, half = u - grp * nh, buf = ucount & 1;
            const int mt = grp * CL + crank;
            const int n0 = half * BN;
            const int r = mt * BM + quarter * 32 + lane;
            const int rr = r % p.slots;
            const bool live = (r < p.rows_valid) && (rr / p.n1 != 0) && (rr % p.n1 != p.n1 - 1);
            const bool in_range = (mt < rp.num_mtiles); // padding tile of an incomplete group: computed (lockstep), not stored
            mbar_wait(&acc_full[buf], (ucount >> 1) & 1);
            tcgen05_fence_after();
            __half* out_row = p.out + static_cast<size_t>(r) * p.cout + n0;
            const __half* res_row = (p.residual ? p.residual + static_cast<size_t>(r) * p.cout + n0 : nullptr);
#pragma unroll 1
            for (int c = 0; c < BN && in_range; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + c, v);
                uint4 res[4];
                if (res_row && live) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { res[q] = *reinterpret_cast<const uint4*>(res_row + c + q * 8); }
                }
                float4 bias4[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { bias4[q] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c) + q); }
                const float* bias = reinterpret_cast<const float*>(bias4);
                tmem_ld_wait();
                uint4 packed[4];
                uint32_t* pk = reinterpret_cast<uint32_t*>(packed);
                const __half2* rh = reinterpret_cast<const __half2*>(res);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x0 = __uint_as_float(v[2 * j]) + bias[2 * j];
                    float x1 = __uint_as_float(v[2 * j + 1]) + bias[2 * j + 1];
                    if (res_row && live) {
                        const float2 rf = __half22float2(rh[j]);
                        x0 += rf.x, x1 += rf.y;
                    }
                    if (!live) { x0 = 0.0f, x1 = 0.0f; }
                    pk[j] = pack_act_f16x2(x0, x1, p.relu); // ReLU + saturation at the fp16 range + rounding in one F2FP
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { *reinterpret_cast<uint4*>(out_row + c + q * 8) = packed[q]; }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&acc_empty[buf]); }
        }
    }
    __syncthreads();
    if constexpr (CL > 1) { cluster_sync_all(); } // no CTA leaves while a peer may still multicast into it or signal its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster issue ONE M=256 MMA over their two row tiles. Each CTA
// keeps its own resident input block and only HALF of every weight tile (the tensor cores of both SMs read both halves),
// so per SM the shared-memory traffic per MMA drops from A + B (at N = 128 exactly the 128 B/clk the SM can deliver,
// before the TMA writes are added) to A + B/2, and the weight bytes arriving per SM halve. The leader CTA (cluster rank 0)
// issues all MMAs; both CTAs run a producer warp (their loads signal the LEADER's barriers) and epilogue warps (each
// drains its own TMEM half); stage / block / accumulator releases are committed to both CTAs at once.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_f16_lohi_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm_u32(uint32_t bar_addr)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, uint64_t map_ptr, uint32_t leader_bar, int32_t c0, int32_t c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst), "l"(map_ptr),
                 "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar_addr, uint32_t cta_rank)
{
    asm volatile(
        "{\n\t.reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t}" ::"r"(local_bar_addr),
        "r"(cta_rank)
        : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w_half, const ConvResParams rp)
{
    const ConvParams& p = rp.c;
    constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
    constexpr uint32_t kPeerMask = 0xFEFFFFFFu; // cute/arch/copy_sm100_tma.hpp Sm100MmaPeerBitMask: address the leader CTA's barrier
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int a_kb = p.cin / BK;
    const int a_kb_bytes = rp.rows_ext * 128;
    const int a_bytes = a_kb * a_kb_bytes; // one input block; two of them: the next row-tile group is prefetched during the current one
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + 2 * a_bytes;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_HALF_BYTES);
    uint64_t* b_empty = b_full + STAGES;
    uint64_t* a_full = b_empty + STAGES; // [2]
    uint64_t* a_empty = a_full + 2;      // [2]
    uint64_t* acc_full = a_empty + 2;    // [2]
    uint64_t* acc_empty = acc_full + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    // programmatic dependent launch: the next layer's CTAs may be scheduled (and run their set-up) as soon as SMs free up;
    // they block in griddepcontrol.wait below until this whole grid has completed and its writes are visible
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = p.cout / BN;
    const int crank = static_cast<int>(cluster_ctarank());
    const bool leader = (crank == 0);
    const int cid = blockIdx.x / 2, num_clusters = gridDim.x / 2;
    const int units = ((rp.num_mtiles + 1) / 2) * nh;
    const int u_begin = static_cast<int>((static_cast<long long>(cid) * units) / num_clusters);
    const int u_end = static_cast<int>((static_cast<long long>(cid + 1) * units) / num_clusters);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_in)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_half)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8); // the 4 epilogue warps of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) { // the same warp of both CTAs allocates the pair's TMEM columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t full0 = smem_u32(b_full), empty0 = smem_u32(b_empty);
    const int g_first = u_begin / nh, g_last = (u_end - 1) / nh; // row-tile groups this cluster touches
    // everything above touched no global data of the previous layer (with programmatic dependent launch — MZ_CONV_PDL=1,
    // measured no faster than plain stream order on this path — the set-up overlaps the previous grid's tail)
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        { // ===== TMA producer (both CTAs): own input blocks, own half of every weight tile; completion goes to the leader =====
            int s = 0, grp = g_first, half = u_begin - g_first * nh;
            uint32_t ph = 1;
            const uint32_t b_dst0 = smem_u32(smem_b), a_dst0 = smem_u32(smem_a);
            const uint64_t map_w_ptr = reinterpret_cast<uint64_t>(&map_w_half), map_in_ptr = reinterpret_cast<uint64_t>(&map_in);
            auto load_block = [&](int g) { // input block of group g into buffer (g - g_first) & 1
                const int gi = g - g_first, buf = gi & 1;
                mbar_wait_u32(smem_u32(&a_empty[buf]), ((gi >> 1) & 1) ^ 1); // the MMAs that read this buffer two groups ago are done
                if (elect_one_sync()) {
                    if (leader) { mbar_arrive_expect_tx(&a_full[buf], 2 * a_bytes); }
                    const uint32_t bar = smem_u32(&a_full[buf]) & kPeerMask;
                    for (int kb = 0; kb < a_kb; ++kb) { tma_load_2d_2sm(a_dst0 + buf * a_bytes + kb * a_kb_bytes, map_in_ptr, bar, kb * BK, (g * 2 + crank) * BM - rp.halo); }
                }
                __syncwarp();
            };
            if (u_begin < u_end) { load_block(g_first); }
            int prefetched = g_first;
            for (int u = u_begin; u < u_end; ++u) {
                int wrow = half * BN + crank * (BN / 2);
                for (int tap = 0; tap < 9; ++tap, wrow += p.cout) {
                    for (int kc = 0; kc < p.cin; kc += BK) {
                        mbar_wait_u32(empty0 + s * 8, ph);
                        if (elect_one_sync()) {
                            if (leader) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + s * 8), "r"(2 * B_HALF_BYTES) : "memory"); }
                            tma_load_2d_2sm(b_dst0 + s * B_HALF_BYTES, map_w_ptr, (full0 + s * 8) & kPeerMask, kc, wrow);
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0, ph ^= 1; }
                    }
                }
                // the weights of this unit are on their way: now (the MMAs are at most STAGES K-blocks behind, far past the
                // previous group) fetch the NEXT group's input block into the other buffer
                if (prefetched == grp && grp < g_last) {
                    load_block(grp + 1);
                    prefetched = grp + 1;
                }
                if (++half == nh) { half = 0, ++grp; }
            }
        }
    } else if (warp == 1) {
        if (leader) { // ===== MMA issuer (leader CTA only): M = 256 over both CTAs' row tiles =====
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t a_kb_step = static_cast<uint32_t>(a_kb_bytes) >> 4, a_buf_step = static_cast<uint32_t>(a_bytes) >> 4;
            int s = 0, grp = g_first, half = u_begin - g_first * nh, cur_grp = -1, buf = 0;
            uint32_t ph = 0, acc_ph0 = 1, acc_ph1 = 1;
            for (int u = u_begin; u < u_end; ++u) {
                const int gi = grp - g_first, abuf = gi & 1;
                if (grp != cur_grp) {
                    mbar_wait_u32(smem_u32(&a_full[abuf]), (gi >> 1) & 1);
                    cur_grp = grp;
                }
                if (buf == 0) {
                    mbar_wait_u32(smem_u32(&acc_empty[0]), acc_ph0);
                    acc_ph0 ^= 1;
                } else {
                    mbar_wait_u32(smem_u32(&acc_empty[1]), acc_ph1);
                    acc_ph1 ^= 1;
                }
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                uint32_t accumulate = 0;
                int row0 = rp.halo - p.n1 - 1;
                for (int ty = 0; ty < 3; ++ty, row0 += p.n1 - 3) {
                    for (int tx = 0; tx < 3; ++tx, ++row0) {
                        uint32_t a_lo = a_lo0 + static_cast<uint32_t>(abuf) * a_buf_step + static_cast<uint32_t>(row0) * 8u;
                        for (int kc = 0; kc < p.cin; kc += BK, a_lo += a_kb_step) {
                            mbar_wait_u32(full0 + s * 8, ph);
                            tcgen05_fence_after();
                            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(s) * (B_HALF_BYTES >> 4);
                            if (elect_one_sync()) {
                                umma_f16_lohi_2sm(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate);
                                umma_f16_lohi_2sm(tmem_d, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
                                umma_f16_lohi_2sm(tmem_d, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
                                umma_f16_lohi_2sm(tmem_d, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
                                tcgen05_commit_2sm_u32(empty0 + s * 8); // frees the stage in BOTH CTAs
                            }
                            __syncwarp();
                            accumulate = 1u;
                            if (++s == STAGES) { s = 0, ph ^= 1; }
                        }
                    }
                }
                if (++half == nh) { half = 0, ++grp; }
                if (elect_one_sync()) {
                    tcgen05_commit_2sm_u32(smem_u32(&acc_full[buf]));
                    if (grp != cur_grp || u + 1 == u_end) { tcgen05_commit_2sm_u32(smem_u32(&a_empty[abuf])); } // this input block is free again
                }
                __syncwarp();
                buf ^= 1;
            }
        }
    } else { // ===== epilogue (both CTAs): own 128 rows of the pair's accumulator =====
        const int quarter = warp & 3;
        int ucount = 0;
        for (int u = u_begin; u < u_end; ++u, ++ucount) {
            const int grp = u / nh, half = u - grp * nh, buf = ucount & 1;
            const int mt = grp * 2 + crank;
            const int n0 = half * BN;
            const int r = mt * BM + quarter * 32 + lane;
            const int rr = r % p.slots;
            const bool live = (r < p.rows_valid) && (rr / p.n1 != 0) && (rr % p.n1 != p.n1 - 1);
            const bool in_range = (mt < rp.num_mtiles);
            mbar_wait(&acc_full[buf], (ucount >> 1) & 1);
            tcgen05_fence_after();
            __half* out_row = p.out + static_cast<size_t>(r) * p.cout + n0;
            const __half* res_row = (p.residual ? p.residual + static_cast<size_t>(r) * p.cout + n0 : nullptr);
#pragma unroll 1
            for (int c = 0; c < BN && in_range; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + c, v);
                uint4 res[4];
                if (res_row && live) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { res[q] = *reinterpret_cast<const uint4*>(res_row + c + q * 8); }
                }
                float4 bias4[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) { bias4[q] = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c) + q); }
                const float* bias = reinterpret_cast<const float*>(bias4);
                tmem_ld_wait();
                uint4 packed[4];
                uint32_t* pk = reinterpret_cast<uint32_t*>(packed);
                const __half2* rh = reinterpret_cast<const __half2*>(res);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x0 = __uint_as_float(v[2 * j]) + bias[2 * j];
                    float x1 = __uint_as_float(v[2 * j + 1]) + bias[2 * j + 1];
                    if (res_row && live) {
                        const float2 rf = __half22float2(rh[j]);
                        x0 += rf.x, x1 += rf.y;
                    }
                    if (!live) { x0 = 0.0f, x1 = 0.0f; }
                    pk[j] = pack_act_f16x2(x0, x1, p.relu); // ReLU + saturation at the fp16 range + rounding in one F2FP
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { *reinterpret_cast<uint4*>(out_row + c + q * 8) = packed[q]; }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u); } // the leader's barrier counts both CTAs
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Whole-tower variant: ALL 3x3 conv layers of the network (stem + 2 per residual block) in ONE persistent launch of CTA
// pairs. Per launch the separate-kernel version pays launch gap + barrier / TMEM set-up + pipeline fill + a tail where the
// CTAs with fewer units idle (measured: the tensor pipe is busy 45 % of a layer's wall time); here the set-up happens once,
// the per-layer unit ranges rotate over the clusters so that nobody is systematically short of work, and a cluster moves on
// to the next layer as soon as ITS inputs are complete: a unit (layer l, row-tile group g) only needs groups g-1, g, g+1 of
// layer l-1 (the 3x3 halo), which it learns from per-(layer, group) completion counters in global memory
// (release: the epilogue's TMA stores -> cp.async.bulk.wait_group 0 -> one red.release.gpu per warp [lane-per-row epilogue of
// the many-unit Atari stages: st.global -> __threadfence -> atomicAdd]; acquire: producer polls with ld.acquire.gpu, then a
// generic->async proxy fence before the TMA loads). All CTAs are co-resident (one per SM), and every wait points at a
// strictly earlier layer, so the waits cannot cycle.
// ---------------------------------------------------------------------------------------------
constexpr int TOWER_MAX_LAYERS = 48;
constexpr int TOWER_THREADS = 224; // warp 0: weight TMA, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue, warp 6: input-block TMA + dependency waits

struct alignas(64) TowerLayer {
    CUtensorMap map_in;     // input activations, box = resident block
    CUtensorMap map_w;      // weights [9 * cout][cin], box = half tile
    __half* out;
    const __half* residual; // or null
    const float* bias;
    int cin;
    int relu;
    int cin_off;  // first input channel of this layer within the rows of map_in (a layer may read a channel slice: the four
                  // sub-position layers of a stride-2 convolution share one space-to-depth input)
    int tap_mask; // bit t set: tap t = (ky * 3 + kx) takes part (0x1ff = a full 3x3; a stride-2 3x3 convolution over a
                  // space-to-depth input only has taps that reach up / left: see engine.cu, "stride 2")
    int out_map;  // wide tower: which of TowerParams::map_out covers `out` (its epilogue stores through TMA)
    int res_layer; // wide tower: the layer whose output `residual` is (its completion counters let the epilogue fetch the residual
                   // rows before the accumulators are ready), or -1 when `residual` was written before the launch
};

struct TowerParams {
    TowerLayer layer[TOWER_MAX_LAYERS];
    int num_layers;
    int rows_valid, n1, slots, cout, rows_ext, halo, num_mtiles;
    int cin_max;  // widest layer input (the dynamics stem of a MuZero network reads hidden + action planes): sizes the input-block buffers
    int rotate;   // cluster offset per layer for the unit ranges
    int shift;    // unit offset per layer: position q of layer l is unit (q + l * shift) mod units
    int zigzag;   // odd layers process the cluster's range in reverse order
    int strided;  // units dealt round-robin to the clusters instead of in contiguous ranges
    int tap_rot;  // 1: a unit starts its tap loop at tap (row group % 9) instead of tap 0, so that the CTA pairs working on the same (layer,
                  // channel half) do not all ask the L2 for the same weight tile at the same moment. The order is a function of the row
                  // group alone, never of the grid: a position's result does not depend on the batch it is evaluated in
    int fence_mode; // lane-per-row epilogue (epi_bufs == 0) only: how a warp publishes its rows: 0 = __threadfence by every lane, then one atomicAdd; 1 = __syncwarp, then one red.release.gpu
    int pdl;      // launched with programmatic stream serialization: the grid may start while the tree step before it is still running;
                  // only the first layer's input rows depend on that kernel, and the input producer waits for it (griddepcontrol.wait)
    int* done;    // [num_layers][num_groups] completion counters, zeroed before every launch
    unsigned long long* dbg; // optional [grid][8] cycle counters (profiling)
    int epi_bufs;   // narrow tower: staging tiles per epilogue warp (1 or 2; 0 = lane-per-row stores), see tower_smem_bytes
    CUtensorMap map_out[3]; // wide tower: the three activation buffers as TMA store targets, box = 32 channels x 32 rows (one TMEM load of an
                            // epilogue warp), 64-byte swizzle
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Wait until a completion counter of the previous layer reaches `need`. The tower relies on every CTA of the grid being resident (it is launched
// cooperatively, so the driver guarantees that); should a counter nevertheless stay short for 10 seconds of wall time — a launch that lost its
// co-residency, counters that were not cleared — the kernel traps: the stream reports an error instead of spinning for ever.
__device__ __forceinline__ void wait_counter(const int* p, int need)
{
    unsigned spins = 0;
    unsigned long long t_first = 0;
    while (ld_acquire_gpu(p) < need) { // every poll is an L2 round trip: no back-off needed between them
        if ((++spins & 0x3FFFFu) == 0u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t_first == 0) {
                t_first = now;
            } else if (now - t_first > 10000000000ull) {
                __trap();
            }
        }
    }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int32_t c0, int32_t c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
// explicit shared-space accesses (through a generic pointer the compiler emits LD / ST, which wait on the long scoreboard like a global access)
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4& v) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f2(uint32_t addr, const float2& v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } // every store has read its source tile
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); } // all but the newest store have read their source tiles
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }       // the stores are complete (visible)
// shared memory of a narrow tower CTA (the 1024 bytes at the end pay for aligning the base). epi_bufs: staging tiles per epilogue warp (32 rows x 32 fp16
// channels, 64-byte swizzle, 2 KB each): 2 where they fit, 1 where the input blocks leave no room (Atari dynamics: 320 input channels), 0 = the
// lane-per-row st.global epilogue (towers with many units per CTA pair and layer: nothing waits for a single unit's rows, and with 128 input
// channels a unit's MMAs are shorter than four staged chunks + a release — measured on the 48 x 48 Atari stage: 517 us against 603 us)
__host__ __device__ constexpr size_t tower_smem_bytes(int cin_max, int rows_ext, int stages, int bn, int epi_bufs)
{
    return static_cast<size_t>(bn == 128 && epi_bufs > 0 ? 4 * epi_bufs * 2048 + 4 * bn * 4 : 0) + 2 * static_cast<size_t>(cin_max / BK) * rows_ext * 128 +
           static_cast<size_t>(stages) * (bn / 2) * BK * 2 + 24 * 8 + 16 + 1024;
}

__device__ __forceinline__ void sts_f4(uint32_t addr, const float4& v) { asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

template <int BN, int STAGES, bool DBG>
__global__ void __launch_bounds__(TOWER_THREADS, 1)
conv_tower_kernel(const __grid_constant__ TowerParams tp)
{
    constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
    constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int a_kb_bytes = tp.rows_ext * 128;
    const int a_bytes_max = (tp.cin_max / BK) * a_kb_bytes; // hidden layers have cin == cout; an AlphaZero stem is narrower, a MuZero dynamics stem wider
    const int epi_bytes = (BN == 128 && tp.epi_bufs > 0 ? 4 * tp.epi_bufs * 2048 + 4 * BN * 4 : 0); // staging tiles + per-warp bias rows of the epilogue
    uint8_t* smem_epi = smem;
    float* smem_bias = reinterpret_cast<float*>(smem + 4 * tp.epi_bufs * 2048);
    uint8_t* smem_a = smem + epi_bytes;
    uint8_t* smem_b = smem_a + 2 * a_bytes_max;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_HALF_BYTES);
    uint64_t* b_empty = b_full + STAGES;
    uint64_t* a_full = b_empty + STAGES; // [2]
    uint64_t* a_empty = a_full + 2;      // [2]
    uint64_t* acc_full = a_empty + 2;    // [2]
    uint64_t* acc_empty = acc_full + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = tp.cout / BN;
    const int crank = static_cast<int>(cluster_ctarank());
    const bool leader = (crank == 0);
    const int cid = blockIdx.x / 2, nc = gridDim.x / 2;
    const int num_groups = (tp.num_mtiles + 1) / 2;
    const int units = num_groups * nh;
    const int need = 8 * nh; // arrivals per (layer, group): 4 epilogue warps x 2 CTAs x nh halves

    if (warp == 0 && lane == 0) {
        for (int l = 0; l < tp.num_layers; ++l) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tp.layer[l].map_in)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tp.layer[l].map_w)) : "memory");
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t full0 = smem_u32(b_full), empty0 = smem_u32(b_empty);

    // unit range of this cluster in layer l (rotated so that the uneven split does not always hit the same clusters)
    // (ub .. ue are positions in the layer's unit order; position q is unit (q + l * shift) mod units, so the range borders
    //  slide from layer to layer and a range that is long in one layer leans on shorter ones in the next)
    // strided ownership (tp.strided): cluster c owns units c, c + nc, c + 2 nc, ... of every layer. The six units a unit's
    // halo needs (groups g-1 .. g+1, both halves) then belong to six neighbouring clusters in the SAME pass of the previous
    // layer, finished a whole layer-time ago — the epilogue -> counter -> TMA round trip never sits on the critical path.
    // (Cost: the two halves of a group go to different clusters, so every unit loads its own input block.)
    auto range = [&](int l, int& ub, int& ue) {
        if (tp.strided) {
            ub = 0;
            ue = (units - (cid + l * tp.rotate) % nc + nc - 1) / nc;
            return;
        }
        const int cl = (cid + l * tp.rotate) % nc;
        ub = static_cast<int>((static_cast<long long>(cl) * units) / nc);
        ue = static_cast<int>((static_cast<long long>(cl + 1) * units) / nc);
    };
    // zigzag: odd layers walk the cluster's range backwards. With fixed ownership (rotate == 0, shift == 0) a cluster then
    // starts every layer with the group it finished last, and the neighbours' edge groups it needs were the FIRST ones they
    // computed in the previous layer: the halo dependency stops acting as a per-layer barrier.
    auto unit_of = [&](int l, int q) {
        if (tp.strided) { return (cid + l * tp.rotate) % nc + q * nc; } // ownership rotates so that the clusters take turns at the short lists
        if (tp.zigzag && (l & 1)) {
            int ub, ue;
            range(l, ub, ue);
            q = ub + ue - 1 - q;
        }
        return (q + l * tp.shift) % units;
    };

    if (warp == 0) {
        // ===== weight producer (both CTAs): streams this CTA's half of every weight tile, never waits for anything but a free stage =====
        int s = 0;
        uint32_t ph = 1;
        long long t_bempty = 0;
        const long long t_start = (DBG ? clock64() : 0ll);
        const uint32_t b_dst0 = smem_u32(smem_b);
        for (int l = 0; l < tp.num_layers; ++l) {
            const TowerLayer& L = tp.layer[l];
            const uint64_t map_w_ptr = reinterpret_cast<uint64_t>(&L.map_w);
            int ub, ue;
            range(l, ub, ue);
            for (int q = ub; q < ue; ++q) {
                const int u = unit_of(l, q);
                const int half = u % nh;
                const int wrow0 = half * BN + crank * (BN / 2), tap0 = (tp.tap_rot ? (u / nh) % 9 : 0);
                // K-block outer, tap inner: the same accumulation order as conv_tower_wide_kernel, so that a position's result does not
                // depend on which of the two kernels (i.e. which batch size) evaluates it
                for (int kc = 0; kc < L.cin; kc += BK) {
                    for (int t9 = 0; t9 < 9; ++t9) {
                        const int tap = (t9 + tap0 >= 9 ? t9 + tap0 - 9 : t9 + tap0), wrow = wrow0 + tap * tp.cout;
                        if (!((L.tap_mask >> tap) & 1)) { continue; }
                        const long long te = (DBG ? clock64() : 0ll);
                        mbar_wait_u32(empty0 + s * 8, ph);
                        t_bempty += (DBG ? clock64() : 0ll) - te;
                        if (elect_one_sync()) {
                            if (leader) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + s * 8), "r"(2 * B_HALF_BYTES) : "memory"); }
                            tma_load_2d_2sm(b_dst0 + s * B_HALF_BYTES, map_w_ptr, (full0 + s * 8) & kPeerMask, kc, wrow);
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0, ph ^= 1; }
                    }
                }
            }
        }
        if (DBG && tp.dbg && lane == 0) {
            tp.dbg[blockIdx.x * 8 + 0] = (DBG ? clock64() : 0ll) - t_start;
            tp.dbg[blockIdx.x * 8 + 2] = t_bempty;
        }
    } else if (warp == 6) {
        // ===== input-block producer (both CTAs): one block per group of consecutive units, as far ahead as the two buffers allow;
        //       it alone waits for the previous layer's completion counters =====
        int gcount = 0, cur_key = -1;
        long long t_dep = 0;
        const uint32_t a_dst0 = smem_u32(smem_a);
        for (int l = 0; l < tp.num_layers; ++l) {
            const TowerLayer& L = tp.layer[l];
            const int a_kb = L.cin / BK;
            int ub, ue;
            range(l, ub, ue);
            for (int q = ub; q < ue; ++q) {
                const int g = unit_of(l, q) / nh;
                if (l * 65536 + g == cur_key) { continue; }
                cur_key = l * 65536 + g;
                const int buf = gcount & 1;
                mbar_wait_u32(smem_u32(&a_empty[buf]), ((gcount >> 1) & 1) ^ 1); // the MMAs that read this buffer two blocks ago are done
                if (l == 0 && gcount == 0 && tp.pdl) { // everything before this point (barriers, TMEM, weight prefetch) overlapped the previous kernel
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                if (l > 0) { // the 3x3 halo reaches into the neighbouring groups of the previous layer
                    const long long td = (DBG ? clock64() : 0ll);
                    if (lane < 3) { // three lanes, three counters: one L2 round trip when the rows are long complete, not three
                        const int gg = g - 1 + lane;
                        if (gg >= 0 && gg < num_groups) { wait_counter(tp.done + (l - 1) * num_groups + gg, need); }
                    }
                    __syncwarp();
                    asm volatile("fence.proxy.async;" ::: "memory"); // generic-proxy writes of other SMs -> this warp's TMA (async proxy) reads
                    t_dep += (DBG ? clock64() : 0ll) - td;
                }
                if (elect_one_sync()) {
                    if (leader) { mbar_arrive_expect_tx(&a_full[buf], 2 * a_kb * a_kb_bytes); }
                    const uint32_t bar = smem_u32(&a_full[buf]) & kPeerMask;
                    const uint64_t map_in_ptr = reinterpret_cast<uint64_t>(&L.map_in);
                    for (int kb = 0; kb < a_kb; ++kb) { tma_load_2d_2sm(a_dst0 + buf * a_bytes_max + kb * a_kb_bytes, map_in_ptr, bar, L.cin_off + kb * BK, (g * 2 + crank) * BM - tp.halo); }
                }
                __syncwarp();
                ++gcount;
            }
        }
        if (DBG && tp.dbg && lane == 0) { tp.dbg[blockIdx.x * 8 + 1] = t_dep; }
    } else if (warp == 1) {
        if (leader) { // ===== MMA issuer =====
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t a_kb_step = static_cast<uint32_t>(a_kb_bytes) >> 4, a_buf_step = static_cast<uint32_t>(a_bytes_max) >> 4;
            int s = 0, buf = 0, gcount = -1, cur_key = -1;
            uint32_t ph = 0, acc_ph0 = 1, acc_ph1 = 1;
            long long t_afull = 0, t_acc = 0, t_bfull = 0;
            const long long t_start = (DBG ? clock64() : 0ll);
            for (int l = 0; l < tp.num_layers; ++l) {
                const int cin = tp.layer[l].cin, tap_mask = tp.layer[l].tap_mask;
                int ub, ue;
                range(l, ub, ue);
                for (int q = ub; q < ue; ++q) {
                    const int grp = unit_of(l, q) / nh;
                    const int key = l * 65536 + grp;
                    if (key != cur_key) {
                        ++gcount;
                        const long long ta = (DBG ? clock64() : 0ll);
                        mbar_wait_u32(smem_u32(&a_full[gcount & 1]), (gcount >> 1) & 1);
                        t_afull += (DBG ? clock64() : 0ll) - ta;
                        cur_key = key;
                    }
                    const int abuf = gcount & 1;
                    const long long tc = (DBG ? clock64() : 0ll);
                    if (buf == 0) {
                        mbar_wait_u32(smem_u32(&acc_empty[0]), acc_ph0);
                        acc_ph0 ^= 1;
                    } else {
                        mbar_wait_u32(smem_u32(&acc_empty[1]), acc_ph1);
                        acc_ph1 ^= 1;
                    }
                    t_acc += (DBG ? clock64() : 0ll) - tc;
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * BN;
                    uint32_t accumulate = 0;
                    const int tap0 = (tp.tap_rot ? grp % 9 : 0);
                    {
                        uint32_t a_blk = a_lo0 + static_cast<uint32_t>(abuf) * a_buf_step;
                        for (int kc = 0; kc < cin; kc += BK, a_blk += a_kb_step) {
                            for (int t9 = 0; t9 < 9; ++t9) {
                                const int tap = (t9 + tap0 >= 9 ? t9 + tap0 - 9 : t9 + tap0), ty = tap / 3, tx = tap - 3 * ty;
                                const int row0 = tp.halo - tp.n1 - 1 + ty * tp.n1 + tx;
                                if (!((tap_mask >> tap) & 1)) { continue; }
                                const uint32_t a_lo = a_blk + static_cast<uint32_t>(row0) * 8u;
                                const long long tf = (DBG ? clock64() : 0ll);
                                mbar_wait_u32(full0 + s * 8, ph);
                                t_bfull += (DBG ? clock64() : 0ll) - tf;
                                tcgen05_fence_after();
                                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(s) * (B_HALF_BYTES >> 4);
                                if (elect_one_sync()) {
                                    umma_f16_lohi_2sm(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate);
                                    umma_f16_lohi_2sm(tmem_d, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
                                    umma_f16_lohi_2sm(tmem_d, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
                                    umma_f16_lohi_2sm(tmem_d, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
                                    tcgen05_commit_2sm_u32(empty0 + s * 8);
                                }
                                __syncwarp();
                                accumulate = 1u;
                                if (++s == STAGES) { s = 0, ph ^= 1; }
                            }
                        }
                    }
                    const bool last_of_group = (q + 1 == ue) || (unit_of(l, q + 1) / nh != grp);
                    if (elect_one_sync()) {
                        tcgen05_commit_2sm_u32(smem_u32(&acc_full[buf]));
                        if (last_of_group) { tcgen05_commit_2sm_u32(smem_u32(&a_empty[abuf])); }
                    }
                    __syncwarp();
                    buf ^= 1;
                }
            }
            if (DBG && tp.dbg && lane == 0) {
                tp.dbg[blockIdx.x * 8 + 3] = (DBG ? clock64() : 0ll) - t_start;
                tp.dbg[blockIdx.x * 8 + 4] = t_afull;
                tp.dbg[blockIdx.x * 8 + 5] = t_acc;
                tp.dbg[blockIdx.x * 8 + 6] = t_bfull;
            }
        }
    } else if (warp >= 2 && warp <= 5) { // ===== epilogue (both CTAs) =====
        const int quarter = warp & 3;
        int ucount = 0;
        long long t_epi_work = 0;
        if (BN == 128 && tp.epi_bufs > 0) {
            // Everything that does not need the accumulators happens before they are ready (the bias row in shared memory; the residual rows in
            // registers where the layer that wrote them is known: one acquire of its counter orders the loads); afterwards the warp converts
            // TMEM -> registers -> fp16 rows in a 64-byte-swizzled staging tile -> one TMA store per 32 channels, and publishes with one
            // red.release after the bulk stores have completed. (A lane-per-row st.global touches 32 cache lines per instruction, and a
            // __threadfence by every lane costs a MEMBAR.SC + an L1 invalidate: see conv_tower_wide_kernel.)
            const int ew = warp - 2;
            const uint32_t stage_u32 = smem_u32(smem_epi + ew * tp.epi_bufs * 2048), bias_u32 = smem_u32(smem_bias + ew * BN);
            const int buf_mask = tp.epi_bufs - 1;
            for (int l = 0; l < tp.num_layers; ++l) {
                const TowerLayer& L = tp.layer[l];
                const CUtensorMap* map_out = &tp.map_out[L.out_map];
                int ub, ue;
                range(l, ub, ue);
                for (int q = ub; q < ue; ++q, ++ucount) {
                    const int u = unit_of(l, q);
                    const int grp = u / nh, half = u - grp * nh, buf = ucount & 1;
                    const int mt = grp * 2 + crank;
                    const int n0 = half * BN;
                    const int r = mt * BM + quarter * 32 + lane;
                    const int rr = r % tp.slots;
                    const bool live = (r < tp.rows_valid) && (rr / tp.n1 != 0) && (rr % tp.n1 != tp.n1 - 1);
                    const bool in_range = (mt < tp.num_mtiles);
                    const bool add_res = (L.residual != nullptr) && live && in_range;
                    __syncwarp();
                    sts_f4(bias_u32 + 16 * lane, __ldg(reinterpret_cast<const float4*>(L.bias + n0) + lane));
                    uint4 res[16];
                    const uint4* res_row = reinterpret_cast<const uint4*>(L.residual + static_cast<size_t>(r) * tp.cout + n0);
                    const bool res_early = (L.residual != nullptr) && (L.res_layer != -2);
                    if (res_early) {
                        if (L.res_layer >= 0) { // rows written by other SMs earlier in this launch: complete long ago, but this warp has not synchronised with them yet
                            if (lane == 0) { wait_counter(tp.done + L.res_layer * num_groups + grp, need); }
                            __syncwarp();
                        }
                        if (add_res) { // read through L2, never through this SM's L1
#pragma unroll
                            for (int i = 0; i < 16; ++i) { res[i] = __ldcg(res_row + i); }
                        }
                    }
                    __syncwarp();
                    mbar_wait(&acc_full[buf], (ucount >> 1) & 1);
                    const long long tw = (DBG ? clock64() : 0ll);
                    tcgen05_fence_after();
                    if (add_res && !res_early) { // a layer wiring this kernel has no counter for: ordered by the accumulator barrier, as before
#pragma unroll
                        for (int i = 0; i < 16; ++i) { res[i] = __ldcg(res_row + i); }
                    }
                    if (in_range) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint32_t v[32];
                            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + c * 32, v);
                            float4 b0 = lds_f4(bias_u32 + c * 128), b1 = lds_f4(bias_u32 + c * 128 + 16);
                            tmem_ld_wait();
                            if (lane == 0) { // the staging tile is free again once the store that used it last has read it
                                if (buf_mask) {
                                    tma_store_wait_read1();
                                } else {
                                    tma_store_wait_read0();
                                }
                            }
                            __syncwarp();
                            const __half2* rh = reinterpret_cast<const __half2*>(&res[c * 4]);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                uint4 pk4;
                                uint32_t* pk = reinterpret_cast<uint32_t*>(&pk4);
                                const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                                if (k < 3) { b0 = lds_f4(bias_u32 + c * 128 + (k + 1) * 32), b1 = lds_f4(bias_u32 + c * 128 + (k + 1) * 32 + 16); }
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float x0 = __uint_as_float(v[k * 8 + 2 * j]) + bias[2 * j];
                                    float x1 = __uint_as_float(v[k * 8 + 2 * j + 1]) + bias[2 * j + 1];
                                    if (add_res) {
                                        const float2 rf = __half22float2(rh[k * 4 + j]);
                                        x0 += rf.x, x1 += rf.y;
                                    }
                                    if (!live) { x0 = 0.0f, x1 = 0.0f; }
                                    pk[j] = pack_act_f16x2(x0, x1, L.relu); // ReLU + saturation at the fp16 range + rounding in one F2FP
                                }
                                sts_u4(stage_u32 + (c & buf_mask) * 2048 + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4), pk4);
                            }
                            if (c == 3) { // the accumulators of this unit have been read: the issuer may reuse them
                                tcgen05_fence_before();
                                __syncwarp();
                                if (lane == 0) { mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u); }
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // this lane's row in shared memory -> the TMA engine
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(map_out, stage_u32 + (c & buf_mask) * 2048, n0 + c * 32, mt * BM + quarter * 32);
                                tma_store_commit();
                            }
                        }
                    } else if (lane == 0) { // a phantom tile behind the last row tile: nothing to read
                        mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u);
                    }
                    if (lane == 0) {
                        tma_store_wait_all(); // this warp's rows of (layer l, group grp) are written ...
                        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(tp.done + l * num_groups + grp), "r"(1) : "memory"); // ... before the counter moves
                    }
                    t_epi_work += (DBG ? clock64() : 0ll) - tw;
                }
            }
        } else {
            for (int l = 0; l < tp.num_layers; ++l) {
                const TowerLayer& L = tp.layer[l];
                int ub, ue;
                range(l, ub, ue);
                for (int q = ub; q < ue; ++q, ++ucount) {
                    const int u = unit_of(l, q);
                    const int grp = u / nh, half = u - grp * nh, buf = ucount & 1;
                    const int mt = grp * 2 + crank;
                    const int n0 = half * BN;
                    const int r = mt * BM + quarter * 32 + lane;
                    const int rr = r % tp.slots;
                    const bool live = (r < tp.rows_valid) && (rr / tp.n1 != 0) && (rr % tp.n1 != tp.n1 - 1);
                    const bool in_range = (mt < tp.num_mtiles);
                    mbar_wait(&acc_full[buf], (ucount >> 1) & 1);
                    const long long tw = (DBG ? clock64() : 0ll);
                    tcgen05_fence_after();
                    __half* out_row = L.out + static_cast<size_t>(r) * tp.cout + n0;
                    const __half* res_row = (L.residual ? L.residual + static_cast<size_t>(r) * tp.cout + n0 : nullptr);
    #pragma unroll 1
                    for (int c = 0; c < BN && in_range; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + c, v);
                        uint4 res[4];
                        if (res_row && live) { // written by other SMs earlier in this launch: read through L2, never through this SM's L1
    #pragma unroll
                            for (int q = 0; q < 4; ++q) { res[q] = __ldcg(reinterpret_cast<const uint4*>(res_row + c + q * 8)); }
                        }
                        float4 bias4[8];
    #pragma unroll
                        for (int q = 0; q < 8; ++q) { bias4[q] = __ldg(reinterpret_cast<const float4*>(L.bias + n0 + c) + q); }
                        const float* bias = reinterpret_cast<const float*>(bias4);
                        tmem_ld_wait();
                        uint4 packed[4];
                        uint32_t* pk = reinterpret_cast<uint32_t*>(packed);
                        const __half2* rh = reinterpret_cast<const __half2*>(res);
    #pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x0 = __uint_as_float(v[2 * j]) + bias[2 * j];
                            float x1 = __uint_as_float(v[2 * j + 1]) + bias[2 * j + 1];
                            if (res_row && live) {
                                const float2 rf = __half22float2(rh[j]);
                                x0 += rf.x, x1 += rf.y;
                            }
                            if (!live) { x0 = 0.0f, x1 = 0.0f; }
                            pk[j] = pack_act_f16x2(x0, x1, L.relu); // ReLU + saturation at the fp16 range + rounding in one F2FP
                        }
    #pragma unroll
                        for (int q = 0; q < 4; ++q) { *reinterpret_cast<uint4*>(out_row + c + q * 8) = packed[q]; }
                    }
                    tcgen05_fence_before();
                    if (tp.fence_mode == 0) {
                        __threadfence(); // this warp's rows of (layer l, group grp) are visible device-wide before the counter moves
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u);
                            atomicAdd(tp.done + l * num_groups + grp, 1);
                        }
                    } else {
                        __syncwarp(); // orders the lanes' row stores before lane 0's release (cumulativity carries them along)
                        if (lane == 0) {
                            mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u);
                            asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(tp.done + l * num_groups + grp), "r"(1) : "memory");
                        }
                    }
                    t_epi_work += (DBG ? clock64() : 0ll) - tw;
                }
            }
        }
        if (DBG && tp.dbg && warp == 2 && lane == 0) { tp.dbg[blockIdx.x * 8 + 7] = t_epi_work; }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Wide variant of the tower: every CTA owns TWO row tiles (256 rows, 512 per pair) per unit, so that each weight stage that
// arrives in shared memory feeds twice as many MMAs. Why: the narrow kernel streams a layer's full weights (1.18 MB at 256
// channels) through every CTA pair once per 256 rows, 1.5 GB of L2 -> shared-memory traffic per launch at config 2; with the
// MMAs running at their full rate that is ~39 B / clock / SM, ~5.8 KB / clock chip-wide — the L2's delivery limit — and the
// issuer waits for weights 19 % of its time (DESIGN.md "What bounds the tower"). Here the same weights serve 512 rows: half
// the weight traffic per FLOP. What changes with it:
//   * accumulators: 2 tiles x 128 columns per unit, double-buffered = all 512 TMEM columns;
//   * the input block of a unit (256 + 2 * halo rows x cin per CTA) no longer fits twice, so it is streamed as a ring of
//     WIDE_AK K-blocks (64 input channels each) and the loop order becomes K-block outer, row tile middle, tap inner: the
//     K-block of the NEXT unit loads while the current unit works on its later K-blocks, a K-block's nine weight stages
//     serve tile 0 and then tile 1 (STAGES >= 9), and tile 0's accumulators are complete 36 MMAs before tile 1's;
//   * 8 epilogue warps (two per TMEM lane quarter; each converts 32 of every 64 output channels, for both tiles): bias and
//     residual rows are fetched before the accumulators are ready, the fp16 rows leave through swizzled staging tiles and TMA
//     stores (DESIGN.md "The hand-off between layers");
//   * completion counters per (256-row subgroup = one CTA's rows, block of 64 output channels = one K-block of the next
//     layer), published as soon as both tiles of the block are stored; the consumer waits K-block by K-block.
// Units are dealt round-robin to the pairs (the narrow kernel's "strided" order). Same arithmetic per output element as the
// narrow kernel: both accumulate K-block outer / tap inner, so a position's result does not depend on which kernel ran it.
// ---------------------------------------------------------------------------------------------
constexpr int WIDE_THREADS = 384; // warp 0: weight TMA, warp 1: MMA issuer + TMEM owner, warp 2: input K-block TMA + dependency waits, warp 3: idle, warps 4-11: epilogue
constexpr int WIDE_AK = 3;        // ring slots of input K-blocks (64 input channels each): the producer runs up to three K-blocks = 0.75 units ahead of the MMAs,
                                  // far more than a TMA round trip; a fourth slot bought nothing and costs four weight stages
constexpr int WIDE_EPI_BYTES = 8 * 2 * 32 * 64; // epilogue staging: per warp two tiles of 32 rows x 32 fp16 channels, 64-byte swizzled like the TMA box that stores them
constexpr int WIDE_BIAS_BYTES = 8 * 64 * 4;

// shared memory of a wide tower CTA (the 1024 bytes at the end pay for aligning the base)
__host__ __device__ constexpr size_t wide_smem_bytes(int rows_ext, int stages)
{
    return static_cast<size_t>(WIDE_EPI_BYTES) + WIDE_BIAS_BYTES + static_cast<size_t>(WIDE_AK) * rows_ext * 128 + static_cast<size_t>(stages) * 64 * BK * 2 +
           (2 * stages + 2 * WIDE_AK + 6) * 8 + 16 + 1024;
}


template <int STAGES, bool DBG>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
conv_tower_wide_kernel(const __grid_constant__ TowerParams tp)
{
    static_assert(STAGES >= 9, "a K-block's nine weight stages stay resident while both row tiles use them");
    constexpr int BN = 128;
    constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
    constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const int a_kb_bytes = tp.rows_ext * 128; // rows_ext: 256 + 2 * halo rounded up to 16 (two TMA boxes of rows_ext / 2 rows)
    uint8_t* smem_epi = smem;                                   // 8 epilogue warps x 2 x 2 KB: fp16 rows of (32 rows, 32 channels) on their way to a TMA store
    float* smem_bias = reinterpret_cast<float*>(smem + WIDE_EPI_BYTES); // 8 epilogue warps x 64 floats: the bias of the warp's 64 channels
    uint8_t* smem_a = smem + WIDE_EPI_BYTES + WIDE_BIAS_BYTES;
    uint8_t* smem_b = smem_a + WIDE_AK * a_kb_bytes;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_b + STAGES * B_HALF_BYTES);
    uint64_t* b_empty = b_full + STAGES;
    uint64_t* a_full = b_empty + STAGES;   // [WIDE_AK]
    uint64_t* a_empty = a_full + WIDE_AK;  // [WIDE_AK]
    uint64_t* acc_full = a_empty + WIDE_AK; // [2 accumulator buffers][2 row tiles]
    uint64_t* acc_empty = acc_full + 4;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = tp.cout / BN;
    const int crank = static_cast<int>(cluster_ctarank());
    const bool leader = (crank == 0);
    const int cid = blockIdx.x / 2, nc = gridDim.x / 2;
    const int num_groups = (tp.num_mtiles + 3) / 4; // groups of 512 rows: two tiles per CTA of the pair
    const int num_sg = 2 * num_groups;              // subgroups of 256 rows: the rows one CTA writes
    const int units = num_groups * nh;
    const int nkbo = tp.cout / 64; // completion counters per (layer, subgroup): one per block of 64 output channels = per K-block of the next layer's input
    constexpr int need = 8;        // arrivals per counter: the CTA's 8 epilogue warps (each converts 32 of the 64 channels, for both row tiles)

    if (warp == 0 && lane == 0) {
        for (int l = 0; l < tp.num_layers; ++l) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tp.layer[l].map_in)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tp.layer[l].map_w)) : "memory");
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int i = 0; i < WIDE_AK; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_empty[i], 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t full0 = smem_u32(b_full), empty0 = smem_u32(b_empty);
    const uint32_t afull0 = smem_u32(a_full), aempty0 = smem_u32(a_empty);

    // pair c owns units c', c' + nc, c' + 2 nc, ... of layer l with c' = (c + l * rotate) mod nc (the rotation lets the pairs take turns at the short lists)
    auto first_unit = [&](int l) { return (cid + l * tp.rotate) % nc; };

    if (warp == 0) {
        // ===== weight producer (both CTAs): this CTA's half of every weight tile; waits for nothing but a free stage =====
        int s = 0;
        uint32_t ph = 1;
        long long t_bempty = 0;
        const long long t_start = (DBG ? clock64() : 0ll);
        const uint32_t b_dst0 = smem_u32(smem_b);
        for (int l = 0; l < tp.num_layers; ++l) {
            const TowerLayer& L = tp.layer[l];
            const uint64_t map_w_ptr = reinterpret_cast<uint64_t>(&L.map_w);
            for (int u = first_unit(l); u < units; u += nc) {
                const int wrow0 = (u % nh) * BN + crank * (BN / 2);
                for (int kc = 0; kc < L.cin; kc += BK) {
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!((L.tap_mask >> tap) & 1)) { continue; }
                        const long long te = (DBG ? clock64() : 0ll);
                        mbar_wait_u32(empty0 + s * 8, ph);
                        t_bempty += (DBG ? clock64() : 0ll) - te;
                        if (elect_one_sync()) {
                            if (leader) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + s * 8), "r"(2 * B_HALF_BYTES) : "memory"); }
                            tma_load_2d_2sm(b_dst0 + s * B_HALF_BYTES, map_w_ptr, (full0 + s * 8) & kPeerMask, kc, wrow0 + tap * tp.cout);
                        }
                        __syncwarp();
                        if (++s == STAGES) { s = 0, ph ^= 1; }
                    }
                }
            }
        }
        if (DBG && tp.dbg && lane == 0) {
            tp.dbg[blockIdx.x * 8 + 0] = (DBG ? clock64() : 0ll) - t_start;
            tp.dbg[blockIdx.x * 8 + 2] = t_bempty;
        }
    } else if (warp == 2) {
        // ===== input producer (both CTAs): the K-blocks of every unit's 256 + 2 * halo rows through the ring; it alone waits for the
        //       previous layer's completion counters =====
        int slot = 0;
        uint32_t ph = 1;
        long long t_dep = 0;
        const uint32_t a_dst0 = smem_u32(smem_a);
        const int box_rows = tp.rows_ext / 2;
        bool first = true;
        for (int l = 0; l < tp.num_layers; ++l) {
            const TowerLayer& L = tp.layer[l];
            const uint64_t map_in_ptr = reinterpret_cast<uint64_t>(&L.map_in);
            const int a_kb = L.cin / BK;
            for (int u = first_unit(l); u < units; u += nc) {
                const int sg = 2 * (u / nh) + crank;
                if (first && tp.pdl) { // everything before this point (barriers, TMEM, weight prefetch) overlapped the previous kernel
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                first = false;
                const int row0 = sg * 2 * BM - tp.halo;
                for (int kb = 0; kb < a_kb; ++kb) {
                    if (l > 0) { // K-block kb = 64 channels of the previous layer's output, rows of this subgroup and (3x3 halo) of its two neighbours: their
                                 // counters for THAT channel block only — a unit starts when the first 64 channels of its rows exist, the later blocks are
                                 // waited for as the ring gets to them (a quarter of a unit later each)
                        const long long td = (DBG ? clock64() : 0ll);
                        if (lane < 3) { // three lanes, three counters: one L2 round trip when the rows are long complete, not three
                            const int gg = sg - 1 + lane;
                            if (gg >= 0 && gg < num_sg) { wait_counter(tp.done + (static_cast<size_t>(l - 1) * num_sg + gg) * nkbo + (kb < nkbo ? kb : nkbo - 1), need); }
                        }
                        __syncwarp();
                        asm volatile("fence.proxy.async;" ::: "memory"); // writes of other SMs (TMA stores, ordered by their release) -> this warp's TMA (async proxy) reads
                        t_dep += (DBG ? clock64() : 0ll) - td;
                    }
                    mbar_wait_u32(aempty0 + slot * 8, ph); // the MMAs that read this slot WIDE_AK K-blocks ago are done
                    if (elect_one_sync()) {
                        if (leader) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(afull0 + slot * 8), "r"(2 * a_kb_bytes) : "memory"); }
                        const uint32_t bar = (afull0 + slot * 8) & kPeerMask;
                        const uint32_t dst = a_dst0 + slot * a_kb_bytes;
                        tma_load_2d_2sm(dst, map_in_ptr, bar, L.cin_off + kb * BK, row0);
                        tma_load_2d_2sm(dst + box_rows * 128, map_in_ptr, bar, L.cin_off + kb * BK, row0 + box_rows);
                    }
                    __syncwarp();
                    if (++slot == WIDE_AK) { slot = 0, ph ^= 1; }
                }
            }
        }
        if (DBG && tp.dbg && lane == 0) { tp.dbg[blockIdx.x * 8 + 1] = t_dep; }
    } else if (warp == 1) {
        if (leader) { // ===== MMA issuer =====
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (1u << 16), b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t a_slot_step = static_cast<uint32_t>(a_kb_bytes) >> 4;
            constexpr uint32_t a_tile_step = (BM * 128) >> 4; // the CTA's second tile starts 128 rows further down the block
            int s = 0, slot = 0, buf = 0;
            uint32_t ph = 0, aph = 0, acc_ph0 = 1, acc_ph1 = 1;
            long long t_afull = 0, t_acc = 0, t_bfull = 0;
            const long long t_start = (DBG ? clock64() : 0ll);
            for (int l = 0; l < tp.num_layers; ++l) {
                const int a_kb = tp.layer[l].cin / BK, tap_mask = tp.layer[l].tap_mask;
                for (int u = first_unit(l); u < units; u += nc) {
                    const long long tc = (DBG ? clock64() : 0ll);
                    if (buf == 0) {
                        mbar_wait_u32(smem_u32(&acc_empty[0]), acc_ph0);
                        acc_ph0 ^= 1;
                    } else {
                        mbar_wait_u32(smem_u32(&acc_empty[1]), acc_ph1);
                        acc_ph1 ^= 1;
                    }
                    t_acc += (DBG ? clock64() : 0ll) - tc;
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * (2 * BN);
                    // K-block outer, row tile middle, tap inner: a K-block's weight stages (one per tap) serve tile 0, then tile 1, and are released by
                    // tile 1 — the weights still stream once per 512 rows, but tile 0's accumulators are complete nine taps (36 MMAs, ~2.3 k cycles)
                    // before tile 1's, so half of the unit's epilogue runs under the last MMAs instead of after them. (Needs STAGES >= 9.)
                    for (int kb = 0; kb < a_kb; ++kb) {
                        const long long ta = (DBG ? clock64() : 0ll);
                        mbar_wait_u32(afull0 + slot * 8, aph);
                        t_afull += (DBG ? clock64() : 0ll) - ta;
                        tcgen05_fence_after();
                        const uint32_t a_blk = a_lo0 + static_cast<uint32_t>(slot) * a_slot_step;
                        const int s_kb = s;
                        const bool last_kb = (kb + 1 == a_kb);
                        const uint32_t acc0 = (kb > 0 ? 1u : 0u); // a unit's first MMA into a tile overwrites it
                        const uint32_t tap_row0 = static_cast<uint32_t>(tp.halo - tp.n1 - 1);
                        { // ---- tile 0: waits for the weight stages as they arrive
                            uint32_t accumulate = acc0;
                            for (int tap = 0; tap < 9; ++tap) {
                                if (!((tap_mask >> tap) & 1)) { continue; }
                                const int ty = tap / 3, tx = tap - 3 * ty;
                                const uint32_t a_lo = a_blk + (tap_row0 + static_cast<uint32_t>(ty * tp.n1 + tx)) * 8u;
                                const long long tf = (DBG ? clock64() : 0ll);
                                mbar_wait_u32(full0 + s * 8, ph);
                                t_bfull += (DBG ? clock64() : 0ll) - tf;
                                tcgen05_fence_after();
                                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(s) * (B_HALF_BYTES >> 4);
                                if (elect_one_sync()) {
                                    umma_f16_lohi_2sm(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate);
                                    umma_f16_lohi_2sm(tmem_d, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
                                    umma_f16_lohi_2sm(tmem_d, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
                                    umma_f16_lohi_2sm(tmem_d, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
                                }
                                __syncwarp();
                                accumulate = 1u;
                                if (++s == STAGES) { s = 0, ph ^= 1; }
                            }
                        }
                        // ---- tile 1: every stage of the K-block is in place — one thread issues the whole pass (no per-tap election or barrier
                        //      traffic: with four MMAs per tap the issue loop must stay well under their 256 tensor cycles) and releases the stages
                        if (elect_one_sync()) {
                            if (last_kb) { tcgen05_commit_2sm_u32(smem_u32(&acc_full[buf * 2 + 0])); } // tile 0's accumulators are complete
                            uint32_t accumulate = acc0;
                            int s1 = s_kb;
                            const uint32_t a_t1 = a_blk + a_tile_step;
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                if (!((tap_mask >> tap) & 1)) { continue; }
                                const int ty = tap / 3, tx = tap - 3 * ty;
                                const uint32_t a_lo = a_t1 + (tap_row0 + static_cast<uint32_t>(ty * tp.n1 + tx)) * 8u;
                                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(s1) * (B_HALF_BYTES >> 4);
                                umma_f16_lohi_2sm(tmem_d + BN, a_lo, b_lo, desc_hi, idesc, accumulate);
                                umma_f16_lohi_2sm(tmem_d + BN, a_lo + 2, b_lo + 2, desc_hi, idesc, 1u);
                                umma_f16_lohi_2sm(tmem_d + BN, a_lo + 4, b_lo + 4, desc_hi, idesc, 1u);
                                umma_f16_lohi_2sm(tmem_d + BN, a_lo + 6, b_lo + 6, desc_hi, idesc, 1u);
                                tcgen05_commit_2sm_u32(empty0 + s1 * 8);
                                accumulate = 1u;
                                if (++s1 == STAGES) { s1 = 0; }
                            }
                            if (last_kb) { tcgen05_commit_2sm_u32(smem_u32(&acc_full[buf * 2 + 1])); }
                        }
                        __syncwarp();
                        if (elect_one_sync()) { tcgen05_commit_2sm_u32(aempty0 + slot * 8); }
                        __syncwarp();
                        if (++slot == WIDE_AK) { slot = 0, aph ^= 1; }
                    }
                    buf ^= 1;
                }
            }
            if (DBG && tp.dbg && lane == 0) {
                tp.dbg[blockIdx.x * 8 + 3] = (DBG ? clock64() : 0ll) - t_start;
                tp.dbg[blockIdx.x * 8 + 4] = t_afull;
                tp.dbg[blockIdx.x * 8 + 5] = t_acc;
                tp.dbg[blockIdx.x * 8 + 6] = t_bfull;
            }
        }
    } else if (warp >= 4) { // ===== epilogue (both CTAs): warp & 3 = TMEM lane quarter, (warp - 4) / 4 = which 64 of a tile's 128 columns =====
        // The time from a unit's last MMA to its rows being published is on the critical path of every unit of the next layer that needs them
        // (at config 2 a layer holds 1.35 units per pair: most units wait for rows finished just before). So everything that does not need the
        // accumulators happens BEFORE they are ready — the bias of the warp's 64 channels goes to shared memory, the residual rows (complete
        // since layer l - 2; checked with one acquire of that layer's counter) into registers — and afterwards the warp only converts: TMEM ->
        // registers -> fp16 rows in a 128-byte-swizzled staging tile -> ONE TMA store per (row tile, 64 channels). A lane-per-row st.global
        // touches 32 cache lines per instruction (measured: 12 k cycles per unit in the epilogue, a third of them in __threadfence's
        // MEMBAR.SC + L1 invalidate); the TMA store leaves the LSU out of it, and publishing is: wait for the bulk stores, one red.release.
        const int quarter = warp & 3, x = (warp - 4) >> 2, ew = warp - 4; // x: which 32 of every 64-channel block this warp converts
        const uint32_t stage_u32 = smem_u32(smem_epi + ew * 4096); // two tiles of 2 KB: piece 0 / piece 1
        const uint32_t bias_u32 = smem_u32(smem_bias + ew * 64);
        int ucount = 0;
        long long t_epi_work = 0;
        for (int l = 0; l < tp.num_layers; ++l) {
            const TowerLayer& L = tp.layer[l];
            const CUtensorMap* map_out = &tp.map_out[L.out_map];
            for (int u = first_unit(l); u < units; u += nc, ++ucount) {
                const int grp = u / nh, half = u - grp * nh, buf = ucount & 1;
                const int mt0 = grp * 4 + crank * 2, sg = 2 * grp + crank;
                const int n0 = half * BN + x * 32; // this warp's channels: n0 .. n0 + 32 (piece 0, channel block 2 * half) and n0 + 64 .. n0 + 96 (piece 1, block 2 * half + 1)
                int r[2];
                bool live[2], in_range[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    r[t] = (mt0 + t) * BM + quarter * 32 + lane;
                    const int rr = r[t] % tp.slots;
                    live[t] = (r[t] < tp.rows_valid) && (rr / tp.n1 != 0) && (rr % tp.n1 != tp.n1 - 1);
                    in_range[t] = (mt0 + t < tp.num_mtiles);
                }
                // ---- before the accumulators are ready: bias slice and residual rows
                __syncwarp();
                sts_f1(bias_u32 + 4 * lane, __ldg(L.bias + n0 + lane)), sts_f1(bias_u32 + 128 + 4 * lane, __ldg(L.bias + n0 + 64 + lane)); // piece p, channel j: float 32 p + j
                uint4 res[2][8];
                if (L.residual) {
                    if (L.res_layer >= 0) { // rows of an earlier layer of this launch, written by other SMs: complete long ago, but this warp has not synchronised with them yet
                        if (lane < 2) { wait_counter(tp.done + (static_cast<size_t>(L.res_layer) * num_sg + sg) * nkbo + 2 * half + lane, need); }
                        __syncwarp();
                    }
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (live[t] && in_range[t]) { // read through L2, never through this SM's L1
                            const uint4* res_row = reinterpret_cast<const uint4*>(L.residual + static_cast<size_t>(r[t]) * tp.cout + n0);
#pragma unroll
                            for (int q = 0; q < 4; ++q) { res[t][q] = __ldcg(res_row + q), res[t][4 + q] = __ldcg(res_row + 8 + q); }
                        }
                    }
                }
                __syncwarp();
                long long tw = 0;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (!in_range[t]) { continue; } // warp-uniform
                    mbar_wait(&acc_full[buf * 2 + t], (ucount >> 1) & 1); // tile 0 is complete nine taps before tile 1: its rows are converted under the last MMAs
                    const bool last_tile = (t == 1 || !in_range[1]);
                    if (last_tile) { tw = (DBG ? clock64() : 0ll); }
                    tcgen05_fence_after();
                    const bool add_res = (L.residual != nullptr) && live[t];
                    uint32_t vv[2][32]; // both pieces of the tile at once: one TMEM round trip, and the accumulators are free before the first store
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * (2 * BN) + t * BN + x * 32, vv[0]);
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * (2 * BN) + t * BN + 64 + x * 32, vv[1]);
                    tmem_ld_wait();
                    if (last_tile) { // the accumulators of this unit have been read: the issuer may reuse them
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u); }
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) { // piece c: channels n0 + 64 c .. + 32
                        const uint32_t* v = vv[c];
                        if (lane == 0) { tma_store_wait_read1(); } // staging tile c is free again once the store before the last one has read it
                        __syncwarp();
                        const __half2* rh = reinterpret_cast<const __half2*>(&res[t][c * 4]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 pk4;
                            uint32_t* pk = reinterpret_cast<uint32_t*>(&pk4);
                            const float4 b0 = lds_f4(bias_u32 + c * 128 + q * 32), b1 = lds_f4(bias_u32 + c * 128 + q * 32 + 16);
                            const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float x0 = __uint_as_float(v[q * 8 + 2 * j]) + bias[2 * j];
                                float x1 = __uint_as_float(v[q * 8 + 2 * j + 1]) + bias[2 * j + 1];
                                if (add_res) {
                                    const float2 rf = __half22float2(rh[q * 4 + j]);
                                    x0 += rf.x, x1 += rf.y;
                                }
                                if (!live[t]) { x0 = 0.0f, x1 = 0.0f; }
                                pk[j] = pack_act_f16x2(x0, x1, L.relu); // ReLU + saturation at the fp16 range + rounding in one F2FP
                            }
                            // row `lane` of staging tile c, 16-byte chunk q, at the position the 64-byte swizzle gives it (conflict-free: 8 lanes cover 8 chunk columns)
                            sts_u4(stage_u32 + c * 2048 + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), pk4);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // this lane's row in shared memory -> the TMA engine
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(map_out, stage_u32 + c * 2048, n0 + c * 64, (mt0 + t) * BM + quarter * 32);
                            tma_store_commit();
                            if (last_tile) { // channel block 2 * half + c of this subgroup: both tiles' rows of this warp are on their way — publish it on its own,
                                             // the next layer's K-block 2 * half + c waits for nothing else
                                tma_store_wait_all(); // ... written ...
                                asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(tp.done + (static_cast<size_t>(l) * num_sg + sg) * nkbo + 2 * half + c), "r"(1)
                                             : "memory"); // ... before the counter moves
                            }
                        }
                    }
                }
                if (!in_range[0] && lane == 0) { // a phantom unit behind the last row tile: nothing to read, but the counters are waited for
                    mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0u);
                    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(tp.done + (static_cast<size_t>(l) * num_sg + sg) * nkbo + 2 * half), "r"(1) : "memory");
                    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(tp.done + (static_cast<size_t>(l) * num_sg + sg) * nkbo + 2 * half + 1), "r"(1) : "memory");
                }
                t_epi_work += (DBG ? clock64() : 0ll) - tw;
            }
        }
        if (DBG && tp.dbg && warp == 4 && lane == 0) { tp.dbg[blockIdx.x * 8 + 7] = t_epi_work; }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// heads: one CTA per board. conv1x1 (+folded BN) + ReLU for the policy and value planes, then the
// fully connected layers, softmax over the policy logits and tanh on the value. fp32 SIMT: 0.2 MFLOP / board.
// ---------------------------------------------------------------------------------------------
struct HeadParams {
    const __half* act;  // [rows][c] final tower activations
    const float* w_pc;  // [pol_ch][c] policy conv (BN folded)   b_pc [pol_ch]
    const float* b_pc;
    const float* w_pf;  // [pol_ch * hw][A] policy fc, transposed  b_pf [A]
    const float* b_pf;
    const float* w_vc;  // [c] value conv (BN folded)            b_vc [1]
    const float* b_vc;
    const float* w_v1;  // [hw][vh] transposed   b_v1 [vh]
    const float* b_v1;
    const float* w_v2;  // [vh]       b_v2 [1]
    const float* b_v2;
    float* policy;      // [B][A]
    float* logits;      // [B][A]
    float* value;       // [B]
    int c, n, slots, pol_ch, actions, vh;
    int batch;          // boards; CTAs loop over them
    int fc_in_smem;     // 1: the transposed FC weights are staged in shared memory once per CTA
    int* clear;         // completion counters of the tower launch that produced `act` (or null): zeroed here for its next launch,
    int clear_count;    // which saves a memset node between every two kernels of the search graph
};

template <int NP1> // NP1 = policy planes + 1 value plane, a compile-time constant so that the plane loops carry no predicates
__global__ void __launch_bounds__(1024) heads_kernel(const HeadParams p)
{
    extern __shared__ float sm[];
    const int hw = p.n * p.n, n1 = p.n + 1;
    float* wc = sm;                      // [NP1][c] 1x1 conv weights (first: float2 reads need 8-byte alignment)
    float* planes = wc + NP1 * p.c;      // [NP1][hw]: policy planes then the value plane
    float* vhid = planes + NP1 * hw;     // [vh]
    float* lg = vhid + p.vh;             // [A]
    float* red = lg + p.actions;         // [32]
    int* rowoff = reinterpret_cast<int*>(red + 32); // [hw] element offset of every cell's activation row
    float* partial = reinterpret_cast<float*>(rowoff + hw); // [parts][A + vh]
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
    const int g = blockIdx.x;
    const __half* act = p.act + static_cast<size_t>(g) * p.slots * p.c;
    if (p.clear) { // the tower has finished (stream order); its next launch comes after this kernel
        for (int i = g * nthr + tid; i < p.clear_count; i += gridDim.x * nthr) { p.clear[i] = 0; }
    }
    for (int i = tid; i < NP1 * p.c; i += nthr) { wc[i] = (i < p.pol_ch * p.c ? p.w_pc[i] : p.w_vc[i - p.pol_ch * p.c]); }
    for (int cell = tid; cell < hw; cell += nthr) { rowoff[cell] = ((cell / p.n + 1) * n1 + cell % p.n) * p.c; }
    __syncthreads();
    // 1x1 convolutions: one warp per cell; lane l owns channel pairs {2l + 64i}: every load instruction is one contiguous
    // 128-byte row segment and the weight reads from shared memory are conflict-free
    const int npair = p.c / 64;
    for (int cell = warp; cell < hw; cell += nwarp) {
        const __half2* row = reinterpret_cast<const __half2*>(act + rowoff[cell]);
        float acc[NP1];
#pragma unroll
        for (int o = 0; o < NP1; ++o) { acc[o] = 0.0f; }
        for (int i = 0; i < npair; ++i) {
            const float2 a = __half22float2(row[lane + 32 * i]);
            const float* wp = wc + 2 * lane + 64 * i;
#pragma unroll
            for (int o = 0; o < NP1; ++o) {
                const float2 wv = *reinterpret_cast<const float2*>(wp + o * p.c);
                acc[o] = fmaf(a.x, wv.x, fmaf(a.y, wv.y, acc[o]));
            }
        }
#pragma unroll
        for (int o = 0; o < NP1; ++o) {
            float v = acc[o];
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, sft); }
            acc[o] = v;
        }
        if (lane < NP1) {
            float v = acc[0];
#pragma unroll
            for (int o = 1; o < NP1; ++o) { v = (lane == o ? acc[o] : v); }
            planes[lane * hw + cell] = fmaxf(v + (lane < p.pol_ch ? p.b_pc[lane] : p.b_vc[0]), 0.0f);
        }
    }
    __syncthreads();
    // policy fc and value fc1 over TRANSPOSED weights [in][out] (coalesced across threads); every output's input range is
    // split over `parts` threads so that enough independent loads are in flight; partial sums meet in shared memory
    const int nout = p.actions + p.vh;
    const int parts = (nthr >= 4 * nout ? 4 : (nthr >= 2 * nout ? 2 : 1));
    for (int t = tid; t < parts * nout; t += nthr) {
        const int part = t / nout, o = t - part * nout;
        float acc = 0.0f;
        if (o < p.actions) {
            const int nin = p.pol_ch * hw, i0 = (nin * part) / parts, i1 = (nin * (part + 1)) / parts;
            const float* wp = p.w_pf + o;
#pragma unroll 16
            for (int i = i0; i < i1; ++i) { acc = fmaf(planes[i], __ldg(wp + static_cast<size_t>(i) * p.actions), acc); }
        } else {
            const int i0 = (hw * part) / parts, i1 = (hw * (part + 1)) / parts;
            const float* vp = planes + p.pol_ch * hw;
            const float* wp = p.w_v1 + (o - p.actions);
#pragma unroll 16
            for (int i = i0; i < i1; ++i) { acc = fmaf(vp[i], __ldg(wp + static_cast<size_t>(i) * p.vh), acc); }
        }
        partial[t] = acc;
    }
    __syncthreads();
    for (int o = tid; o < nout; o += nthr) {
        float acc = 0.0f;
        for (int part = 0; part < parts; ++part) { acc += partial[part * nout + o]; }
        if (o < p.actions) {
            lg[o] = acc + p.b_pf[o];
        } else {
            vhid[o - p.actions] = fmaxf(acc + p.b_v1[o - p.actions], 0.0f);
        }
    }
    __syncthreads();
    // softmax over the logits and value fc2 + tanh: warp 0 reduces, everybody normalises
    if (warp == 0) {
        float mx = -3.402823466e+38f;
        for (int a = lane; a < p.actions; a += 32) { mx = fmaxf(mx, lg[a]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        float sum = 0.0f;
        for (int a = lane; a < p.actions; a += 32) { sum += expf(lg[a] - mx); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); }
        if (lane == 0) { red[0] = mx, red[1] = sum; }
    } else if (warp == 1) {
        float acc = 0.0f;
        for (int j = lane; j < p.vh; j += 32) { acc = fmaf(vhid[j], p.w_v2[j], acc); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); }
        if (lane == 0) { p.value[g] = tanhf(acc + p.b_v2[0]); }
    }
    __syncthreads();
    const float mx = red[0], inv = 1.0f / red[1];
    for (int a = tid; a < p.actions; a += nthr) {
        p.logits[static_cast<size_t>(g) * p.actions + a] = lg[a];
        p.policy[static_cast<size_t>(g) * p.actions + a] = expf(lg[a] - mx) * inv;
    }
}


// ---------------------------------------------------------------------------------------------
// MuZero: MuZeroNetwork.scale_hidden_state (network/py/muzero_network.py:150-160) on the tower's output rows, one CTA per
// board: min / max over the board's real channels and cells, x <- (x - min) / scale with scale = max - min (+1e-5 when
// below 1e-5). The scaled state is written back in place (the prediction heads read it there) and into the hidden-state
// slot of the node being evaluated (hid[g][slot[g]], compact [cell][c] rows; padded channels stay zero).
// ---------------------------------------------------------------------------------------------
// think() lanes (trees > 0): board g = lane * trees + tree stores into ITS TREE's slots, and only when the lane holds a leaf to evaluate (path_len > 0;
// a duplicate or unused lane carries stale rows and a stale slot index)
__global__ void __launch_bounds__(256) scale_hidden_kernel(__half* __restrict__ act, __half* __restrict__ hid, const int32_t* __restrict__ slot, int n, int slots, int c,
                                                          int c_real, int num_slots, int trees, const int32_t* __restrict__ path_len)
{
    __shared__ float red_mn[8], red_mx[8];
    const int g = blockIdx.x, tid = threadIdx.x, hw = n * n, n1 = n + 1;
    if (path_len && path_len[g] <= 0) { return; } // uniform for the block
    __half* rows = act + static_cast<size_t>(g) * slots * c;
    __half* dst = hid + (static_cast<size_t>(trees > 0 ? g % trees : g) * num_slots + slot[g]) * hw * c;
    const bool vec = (c_real % 8 == 0); // 16-byte accesses: 8 channels per thread and step
    const int per_cell = c / 8, real_per_cell = c_real / 8;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    if (vec && c == c_real && hw * per_cell <= 4 * static_cast<int>(blockDim.x)) {
        // the common shape (no padded channels, at most 4 x 16 bytes per thread: 8 x 8 cells x 128 channels at 256 threads): the board is read ONCE, all loads
        // in flight together, kept in registers across the reduction, and leaves as two coalesced stores — one round trip to L2 instead of two
        const int items = hw * per_cell;
        uint4 v[4];
        __half* src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = tid + u * blockDim.x;
            src[u] = nullptr;
            if (i < items) {
                const int cell = i / per_cell, k = i - cell * per_cell;
                src[u] = rows + static_cast<size_t>((cell / n + 1) * n1 + cell % n) * c + 8 * k;
                v[u] = *reinterpret_cast<const uint4*>(src[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (src[u]) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    mn = fminf(mn, fminf(f.x, f.y)), mx = fmaxf(mx, fmaxf(f.x, f.y));
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)), mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if ((tid & 31) == 0) { red_mn[tid >> 5] = mn, red_mx[tid >> 5] = mx; }
        __syncthreads();
        mn = red_mn[0], mx = red_mx[0];
        for (int i = 1; i < (blockDim.x >> 5); ++i) { mn = fminf(mn, red_mn[i]), mx = fmaxf(mx, red_mx[i]); }
        float scale = mx - mn;
        if (scale < 1e-5f) { scale += 1e-5f; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (src[u]) {
                uint4 out;
                const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
                __half2* o = reinterpret_cast<__half2*>(&out);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    o[j] = __floats2half2_rn((f.x - mn) / scale, (f.y - mn) / scale); // same operations as the general path below
                }
                *reinterpret_cast<uint4*>(src[u]) = out;
                *reinterpret_cast<uint4*>(dst + static_cast<size_t>(tid + u * blockDim.x) * 8) = out;
            }
        }
        return;
    }
    if (vec) {
        for (int i = tid; i < hw * real_per_cell; i += blockDim.x) {
            const int cell = i / real_per_cell, k = i - cell * real_per_cell;
            const uint4 v = *reinterpret_cast<const uint4*>(rows + static_cast<size_t>((cell / n + 1) * n1 + cell % n) * c + 8 * k);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                mn = fminf(mn, fminf(f.x, f.y)), mx = fmaxf(mx, fmaxf(f.x, f.y));
            }
        }
    } else {
        for (int i = tid; i < hw * c_real; i += blockDim.x) {
            const int cell = i / c_real, ch = i - cell * c_real;
            const float v = __half2float(rows[static_cast<size_t>((cell / n + 1) * n1 + cell % n) * c + ch]);
            mn = fminf(mn, v), mx = fmaxf(mx, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)), mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((tid & 31) == 0) { red_mn[tid >> 5] = mn, red_mx[tid >> 5] = mx; }
    __syncthreads();
    mn = red_mn[0], mx = red_mx[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) { mn = fminf(mn, red_mn[i]), mx = fmaxf(mx, red_mx[i]); }
    float scale = mx - mn;
    if (scale < 1e-5f) { scale += 1e-5f; }
    if (vec) {
        for (int i = tid; i < hw * per_cell; i += blockDim.x) {
            const int cell = i / per_cell, k = i - cell * per_cell;
            uint4 out = make_uint4(0u, 0u, 0u, 0u); // padded channels stay zero
            if (k < real_per_cell) {
                uint4* p = reinterpret_cast<uint4*>(rows + static_cast<size_t>((cell / n + 1) * n1 + cell % n) * c + 8 * k);
                const uint4 v = *p;
                const __half2* h = reinterpret_cast<const __half2*>(&v);
                __half2* o = reinterpret_cast<__half2*>(&out);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    o[j] = __floats2half2_rn((f.x - mn) / scale, (f.y - mn) / scale);
                }
                *p = out;
            }
            *reinterpret_cast<uint4*>(dst + static_cast<size_t>(i) * 8) = out;
        }
    } else {
        for (int i = tid; i < hw * c; i += blockDim.x) {
            const int cell = i / c, ch = i - cell * c;
            __half* p = rows + static_cast<size_t>((cell / n + 1) * n1 + cell % n) * c + ch;
            __half out = __float2half_rn(0.0f);
            if (ch < c_real) {
                out = __float2half_rn((__half2float(*p) - mn) / scale);
                *p = out;
            }
            dst[i] = out;
        }
    }
}

// MuZeroNetwork::pushBackRecurrentData layout (network/muzero_network.h:78-93) -> rows of the dynamics network's input:
// hidden [n][c_real][H][W] fp32 + action ids -> [rows][dyn_c] fp16 with the one-hot action plane at column act_col (parity hook)
__global__ void pack_hidden_kernel(const float* __restrict__ hidden, const int32_t* __restrict__ actions, __half* __restrict__ rows, int batch, int c_real, int n,
                                   int slots, int dyn_c, int act_col, int act_planes)
{
    const int hw = n * n, total = batch * hw * dyn_c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = i / (hw * dyn_c), cell = (i / dyn_c) % hw, ch = i % dyn_c;
        float v = 0.0f;
        if (ch < c_real) {
            v = hidden[(static_cast<size_t>(g) * c_real + ch) * hw + cell];
        } else if (act_planes == 1 && ch == act_col) {
            v = (actions[g] == cell ? 1.0f : 0.0f);
        } else if (act_planes > 1 && ch >= act_col && ch < act_col + act_planes) { // Atari: plane `action` of the block is all ones (atari.cpp:124-130)
            v = (actions[g] == ch - act_col ? 1.0f : 0.0f);
        }
        rows[(static_cast<size_t>(g) * slots + (cell / n + 1) * (n + 1) + cell % n) * dyn_c + ch] = __float2half_rn(v);
    }
}

// stored hidden state (slot `which` of every game, compact [cell][c] fp16) -> [n][c_real][H][W] fp32 (parity hook)
__global__ void unpack_hidden_kernel(const __half* __restrict__ hid, float* __restrict__ hidden, int batch, int c_real, int n, int c, int num_slots, int which)
{
    const int hw = n * n, total = batch * c_real * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = i / (c_real * hw), ch = (i / hw) % c_real, cell = i % hw;
        hidden[i] = __half2float(hid[((static_cast<size_t>(g) * num_slots + which) * hw + cell) * c + ch]);
    }
}

// float NCHW feature planes (host layout of the reference, alphazero_network.h:48-61) -> fp16 shared-halo rows
__global__ void pack_features_kernel(const float* __restrict__ feats, __half* __restrict__ rows, int batch, int c, int n, int slots, int cpad)
{
    const int hw = n * n, total = batch * c * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = i / (c * hw), ch = (i / hw) % c, cell = i % hw;
        rows[(static_cast<size_t>(g) * slots + (cell / n + 1) * (n + 1) + cell % n) * cpad + ch] = __float2half_rn(feats[i]);
    }
}

// fp16 shared-halo rows -> float NCHW planes (parity hook for the feature planes the search produced)
__global__ void unpack_features_kernel(const __half* __restrict__ rows, float* __restrict__ feats, int batch, int c, int n, int slots, int cpad)
{
    const int hw = n * n, total = batch * c * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = i / (c * hw), ch = (i / hw) % c, cell = i % hw;
        feats[i] = __half2float(rows[(static_cast<size_t>(g) * slots + (cell / n + 1) * (n + 1) + cell % n) * cpad + ch]);
    }
}

} // namespace mznn
