// Kernels around the conv tower for the Atari MuZero network (network/py/muzero_atari_network.py): screen-ring -> network input,
// space-to-depth (the stride-2 convolutions run as stride-1 towers over a space-to-depth input, see engine.cu "stride 2"),
// 3x3 / stride 2 average pooling, and the discrete (601-bin) value / reward heads with the expectation and invertValue fused.
//
// Activation layout as in nn_kernels.cuh ("shared-halo rows"): a feature map of n x n cells owns (n + 1)^2 consecutive rows of
// `c` fp16 channels; cell (x, y) is row (y + 1) * (n + 1) + x; rows with yy == 0 or xx == n are zero.
#pragma once
#include "nn_kernels.cuh"

namespace mzat {

constexpr int RES = 96;               // kAtariResolution, atari.h:24
constexpr int HIST = 8;               // kAtariFeatureHistorySize, atari.h:27
constexpr int PLANES = 4 * HIST;      // per history entry: action plane, R, G, B (atari.cpp:106-116)
constexpr int FRAME = 3 * RES * RES;  // bytes of one screen, CHW

__device__ __forceinline__ size_t cell_row(int g, int n, int y, int x) { return static_cast<size_t>(g) * (n + 1) * (n + 1) + static_cast<size_t>(y + 1) * (n + 1) + x; }

// value of plane `c` (0 .. 31) at pixel `pix` of game g's current observation history: AtariEnv::getFeatures (atari.cpp:106-116)
// over the ring (oldest entry at meta[0]); action planes are id * 1.0f / 18 (atari.cpp:82), screens byte / 255.0f (:152-156)
__device__ __forceinline__ float plane_value(const uint8_t* __restrict__ frames, const int32_t* __restrict__ meta, int g, int c, int pix)
{
    const int32_t* m = meta + static_cast<size_t>(g) * 16;
    const int slot = (m[0] + (c >> 2)) & (HIST - 1), kind = c & 3;
    if (kind == 0) {
        const int a = m[1 + slot];
        return a < 0 ? 0.0f : __fdiv_rn(__fmul_rn(static_cast<float>(a), 1.0f), 18.0f);
    }
    if (!((m[9] >> slot) & 1)) { return 0.0f; }
    return __fdiv_rn(static_cast<float>(frames[(static_cast<size_t>(g) * HIST + slot) * FRAME + (kind - 1) * RES * RES + pix]), 255.0f);
}

// observation history -> float planes [n][32][96][96] (parity hook: exactly what the reference pushes to the network)
__global__ void planes_f32_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ meta, float* __restrict__ out, int batch)
{
    const size_t total = static_cast<size_t>(batch) * PLANES * RES * RES;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int pix = static_cast<int>(i % (RES * RES)), c = static_cast<int>((i / (RES * RES)) % PLANES), g = static_cast<int>(i / (static_cast<size_t>(PLANES) * RES * RES));
        out[i] = plane_value(frames, meta, g, c, pix);
    }
}

// network input of the representation tower: space-to-depth(2) of the 32 planes, fp16 rows of a 48 x 48 map with 4 * 32 = 128
// channels; channel (dy * 2 + dx) * 32 + c of cell (X, Y) = plane c at pixel (2X + dx, 2Y + dy). One thread per (cell, 8 channels).
// FROM_F32: the planes come as floats [n][32][96][96] (network parity hook), else from the screen ring.
template <bool FROM_F32>
__global__ void pack_input_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ meta, const float* __restrict__ planes, __half* __restrict__ rows, int batch)
{
    constexpr int NS = RES / 2, CH = 4 * PLANES;
    const size_t total = static_cast<size_t>(batch) * NS * NS * (CH / 8);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(i % (CH / 8));
        const int cell = static_cast<int>((i / (CH / 8)) % (NS * NS)), g = static_cast<int>(i / (static_cast<size_t>(CH / 8) * NS * NS));
        const int X = cell % NS, Y = cell / NS, sub = chunk / (PLANES / 8), c0 = (chunk % (PLANES / 8)) * 8;
        const int pix = (2 * Y + (sub >> 1)) * RES + 2 * X + (sub & 1);
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a, b;
            if (FROM_F32) {
                a = planes[(static_cast<size_t>(g) * PLANES + c0 + 2 * j) * RES * RES + pix], b = planes[(static_cast<size_t>(g) * PLANES + c0 + 2 * j + 1) * RES * RES + pix];
            } else {
                a = plane_value(frames, meta, g, c0 + 2 * j, pix), b = plane_value(frames, meta, g, c0 + 2 * j + 1, pix);
            }
            h[j] = __floats2half2_rn(a, b);
        }
        *reinterpret_cast<uint4*>(rows + cell_row(g, NS, Y, X) * CH + chunk * 8) = v;
    }
}

// space-to-depth(2) between two towers: in = n_in x n_in cells of c_in channels, out = (n_in / 2)^2 cells of 4 * c_in channels,
// channel (dy * 2 + dx) * c_in + c of cell (X, Y) = channel c of cell (2X + dx, 2Y + dy)
__global__ void space_to_depth_kernel(const __half* __restrict__ in, __half* __restrict__ out, int batch, int n_in, int c_in)
{
    const int n_out = n_in / 2, per_cell = 4 * c_in / 8, per_sub = c_in / 8;
    const size_t total = static_cast<size_t>(batch) * n_out * n_out * per_cell;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % per_cell), cell = static_cast<int>((i / per_cell) % (n_out * n_out)), g = static_cast<int>(i / (static_cast<size_t>(per_cell) * n_out * n_out));
        const int X = cell % n_out, Y = cell / n_out, sub = k / per_sub, kc = k % per_sub;
        const uint4 v = *reinterpret_cast<const uint4*>(in + cell_row(g, n_in, 2 * Y + (sub >> 1), 2 * X + (sub & 1)) * c_in + kc * 8);
        *reinterpret_cast<uint4*>(out + cell_row(g, n_out, Y, X) * (4 * c_in) + k * 8) = v;
    }
}

// nn.AvgPool2d(kernel_size=3, stride=2, padding=1) (muzero_atari_network.py:16,18): the padding counts in the divisor (always 9);
// the zero halo rows of the layout are that padding. fp32 sum of the fp16 inputs, fp16 out.
__global__ void avgpool_kernel(const __half* __restrict__ in, __half* __restrict__ out, int batch, int n_in, int c)
{
    const int n_out = n_in / 2, per_cell = c / 8, n1 = n_in + 1;
    const size_t total = static_cast<size_t>(batch) * n_out * n_out * per_cell;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % per_cell), cell = static_cast<int>((i / per_cell) % (n_out * n_out)), g = static_cast<int>(i / (static_cast<size_t>(per_cell) * n_out * n_out));
        const int X = cell % n_out, Y = cell / n_out;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j] = 0.0f; }
        const __half* centre = in + cell_row(g, n_in, 2 * Y, 2 * X) * c + k * 8;
#pragma unroll
        for (int ky = -1; ky <= 1; ++ky) {
#pragma unroll
            for (int kx = -1; kx <= 1; ++kx) {
                if (2 * Y + ky >= n_in || 2 * X + kx >= n_in) { continue; } // below / right of the map (never with an even n_in; kept for safety)
                if (g == 0 && Y == 0 && X == 0 && ky < 0 && kx < 0) { continue; }      // the one tap that would lie before the first row of the buffer
                const uint4 v = *reinterpret_cast<const uint4*>(centre + static_cast<ptrdiff_t>(ky * n1 + kx) * c); // x = -1 is the previous row's zero column
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
                }
            }
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) { oh[j] = __floats2half2_rn(acc[2 * j] / 9.0f, acc[2 * j + 1] / 9.0f); }
        *reinterpret_cast<uint4*>(out + cell_row(g, n_out, Y, X) * c + k * 8) = o;
    }
}

// utils::invertValue (utils/utils.h:102-108)
__device__ __forceinline__ float invert_value(float value)
{
    const float epsilon = 0.001f;
    const float sign_value = (value > 0.0f ? 1.0f : (value == 0.0f ? 0.0f : -1.0f));
    const float r = (sqrtf(1.0f + 4.0f * epsilon * (fabsf(value) + 1.0f + epsilon)) - 1.0f) / (2.0f * epsilon);
    return sign_value * (r * r - 1.0f);
}

// ---------------------------------------------------------------------------------------------
// Heads of the Atari network. PolicyNetwork (network_unit.py:26-42): conv1x1-BN-ReLU-fc, softmax. DiscreteValueNetwork
// (network_unit.py:68-87): conv1x1-BN-ReLU-fc1-ReLU-fc2 -> 601 logits; MuZeroNetwork::forward (muzero_network.h:157-171) takes the
// softmax, the expectation sum_i p_i * (i - 300) and utils::invertValue. The value and policy heads read the SCALED hidden state,
// the reward head the dynamics output before scaling (muzero_atari_network.py:53-54,176-178). Five launches per evaluation:
//   fc_gemm_kernel         the 1x1 convolutions of all three heads as ONE GEMM over the tower's output rows (unscaled)
//   hidden_planes_kernel   per board: min / max scaling of the hidden state (scale_hidden_state, :185-193) into the node's hidden
//                          slot; bias + ReLU on the convolutions (rescaled by linearity for the value / policy planes) -> FC inputs
//   fc_gemm_kernel x 2     fc1 (+ ReLU) and fc2 of the value and reward heads as batched GEMMs over the boards on tensor cores
//                          (mma.sync m16n8k16, fp16 in / fp32 accumulate; M = boards: far too small for a tcgen05 pipeline)
//   discrete_finalize_kernel  per board: softmax + expectation + invertValue of both heads, policy fc + softmax
// ---------------------------------------------------------------------------------------------
struct PlanesParams {
    const __half* act;   // [batch * slots][c] tower output rows (unscaled)
    const float* raw;    // [batch * slots][ld_raw] 1x1 convolutions of all heads on the UNSCALED rows (fc_gemm_kernel): policy planes, value planes, reward planes
    int ld_raw;
    __half* hid;         // [batch][num_slots][hw][c] hidden states of the search
    const int32_t* slot; // [batch] slot of this evaluation
    int n, slots, c, c_real, num_slots;
    const float* bias;   // [planes] folded BatchNorm shifts of the 1x1 convolutions, in plane order
    const float* wsum;   // [planes] sum over the channels of every plane's (fp16-rounded) weights
    int pol_ch, hc, with_reward;
    __half* a_val;       // [m_pad][k_pad] flattened value planes (plane-major, as x.view(-1, hw * hc) of an NCHW tensor)
    __half* a_rew;
    float* pol_planes;   // [batch][pol_ch * hw]
    int k_pad;
};

// per board: min / max of the hidden state, the scaled state into the node's hidden slot, and the heads' planes from the raw 1x1
// convolutions: a convolution is linear, so conv((x - min) / scale) = (conv(x) - min * sum(w)) / scale for the value and policy
// planes (which read the scaled state); the reward planes read the unscaled one
__global__ void __launch_bounds__(256) hidden_planes_kernel(const PlanesParams p)
{
    __shared__ float red_mn[8], red_mx[8];
    const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, hw = p.n * p.n, n1 = p.n + 1;
    const __half* rows = p.act + static_cast<size_t>(g) * p.slots * p.c;
    const int per_cell = p.c / 8, real_per_cell = p.c_real / 8; // 16-byte chunks (c_real is a multiple of 16)
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (int i = tid; i < hw * real_per_cell; i += blockDim.x) {
        const int cell = i / real_per_cell, k = i - cell * real_per_cell;
        const uint4 v = *reinterpret_cast<const uint4*>(rows + static_cast<size_t>((cell / p.n + 1) * n1 + cell % p.n) * p.c + 8 * k);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            mn = fminf(mn, fminf(f.x, f.y)), mx = fmaxf(mx, fmaxf(f.x, f.y));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)), mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { red_mn[warp] = mn, red_mx[warp] = mx; }
    __syncthreads();
    mn = red_mn[0], mx = red_mx[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) { mn = fminf(mn, red_mn[i]), mx = fmaxf(mx, red_mx[i]); }
    float scale = mx - mn;
    if (scale < 1e-5f) { scale += 1e-5f; }
    // the scaled state -> the node's hidden slot (fp16, padded channels zero): what the next recurrent inference gathers
    __half* dst = p.hid + (static_cast<size_t>(g) * p.num_slots + p.slot[g]) * hw * p.c;
    for (int i = tid; i < hw * per_cell; i += blockDim.x) {
        const int cell = i / per_cell, k = i - cell * per_cell;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (k < real_per_cell) {
            const uint4 v = *reinterpret_cast<const uint4*>(rows + static_cast<size_t>((cell / p.n + 1) * n1 + cell % p.n) * p.c + 8 * k);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
            __half2* o = reinterpret_cast<__half2*>(&out);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                o[j] = __floats2half2_rn((f.x - mn) / scale, (f.y - mn) / scale);
            }
        }
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(i) * 8) = out;
    }
    // planes: folded-BN bias + ReLU on the (rescaled) raw convolutions
    const int planes = p.pol_ch + p.hc + (p.with_reward ? p.hc : 0);
    for (int o = tid; o < planes * hw; o += blockDim.x) {
        const int plane = o / hw, cell = o - plane * hw;
        float v = p.raw[(static_cast<size_t>(g) * p.slots + (cell / p.n + 1) * n1 + cell % p.n) * p.ld_raw + plane];
        if (plane < p.pol_ch + p.hc) { v = (v - mn * p.wsum[plane]) / scale; }
        v = fmaxf(v + p.bias[plane], 0.0f);
        if (plane < p.pol_ch) {
            p.pol_planes[static_cast<size_t>(g) * p.pol_ch * hw + o] = v;
        } else if (plane < p.pol_ch + p.hc) {
            p.a_val[static_cast<size_t>(g) * p.k_pad + (o - p.pol_ch * hw)] = __float2half_rn(v);
        } else {
            p.a_rew[static_cast<size_t>(g) * p.k_pad + (o - (p.pol_ch + p.hc) * hw)] = __float2half_rn(v);
        }
    }
}

// out[m][n] = act(sum_k a[m][k] * w[n][k] + bias[n]) for up to two independent problems (blockIdx.z): torch's Linear with its
// weight in its own [out][in] layout. 64 x 64 tile per CTA of 4 warps (16 rows each), K in steps of 64 through a two-stage
// cp.async pipeline, mma.sync.m16n8k16 fp16 -> fp32. m, n, k are padded to multiples of 64 by the caller (zero weights / rows).
struct FcGemmParams {
    const __half* a[2];
    const __half* w[2];
    const float* bias[2];
    void* out[2];
    int n[2], k[2], lda[2], ldw[2], ldc[2];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mznn::smem_u32(smem)), "l"(gmem) : "memory");
}

constexpr int FC_TILE = 64, FC_LD = FC_TILE + 8, FC_STAGES = 4; // +8 halves: rows 144 bytes apart, fragment loads hit distinct banks
constexpr int FC_SMEM = FC_STAGES * 2 * FC_TILE * FC_LD * 2;   // bytes of dynamic shared memory

template <bool HALF_RELU_OUT>
__global__ void __launch_bounds__(128) fc_gemm_kernel(const FcGemmParams p)
{
    constexpr int T = FC_TILE, LD = FC_LD;
    extern __shared__ __align__(16) uint8_t fc_smem[];
    typedef __half Tile[T][LD];
    Tile* As = reinterpret_cast<Tile*>(fc_smem);
    Tile* Ws = As + FC_STAGES;
    const int h = blockIdx.z, n0 = blockIdx.x * T, m0 = blockIdx.y * T;
    if (n0 >= p.n[h]) { return; }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const __half* a = p.a[h] + static_cast<size_t>(m0) * p.lda[h];
    const __half* w = p.w[h] + static_cast<size_t>(n0) * p.ldw[h];
    const int kt = p.k[h] / T;
    auto load_tile = [&](int it) { // always commits a group, so that the group count tracks the tile count
        if (it < kt) {
            const int buf = it % FC_STAGES, k0 = it * T;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int chunk = tid + i * 128, row = chunk >> 3, col = (chunk & 7) * 8;
                cp_async16(&As[buf][row][col], a + static_cast<size_t>(row) * p.lda[h] + k0 + col);
                cp_async16(&Ws[buf][row][col], w + static_cast<size_t>(row) * p.ldw[h] + k0 + col);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
#pragma unroll
    for (int i = 0; i < FC_STAGES - 1; ++i) { load_tile(i); }
    for (int it = 0; it < kt; ++it) {
        const int buf = it % FC_STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(FC_STAGES - 2) : "memory"); // tile `it` has landed
        __syncthreads();                                                         // ... for everybody, and tile it - 1 is no longer being read
        load_tile(it + FC_STAGES - 1);                                           // into the buffer tile it - 1 used
#pragma unroll
        for (int kk = 0; kk < T; kk += 16) {
            const int r = warp * 16 + g;
            const uint32_t a0 = *reinterpret_cast<const uint32_t*>(&As[buf][r][kk + 2 * t]), a1 = *reinterpret_cast<const uint32_t*>(&As[buf][r + 8][kk + 2 * t]);
            const uint32_t a2 = *reinterpret_cast<const uint32_t*>(&As[buf][r][kk + 2 * t + 8]), a3 = *reinterpret_cast<const uint32_t*>(&As[buf][r + 8][kk + 2 * t + 8]);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ws[buf][nt * 8 + g][kk + 2 * t]), b1 = *reinterpret_cast<const uint32_t*>(&Ws[buf][nt * 8 + g][kk + 2 * t + 8]);
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            }
        }
    }
    const int row = m0 + warp * 16 + g;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = n0 + nt * 8 + 2 * t;
        const float b0 = (p.bias[h] ? p.bias[h][col] : 0.0f), b1 = (p.bias[h] ? p.bias[h][col + 1] : 0.0f);
        if (HALF_RELU_OUT) {
            __half* o = static_cast<__half*>(p.out[h]);
            *reinterpret_cast<__half2*>(o + static_cast<size_t>(row) * p.ldc[h] + col) = __floats2half2_rn(fmaxf(acc[nt][0] + b0, 0.0f), fmaxf(acc[nt][1] + b1, 0.0f));
            *reinterpret_cast<__half2*>(o + static_cast<size_t>(row + 8) * p.ldc[h] + col) = __floats2half2_rn(fmaxf(acc[nt][2] + b0, 0.0f), fmaxf(acc[nt][3] + b1, 0.0f));
        } else {
            float* o = static_cast<float*>(p.out[h]);
            *reinterpret_cast<float2*>(o + static_cast<size_t>(row) * p.ldc[h] + col) = make_float2(acc[nt][0] + b0, acc[nt][1] + b1);
            *reinterpret_cast<float2*>(o + static_cast<size_t>(row + 8) * p.ldc[h] + col) = make_float2(acc[nt][2] + b0, acc[nt][3] + b1);
        }
    }
}

struct FinalizeParams {
    const float* lg_val; // [m_pad][ld] logits of the value head's 601 bins
    const float* lg_rew; // or null
    int ld, dv;
    const float* pol_planes; // [batch][pol_ch * hw]
    const float *w_pf, *b_pf; // [pol_ch * hw][actions] transposed, [actions]
    int pol_in, actions;
    float *value, *reward, *policy, *logits;
};

// warp 0: value, warp 1: reward (softmax over the bins, expectation of (i - dv / 2), invertValue); warps 2-3: policy fc + softmax
__global__ void __launch_bounds__(128) discrete_finalize_kernel(const FinalizeParams p)
{
    const int g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
        const float* l = (warp == 0 ? p.lg_val : p.lg_rew);
        if (!l) { return; }
        l += static_cast<size_t>(g) * p.ld;
        float mx = -3.402823466e+38f;
        for (int i = lane; i < p.dv; i += 32) { mx = fmaxf(mx, l[i]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        float sum = 0.0f, ex = 0.0f;
        const int shift = p.dv / 2;
        for (int i = lane; i < p.dv; i += 32) {
            const float e = expf(l[i] - mx);
            sum += e, ex = fmaf(e, static_cast<float>(i - shift), ex);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o), ex += __shfl_xor_sync(0xffffffffu, ex, o); }
        if (lane == 0) { (warp == 0 ? p.value : p.reward)[g] = invert_value(ex / sum); }
        return;
    }
    if (warp == 2) { // (actions <= 32 for every Atari game: 18)
        float acc = -3.402823466e+38f;
        if (lane < p.actions) {
            acc = p.b_pf[lane];
            for (int i = 0; i < p.pol_in; ++i) { acc = fmaf(p.pol_planes[static_cast<size_t>(g) * p.pol_in + i], __ldg(p.w_pf + static_cast<size_t>(i) * p.actions + lane), acc); }
        }
        float mx = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        const float e = (lane < p.actions ? expf(acc - mx) : 0.0f);
        float sum = e;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); }
        if (lane < p.actions) {
            p.logits[static_cast<size_t>(g) * p.actions + lane] = acc;
            p.policy[static_cast<size_t>(g) * p.actions + lane] = e / sum;
        }
    }
}

} // namespace mzat
