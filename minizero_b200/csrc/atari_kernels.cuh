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
// softmax, the expectation sum_i p_i * (i - 300) and utils::invertValue: all fused here, only the scalar leaves the kernel.
// One launch computes one discrete head (the value head together with the policy head on the scaled hidden state; the reward head
// on the dynamics output BEFORE scaling, muzero_atari_network.py:53-54,176-178). BPC boards per CTA share every weight they read.
// ---------------------------------------------------------------------------------------------
struct DiscreteHeadParams {
    const __half* act; // [batch * slots][c] rows of an n x n map
    int c, n, slots, batch;
    int do_policy;
    const float* w_pc; // [pol_ch][c] policy conv (BN folded)      b_pc [pol_ch]
    const float* b_pc;
    const float* w_pf; // [pol_ch * hw][actions] transposed          b_pf [actions]
    const float* b_pf;
    int pol_ch, actions;
    float* policy;     // [batch][actions]
    float* logits;     // [batch][actions]
    const float* w_dc; // [hc][c] conv of the discrete head (BN folded)   b_dc [hc]
    const float* b_dc;
    const float* w_d1; // [hc * hw][vh] transposed                        b_d1 [vh]
    const float* b_d1;
    const float* w_d2; // [vh][dv] transposed                             b_d2 [dv]
    const float* b_d2;
    int hc, vh, dv;
    float* out;        // [batch] invertValue(expectation)
};

template <int BPC>
__global__ void __launch_bounds__(512) discrete_head_kernel(const DiscreteHeadParams p)
{
    extern __shared__ float sm[];
    const int hw = p.n * p.n, n1 = p.n + 1;
    const int np = (p.do_policy ? p.pol_ch : 0) + p.hc; // planes per board: policy planes first
    float* wc = sm;                         // [np][c]
    float* planes = wc + np * p.c;          // [BPC][np * hw]
    float* hid = planes + BPC * np * hw;    // [BPC][vh]
    float* lg = hid + BPC * p.vh;           // [BPC][max(dv, actions)]
    float* partial = lg + BPC * (p.dv > p.actions ? p.dv : p.actions); // [parts][BPC][vh]
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
    const int g0 = blockIdx.x * BPC;
    const int pol_planes = (p.do_policy ? p.pol_ch : 0);
    for (int i = tid; i < np * p.c; i += nthr) { wc[i] = (i < pol_planes * p.c ? p.w_pc[i] : p.w_dc[i - pol_planes * p.c]); }
    __syncthreads();
    // 1x1 convolutions: one warp per (board, cell); lane l owns channel pairs {2l + 64i}; output planes in groups of 6
    const int npair = p.c / 64;
    for (int bc = warp; bc < BPC * hw; bc += nwarp) {
        const int b = bc / hw, cell = bc - b * hw;
        if (g0 + b >= p.batch) { continue; }
        const __half2* row = reinterpret_cast<const __half2*>(p.act + (static_cast<size_t>(g0 + b) * p.slots + (cell / p.n + 1) * n1 + cell % p.n) * p.c);
        for (int o0 = 0; o0 < np; o0 += 6) {
            float acc[6];
#pragma unroll
            for (int o = 0; o < 6; ++o) { acc[o] = 0.0f; }
            for (int i = 0; i < npair; ++i) {
                const float2 a = __half22float2(row[lane + 32 * i]);
                const float* wp = wc + 2 * lane + 64 * i;
#pragma unroll
                for (int o = 0; o < 6; ++o) {
                    if (o0 + o < np) {
                        const float2 wv = *reinterpret_cast<const float2*>(wp + (o0 + o) * p.c);
                        acc[o] = fmaf(a.x, wv.x, fmaf(a.y, wv.y, acc[o]));
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 6; ++o) {
                float v = acc[o];
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, sft); }
                if (lane == 0 && o0 + o < np) {
                    const int pl = o0 + o;
                    const float bias = (pl < pol_planes ? p.b_pc[pl] : p.b_dc[pl - pol_planes]);
                    planes[(b * np + pl) * hw + cell] = fmaxf(v + bias, 0.0f);
                }
            }
        }
    }
    __syncthreads();
    // fc1 of the discrete head: vh outputs over hc * hw inputs (transposed weights, coalesced across threads), the input range split
    // over `parts` thread groups; every weight feeds BPC boards
    const int nin = p.hc * hw;
    const int parts = (nthr / p.vh > 0 ? nthr / p.vh : 1);
    for (int t = tid; t < parts * p.vh; t += nthr) {
        const int part = t / p.vh, o = t - part * p.vh;
        const int i0 = (nin * part) / parts, i1 = (nin * (part + 1)) / parts;
        float acc[BPC];
#pragma unroll
        for (int b = 0; b < BPC; ++b) { acc[b] = 0.0f; }
        const float* wp = p.w_d1 + o;
#pragma unroll 8
        for (int i = i0; i < i1; ++i) {
            const float w = __ldg(wp + static_cast<size_t>(i) * p.vh);
#pragma unroll
            for (int b = 0; b < BPC; ++b) { acc[b] = fmaf(planes[(b * np + pol_planes) * hw + i], w, acc[b]); }
        }
#pragma unroll
        for (int b = 0; b < BPC; ++b) { partial[(part * BPC + b) * p.vh + o] = acc[b]; }
    }
    __syncthreads();
    for (int t = tid; t < BPC * p.vh; t += nthr) {
        const int b = t / p.vh, o = t - b * p.vh;
        float acc = p.b_d1[o];
        for (int q = 0; q < parts; ++q) { acc += partial[(q * BPC + b) * p.vh + o]; }
        hid[b * p.vh + o] = fmaxf(acc, 0.0f);
    }
    __syncthreads();
    // fc2: dv logits over vh inputs
    for (int o = tid; o < p.dv; o += nthr) {
        float acc[BPC];
#pragma unroll
        for (int b = 0; b < BPC; ++b) { acc[b] = 0.0f; }
        const float* wp = p.w_d2 + o;
#pragma unroll 8
        for (int i = 0; i < p.vh; ++i) {
            const float w = __ldg(wp + static_cast<size_t>(i) * p.dv);
#pragma unroll
            for (int b = 0; b < BPC; ++b) { acc[b] = fmaf(hid[b * p.vh + i], w, acc[b]); }
        }
        const float bias = p.b_d2[o];
#pragma unroll
        for (int b = 0; b < BPC; ++b) { lg[b * p.dv + o] = acc[b] + bias; }
    }
    __syncthreads();
    // softmax, expectation over the bins (i - dv / 2), invertValue: one warp per board
    if (warp < BPC && g0 + warp < p.batch) {
        const float* l = lg + warp * p.dv;
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
        if (lane == 0) { p.out[g0 + warp] = invert_value(ex / sum); }
    }
    if (!p.do_policy) { return; }
    __syncthreads();
    // policy fc + softmax (tiny: pol_ch * hw inputs, `actions` outputs)
    for (int t = tid; t < BPC * p.actions; t += nthr) {
        const int b = t / p.actions, o = t - b * p.actions;
        float acc = p.b_pf[o];
        for (int i = 0; i < p.pol_ch * hw; ++i) { acc = fmaf(planes[b * np * hw + i], __ldg(p.w_pf + static_cast<size_t>(i) * p.actions + o), acc); }
        lg[b * p.actions + o] = acc;
    }
    __syncthreads();
    if (warp < BPC && g0 + warp < p.batch) {
        const float* l = lg + warp * p.actions;
        float mx = -3.402823466e+38f;
        for (int a = lane; a < p.actions; a += 32) { mx = fmaxf(mx, l[a]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        float sum = 0.0f;
        for (int a = lane; a < p.actions; a += 32) { sum += expf(l[a] - mx); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); }
        const float inv = 1.0f / sum;
        for (int a = lane; a < p.actions; a += 32) {
            p.logits[static_cast<size_t>(g0 + warp) * p.actions + a] = l[a];
            p.policy[static_cast<size_t>(g0 + warp) * p.actions + a] = expf(l[a] - mx) * inv;
        }
    }
}

} // namespace mzat
