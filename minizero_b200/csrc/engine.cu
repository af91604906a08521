// libmzb200.so — engine state, kernels launch logic and the C ABI of include/mz_b200.h.
// No CPU path: every entry point needs a CUDA device of compute capability 10.x.
#include "../../include/mz_b200.h"
#include "nn_kernels.cuh"
#include "atari_kernels.cuh"
#include "search_core.cuh"

#include <nvtx3/nvToolsExt.h> // header-only; ranges cost nothing unless a profiler injects its library
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

// Experiment switches (tile shapes, stage counts, debug counters ...) exist only in a library built with -DMZ_EXPERIMENT
// (MZ_BUILD_EXPERIMENT=1 python -c "import minizero_b200; minizero_b200.build_library(force=True)"); in the release build no
// environment variable can change what the library computes or how a benchmark runs.
#ifdef MZ_EXPERIMENT
const char* knob(const char* name) { return std::getenv(name); }
#else
const char* knob(const char*) { return nullptr; }
#endif

int fail(int code, const std::string& msg)
{
    g_error = msg;
    return code;
}

#define CUDA_OK(expr)                                                                                                        \
    do {                                                                                                                     \
        cudaError_t err__ = (expr);                                                                                          \
        if (err__ != cudaSuccess) { return fail(MZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); }       \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// kernels around search_core.cuh: one block per game (16 warps for the per-cycle tree step, one warp for the per-move kernels)
// ---------------------------------------------------------------------------------------------------
// NVTX range for the duration of a C-ABI call: a timeline (nsys / ncu --nvtx) shows search, move and model-load phases per engine
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

constexpr int STEP_AFTER = 1, STEP_BEFORE = 2;

constexpr int STEP_WARPS = 12; // warps per game in k_step: the previous path is re-evaluated level-parallel (search_core.cuh, mz_select). 12 warps at two blocks per
                               // SM leave 80 registers per thread (16 warps: 64, and 380 bytes of spill loads on the serial paths): tree step 37.0 -> 34.3 us at
                               // config 2, 73.3 -> 69.7 us at config 4 (profiles/r2_stepwarps_ab.log); 8 warps (114 registers, no spills) measure the same as 12

__global__ void __launch_bounds__(32 * STEP_WARPS, 2) k_step(const mz_dims d, const mz_state s, const int flags)
{
    __shared__ mz_scratch w;
    extern __shared__ uint64_t dyn_smem[]; // lvl_h [S + 2] 16 B | lvl_v [S + 2] 16 B | path_hashes [S + 2] u64 | sel [S + 2] i32 | q_warp [STEP_WARPS][A] f32
    const int g = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        w.lvl_h = reinterpret_cast<mz_hot*>(dyn_smem);
        w.lvl_v = reinterpret_cast<mz_vis*>(dyn_smem + 2 * (d.S + 2));
        w.path_hashes = dyn_smem + 4 * (d.S + 2);
        w.sel = reinterpret_cast<int32_t*>(w.path_hashes + (d.S + 2));
        w.q_warp = reinterpret_cast<float*>(w.sel + (d.S + 2));
    }
    __syncthreads();
    // the network kernel that follows may be launched while this grid is still running (programmatic dependent launch): it only sets
    // itself up and prefetches weights until griddepcontrol.wait lets it read this kernel's output
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (flags & STEP_AFTER) {
        const long long ta = clock64();
        mz_after_nn(d, s, g, &w, threadIdx.x, blockDim.x);
        if (s.dbg && threadIdx.x == 0) { s.dbg[(size_t)g * 16 + 4] += (unsigned long long)(clock64() - ta); }
    }
    if (flags & STEP_BEFORE) {
        __threadfence_block();
        __syncthreads();
        mz_before_nn(d, s, g, &w, lane, wid, STEP_WARPS);
    }
}

__global__ void __launch_bounds__(32) k_reset(const mz_dims d, const mz_state s, const int only_game)
{
    __shared__ mz_scratch w;
    const int g = (only_game >= 0 ? only_game : static_cast<int>(blockIdx.x)), lane = threadIdx.x; // one game: a grid of one block
    mz_game_reset(d, s, g, &w, lane);
}

__global__ void __launch_bounds__(32) k_play(const mz_dims d, const mz_state s, const int32_t* __restrict__ actions, int32_t* __restrict__ out, float* __restrict__ score)
{
    __shared__ mz_scratch w;
    const int g = blockIdx.x, lane = threadIdx.x;
    const int a = actions[g];
    if (a < 0) {
        if (lane == 0) { out[g * 4 + 0] = 0, out[g * 4 + 1] = 0, out[g * 4 + 2] = 0, out[g * 4 + 3] = s.root_meta[g * 4 + 0], score[g] = 0.0f; }
        return;
    }
    mz_play(d, s, g, a, &w, out + g * 4, score + g, lane);
}

// actor_select_action_by_count decided on the device: play the most visited root child of every game; a game that
// ends is reset in place when auto_reset is set (records are then the caller's loss: throughput runs only)
__global__ void __launch_bounds__(32) k_play_max_count(const mz_dims d, const mz_state s, int32_t* __restrict__ actions_out, int32_t* __restrict__ out,
                                                       float* __restrict__ score, const int auto_reset)
{
    __shared__ mz_scratch w;
    const int g = blockIdx.x, lane = threadIdx.x;
    int a;
    if (d.gumbel) { // GumbelZero::decideActionNode, gumbel_zero.cpp:61-66
        a = mz_root_gumbel_action(d, s, g, lane);
        a = __shfl_sync(MZ_FULL, a, 0);
    } else {
        a = mz_root_max_count_action(d, s, g, lane);
    }
    if (lane == 0) { actions_out[g] = a; }
    if (a < 0) {
        if (lane == 0) { out[g * 4 + 0] = 0, out[g * 4 + 1] = 0, out[g * 4 + 2] = 0, out[g * 4 + 3] = s.root_meta[g * 4 + 0], score[g] = 0.0f; }
        return;
    }
    mz_play(d, s, g, a, &w, out + g * 4, score + g, lane);
    __syncwarp();
    if (auto_reset && out[g * 4 + 1]) { mz_game_reset(d, s, g, &w, lane); }
}

// root child table of every game, children in stored order (MCTSNode getters, actor/mcts.h:44-52)
__global__ void __launch_bounds__(32) k_gather_roots(const mz_dims d, const mz_state s, float* __restrict__ info, int32_t* __restrict__ action, float* __restrict__ count,
                                                     float* __restrict__ mean, float* __restrict__ policy, float* __restrict__ logit, float* __restrict__ noise,
                                                     float* __restrict__ value)
{
    const int g = blockIdx.x, lane = threadIdx.x;
    const mz_hot* hot = s.hot + (size_t)g * d.NP;
    const mz_hot root = mz_load_hot(hot);
    const int nc = (int)(root.link >> MZ_LINK_SHIFT), fc = (int)(root.link & ((1u << MZ_LINK_SHIFT) - 1u));
    if (lane == 0) {
        info[g * 4 + 0] = __int_as_float(nc);
        info[g * 4 + 1] = root.count, info[g * 4 + 2] = root.mean, info[g * 4 + 3] = s.value[(size_t)g * d.NP];
    }
    for (int i = lane; i < d.A; i += 32) {
        const size_t o = (size_t)g * d.A + i;
        if (i < nc) {
            const mz_hot c = mz_load_hot(hot + fc + i);
            const size_t n = (size_t)g * d.NP + fc + i;
            action[o] = s.action[n], count[o] = c.count, mean[o] = c.mean, policy[o] = c.policy;
            logit[o] = s.logit[n], noise[o] = s.root_noise[o], value[o] = s.value[n];
        } else {
            action[o] = -1, count[o] = 0.0f, mean[o] = 0.0f, policy[o] = 0.0f, logit[o] = 0.0f, noise[o] = 0.0f, value[o] = 0.0f;
        }
    }
}

// rewards of the root children and the tree's value bounds (MCTSNode::getReward, MCTS::getTreeValueBound)
__global__ void __launch_bounds__(32) k_gather_root_rewards(const mz_dims d, const mz_state s, float* __restrict__ reward, int32_t* __restrict__ bsize, float* __restrict__ blo,
                                                            float* __restrict__ bhi)
{
    const int g = blockIdx.x, lane = threadIdx.x;
    const mz_hot root = mz_load_hot(s.hot + (size_t)g * d.NP);
    const int nc = (int)(root.link >> MZ_LINK_SHIFT), fc = (int)(root.link & ((1u << MZ_LINK_SHIFT) - 1u));
    for (int i = lane; i < d.A; i += 32) { reward[(size_t)g * d.A + i] = (i < nc && s.reward ? s.reward[(size_t)g * d.NP + fc + i] : 0.0f); }
    const mz_qb qb = mz_vb_bounds(d, s, g, lane);
    if (lane == 0) { bsize[g] = qb.n, blo[g] = (qb.n ? qb.lo : 0.0f), bhi[g] = (qb.n ? qb.hi : 0.0f); }
}

// AtariEnv::reset's first screen / AtariEnv::act's history update for every game with actions[g] > -2 (atari.cpp:52-56,82-85)
__global__ void __launch_bounds__(256) k_atari_observe(const mz_dims d, const mz_state s, const int32_t* __restrict__ actions, const uint8_t* __restrict__ frames_in)
{
    const int g = blockIdx.x;
    const int a = actions[g];
    if (a < -1) { return; } // uniform for the block
    const int slot = mz_atari_push(s, g, a, threadIdx.x);
    const uint4* src = reinterpret_cast<const uint4*>(frames_in + (size_t)g * MZ_ATARI_FRAME);
    uint4* dst = reinterpret_cast<uint4*>(s.at_frames + ((size_t)g * MZ_HIST + slot) * MZ_ATARI_FRAME);
    for (int i = threadIdx.x; i < MZ_ATARI_FRAME / 16; i += blockDim.x) { dst[i] = src[i]; }
}

// BaseEnvLoader::getFeatures (environment/base/base_env.h:235-241), the learner's feature reconstruction: sample g replays the first
// positions[g] actions of its record from the initial position and emits the planes of the position reached, rotated. One warp per
// sample; the planes land in the network-input rows of "game" g (the caller unpacks them).
__global__ void __launch_bounds__(32) k_replay_features(const mz_dims d, const mz_state s, const int32_t* __restrict__ actions, int max_len, const int32_t* __restrict__ positions,
                                                        const uint8_t* __restrict__ rotations, int n)
{
    __shared__ mz_scratch w;
    const int g = blockIdx.x, lane = threadIdx.x;
    if (g >= n) { return; }
    mz_env_reset(d, &w, lane);
    const int upto = (positions[g] < max_len ? positions[g] : max_len);
    for (int i = 0; i < upto; ++i) {
        const int a = actions[(size_t)g * max_len + i];
        if (a < 0) { break; } // record shorter than the position asked for: the final position (base_env.h:239)
        mz_env_act(d, s, &w, a, w.turn, lane);
    }
    mz_env_features(d, s, g, &w, rotations ? rotations[g] : 0, lane, 32);
}

// action ids along the selected path of every game (-1 padded): parity hook for the MuZero / Gumbel selection
__global__ void k_path_actions(const mz_dims d, const mz_state s, int32_t* __restrict__ out)
{
    const int g = blockIdx.x;
    const int len = s.path_len[g];
    const int32_t* path = s.path + (size_t)g * (d.S + 2);
    for (int i = threadIdx.x; i < d.S + 2; i += blockDim.x) { out[(size_t)g * (d.S + 2) + i] = (i < len ? (int32_t)s.action[(size_t)g * d.NP + path[i]] : -1); }
}

__global__ void __launch_bounds__(32) k_gumbel_best(const mz_dims d, const mz_state s, int32_t* __restrict__ out)
{
    const int a = mz_root_gumbel_action(d, s, blockIdx.x, threadIdx.x);
    if (threadIdx.x == 0) { out[blockIdx.x] = a; }
}

// ---------------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct ConvLayer {
    int cin = 0, cout = 0, relu = 1;
    size_t w_off = 0, b_off = 0; // offsets into the blob
    CUtensorMap map_w;
    CUtensorMap map_w_mc; // box = 1 / conv_cluster of the weight tile (multicast slices)
    CUtensorMap map_w_tower; // box = half of the fused tower's output-channel tile (tower_bn / 2 rows)
    // ConvStage layers only
    int cin_off = 0, tap_mask = 0x1ff;
    int in_buf = 0, out_buf = 0, res_buf = -2; // activation buffer indices of the stage; -1 = the stage's input rows, -2 = none
};

// A run of 3x3 conv layers at one resolution executed as ONE launch of the fused tower kernel: the Atari representation network
// changes resolution three times (muzero_atari_network.py:7-40), each resolution is a stage with its own geometry and buffers.
struct ConvStage {
    int n = 0, slots = 0, rows_alloc = 0, rows_ext = 0, cout = 0, cin_in = 0, cin_max = 0, stages = 8;
    __half* in = nullptr; // [rows_alloc][cin_in] rows written by the kernel before the stage, or null when the input is act[0]
    __half* act[3] = {nullptr, nullptr, nullptr};
    CUtensorMap map_in_ext, map_act_ext[3];
    std::vector<ConvLayer> convs;
    mznn::TowerParams* params = nullptr;
    int* d_done = nullptr;
    int out_buf = 0;
};

// one residual tower: AlphaZero's, or MuZero's representation (0) / dynamics (1) network
struct NetTower {
    std::vector<ConvLayer> convs;
    std::string prefix;                  // state_dict prefix of its modules
    int cin0_real = 0, cin0 = 0;         // real / padded input channels of its first conv
    void* in = nullptr;                  // fp16 input rows [rows_alloc][cin0]
    CUtensorMap map_in, map_in_ext;      // box = one row tile / the resident block
    CUtensorMap map_in_wide;             // box = half of the wide tower's input block (rows_ext_wide / 2 rows)
    int done_count = 0;                  // completion counters of one launch (layers x row groups of the kernel variant in use)
    mznn::TowerParams* params = nullptr; // host copy of the fused-tower launch parameters (conv_mode 3)
    int* d_done = nullptr;
    int out_buf = 0;                     // index of the activation buffer holding its output
    bool has_stem = true;                // false: the tower starts with a residual block on act[0] (last stage of the Atari representation network)
};

struct Blob {
    size_t size = 0;
    size_t take(size_t bytes)
    {
        const size_t off = size;
        size += (bytes + 255) & ~static_cast<size_t>(255);
        return off;
    }
};

} // namespace

struct SearchGraph {
    cudaGraphExec_t exec;
    int64_t kernels; // kernel nodes of the graph, counted at capture
};

struct mz_engine {
    mz_config cfg{};
    mz_dims d{};
    mz_state s{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::vector<void*> allocs;
    int64_t launches = 0;

    // per-search inputs / outputs
    uint8_t* d_rot_all = nullptr; // [(S+1)][B]
    float* d_noise = nullptr;     // [B][A]
    float* d_bias_table = nullptr;
    double* d_sqrt_table = nullptr;
    uint64_t* d_keys = nullptr;
    int32_t* d_actions = nullptr;
    int32_t* d_play_out = nullptr;
    float* d_play_score = nullptr;
    float* d_root_info = nullptr;
    int32_t* d_root_action = nullptr;
    float* d_root_f[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* d_feat_f32 = nullptr; // [B][C*H*W] staging for the NCHW parity hooks
    bool noise_enabled = false, rot_enabled = false;

    // network
    bool dims_set = false, net_ready = false;
    mz_net_dims nd{};
    int cpad = 0, pol_ch = 0, rows_alloc = 0, bn_tile = 0;
    std::map<std::string, std::vector<float>> tensors;
    NetTower tw[2];
    int num_towers = 1;
    Blob blob;
    uint8_t* d_blob = nullptr;
    size_t off_head[10] = {0};
    __half* act[3] = {nullptr, nullptr, nullptr};
    CUtensorMap map_act[3];
    CUtensorMap map_act_ext[3]; // same buffers, box = the resident block of conv3x3_resident_kernel
    CUtensorMap map_act_wide[3]; // same buffers, box = half of the wide tower's input block
    CUtensorMap map_act_store[3]; // same buffers as the targets of the towers' epilogues (TMA stores): box = 32 channels x 32 rows, 64-byte swizzle
    int tower_rot_override = -1; // MZ_TOWER_ROT (experiment builds)
    bool tower_coop = false;     // the fused tower is launched cooperatively (opt-in: mz_set_tower_cooperative)
    int think_steps = 0;         // batched steps the last think() search took
    int think_trees = 0;         // think mode (mz_config.think_batch_size > 1): number of trees; d.B = think_trees * d.think_k lanes
    bool tower_wide = false;    // conv_tower_wide_kernel (two row tiles per CTA) instead of conv_tower_kernel
    int rows_ext_wide = 0, tower_wide_stages = 8, tower_epi_bufs = 2;
    // MuZero
    float* d_hidden_f32 = nullptr;   // [B][Ch * H * W] staging of the parity hooks
    int32_t* d_path_actions = nullptr; // [B][S + 2]
    int cin_max = 0;
    int tower_bn = 128;
    int conv_mode = 1, rows_ext = 0, base_off_mode = 0, num_sms = 148, krot = 0, conv_cluster = 1, conv_pdl = 0, tower_sms = 148, tower_stages = 8, tower_pdl = 0;
    encode_tiled_fn encode = nullptr;

    // Atari MuZero (MZ_GAME_ATARI)
    bool atari = false;
    ConvStage ast[3];              // representation stages at 48 x 48, 24 x 24, 12 x 12 (the 6 x 6 stage is tw[0], the dynamics network tw[1])
    int at_c1 = 0;                 // padded channels of the first stage (num_hidden_channels / 2)
    size_t off_dhead[3][6] = {{0}}; // value / reward heads: conv w (fp32), conv b, fc1 w (fp16 [out_pad][in_pad]), fc1 b, fc2 w (fp16), fc2 b; [2] = policy: conv w, conv b, fc w^T, fc b (fp32)
    int dh_planes = 0;             // planes of a discrete head: ceil(discrete_value_size / 36)
    int dh_k1 = 0, dh_n1[2] = {0, 0}, dh_n2 = 0, dh_mpad = 0; // padded GEMM sizes of the heads' FC layers: fc1 in, fc1 out (value / reward), fc2 out, boards
    __half *d_a_head[2] = {nullptr, nullptr}, *d_h_head[2] = {nullptr, nullptr}; // FC inputs / hidden activations [mpad][k]
    float* d_l_head[2] = {nullptr, nullptr};                                       // bin logits [mpad][n2]
    float* d_pol_planes = nullptr;
    float* d_planes_raw = nullptr;         // [rows_alloc][64] 1x1 convolutions of the three heads on the unscaled tower output
    size_t off_planes[3] = {0, 0, 0};      // combined 1x1 conv weights fp16 [64][cpad] (policy, value, reward planes), folded biases [64], channel sums of the weights [64]
    uint8_t* d_at_frames_in = nullptr; // [B][3][96][96] staging of mz_atari_observe
    float* d_root_reward = nullptr;    // [B][A]
    int32_t* d_bound_size = nullptr;
    float *d_bound_lo = nullptr, *d_bound_hi = nullptr;
    float* d_planes_f32 = nullptr;     // [B][32][96][96] parity hooks (allocated on first use)
    int64_t memsets = 0;               // memset nodes issued (graph accounting)

    // graphs keyed by (num_evals, noise, rotations)
    std::map<int, SearchGraph> graphs;

    template <class T>
    int dalloc(T** p, size_t n)
    {
        CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
        CUDA_OK(cudaMemsetAsync(*p, 0, n * sizeof(T), stream));
        allocs.push_back(*p);
        return MZ_OK;
    }
};

namespace {

int make_map_2d(mz_engine* e, CUtensorMap* map, void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * sizeof(__half)};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = e->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { return fail(MZ_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r))); }
    return MZ_OK;
}

template <int BN, int STAGES>
int launch_conv(mz_engine* e, const CUtensorMap& in, const ConvLayer& L, __half* out, const __half* residual)
{
    using Smem = mznn::ConvSmem<BN, STAGES>;
    mznn::ConvParams p;
    p.out = out, p.residual = residual, p.bias = reinterpret_cast<const float*>(e->d_blob + L.b_off);
    p.rows_valid = e->d.B * e->d.slots, p.n1 = e->d.N + 1, p.slots = e->d.slots, p.cin = L.cin, p.cout = L.cout, p.relu = L.relu, p.krot = e->krot;
    dim3 grid(e->rows_alloc / mznn::BM, L.cout / BN);
    mznn::conv3x3_tcgen05_kernel<BN, STAGES><<<grid, mznn::CONV_THREADS, Smem::TOTAL, e->stream>>>(in, L.map_w, p);
    e->launches++;
    return MZ_OK;
}

template <int BN, int STAGES>
size_t resident_smem(const mz_engine* e, int cin)
{
    return static_cast<size_t>(cin / mznn::BK) * e->rows_ext * 128 + static_cast<size_t>(STAGES) * BN * mznn::BK * 2 + (2 * STAGES + 6) * 8 + 16 + 1024;
}

template <int BN, int STAGES, int CL>
int launch_conv_resident(mz_engine* e, const CUtensorMap& in_ext, const ConvLayer& L, __half* out, const __half* residual)
{
    mznn::ConvResParams rp;
    mznn::ConvParams& p = rp.c;
    p.out = out, p.residual = residual, p.bias = reinterpret_cast<const float*>(e->d_blob + L.b_off);
    p.rows_valid = e->d.B * e->d.slots, p.n1 = e->d.N + 1, p.slots = e->d.slots, p.cin = L.cin, p.cout = L.cout, p.relu = L.relu, p.krot = e->krot;
    rp.rows_ext = e->rows_ext, rp.halo = e->d.N + 2, rp.num_mtiles = e->rows_alloc / mznn::BM, rp.base_off_mode = e->base_off_mode;
    const int units = ((rp.num_mtiles + CL - 1) / CL) * (L.cout / BN);
    int clusters = e->num_sms / CL;
    if (units < clusters) { clusters = units; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * CL), cfg.blockDim = dim3(mznn::CONV_THREADS), cfg.dynamicSmemBytes = resident_smem<BN, STAGES>(e, L.cin), cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    const CUtensorMap& wmap = (CL == 1 ? L.map_w : L.map_w_mc);
    CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv3x3_resident_kernel<BN, STAGES, CL>, in_ext, wmap, rp));
    e->launches++;
    return MZ_OK;
}

template <int BN, int STAGES>
int launch_conv_pair(mz_engine* e, const CUtensorMap& in_ext, const ConvLayer& L, __half* out, const __half* residual)
{
    mznn::ConvResParams rp;
    mznn::ConvParams& p = rp.c;
    p.out = out, p.residual = residual, p.bias = reinterpret_cast<const float*>(e->d_blob + L.b_off);
    p.rows_valid = e->d.B * e->d.slots, p.n1 = e->d.N + 1, p.slots = e->d.slots, p.cin = L.cin, p.cout = L.cout, p.relu = L.relu, p.krot = 0;
    rp.rows_ext = e->rows_ext, rp.halo = e->d.N + 2, rp.num_mtiles = e->rows_alloc / mznn::BM, rp.base_off_mode = 0;
    const int units = ((rp.num_mtiles + 1) / 2) * (L.cout / BN);
    int clusters = e->num_sms / 2;
    if (units < clusters) { clusters = units; }
    const size_t smem = 2 * static_cast<size_t>(L.cin / mznn::BK) * e->rows_ext * 128 + static_cast<size_t>(STAGES) * (BN / 2) * mznn::BK * 2 + (2 * STAGES + 8) * 8 + 16 + 1024;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 2), cfg.blockDim = dim3(mznn::CONV_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = e->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; // the kernel orders itself with griddepcontrol.wait
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = (e->conv_pdl ? 2 : 1);
    CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv3x3_pair_kernel<BN, STAGES>, in_ext, L.map_w_mc, rp));
    e->launches++;
    return MZ_OK;
}

int configure_conv_kernels()
{
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_tcgen05_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mznn::ConvSmem<64, 4>::TOTAL));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_tcgen05_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mznn::ConvSmem<128, 3>::TOTAL));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_tcgen05_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mznn::ConvSmem<256, 4>::TOTAL));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_resident_kernel<64, 6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_resident_kernel<128, 9, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_resident_kernel<128, 9, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_resident_kernel<128, 9, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv3x3_pair_kernel<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<128, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<128, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<128, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<128, 5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<128, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<256, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_kernel<256, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_wide_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_wide_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_wide_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(mznn::conv_tower_wide_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    return MZ_OK;
}

int conv(mz_engine* e, const CUtensorMap& in, const CUtensorMap& in_ext, const ConvLayer& L, __half* out, const __half* residual)
{
    if (e->conv_mode == 2) { return launch_conv_pair<128, 8>(e, in_ext, L, out, residual); }
    if (e->conv_mode == 1) {
        if (e->bn_tile == 64) { return launch_conv_resident<64, 6, 1>(e, in_ext, L, out, residual); }
        if (e->conv_cluster == 2) { return launch_conv_resident<128, 9, 2>(e, in_ext, L, out, residual); }
        if (e->conv_cluster == 4) { return launch_conv_resident<128, 9, 4>(e, in_ext, L, out, residual); }
        return launch_conv_resident<128, 9, 1>(e, in_ext, L, out, residual);
    }
    switch (e->bn_tile) {
        case 64: return launch_conv<64, 4>(e, in, L, out, residual);
        case 128: return launch_conv<128, 3>(e, in, L, out, residual);
        default: return launch_conv<256, 4>(e, in, L, out, residual);
    }
}

int launch_heads(mz_engine* e, const __half* act, int* clear = nullptr, int clear_count = 0)
{
    mznn::HeadParams p;
    p.clear = clear, p.clear_count = clear_count;
    auto f = [&](int i) { return reinterpret_cast<const float*>(e->d_blob + e->off_head[i]); };
    p.act = act;
    p.w_pc = f(0), p.b_pc = f(1), p.w_pf = f(2), p.b_pf = f(3), p.w_vc = f(4), p.b_vc = f(5), p.w_v1 = f(6), p.b_v1 = f(7), p.w_v2 = f(8), p.b_v2 = f(9);
    p.policy = e->s.policy, p.logits = e->s.logits, p.value = e->s.nn_value;
    p.c = e->cpad, p.n = e->d.N, p.slots = e->d.slots, p.pol_ch = e->pol_ch, p.actions = e->d.A, p.vh = e->nd.num_value_hidden_channels;
    const int hw = e->d.N * e->d.N, np1 = p.pol_ch + 1;
    p.batch = e->d.B, p.fc_in_smem = 0;
    const size_t smem = sizeof(float) * (np1 * p.c + np1 * hw + p.vh + p.actions + 32 + hw + 4 * (p.actions + p.vh));
    static const int threads_env = [] {
        const char* env = knob("MZ_HEADS_THREADS");
        const int v = (env ? std::atoi(env) : 0);
        return (v == 256 || v == 512 || v == 1024) ? v : 0;
    }();
    // measured (profiles/): 256 hidden channels 31.0 / 22.6 / 18.5 us at 256 / 512 / 1024 threads; 128 channels are fastest at 512
    const int threads = (threads_env ? threads_env : (e->cpad >= 256 ? 1024 : 512));
    switch (np1) {
        case 2: mznn::heads_kernel<2><<<e->d.B, threads, smem, e->stream>>>(p); break;
        case 3: mznn::heads_kernel<3><<<e->d.B, threads, smem, e->stream>>>(p); break;
        case 4: mznn::heads_kernel<4><<<e->d.B, threads, smem, e->stream>>>(p); break;
        default: return fail(MZ_ERR_ARG, "policy head with more than 3 planes is not supported");
    }
    e->launches++;
    return MZ_OK;
}

int launch_tower(mz_engine* e, int which, bool clear_counters = true, bool pdl = false);
int launch_tower_params(mz_engine* e, mznn::TowerParams* params, int* d_done, int cout, int cin_max, int rows_ext, int stages, bool clear_counters, bool pdl, int bn = 128);

// hidden-state scaling + the three heads of the Atari network on the tower output `act` (see atari_kernels.cuh): 4 launches
int launch_atari_heads(mz_engine* e, const __half* act, bool with_reward)
{
    auto f = [&](int h, int i) { return e->d_blob + e->off_dhead[h][i]; };
    const int hw = e->d.N * e->d.N, Ch = e->nd.num_hidden_channels;
    if (Ch > 512 || e->d.A > 32) { return fail(MZ_ERR_ARG, "Atari heads support up to 512 hidden channels and 32 actions"); }
    mzat::FcGemmParams g0{};
    g0.a[0] = act, g0.w[0] = reinterpret_cast<const __half*>(e->d_blob + e->off_planes[0]), g0.bias[0] = nullptr, g0.out[0] = e->d_planes_raw;
    g0.n[0] = 64, g0.k[0] = e->cpad, g0.lda[0] = e->cpad, g0.ldw[0] = e->cpad, g0.ldc[0] = 64;
    mzat::fc_gemm_kernel<false><<<dim3(1, e->rows_alloc / 64, 1), 128, mzat::FC_SMEM, e->stream>>>(g0);
    e->launches++;
    mzat::PlanesParams pp;
    pp.act = act, pp.raw = e->d_planes_raw, pp.ld_raw = 64, pp.hid = reinterpret_cast<__half*>(e->s.hid), pp.slot = e->s.eval_slot;
    pp.n = e->d.N, pp.slots = e->d.slots, pp.c = e->cpad, pp.c_real = Ch, pp.num_slots = e->d.S + 1;
    pp.bias = reinterpret_cast<const float*>(e->d_blob + e->off_planes[1]), pp.wsum = reinterpret_cast<const float*>(e->d_blob + e->off_planes[2]);
    pp.pol_ch = e->pol_ch, pp.hc = e->dh_planes, pp.with_reward = (with_reward ? 1 : 0);
    pp.a_val = e->d_a_head[0], pp.a_rew = e->d_a_head[1], pp.pol_planes = e->d_pol_planes, pp.k_pad = e->dh_k1;
    mzat::hidden_planes_kernel<<<e->d.B, 256, 0, e->stream>>>(pp);
    e->launches++;
    const int nh = (with_reward ? 2 : 1);
    mzat::FcGemmParams g1{}, g2{};
    int n1max = 0;
    for (int h = 0; h < nh; ++h) {
        g1.a[h] = e->d_a_head[h], g1.w[h] = reinterpret_cast<const __half*>(f(h, 2)), g1.bias[h] = reinterpret_cast<const float*>(f(h, 3)), g1.out[h] = e->d_h_head[h];
        g1.n[h] = e->dh_n1[h], g1.k[h] = e->dh_k1, g1.lda[h] = e->dh_k1, g1.ldw[h] = e->dh_k1, g1.ldc[h] = e->dh_n1[h];
        g2.a[h] = e->d_h_head[h], g2.w[h] = reinterpret_cast<const __half*>(f(h, 4)), g2.bias[h] = reinterpret_cast<const float*>(f(h, 5)), g2.out[h] = e->d_l_head[h];
        g2.n[h] = e->dh_n2, g2.k[h] = e->dh_n1[h], g2.lda[h] = e->dh_n1[h], g2.ldw[h] = e->dh_n1[h], g2.ldc[h] = e->dh_n2;
        n1max = std::max(n1max, e->dh_n1[h]);
    }
    mzat::fc_gemm_kernel<true><<<dim3(n1max / 64, e->dh_mpad / 64, nh), 128, mzat::FC_SMEM, e->stream>>>(g1);
    e->launches++;
    mzat::fc_gemm_kernel<false><<<dim3(e->dh_n2 / 64, e->dh_mpad / 64, nh), 128, mzat::FC_SMEM, e->stream>>>(g2);
    e->launches++;
    mzat::FinalizeParams fp;
    fp.lg_val = e->d_l_head[0], fp.lg_rew = (with_reward ? e->d_l_head[1] : nullptr), fp.ld = e->dh_n2, fp.dv = e->nd.discrete_value_size;
    fp.pol_planes = e->d_pol_planes, fp.w_pf = reinterpret_cast<const float*>(f(2, 2)), fp.b_pf = reinterpret_cast<const float*>(f(2, 3)), fp.pol_in = e->pol_ch * hw, fp.actions = e->d.A;
    fp.value = e->s.nn_value, fp.reward = e->s.nn_reward, fp.policy = e->s.policy, fp.logits = e->s.logits;
    mzat::discrete_finalize_kernel<<<e->d.B, 128, 0, e->stream>>>(fp);
    e->launches++;
    return MZ_OK;
}

int launch_stage(mz_engine* e, ConvStage& st)
{
    return launch_tower_params(e, st.params, st.d_done, st.cout, st.cin_max, st.rows_ext, st.stages, true, false);
}

// MuZeroAtariNetwork (network/py/muzero_atari_network.py:157-183). which == 0: initial_inference on the space-to-depth planes already
// in the first stage's input rows; which == 1: recurrent_inference on the rows in dyn_in. The reward head reads the dynamics output
// before it is scaled (:53-54), the prediction heads the scaled hidden state.
int forward_atari(mz_engine* e, int which)
{
    int rc;
    const int B = e->d.B, grid = 4 * e->num_sms;
    NetTower& T = e->tw[which];
    if (which == 0) {
        ConvStage &A = e->ast[0], &Bs = e->ast[1], &Cs = e->ast[2];
        if ((rc = launch_stage(e, A))) { return rc; }
        mzat::space_to_depth_kernel<<<grid, 256, 0, e->stream>>>(A.act[A.out_buf], Bs.in, B, A.n, A.cout);
        e->launches++;
        if ((rc = launch_stage(e, Bs))) { return rc; }
        mzat::avgpool_kernel<<<grid, 256, 0, e->stream>>>(Bs.act[Bs.out_buf], Cs.act[0], B, Bs.n, Bs.cout);
        e->launches++;
        if ((rc = launch_stage(e, Cs))) { return rc; }
        mzat::avgpool_kernel<<<grid, 256, 0, e->stream>>>(Cs.act[Cs.out_buf], e->act[0], B, Cs.n, Cs.cout);
        e->launches++;
    }
    if ((rc = launch_tower(e, which, true, false))) { return rc; }
    __half* out = e->act[T.out_buf];
    if (which == 0) {
        CUDA_OK(cudaMemsetAsync(e->s.nn_reward, 0, sizeof(float) * B, e->stream)); // initial_inference has no reward output: reward_ stays 0 (muzero_network.h:25)
        e->memsets++;
    }
    return launch_atari_heads(e, out, which == 1);
}

// AlphaZeroNetwork.forward (network/py/alphazero_network.py:90-113) on the rows already in nn_in; for a MuZero network
// `which` selects initial_inference (0: representation, rows in nn_in) or recurrent_inference (1: dynamics, rows in dyn_in),
// each followed by scale_hidden_state and the prediction heads (network/py/muzero_network.py:136-160)
int forward(mz_engine* e, int which = 0, bool after_tree_step = false)
{
    const bool in_search = after_tree_step; // think(): only a search step knows which lanes hold a leaf (the evaluation hooks fill every board)
    if (!e->net_ready) { return fail(MZ_ERR_STATE, "network not finalized"); }
    if (e->atari) { return forward_atari(e, which); }
    NetTower& T = e->tw[which];
    int rc;
    if (e->conv_mode == 3) {
        if ((rc = launch_tower(e, which, false, after_tree_step && e->tower_pdl))) { return rc; } // counters: zero from allocation, then re-zeroed by every heads launch below
    } else {
        if ((rc = conv(e, T.map_in, T.map_in_ext, T.convs[0], e->act[0], nullptr))) { return rc; }
        int cur = 0;
        for (int b = 0; b < e->nd.num_blocks; ++b) {
            const int t = (cur + 1) % 3, o = (cur + 2) % 3;
            if ((rc = conv(e, e->map_act[cur], e->map_act_ext[cur], T.convs[1 + 2 * b], e->act[t], nullptr))) { return rc; }
            if ((rc = conv(e, e->map_act[t], e->map_act_ext[t], T.convs[2 + 2 * b], e->act[o], e->act[cur]))) { return rc; }
            cur = o;
        }
    }
    __half* out = e->act[T.out_buf];
    if (e->cfg.muzero) {
        mznn::scale_hidden_kernel<<<e->d.B, 256, 0, e->stream>>>(out, reinterpret_cast<__half*>(e->s.hid), e->s.eval_slot, e->d.N, e->d.slots, e->cpad,
                                                                  e->nd.num_hidden_channels, e->d.S + 1, e->d.think_k ? e->think_trees : 0,
                                                                  (e->d.think_k && in_search) ? e->s.path_len : nullptr);
        e->launches++;
    }
    if (e->conv_mode == 3) { return launch_heads(e, out, T.d_done, T.done_count); }
    return launch_heads(e, out);
}

// Ownership rotation of the tower's units: pair c owns units c', c' + nc, ... of layer l with c' = (c + l * rotate) mod nc. A layer of `units`
// units ends with a short pass of units mod nc units; the pairs that have no unit in it go on to the next layer at once, and with
// rotate = nc - units mod nc they find there exactly the units whose 3x3 halo lies in the part of this layer that is already complete
// (full passes), while the pairs still busy with the short pass later take the units next to it. With any other rotation a layer's short
// pass acts as a barrier: measured at config 2, 100 units of the wide kernel on 74 pairs, rotation 22 (the narrow kernel's 200 units): the
// MMA issuer waits for input blocks 39 % of its time.
int tower_rotation(const mz_engine* e, int units, int clusters)
{
    if (e->tower_rot_override >= 0) { return e->tower_rot_override; }
    return (clusters - units % clusters) % clusters;
}

// one launch of the fused tower kernel over `params` (a NetTower or a ConvStage)
int launch_tower_params(mz_engine* e, mznn::TowerParams* params, int* d_done, int cout, int cin_max, int rows_ext, int stages, bool clear_counters, bool pdl, int bn)
{
    const int num_groups = (params->num_mtiles + 1) / 2;
    if (clear_counters) {
        CUDA_OK(cudaMemsetAsync(d_done, 0, sizeof(int) * params->num_layers * num_groups, e->stream));
        e->memsets++;
    }
    const int units = num_groups * (cout / bn);
    int clusters = e->tower_sms / 2;
    if (units < clusters) { clusters = units; }
    params->rotate = tower_rotation(e, units, clusters);
    const size_t smem = mznn::tower_smem_bytes(cin_max, rows_ext, stages, bn, params->epi_bufs);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 2), cfg.blockDim = dim3(mznn::TOWER_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = e->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    params->pdl = (pdl ? 1 : 0);
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (pdl) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    } else if (e->tower_coop) { // the CTAs wait for each other: a cooperative launch makes the driver guarantee that the whole grid is resident
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.numAttrs = 2;
    }
    if (bn == 256) { // output-channel tile of 256: half the input-block reads per FLOP (see DESIGN.md "What bounds the tower")
        if (stages == 4) {
            CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<256, 4, false>, *params));
        } else {
            CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<256, 2, false>, *params));
        }
    } else if (params->dbg && stages == 8) {
        CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<128, 8, true>, *params));
    } else if (stages == 4) { // 185 KB of shared memory: a tree-step block (30 KB) of another engine fits on the same SM
        CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<128, 4, false>, *params));
    } else if (params->dbg && stages == 5) {
        CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<128, 5, true>, *params));
    } else if (stages == 5) {
        CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<128, 5, false>, *params));
    } else {
        CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_kernel<128, 8, false>, *params));
    }
    e->launches++;
    return MZ_OK;
}

// the wide variant (two row tiles per CTA): see conv_tower_wide_kernel
int launch_tower_wide(mz_engine* e, NetTower& T, bool clear_counters, bool pdl)
{
    mznn::TowerParams* params = T.params;
    if (clear_counters) {
        CUDA_OK(cudaMemsetAsync(T.d_done, 0, sizeof(int) * T.done_count, e->stream));
        e->memsets++;
    }
    const int units = ((params->num_mtiles + 3) / 4) * (e->cpad / 128);
    int clusters = e->tower_sms / 2;
    if (units < clusters) { clusters = units; }
    params->rotate = tower_rotation(e, units, clusters);
    const int stages = e->tower_wide_stages;
    const size_t smem = mznn::wide_smem_bytes(e->rows_ext_wide, stages);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * 2), cfg.blockDim = dim3(mznn::WIDE_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = e->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    params->pdl = (pdl ? 1 : 0);
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (pdl) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    } else if (e->tower_coop) { // the CTAs wait for each other: a cooperative launch makes the driver guarantee that the whole grid is resident
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.numAttrs = 2;
    }
    if (stages == 10) {
        if (params->dbg) {
            CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_wide_kernel<10, true>, *params));
        } else {
            CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_wide_kernel<10, false>, *params));
        }
    } else {
        if (params->dbg) {
            CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_wide_kernel<9, true>, *params));
        } else {
            CUDA_OK(cudaLaunchKernelEx(&cfg, mznn::conv_tower_wide_kernel<9, false>, *params));
        }
    }
    e->launches++;
    return MZ_OK;
}

int launch_tower(mz_engine* e, int which, bool clear_counters, bool pdl)
{
    NetTower& T = e->tw[which];
    if (e->tower_wide) { return launch_tower_wide(e, T, clear_counters, pdl); }
    return launch_tower_params(e, T.params, T.d_done, e->cpad, e->cin_max, e->rows_ext, e->tower_stages, clear_counters, pdl, e->tower_bn);
}

size_t step_smem_bytes(const mz_dims& d)
{
    return (16 + 16 + sizeof(uint64_t) + sizeof(int32_t)) * static_cast<size_t>(d.S + 2) + sizeof(float) * STEP_WARPS * d.A;
}

int step(mz_engine* e, int flags, const uint8_t* rotations)
{
    mz_state s = e->s;
    s.rotations = rotations;
    s.noise_in = (e->noise_enabled ? e->d_noise : nullptr);
    if (e->d.think_k) {
        // one batched think() step (zero_actor.cpp:129-157): the lanes' results are applied in selection order, then K selections per tree are made
        // one after the other (each leaves its virtual loss for the next); a pass is one launch over the trees with the lane's view of the state
        const int trees = e->think_trees;
        for (int phase = 0; phase < 2; ++phase) {
            const int f = (phase == 0 ? STEP_AFTER : STEP_BEFORE);
            if (!(flags & f)) { continue; }
            if (phase == 1) {
                CUDA_OK(cudaMemsetAsync(e->s.think_pending, 0, sizeof(int32_t) * trees, e->stream));
                e->memsets++;
            }
            for (int k = 0; k < e->d.think_k; ++k) {
                const mz_state v = mz_lane_view(e->d, s, k, trees);
                k_step<<<trees, 32 * STEP_WARPS, step_smem_bytes(e->d), e->stream>>>(e->d, v, f);
                e->launches++;
            }
        }
        return MZ_OK;
    }
    k_step<<<e->d.B, 32 * STEP_WARPS, step_smem_bytes(e->d), e->stream>>>(e->d, s, flags);
    e->launches++;
    return MZ_OK;
}

std::vector<float> fold_conv(const mz_engine* e, const std::string& conv, const std::string& bn, int cout, int cin, int k, std::vector<float>& bias_out, std::string& err)
{
    auto get = [&](const std::string& name, size_t n) -> const float* {
        auto it = e->tensors.find(name);
        if (it == e->tensors.end()) {
            err = "missing tensor " + name;
            return nullptr;
        }
        if (it->second.size() != n) {
            err = "tensor " + name + " has " + std::to_string(it->second.size()) + " elements, expected " + std::to_string(n);
            return nullptr;
        }
        return it->second.data();
    };
    const size_t wn = static_cast<size_t>(cout) * cin * k * k;
    const float *w = get(conv + ".weight", wn), *b = get(conv + ".bias", cout), *g = get(bn + ".weight", cout), *be = get(bn + ".bias", cout),
                *mu = get(bn + ".running_mean", cout), *var = get(bn + ".running_var", cout);
    std::vector<float> out;
    if (!w || !b || !g || !be || !mu || !var) { return out; }
    out.resize(wn);
    bias_out.resize(cout);
    for (int co = 0; co < cout; ++co) {
        const float sc = g[co] / std::sqrt(var[co] + 1e-5f); // BatchNorm2d eval, eps default
        for (size_t i = 0; i < static_cast<size_t>(cin) * k * k; ++i) { out[co * static_cast<size_t>(cin) * k * k + i] = w[co * static_cast<size_t>(cin) * k * k + i] * sc; }
        bias_out[co] = (b[co] - mu[co]) * sc + be[co];
    }
    return out;
}

// Atari representation stages (muzero_atari_network.py:7-40) and the discrete heads. "stride 2": a 3x3 / stride-2 / pad-1
// convolution is a stride-1 convolution over the space-to-depth(2) input: output cell (X, Y) reads pixels 2Y-1 .. 2Y+1, i.e. cell
// row Y-1 (its lower pixel row, dy = 1) and cell row Y (both pixel rows), so only the taps that reach up / left exist:
//   (tap row ty, sub-row dy) -> kernel row ky:  (0, 1) -> 0,  (1, 0) -> 1,  (1, 1) -> 2,  (0, 0) -> none      (columns alike).
// conv1 (32 planes) runs as ONE layer over all four sub-positions (4 * 32 = 128 input channels, taps {0, 1, 3, 4}); conv2 (C/2
// channels) as FOUR chained layers, one per sub-position, each reading its 128-channel slice of the space-to-depth rows with
// exactly the taps that exist for it (1 + 2 + 2 + 4 = 9 tap-slices: the algorithmic FLOPs, nothing wasted) and adding the
// previous partial sum through the residual path; bias, BatchNorm shift and ReLU belong to the last one.
int stride2_tap_mask(int dy, int dx)
{
    int mask = 0;
    for (int ty = (dy ? 0 : 1); ty <= 1; ++ty) {
        for (int tx = (dx ? 0 : 1); tx <= 1; ++tx) { mask |= 1 << (ty * 3 + tx); }
    }
    return mask;
}
int stride2_kernel_index(int t, int d) { return t == 0 ? (d == 1 ? 0 : -1) : (d == 0 ? 1 : 2); }

int plan_blob_atari(mz_engine* e)
{
    const mz_net_dims& nd = e->nd;
    const int C = e->cpad, C1 = ((nd.num_hidden_channels / 2 + 127) / 128) * 128;
    e->at_c1 = C1;
    auto add = [&](ConvStage& st, int cin, int cout, int cin_off, int taps, int relu, int in_buf, int out_buf, int res_buf) {
        ConvLayer L;
        L.cin = cin, L.cout = cout, L.relu = relu, L.cin_off = cin_off, L.tap_mask = taps, L.in_buf = in_buf, L.out_buf = out_buf, L.res_buf = res_buf;
        L.w_off = e->blob.take(sizeof(__half) * 9 * static_cast<size_t>(cout) * cin);
        L.b_off = e->blob.take(sizeof(float) * cout);
        st.convs.push_back(L);
    };
    ConvStage &A = e->ast[0], &B = e->ast[1], &Cs = e->ast[2];
    A = ConvStage(), B = ConvStage(), Cs = ConvStage();
    A.n = mzat::RES / 2, A.cin_in = 4 * mzat::PLANES, A.cout = C1; // conv1 + residual_blocks1
    add(A, A.cin_in, C1, 0, 0x01b, 1, -1, 0, -2);
    add(A, C1, C1, 0, 0x1ff, 1, 0, 1, -2);
    add(A, C1, C1, 0, 0x1ff, 1, 1, 2, 0);
    A.out_buf = 2;
    B.n = mzat::RES / 4, B.cin_in = 4 * C1, B.cout = C; // conv2 (four sub-position layers) + residual_blocks2
    for (int q = 0; q < 4; ++q) { add(B, C1, C, q * C1, stride2_tap_mask(q >> 1, q & 1), q == 3 ? 1 : 0, -1, q & 1, q == 0 ? -2 : ((q - 1) & 1)); }
    add(B, C, C, 0, 0x1ff, 1, 1, 2, -2);
    add(B, C, C, 0, 0x1ff, 1, 2, 0, 1);
    B.out_buf = 0;
    Cs.n = mzat::RES / 8, Cs.cin_in = 0, Cs.cout = C; // avg_pooling1 writes act[0]; residual_blocks3
    add(Cs, C, C, 0, 0x1ff, 1, 0, 1, -2);
    add(Cs, C, C, 0, 0x1ff, 1, 1, 2, 0);
    Cs.out_buf = 2;
    // heads (network_unit.py:26-42,68-87), fp32: policy conv [pol_ch][C], b, fc^T [pol_ch * 36][A], b; value / reward: conv [hc][C], b,
    // fc1^T [hc * 36][vh], b, fc2^T [vh][dv], b
    const int hw = e->d.N * e->d.N, dv = nd.discrete_value_size, vh = nd.num_value_hidden_channels;
    e->dh_planes = (dv + hw - 1) / hw;
    const int hc = e->dh_planes;
    auto pad64 = [](int v) { return (v + 63) / 64 * 64; };
    e->dh_k1 = pad64(hc * hw), e->dh_n2 = pad64(dv), e->dh_mpad = pad64(e->d.B);
    for (int h = 0; h < 2; ++h) {
        const int fc1_out = (h == 0 ? vh : nd.num_hidden_channels); // the reward head's hidden width is num_channels (muzero_atari_network.py:50)
        e->dh_n1[h] = pad64(fc1_out);
        const size_t sizes[6] = {sizeof(float) * hc * C, sizeof(float) * hc, sizeof(__half) * e->dh_n1[h] * e->dh_k1, sizeof(float) * e->dh_n1[h],
                                 sizeof(__half) * e->dh_n2 * e->dh_n1[h], sizeof(float) * e->dh_n2};
        for (int i = 0; i < 6; ++i) { e->off_dhead[h][i] = e->blob.take(sizes[i]); }
    }
    if (e->pol_ch + 2 * hc > 64) { return fail(MZ_ERR_ARG, "the heads' 1x1 convolutions have more than 64 planes"); }
    e->off_planes[0] = e->blob.take(sizeof(__half) * 64 * C), e->off_planes[1] = e->blob.take(sizeof(float) * 64), e->off_planes[2] = e->blob.take(sizeof(float) * 64);
    const size_t psizes[4] = {static_cast<size_t>(e->pol_ch) * C, static_cast<size_t>(e->pol_ch), static_cast<size_t>(e->pol_ch) * hw * nd.action_size, static_cast<size_t>(nd.action_size)};
    for (int i = 0; i < 4; ++i) { e->off_dhead[2][i] = e->blob.take(sizeof(float) * psizes[i]); }
    return MZ_OK;
}

int plan_blob(mz_engine* e)
{
    const mz_net_dims& nd = e->nd;
    e->cpad = ((nd.num_hidden_channels + 63) / 64) * 64;
    if (e->atari) { e->cpad = ((nd.num_hidden_channels + 127) / 128) * 128; } // every stage of the Atari network runs through the fused tower (128-wide tiles)
    const int head_hw = (e->atari ? e->d.N * e->d.N : nd.input_height * nd.input_width); // heads work on the hidden state's cells
    e->pol_ch = (nd.action_size + head_hw - 1) / head_hw;
    // output-channel tile of the conv kernel: 128 keeps two CTAs resident per SM (epilogue of one overlaps the
    // main loop of the other) and gives 2x the tiles for wave balance; MZ_CONV_BN overrides for experiments
    e->bn_tile = (e->cpad % 128 == 0 ? 128 : 64);
    if (const char* env = knob("MZ_CONV_BN")) {
        const int v = std::atoi(env);
        if ((v == 64 || v == 128 || v == 256) && e->cpad % v == 0) { e->bn_tile = v; }
    }
    e->blob = Blob();
    e->num_towers = (e->cfg.muzero ? 2 : 1);
    e->tw[0].prefix = (e->cfg.muzero ? "representation_network." : "");
    e->tw[0].cin0_real = nd.num_input_channels, e->tw[0].cin0 = MZ_NN_CPAD;
    e->tw[1].prefix = "dynamics_network.";
    e->tw[1].cin0_real = nd.num_hidden_channels + nd.num_action_feature_channels; // torch.cat((hidden_state, action_plane), dim=1), muzero_network.py:31
    e->tw[1].cin0 = ((std::max(e->tw[1].cin0_real, e->cpad) + 63) / 64) * 64; // the gather copies whole (padded) hidden rows, then the action planes over columns Ch ..
    e->cin_max = e->cpad;
    for (int t = 0; t < e->num_towers; ++t) {
        NetTower& T = e->tw[t];
        T.has_stem = !(e->atari && t == 0); // the 6 x 6 stage of the Atari representation network is residual blocks only
        T.convs.assign((T.has_stem ? 1 : 0) + 2 * nd.num_blocks, ConvLayer());
        if (T.has_stem && T.cin0 > e->cin_max) { e->cin_max = T.cin0; }
        for (size_t i = 0; i < T.convs.size(); ++i) {
            ConvLayer& L = T.convs[i];
            L.cin = (i == 0 && T.has_stem ? T.cin0 : e->cpad), L.cout = e->cpad, L.relu = 1;
            L.w_off = e->blob.take(sizeof(__half) * 9 * static_cast<size_t>(L.cout) * L.cin);
            L.b_off = e->blob.take(sizeof(float) * L.cout);
        }
    }
    if (e->atari) { return plan_blob_atari(e); }
    const int hw = nd.input_height * nd.input_width;
    const size_t head_sizes[10] = {static_cast<size_t>(e->pol_ch) * e->cpad, static_cast<size_t>(e->pol_ch), static_cast<size_t>(nd.action_size) * e->pol_ch * hw,
                                   static_cast<size_t>(nd.action_size), static_cast<size_t>(e->cpad), 1, static_cast<size_t>(nd.num_value_hidden_channels) * hw,
                                   static_cast<size_t>(nd.num_value_hidden_channels), static_cast<size_t>(nd.num_value_hidden_channels), 1};
    for (int i = 0; i < 10; ++i) { e->off_head[i] = e->blob.take(sizeof(float) * head_sizes[i]); }
    return MZ_OK;
}

// buffers, tensor maps and launch parameters of the three representation stages
int alloc_atari(mz_engine* e)
{
    int rc;
    for (int si = 0; si < 3; ++si) {
        ConvStage& st = e->ast[si];
        st.slots = (st.n + 1) * (st.n + 1);
        st.rows_alloc = static_cast<int>((static_cast<size_t>(e->d.B) * st.slots + mznn::BM - 1) / mznn::BM * mznn::BM);
        st.rows_ext = (mznn::BM + 2 * (st.n + 2) + 7) / 8 * 8;
        st.cin_max = 0;
        for (const ConvLayer& L : st.convs) { st.cin_max = std::max(st.cin_max, L.cin); }
        st.stages = 0;
        int st_bufs = 2;
        // the TMA-store epilogue shortens the hand-off between layers; a stage with many units per CTA pair and layer is bound by the epilogue's throughput instead
        const int st_units = (st.rows_alloc / mznn::BM + 1) / 2 * (st.cout / 128);
        const bool st_chain_bound = (st_units <= 4 * (e->num_sms / 2));
        for (int bufs : {2, 1, 0}) { // (0 is also the fall-back where the input blocks leave no room for staging tiles: the 24 x 24 stage of a 256-channel network)
            if (bufs != 0 && !st_chain_bound) { continue; }
            for (int cand : {8, 5, 4}) {
                if (st.stages == 0 && mznn::tower_smem_bytes(st.cin_max, st.rows_ext, cand, 128, bufs) <= 227 * 1024) { st.stages = cand, st_bufs = bufs; }
            }
        }
        if (st.stages == 0 || st.rows_ext > 256) { return fail(MZ_ERR_ARG, "an Atari representation stage does not fit the tower kernel's shared memory"); }
        const size_t rows = st.rows_alloc;
        if (st.cin_in > 0) {
            if ((rc = e->dalloc(&st.in, rows * st.cin_in))) { return rc; }
            if ((rc = make_map_2d(e, &st.map_in_ext, st.in, st.cin_in, rows, mznn::BK, st.rows_ext))) { return rc; }
        }
        for (int i = 0; i < 3; ++i) {
            if ((rc = e->dalloc(&st.act[i], rows * st.cout))) { return rc; }
            if ((rc = make_map_2d(e, &st.map_act_ext[i], st.act[i], st.cout, rows, mznn::BK, st.rows_ext))) { return rc; }
        }
        st.params = new mznn::TowerParams();
        mznn::TowerParams& T = *st.params;
        T.num_layers = static_cast<int>(st.convs.size());
        T.rows_valid = e->d.B * st.slots, T.n1 = st.n + 1, T.slots = st.slots, T.cout = st.cout, T.rows_ext = st.rows_ext, T.halo = st.n + 2;
        T.num_mtiles = st.rows_alloc / mznn::BM, T.cin_max = st.cin_max;
        T.rotate = 22, T.shift = 0, T.zigzag = 0, T.strided = 1, T.tap_rot = 0, T.fence_mode = 0, T.pdl = 0, T.dbg = nullptr, T.epi_bufs = st_bufs;
        for (int li = 0; li < T.num_layers; ++li) {
            ConvLayer& L = st.convs[li];
            if ((rc = make_map_2d(e, &L.map_w_mc, e->d_blob + L.w_off, L.cin, 9ull * L.cout, mznn::BK, 64))) { return rc; }
            mznn::TowerLayer& TL = T.layer[li];
            TL.map_in = (L.in_buf < 0 ? st.map_in_ext : st.map_act_ext[L.in_buf]), TL.map_w = L.map_w_mc;
            TL.out = st.act[L.out_buf], TL.residual = (L.res_buf == -2 ? nullptr : (L.res_buf < 0 ? st.in : st.act[L.res_buf]));
            TL.bias = reinterpret_cast<const float*>(e->d_blob + L.b_off), TL.cin = L.cin, TL.relu = L.relu, TL.cin_off = L.cin_off, TL.tap_mask = L.tap_mask;
            // the epilogue stores through TMA; residual rows: the stage's input was written before the launch, an activation buffer by some earlier
            // layer of this launch (-2: the epilogue reads it after the accumulator barrier, which orders it behind that layer like every input row)
            TL.out_map = L.out_buf, TL.res_layer = (L.res_buf == -2 || L.res_buf < 0 ? -1 : -2);
        }
        for (int i = 0; i < 3; ++i) {
            if ((rc = make_map_2d(e, &T.map_out[i], st.act[i], st.cout, rows, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) { return rc; }
        }
        const int num_groups = (T.num_mtiles + 1) / 2;
        if ((rc = e->dalloc(&st.d_done, static_cast<size_t>(T.num_layers) * num_groups))) { return rc; }
        T.done = st.d_done;
    }
    for (int h = 0; h < 2; ++h) {
        if ((rc = e->dalloc(&e->d_a_head[h], static_cast<size_t>(e->dh_mpad) * e->dh_k1))) { return rc; }
        if ((rc = e->dalloc(&e->d_h_head[h], static_cast<size_t>(e->dh_mpad) * e->dh_n1[h]))) { return rc; }
        if ((rc = e->dalloc(&e->d_l_head[h], static_cast<size_t>(e->dh_mpad) * e->dh_n2))) { return rc; }
    }
    if ((rc = e->dalloc(&e->d_pol_planes, static_cast<size_t>(e->d.B) * e->pol_ch * e->d.N * e->d.N))) { return rc; }
    if ((rc = e->dalloc(&e->d_planes_raw, static_cast<size_t>(e->rows_alloc) * 64))) { return rc; }
    CUDA_OK(cudaFuncSetAttribute(mzat::fc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mzat::FC_SMEM));
    CUDA_OK(cudaFuncSetAttribute(mzat::fc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mzat::FC_SMEM));
    return MZ_OK;
}

int alloc_net(mz_engine* e)
{
    if (e->d_blob) { return MZ_OK; }
    int rc;
    if ((rc = configure_conv_kernels())) { return rc; }
    if ((rc = e->dalloc(&e->d_blob, e->blob.size))) { return rc; }
    const size_t rows = e->rows_alloc;
    e->rows_ext = (mznn::BM + 2 * (e->d.N + 2) + 7) / 8 * 8;
    // conv kernel variant: 3 = all layers in one persistent launch of CTA pairs, chained by completion counters (default
    // where the shape allows); 2 = CTA pairs (cta_group::2) over resident input blocks, one launch per layer;
    // 1 = one CTA per tile with a resident input block; 0 = every tap re-loads its shifted A tile
    e->conv_mode = 3;
    if (const char* env = knob("MZ_CONV_MODE")) { e->conv_mode = std::atoi(env); }
    if (const char* env = knob("MZ_CONV_PDL")) { e->conv_pdl = std::atoi(env); }
    const bool want_tower = (e->conv_mode == 3);
    if (want_tower) { e->conv_mode = 2; } // the tower needs everything the pair kernel needs
    if (const char* env = knob("MZ_CONV_BASEOFF")) { e->base_off_mode = std::atoi(env); }
    if (const char* env = knob("MZ_CONV_ROT")) { e->krot = std::atoi(env); }
    e->conv_cluster = 1; // CTAs per cluster sharing every weight tile by TMA multicast (1, 2 or 4)
    if (const char* env = knob("MZ_CONV_CLUSTER")) {
        const int v = std::atoi(env);
        if (v == 1 || v == 2 || v == 4) { e->conv_cluster = v; }
    }
    if (e->bn_tile == 256) { e->conv_mode = 0; }
    const size_t need = (e->bn_tile == 64 ? resident_smem<64, 6>(e, e->cin_max) : resident_smem<128, 9>(e, e->cin_max));
    if (e->bn_tile == 64) { e->conv_cluster = 1; }
    if (e->conv_mode == 2) { // CTA pairs: needs the 128-wide tile and half-tile weight boxes
        // the fused tower runs as fast with 4 or 5 weight stages as with 8 (measured), so a larger resident block (19x19: 176 rows)
        // simply takes fewer stages; the per-layer pair kernel is instantiated for 8 only
        auto pair_need = [&](int stages, int bufs) { return mznn::tower_smem_bytes(e->cin_max, e->rows_ext, stages, 128, bufs); };
        int stages = 0;
        const int narrow_units = (e->rows_alloc / mznn::BM + 1) / 2 * (e->cpad / 128);
        const bool chain_bound = (narrow_units <= 4 * (e->num_sms / 2)); // else: many units per CTA pair and layer, the epilogue's throughput counts (see tower_smem_bytes)
        for (int bufs : {2, 1, 0}) { // staging tiles per epilogue warp: two where they fit
            if (bufs != 0 && !chain_bound) { continue; }
            for (int cand : {8, 5, 4}) {
                if (stages == 0 && pair_need(cand, bufs) <= 227 * 1024 && (want_tower || cand == 8)) { stages = cand, e->tower_epi_bufs = bufs; }
            }
        }
        if (e->bn_tile == 128 && stages != 0) {
            e->conv_cluster = 2;
            if (!knob("MZ_TOWER_STAGES")) { e->tower_stages = stages; }
            if (const char* env = knob("MZ_TOWER_BN")) { // 256-wide output tiles (experiment): stages of 16 KB per CTA, 4 or 2 of them
                if (std::atoi(env) == 256 && want_tower && e->cpad % 256 == 0 && !e->atari) {
                    auto need256 = [&](int st) { return mznn::tower_smem_bytes(e->cin_max, e->rows_ext, st, 256, 2); };
                    const int st = (need256(4) <= 227 * 1024 ? 4 : (need256(2) <= 227 * 1024 ? 2 : 0));
                    if (st) { e->tower_bn = 256, e->tower_stages = st; }
                }
            }
        } else {
            e->conv_mode = 1;
        }
    }
    if (e->conv_mode == 1 && need > 227 * 1024) { e->conv_mode = 0; }
    cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->cfg.device);
    for (int i = 0; i < 3; ++i) {
        if ((rc = e->dalloc(&e->act[i], rows * e->cpad))) { return rc; }
        if ((rc = make_map_2d(e, &e->map_act[i], e->act[i], e->cpad, rows, mznn::BK, mznn::BM))) { return rc; }
        if ((rc = make_map_2d(e, &e->map_act_ext[i], e->act[i], e->cpad, rows, mznn::BK, e->rows_ext))) { return rc; }
        if (e->cpad % 32 == 0 && (rc = make_map_2d(e, &e->map_act_store[i], e->act[i], e->cpad, rows, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) { return rc; }
    }
    if (e->cfg.muzero) {
        // hidden states of the evaluated nodes (slot = simulation index) and the dynamics network's input rows
        const size_t hw = static_cast<size_t>(e->d.N) * e->d.N;
        if ((rc = e->dalloc(&e->s.hid, static_cast<size_t>(e->d.B) * (e->d.S + 1) * hw * e->cpad))) { return rc; }
        if ((rc = e->dalloc(&e->s.dyn_in, rows * e->tw[1].cin0))) { return rc; }
        if ((rc = e->dalloc(&e->d_hidden_f32, static_cast<size_t>(e->d.B) * e->nd.num_hidden_channels * hw))) { return rc; }
        e->d.hid_c = e->cpad, e->d.dyn_c = e->tw[1].cin0, e->d.act_col = e->nd.num_hidden_channels;
    }
    e->tw[0].in = e->s.nn_in, e->tw[1].in = e->s.dyn_in;
    if (const char* env = knob("MZ_PDL")) { e->tower_pdl = std::atoi(env); }
    if (const char* env = knob("MZ_TOWER_STAGES")) {
        const int v = std::atoi(env);
        if (v == 4 || v == 5 || v == 8) { e->tower_stages = v; }
    }
    e->tower_sms = e->num_sms; // SMs the persistent tower may occupy (MZ_TOWER_SMS: experiments with a second engine beside it)
    if (const char* env = knob("MZ_TOWER_SMS")) {
        const int v = std::atoi(env);
        if (v >= 2 && v <= e->num_sms) { e->tower_sms = v & ~1; }
    }
    const bool tower_ok = (want_tower && e->conv_mode == 2 && 1 + 2 * e->nd.num_blocks <= mznn::TOWER_MAX_LAYERS && (!e->atari || e->nd.num_blocks > 0));
    // two row tiles per CTA (half the weight traffic per FLOP) where a layer still has a unit for every CTA pair
    e->rows_ext_wide = (2 * mznn::BM + 2 * (e->d.N + 2) + 15) / 16 * 16;
    {
        auto wide_need = [&](int stages) { return mznn::wide_smem_bytes(e->rows_ext_wide, stages); };
        e->tower_wide_stages = (wide_need(10) <= 227 * 1024 ? 10 : 9); // weight stages beside the input ring and the epilogue's staging tiles: at least the nine taps of a K-block
        const int wide_units = ((e->rows_alloc / mznn::BM + 3) / 4) * (e->cpad / 128);
        e->tower_wide = (tower_ok && !e->atari && e->tower_bn == 128 && e->cpad % 128 == 0 && wide_need(e->tower_wide_stages) <= 227 * 1024 && e->rows_ext_wide / 2 <= 256 &&
                         wide_units >= e->tower_sms / 2);
        // (>= one wide unit per pair and layer. A layer's critical path is one unit's MMAs + the hand-off to the next layer (epilogue -> counter -> poll -> TMA
        //  of the first K-block), and a layer holds units / pairs unit-times of work: the wide kernel wins where its halved weight traffic outweighs the longer
        //  unit. Measured after the epilogue went to TMA stores (profiles/r2_epi_ab.log), wide vs narrow: config 4 (200 units on 74 pairs) 1640 vs 2000 us,
        //  config 2 (100 units) 241-246 vs 263-271 us, config 3 (81 units, dynamics tower) 95.7 vs 102.9 us. Before that the hand-off cost 12 k cycles per
        //  unit in the epilogue alone and config 2 lost: 302 vs 264 us, the issuer waiting 39 % of its time for input.)
        if (const char* env = knob("MZ_TOWER_WIDE")) {
            const int v = std::atoi(env);
            if (v == 0) { e->tower_wide = false; }
            if (v == 1 && wide_units >= e->tower_sms / 2 && tower_ok && !e->atari && e->tower_bn == 128) { e->tower_wide = true; }
        }
    }
    if (e->tower_wide) {
        for (int i = 0; i < 3; ++i) {
            if ((rc = make_map_2d(e, &e->map_act_wide[i], e->act[i], e->cpad, rows, mznn::BK, e->rows_ext_wide / 2))) { return rc; }
        }
    }
    if (e->atari && !tower_ok) { return fail(MZ_ERR_ARG, "the Atari network needs the fused tower (1 <= num_blocks <= 23)"); }
    for (int t = 0; t < e->num_towers; ++t) {
        NetTower& NT = e->tw[t];
        if ((rc = make_map_2d(e, &NT.map_in, NT.in, NT.cin0, rows, mznn::BK, mznn::BM))) { return rc; }
        if ((rc = make_map_2d(e, &NT.map_in_ext, NT.in, NT.cin0, rows, mznn::BK, e->rows_ext))) { return rc; }
        if (e->tower_wide && (rc = make_map_2d(e, &NT.map_in_wide, NT.in, NT.cin0, rows, mznn::BK, e->rows_ext_wide / 2))) { return rc; }
        for (ConvLayer& L : NT.convs) {
            if ((rc = make_map_2d(e, &L.map_w, e->d_blob + L.w_off, L.cin, 9ull * L.cout, mznn::BK, e->bn_tile))) { return rc; }
            if ((rc = make_map_2d(e, &L.map_w_mc, e->d_blob + L.w_off, L.cin, 9ull * L.cout, mznn::BK, e->bn_tile / e->conv_cluster))) { return rc; }
            if ((rc = make_map_2d(e, &L.map_w_tower, e->d_blob + L.w_off, L.cin, 9ull * L.cout, mznn::BK, e->tower_bn / 2))) { return rc; }
        }
        NT.out_buf = (2 * e->nd.num_blocks) % 3; // forward()'s buffer rotation: cur -> t -> o per residual block, o = cur + 2
        if (!tower_ok) { continue; }
        // whole-tower launch: same buffer rotation as forward()
        NT.params = new mznn::TowerParams();
        mznn::TowerParams& T = *NT.params;
        T.num_layers = static_cast<int>(NT.convs.size());
        T.rows_valid = e->d.B * e->d.slots, T.n1 = e->d.N + 1, T.slots = e->d.slots, T.cout = e->cpad, T.rows_ext = (e->tower_wide ? e->rows_ext_wide : e->rows_ext), T.halo = e->d.N + 2;
        T.num_mtiles = e->rows_alloc / mznn::BM;
        T.cin_max = e->cin_max;
        T.rotate = 22, T.shift = 0, T.zigzag = 0, T.strided = 1, T.tap_rot = 0, T.fence_mode = 0, T.epi_bufs = e->tower_epi_bufs;
        if (const char* env = knob("MZ_TOWER_TAPROT")) { T.tap_rot = std::atoi(env); }
        if (const char* env = knob("MZ_TOWER_FENCE")) { T.fence_mode = std::atoi(env); }
        if (const char* env = knob("MZ_TOWER_STRIDED")) { T.strided = std::atoi(env); }
        if (const char* env = knob("MZ_TOWER_ZIGZAG")) { T.zigzag = std::atoi(env); }
        if (const char* env = knob("MZ_TOWER_ROT")) { e->tower_rot_override = std::atoi(env); }
        if (const char* env = knob("MZ_TOWER_SHIFT")) { T.shift = std::atoi(env); }
        for (int i = 0; i < 3; ++i) { T.map_out[i] = e->map_act_store[i]; }
        auto set = [&](int li, const CUtensorMap& in, __half* out, const __half* residual) {
            mznn::TowerLayer& L = T.layer[li];
            L.map_in = in, L.map_w = NT.convs[li].map_w_tower, L.out = out, L.residual = residual;
            L.bias = reinterpret_cast<const float*>(e->d_blob + NT.convs[li].b_off), L.cin = NT.convs[li].cin, L.relu = NT.convs[li].relu;
            L.cin_off = 0, L.tap_mask = 0x1ff;
            L.out_map = 0, L.res_layer = -1;
            for (int i = 0; i < 3; ++i) {
                if (out == e->act[i]) { L.out_map = i; }
            }
            // the residual of a block's second convolution is the block's input: the output of layer li - 2, or the tower's input rows (written before the launch)
            if (residual != nullptr && li >= 2) { L.res_layer = li - 2; }
        };
        const int base = (NT.has_stem ? 1 : 0);
        const CUtensorMap* act_maps = (e->tower_wide ? e->map_act_wide : e->map_act_ext);
        if (NT.has_stem) { set(0, e->tower_wide ? NT.map_in_wide : NT.map_in_ext, e->act[0], nullptr); }
        int cur = 0;
        for (int b = 0; b < e->nd.num_blocks; ++b) {
            const int tt = (cur + 1) % 3, o = (cur + 2) % 3;
            set(base + 2 * b, act_maps[cur], e->act[tt], nullptr);
            set(base + 1 + 2 * b, act_maps[tt], e->act[o], e->act[cur]);
            cur = o;
        }
        // completion counters: narrow tower one per (layer, 256-row group); wide tower one per (layer, 256-row subgroup incl. the phantom one of an odd tail,
        // block of 64 output channels)
        const int num_groups = (e->tower_wide ? 2 * ((T.num_mtiles + 3) / 4) * (e->cpad / 64) : (T.num_mtiles + 1) / 2);
        NT.done_count = T.num_layers * num_groups;
        if ((rc = e->dalloc(&NT.d_done, static_cast<size_t>(T.num_layers) * num_groups))) { return rc; }
        T.done = NT.d_done;
        T.dbg = nullptr;
        if (const char* env = knob("MZ_DEBUG_TOWER")) {
            if (std::atoi(env) != 0) {
                unsigned long long* buf = nullptr;
                if ((rc = e->dalloc(&buf, static_cast<size_t>(e->num_sms) * 8))) { return rc; }
                T.dbg = buf;
            }
        }
    }
    if (tower_ok) { e->conv_mode = 3; }
    if (e->atari && (rc = alloc_atari(e))) { return rc; }
    if (tower_ok) {
        // the tower's CTAs wait for each other: the whole grid must be able to be resident at once on this device / context (SM-restricted contexts,
        // MPS thread percentages ...). Checked here once; a second kernel competing for the SMs at run time is what the cooperative mode is for
        for (int t = 0; t < e->num_towers; ++t) {
            const mznn::TowerParams& T = *e->tw[t].params;
            const int units = (e->tower_wide ? ((T.num_mtiles + 3) / 4) : ((T.num_mtiles + 1) / 2)) * (e->cpad / e->tower_bn);
            const int clusters = std::min(units, e->tower_sms / 2);
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(clusters * 2), cfg.attrs = attr, cfg.numAttrs = 1;
            int max_clusters = 0;
            cudaError_t err;
            if (e->tower_wide) {
                cfg.blockDim = dim3(mznn::WIDE_THREADS);
                cfg.dynamicSmemBytes = mznn::wide_smem_bytes(e->rows_ext_wide, e->tower_wide_stages);
                err = (e->tower_wide_stages == 10 ? cudaOccupancyMaxActiveClusters(&max_clusters, mznn::conv_tower_wide_kernel<10, false>, &cfg)
                                                 : cudaOccupancyMaxActiveClusters(&max_clusters, mznn::conv_tower_wide_kernel<9, false>, &cfg));
            } else {
                cfg.blockDim = dim3(mznn::TOWER_THREADS);
                cfg.dynamicSmemBytes = mznn::tower_smem_bytes(e->cin_max, e->rows_ext, e->tower_stages, e->tower_bn, e->tower_epi_bufs);
                err = (e->tower_bn == 256 ? cudaOccupancyMaxActiveClusters(&max_clusters, mznn::conv_tower_kernel<256, 4, false>, &cfg)
                       : e->tower_stages == 4 ? cudaOccupancyMaxActiveClusters(&max_clusters, mznn::conv_tower_kernel<128, 4, false>, &cfg)
                       : e->tower_stages == 5 ? cudaOccupancyMaxActiveClusters(&max_clusters, mznn::conv_tower_kernel<128, 5, false>, &cfg)
                                              : cudaOccupancyMaxActiveClusters(&max_clusters, mznn::conv_tower_kernel<128, 8, false>, &cfg));
            }
            if (err != cudaSuccess) {
                (void)cudaGetLastError();
                return fail(MZ_ERR_CUDA, std::string("cudaOccupancyMaxActiveClusters: ") + cudaGetErrorString(err));
            }
            if (max_clusters < clusters) {
                return fail(MZ_ERR_CUDA, "the fused tower needs " + std::to_string(clusters) + " co-resident CTA pairs, this device / context can hold " + std::to_string(max_clusters));
            }
        }
    }
    return MZ_OK;
}

// weights of the Atari representation stages and of the discrete heads into the host copy of the blob
int pack_atari_weights(mz_engine* e, std::vector<uint8_t>& host)
{
    const mz_net_dims& nd = e->nd;
    const int Ch = nd.num_hidden_channels, Ch1 = Ch / 2, hw = e->d.N * e->d.N;
    std::string err;
    const std::string rp = "representation_network.";
    // a plain 3x3 layer: [tap][cout_pad][cin_pad] fp16
    auto plain = [&](const ConvLayer& L, const std::string& cname, const std::string& bname, int cout_real, int cin_real) -> int {
        std::vector<float> bias, w = fold_conv(e, cname, bname, cout_real, cin_real, 3, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        __half* wd = reinterpret_cast<__half*>(host.data() + L.w_off);
        float* bd = reinterpret_cast<float*>(host.data() + L.b_off);
        for (int co = 0; co < cout_real; ++co) {
            bd[co] = bias[co];
            for (int ci = 0; ci < cin_real; ++ci) {
                for (int tap = 0; tap < 9; ++tap) { wd[(static_cast<size_t>(tap) * L.cout + co) * L.cin + ci] = __float2half_rn(w[(static_cast<size_t>(co) * cin_real + ci) * 9 + tap]); }
            }
        }
        return MZ_OK;
    };
    int rc;
    ConvStage &A = e->ast[0], &B = e->ast[1], &Cs = e->ast[2];
    { // conv1: stride 2 over the 32 planes, all four sub-positions in one layer (see plan_blob_atari)
        const ConvLayer& L = A.convs[0];
        std::vector<float> bias, w = fold_conv(e, rp + "conv1", rp + "bn1", Ch1, nd.num_input_channels, 3, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        __half* wd = reinterpret_cast<__half*>(host.data() + L.w_off);
        float* bd = reinterpret_cast<float*>(host.data() + L.b_off);
        const int cin_real = nd.num_input_channels;
        for (int co = 0; co < Ch1; ++co) {
            bd[co] = bias[co];
            for (int ty = 0; ty < 2; ++ty) {
                for (int tx = 0; tx < 2; ++tx) {
                    for (int sub = 0; sub < 4; ++sub) {
                        const int ky = stride2_kernel_index(ty, sub >> 1), kx = stride2_kernel_index(tx, sub & 1);
                        if (ky < 0 || kx < 0) { continue; }
                        for (int c = 0; c < cin_real; ++c) {
                            wd[(static_cast<size_t>(ty * 3 + tx) * L.cout + co) * L.cin + sub * mzat::PLANES + c] = __float2half_rn(w[(static_cast<size_t>(co) * cin_real + c) * 9 + ky * 3 + kx]);
                        }
                    }
                }
            }
        }
    }
    if ((rc = plain(A.convs[1], rp + "residual_blocks1.0.conv1", rp + "residual_blocks1.0.bn1", Ch1, Ch1))) { return rc; }
    if ((rc = plain(A.convs[2], rp + "residual_blocks1.0.conv2", rp + "residual_blocks1.0.bn2", Ch1, Ch1))) { return rc; }
    { // conv2: stride 2 over Ch / 2 channels, one layer per sub-position; the folded bias rides on the first
        std::vector<float> bias, w = fold_conv(e, rp + "conv2", rp + "bn2", Ch, Ch1, 3, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        for (int sub = 0; sub < 4; ++sub) {
            const ConvLayer& L = B.convs[sub];
            __half* wd = reinterpret_cast<__half*>(host.data() + L.w_off);
            float* bd = reinterpret_cast<float*>(host.data() + L.b_off);
            for (int co = 0; co < Ch; ++co) {
                if (sub == 0) { bd[co] = bias[co]; }
                for (int ty = 0; ty < 2; ++ty) {
                    for (int tx = 0; tx < 2; ++tx) {
                        const int ky = stride2_kernel_index(ty, sub >> 1), kx = stride2_kernel_index(tx, sub & 1);
                        if (ky < 0 || kx < 0) { continue; }
                        for (int c = 0; c < Ch1; ++c) {
                            wd[(static_cast<size_t>(ty * 3 + tx) * L.cout + co) * L.cin + c] = __float2half_rn(w[(static_cast<size_t>(co) * Ch1 + c) * 9 + ky * 3 + kx]);
                        }
                    }
                }
            }
        }
    }
    if ((rc = plain(B.convs[4], rp + "residual_blocks2.0.conv1", rp + "residual_blocks2.0.bn1", Ch, Ch))) { return rc; }
    if ((rc = plain(B.convs[5], rp + "residual_blocks2.0.conv2", rp + "residual_blocks2.0.bn2", Ch, Ch))) { return rc; }
    if ((rc = plain(Cs.convs[0], rp + "residual_blocks3.0.conv1", rp + "residual_blocks3.0.bn1", Ch, Ch))) { return rc; }
    if ((rc = plain(Cs.convs[1], rp + "residual_blocks3.0.conv2", rp + "residual_blocks3.0.bn2", Ch, Ch))) { return rc; }
    // heads
    auto raw = [&](const std::string& name, size_t n) -> const float* {
        auto it = e->tensors.find(name);
        if (it == e->tensors.end() || it->second.size() != n) {
            err = "missing or mis-sized tensor " + name;
            return nullptr;
        }
        return it->second.data();
    };
    auto conv1x1 = [&](const std::string& name, int planes, size_t w_off, size_t b_off) -> int {
        std::vector<float> bias, w = fold_conv(e, name + ".conv", name + ".bn", planes, Ch, 1, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        float* wd = reinterpret_cast<float*>(host.data() + w_off);
        for (int o = 0; o < planes; ++o) {
            for (int c = 0; c < Ch; ++c) { wd[static_cast<size_t>(o) * e->cpad + c] = w[static_cast<size_t>(o) * Ch + c]; }
        }
        std::memcpy(host.data() + b_off, bias.data(), sizeof(float) * planes);
        return MZ_OK;
    };
    auto fc_half = [&](const std::string& name, int nout, int nin, int nin_pad, size_t w_off, size_t b_off) -> int { // torch [out][in] -> fp16 [out_pad][in_pad]
        const float *w = raw(name + ".weight", static_cast<size_t>(nout) * nin), *b = raw(name + ".bias", nout);
        if (!w || !b) { return fail(MZ_ERR_ARG, err); }
        __half* wd = reinterpret_cast<__half*>(host.data() + w_off);
        for (int o = 0; o < nout; ++o) {
            for (int i = 0; i < nin; ++i) { wd[static_cast<size_t>(o) * nin_pad + i] = __float2half_rn(w[static_cast<size_t>(o) * nin + i]); }
        }
        std::memcpy(host.data() + b_off, b, sizeof(float) * nout);
        return MZ_OK;
    };
    auto fc_t = [&](const std::string& name, int nout, int nin, size_t w_off, size_t b_off) -> int { // torch [out][in] -> [in][out]
        const float *w = raw(name + ".weight", static_cast<size_t>(nout) * nin), *b = raw(name + ".bias", nout);
        if (!w || !b) { return fail(MZ_ERR_ARG, err); }
        float* wd = reinterpret_cast<float*>(host.data() + w_off);
        for (int o = 0; o < nout; ++o) {
            for (int i = 0; i < nin; ++i) { wd[static_cast<size_t>(i) * nout + o] = w[static_cast<size_t>(o) * nin + i]; }
        }
        std::memcpy(host.data() + b_off, b, sizeof(float) * nout);
        return MZ_OK;
    };
    const int hc = e->dh_planes, dv = nd.discrete_value_size;
    const std::string names[2] = {"prediction_network.value", "dynamics_network.reward_network"};
    for (int h = 0; h < 2; ++h) {
        const int fc1_out = (h == 0 ? nd.num_value_hidden_channels : Ch);
        if ((rc = conv1x1(names[h], hc, e->off_dhead[h][0], e->off_dhead[h][1]))) { return rc; }
        if ((rc = fc_half(names[h] + ".fc1", fc1_out, hc * hw, e->dh_k1, e->off_dhead[h][2], e->off_dhead[h][3]))) { return rc; }
        if ((rc = fc_half(names[h] + ".fc2", dv, fc1_out, e->dh_n1[h], e->off_dhead[h][4], e->off_dhead[h][5]))) { return rc; }
    }
    if ((rc = conv1x1("prediction_network.policy", e->pol_ch, e->off_dhead[2][0], e->off_dhead[2][1]))) { return rc; }
    { // the three 1x1 convolutions as one fp16 weight matrix [plane][channel] for the planes GEMM, with their folded biases and channel sums
        __half* wd = reinterpret_cast<__half*>(host.data() + e->off_planes[0]);
        float *bd = reinterpret_cast<float*>(host.data() + e->off_planes[1]), *sd = reinterpret_cast<float*>(host.data() + e->off_planes[2]);
        const int src[3] = {2, 0, 1}, cnt[3] = {e->pol_ch, hc, hc}; // plane order: policy, value, reward
        int plane = 0;
        for (int k = 0; k < 3; ++k) {
            const float* w = reinterpret_cast<const float*>(host.data() + e->off_dhead[src[k]][0]);
            const float* b = reinterpret_cast<const float*>(host.data() + e->off_dhead[src[k]][1]);
            for (int o = 0; o < cnt[k]; ++o, ++plane) {
                float sum = 0.0f;
                for (int c = 0; c < Ch; ++c) {
                    const __half hv = __float2half_rn(w[static_cast<size_t>(o) * e->cpad + c]);
                    wd[static_cast<size_t>(plane) * e->cpad + c] = hv;
                    sum += __half2float(hv);
                }
                bd[plane] = b[o], sd[plane] = sum;
            }
        }
    }
    if ((rc = fc_t("prediction_network.policy.fc", nd.action_size, e->pol_ch * hw, e->off_dhead[2][2], e->off_dhead[2][3]))) { return rc; }
    return MZ_OK;
}

} // namespace

extern "C" {

const char* mz_last_error(void) { return g_error.c_str(); }

int mz_create(const mz_config* cfg, mz_engine** out)
{
    if (!cfg || !out) { return fail(MZ_ERR_ARG, "null argument"); }
    *out = nullptr;
    if (cfg->game != MZ_GAME_GO && cfg->game != MZ_GAME_TICTACTOE && cfg->game != MZ_GAME_OTHELLO && cfg->game != MZ_GAME_NOGO && cfg->game != MZ_GAME_GOMOKU && cfg->game != MZ_GAME_HEX &&
        cfg->game != MZ_GAME_ATARI && cfg->game != MZ_GAME_KILLALLGO) {
        return fail(MZ_ERR_ARG, "unsupported game");
    }
    if (cfg->game == MZ_GAME_KILLALLGO && cfg->board_size != 7) { return fail(MZ_ERR_ARG, "KillAllGo is played on 7 x 7 (killallgo.h:12,23)"); }
    if (cfg->game == MZ_GAME_KILLALLGO && cfg->muzero) { return fail(MZ_ERR_ARG, "KillAllGo is built for AlphaZero networks (its terminal test needs the position)"); }
    const bool atari = (cfg->game == MZ_GAME_ATARI);
    if (atari && !cfg->muzero) { return fail(MZ_ERR_ARG, "Atari is searched with a MuZero network only (the emulator cannot be copied into the tree)"); }
    if (atari && (cfg->atari_legal_mask == 0 || (cfg->atari_legal_mask >> 18) != 0)) { return fail(MZ_ERR_ARG, "atari_legal_mask must name at least one of the 18 actions"); }
    const int N = (cfg->game == MZ_GAME_TICTACTOE ? 3 : (atari ? 6 : cfg->board_size)); // Atari: side of the hidden state (atari.h:25-26)
    if (N < 2 || N > MZ_MAXN) { return fail(MZ_ERR_ARG, "board_size must be in [2, 19]"); }
    if (cfg->game == MZ_GAME_OTHELLO && (N < 4 || N > 16 || (N & 1))) { return fail(MZ_ERR_ARG, "othello board_size must be even and in [4, 16]"); }
    if (cfg->use_gumbel && cfg->gumbel_sample_size < 2) { return fail(MZ_ERR_ARG, "actor_gumbel_sample_size must be at least 2"); }
    if (cfg->num_games < 1 || cfg->num_simulation < 1) { return fail(MZ_ERR_ARG, "num_games and num_simulation must be positive"); }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { return fail(MZ_ERR_CUDA, "no CUDA device: libmzb200 has no CPU path"); }
    if (cfg->device < 0 || cfg->device >= ndev) { return fail(MZ_ERR_ARG, "bad device ordinal"); }
    CUDA_OK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) { return fail(MZ_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100: libmzb200 is built for sm_100a only"); }

    mz_engine* e = new mz_engine();
    e->cfg = *cfg;
    mz_dims& d = e->d;
    d.game = (cfg->game == MZ_GAME_KILLALLGO ? MZ_GAME_GO : cfg->game), d.killall = (cfg->game == MZ_GAME_KILLALLGO ? 1 : 0), d.N = N, d.A = (cfg->game == MZ_GAME_TICTACTOE ? 9 : ((cfg->game == MZ_GAME_GOMOKU || cfg->game == MZ_GAME_HEX) ? N * N : N * N + 1)), d.C = ((MZ_GO_FAMILY(cfg->game) || cfg->game == MZ_GAME_KILLALLGO) ? 18 : 4);
    d.hex_swap_rule = (cfg->hex_swap_rule != 0);
    d.num_players = 2, d.act_planes = 1, d.value_rescale = (cfg->value_rescale != 0);
    e->atari = atari;
    if (atari) { // atari.h:18-26, mcts.cpp:211-213
        d.A = 18, d.C = mzat::PLANES, d.num_players = 1, d.atari_init_q = 1, d.has_reward = 1, d.act_planes = 18, d.legal_mask = cfg->atari_legal_mask;
    }
    d.gomoku_exactly_five = (cfg->gomoku_exactly_five != 0), d.gomoku_outer_open = (cfg->gomoku_outer_open != 0);
    d.muzero = (cfg->muzero != 0), d.gumbel = (cfg->use_gumbel != 0), d.gumbel_noise = (cfg->gumbel_noise != 0), d.gumbel_m = cfg->gumbel_sample_size;
    d.sigma_visit_c = cfg->gumbel_sigma_visit_c, d.sigma_scale_c = cfg->gumbel_sigma_scale_c;
    if (d.gumbel) { // simulation budgets in the reference's double arithmetic (gumbel_zero.cpp:99,109)
        const double lg = std::log2(static_cast<double>(d.gumbel_m));
        d.gumbel_budget0 = static_cast<int>(std::max(1.0, std::floor(cfg->num_simulation / (lg * d.gumbel_m))));
        for (int l = 0; l < MZ_GUMBEL_LEVELS; ++l) {
            // divisor evaluated in double like the reference: log2(m) * sample_size_ / 2 (an odd sample size divides by x.5, gumbel_zero.cpp:109)
            const int size = (d.gumbel_m >> l);
            d.gumbel_next[l] = (size > 0 ? static_cast<int>(std::floor(cfg->num_simulation / (lg * size / 2))) : 0);
        }
    }
    d.S = cfg->num_simulation, d.B = cfg->num_games;
    // console think() with a selection batch (zero_actor.cpp:129-157): K lanes per tree. Every array is sized for trees x K "games"; the trees use the
    // first `trees` entries of the per-tree arrays, the lanes the sections of the per-leaf arrays (mz_lane_view), the network sees trees x K positions
    const int think_k = (cfg->think_batch_size > 1 ? cfg->think_batch_size : 0);
    if (think_k) {
        if (cfg->use_gumbel || atari || cfg->value_rescale) {
            delete e;
            return fail(MZ_ERR_ARG, "think_batch_size > 1 is built for PUCT selection on the board games (no Gumbel / Atari / value rescale)");
        }
        d.think_k = think_k, e->think_trees = cfg->num_games, d.B = cfg->num_games * think_k;
    }
    d.NP = 1 + (d.S + 1) * d.A; // actor_group.cpp:183, tree.h:66
    d.vb_cap = d.S + 2;
    if (d.NP >= (1 << MZ_LINK_SHIFT)) {
        delete e;
        return fail(MZ_ERR_ARG, "node pool per game exceeds 2^20 nodes");
    }
    d.slots = (N + 1) * (N + 1);
    d.max_hashes = 2 * N * N + 4;
    d.puct_init = cfg->puct_init, d.puct_base = cfg->puct_base, d.discount = cfg->reward_discount, d.komi = cfg->komi, d.eps = cfg->dirichlet_epsilon;

    // Zobrist keys exactly as go.cpp:19-32 draws them: mt19937_64(0): turn key, then per position empty / black / white
    std::mt19937_64 gen(0);
    std::vector<uint64_t> keys(2 * 361);
    const uint64_t turn_key = gen();
    for (int pos = 0; pos < 361; ++pos) {
        (void)gen();
        keys[0 * 361 + pos] = gen();
        keys[1 * 361 + pos] = gen();
    }
    d.turn_key = (cfg->ko_situational ? turn_key : 0);
    // puct_bias[n] = (float)(init + log((1 + n + base) / base)) with the reference's float / double mix (mcts.cpp:57)
    std::vector<float> bias(d.S + 2 + think_k); // under virtual loss a node's total reaches S + K
    for (int n = 0; n < d.S + 2 + think_k; ++n) {
        float t = static_cast<float>(1 + n) + cfg->puct_base;
        t = t / cfg->puct_base;
        bias[n] = static_cast<float>(static_cast<double>(cfg->puct_init) + std::log(static_cast<double>(t)));
    }

    int rc = MZ_OK;
    auto guard = [&](int r) {
        if (r && !rc) { rc = r; }
    };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
        cudaEventCreate(&e->ev2) != cudaSuccess || cudaEventCreate(&e->ev3) != cudaSuccess) {
        delete e;
        return fail(MZ_ERR_CUDA, "stream / event creation failed");
    }
    mz_state& s = e->s;
    const size_t B = d.B, np = B * d.NP, BA = B * d.A;
    e->rows_alloc = static_cast<int>((B * d.slots + mznn::BM - 1) / mznn::BM * mznn::BM);
    guard(e->dalloc(&s.hot, np)), guard(e->dalloc(&s.action, np)), guard(e->dalloc(&s.logit, np)), guard(e->dalloc(&s.value, np));
    guard(e->dalloc(&s.root_noise, BA)), guard(e->dalloc(&s.cursor, B));
    guard(e->dalloc(&s.last_child, np));
    if (!knob("MZ_NO_VIS") && !think_k) { guard(e->dalloc(&s.vis, np)); } // MZ_NO_VIS=1: selection always scans (A/B timing of the visited lists)
    if (think_k) { guard(e->dalloc(&s.vloss, np)), guard(e->dalloc(&s.think_pending, B)); }
    guard(e->dalloc(&s.node_slot, np)), guard(e->dalloc(&s.slot_st, B * (d.S + 1) * 2 * N)), guard(e->dalloc(&s.slot_hash, B * (d.S + 1)));
    guard(e->dalloc(&s.slot_meta, B * (d.S + 1) * 4));
    guard(e->dalloc(&s.root_st, B * 2 * MZ_ROWS)), guard(e->dalloc(&s.root_hist, B * MZ_HIST * 2 * MZ_ROWS)), guard(e->dalloc(&s.root_hash, B));
    guard(e->dalloc(&s.root_meta, B * 4)), guard(e->dalloc(&s.hashes, B * d.max_hashes));
    guard(e->dalloc(&s.spec_len, B));
    guard(e->dalloc(&s.eval_slot, B)), guard(e->dalloc(&s.leaf_parent, B * 2)), guard(e->dalloc(&s.gum_cand, BA)), guard(e->dalloc(&s.gum_meta, B * 4));
    guard(e->dalloc(&e->d_path_actions, B * (d.S + 2)));
    guard(e->dalloc(&s.path, B * (d.S + 2))), guard(e->dalloc(&s.path_len, B)), guard(e->dalloc(&s.leaf_legal, B * MZ_LEGAL_WORDS));
    guard(e->dalloc(&s.leaf_meta, B * 4)), guard(e->dalloc(&s.leaf_score, B));
    guard(e->dalloc(&s.nn_in, static_cast<size_t>(e->rows_alloc) * MZ_NN_CPAD));
    guard(e->dalloc(&s.policy, BA)), guard(e->dalloc(&s.logits, BA)), guard(e->dalloc(&s.nn_value, B));
    guard(e->dalloc(&e->d_rot_all, B * (d.S + 1))), guard(e->dalloc(&e->d_noise, BA));
    guard(e->dalloc(&e->d_bias_table, bias.size())), guard(e->dalloc(&e->d_keys, keys.size()));
    guard(e->dalloc(&e->d_sqrt_table, bias.size()));
    guard(e->dalloc(&e->d_actions, B)), guard(e->dalloc(&e->d_play_out, B * 4)), guard(e->dalloc(&e->d_play_score, B));
    guard(e->dalloc(&e->d_root_info, B * 4)), guard(e->dalloc(&e->d_root_action, BA));
    for (int i = 0; i < 6; ++i) { guard(e->dalloc(&e->d_root_f[i], BA)); }
    guard(e->dalloc(&e->d_feat_f32, B * d.C * N * N));
    if (d.has_reward) { guard(e->dalloc(&s.reward, np)); }
    guard(e->dalloc(&s.nn_reward, B));
    if (d.value_rescale) { guard(e->dalloc(&s.vb_key, B * d.vb_cap)), guard(e->dalloc(&s.vb_cnt, B * d.vb_cap)), guard(e->dalloc(&s.vb_n, B)); }
    guard(e->dalloc(&e->d_root_reward, BA)), guard(e->dalloc(&e->d_bound_size, B)), guard(e->dalloc(&e->d_bound_lo, B)), guard(e->dalloc(&e->d_bound_hi, B));
    if (atari) {
        guard(e->dalloc(&s.at_frames, B * MZ_HIST * MZ_ATARI_FRAME)), guard(e->dalloc(&s.at_meta, B * 16));
        guard(e->dalloc(&e->d_at_frames_in, B * MZ_ATARI_FRAME));
    }
    if (const char* env = knob("MZ_DEBUG_TREE")) {
        if (std::atoi(env) != 0) { guard(e->dalloc(&s.dbg, B * 16)); }
    }
    if (rc) {
        mz_destroy(e);
        return rc;
    }
    if (cudaMemcpyAsync(e->d_bias_table, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
        cudaMemcpyAsync(e->d_keys, keys.data(), keys.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream) != cudaSuccess) {
        mz_destroy(e);
        return fail(MZ_ERR_CUDA, "table upload failed");
    }
    {
        std::vector<double> sq(bias.size());
        for (size_t n = 0; n < sq.size(); ++n) { sq[n] = std::sqrt(static_cast<double>(n)); } // sqrt(total_simulation), mcts.cpp:58
        if (cudaMemcpyAsync(e->d_sqrt_table, sq.data(), sq.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) {
            mz_destroy(e);
            return fail(MZ_ERR_CUDA, "table upload failed");
        }
    }
    s.puct_bias = e->d_bias_table, s.keys = e->d_keys, s.sqrt_table = e->d_sqrt_table;
    // driver entry point for tensor maps (no link-time dependency on libcuda)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        mz_destroy(e);
        return fail(MZ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    }
    e->encode = reinterpret_cast<encode_tiled_fn>(fn);
    // the attribute belongs to the function, not to the engine: only ever raise it (several engines may coexist)
    static size_t step_smem_max = 0;
    if (step_smem_bytes(d) > step_smem_max) { step_smem_max = step_smem_bytes(d); }
    if (const char* env = knob("MZ_CARVEOUT")) {
        // same shared-memory carve-out as the tower kernel, so that tree-step / heads blocks of one engine can be co-resident
        // with the tower CTAs of another engine on the same SM (an SM runs one carve-out configuration at a time)
        if (std::atoi(env) != 0) {
            cudaFuncSetAttribute(k_step, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(mznn::heads_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(mznn::heads_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(mznn::heads_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
    }
    if (cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(step_smem_max)) != cudaSuccess) {
        mz_destroy(e);
        return fail(MZ_ERR_CUDA, "k_step shared memory request refused");
    }
    k_reset<<<d.B, 32, 0, e->stream>>>(d, s, -1);
    e->launches++;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
        std::string msg = cudaGetErrorString(cudaGetLastError());
        mz_destroy(e);
        return fail(MZ_ERR_CUDA, "engine initialisation failed: " + msg);
    }
    *out = e;
    return MZ_OK;
}

void mz_destroy(mz_engine* e)
{
    if (!e) { return; }
    cudaSetDevice(e->cfg.device);
    if (e->stream) { cudaStreamSynchronize(e->stream); }
    for (auto& kv : e->graphs) { cudaGraphExecDestroy(kv.second.exec); }
    for (void* p : e->allocs) { cudaFree(p); }
    delete e->tw[0].params;
    delete e->tw[1].params;
    for (ConvStage& st : e->ast) { delete st.params; }
    if (e->ev0) { cudaEventDestroy(e->ev0); }
    if (e->ev1) { cudaEventDestroy(e->ev1); }
    if (e->ev2) { cudaEventDestroy(e->ev2); }
    if (e->ev3) { cudaEventDestroy(e->ev3); }
    if (e->stream) { cudaStreamDestroy(e->stream); }
    delete e;
}

int mz_action_size(const mz_engine* e) { return e ? e->d.A : MZ_ERR_ARG; }
int mz_num_features(const mz_engine* e) { return e ? (e->atari ? mzat::PLANES * mzat::RES * mzat::RES : e->d.C * e->d.N * e->d.N) : MZ_ERR_ARG; }
int64_t mz_launch_count(const mz_engine* e) { return e ? e->launches : 0; }
int mz_think_steps(const mz_engine* e) { return e ? e->think_steps : MZ_ERR_ARG; }
int mz_set_tower_cooperative(mz_engine* e, int32_t on)
{
    if (!e || !e->net_ready) { return fail(MZ_ERR_STATE, "network not finalized"); }
    if (e->conv_mode != 3) { return fail(MZ_ERR_STATE, "this network does not run through the fused tower"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    for (auto& kv : e->graphs) { cudaGraphExecDestroy(kv.second.exec); } // the launch attribute is part of the captured kernel nodes
    e->graphs.clear();
    e->tower_coop = (on != 0);
    if (!e->tower_coop) { return MZ_OK; }
    // one launch of every tower, outside any capture: does this driver / context accept a cooperative cluster launch of that grid?
    cudaError_t err = cudaSuccess;
    for (int t = 0; t < e->num_towers && err == cudaSuccess; ++t) {
        if (launch_tower(e, t, true, false) != MZ_OK) {
            err = cudaErrorUnknown;
            (void)cudaGetLastError();
        }
    }
    if (err == cudaSuccess) { err = cudaStreamSynchronize(e->stream); }
    for (int t = 0; t < e->num_towers; ++t) { cudaMemsetAsync(e->tw[t].d_done, 0, sizeof(int) * e->tw[t].done_count, e->stream); }
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        e->tower_coop = false;
        return fail(MZ_ERR_CUDA, std::string("cooperative launch of the tower refused: ") + cudaGetErrorString(err));
    }
    return MZ_OK;
}
int mz_tower_is_wide(const mz_engine* e) { return (e && e->net_ready) ? ((e->conv_mode == 3 && e->tower_wide) ? 1 : 0) : MZ_ERR_STATE; }
int mz_tower_is_cooperative(const mz_engine* e) { return (e && e->net_ready) ? ((e->conv_mode == 3 && e->tower_coop) ? 1 : 0) : MZ_ERR_STATE; }
int mz_conv_layers_per_launch(const mz_engine* e) { return (e && e->net_ready) ? (e->conv_mode == 3 ? static_cast<int>(e->tw[0].convs.size()) : 1) : MZ_ERR_STATE; }

int mz_net_configure(mz_engine* e, const mz_net_dims* dims)
{
    if (!e || !dims) { return fail(MZ_ERR_ARG, "null argument"); }
    if (e->atari) { // muzero_atari: 32 x 96 x 96 planes, discrete heads, 18 action planes (atari.h:19-27,68-76)
        if (dims->num_input_channels != mzat::PLANES || dims->input_height != mzat::RES || dims->input_width != mzat::RES || dims->action_size != e->d.A ||
            dims->num_action_feature_channels != 18 || dims->discrete_value_size < 3 || !dims->is_muzero) {
            return fail(MZ_ERR_ARG, "network dimensions do not match the Atari game (muzero_atari network expected)");
        }
        if (dims->num_hidden_channels % 16 != 0) { return fail(MZ_ERR_ARG, "Atari networks need num_hidden_channels divisible by 16"); }
    } else {
    if (dims->discrete_value_size != 1) { return fail(MZ_ERR_ARG, "discrete value heads exist for the Atari (muzero_atari) network only"); }
    if (dims->num_input_channels != e->d.C || dims->input_height != e->d.N || dims->input_width != e->d.N || dims->action_size != e->d.A) {
        return fail(MZ_ERR_ARG, "network dimensions do not match the game");
    }
    if ((dims->action_size + dims->input_height * dims->input_width - 1) / (dims->input_height * dims->input_width) > 3) {
        return fail(MZ_ERR_ARG, "policy head with more than 3 planes is not supported");
    }
    }
    if ((!e->atari && dims->num_input_channels > MZ_NN_CPAD) || dims->num_hidden_channels < 1 || dims->num_blocks < 0) { return fail(MZ_ERR_ARG, "unsupported network size"); }
    if ((dims->is_muzero != 0) != (e->cfg.muzero != 0)) { return fail(MZ_ERR_ARG, "network type (alphazero / muzero) does not match the engine's nn_type_name"); }
    if (!e->atari && dims->is_muzero && dims->num_action_feature_channels != 1) { return fail(MZ_ERR_ARG, "board-game MuZero networks have one action plane"); }
    if (e->d_blob && std::memcmp(&e->nd, dims, sizeof(*dims)) != 0) {
        return fail(MZ_ERR_STATE, "a network of a different shape is already allocated for this engine");
    }
    if (e->d_blob) { // load_model of the next iteration: same shape, new values; buffers, tensor maps and graphs stay valid
        e->tensors.clear();
        return MZ_OK;
    }
    e->nd = *dims;
    e->dims_set = true;
    e->net_ready = false;
    e->tensors.clear();
    return plan_blob(e);
}

int mz_net_set_tensor(mz_engine* e, const char* name, const float* data, int64_t numel)
{
    if (!e || !name || !data || numel < 0) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (!e->dims_set) { return fail(MZ_ERR_STATE, "mz_net_configure first"); }
    e->tensors[name].assign(data, data + numel);
    return MZ_OK;
}

int mz_net_finalize_empty(mz_engine* e)
{
    if (!e || !e->dims_set) { return fail(MZ_ERR_STATE, "mz_net_configure first"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    int rc = alloc_net(e);
    if (rc) { return rc; }
    for (auto& kv : e->graphs) { cudaGraphExecDestroy(kv.second.exec); }
    e->graphs.clear();
    e->net_ready = true;
    return MZ_OK;
}

int mz_net_finalize(mz_engine* e)
{
    NvtxRange nvtx_range("mz_net_finalize");
    if (!e || !e->dims_set) { return fail(MZ_ERR_STATE, "mz_net_configure first"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_net_dims& nd = e->nd;
    const int hw = nd.input_height * nd.input_width, Ch = nd.num_hidden_channels, cp = e->cpad;
    std::vector<uint8_t> host(e->blob.size, 0);
    std::string err;
    // 3x3 convolutions: [tap][cout_pad][cin_pad] fp16, BN folded
    for (int t = 0; t < e->num_towers; ++t) {
    const NetTower& T = e->tw[t];
    for (size_t li = 0; li < T.convs.size(); ++li) {
        const ConvLayer& L = T.convs[li];
        std::string cname, bname;
        int cin_real;
        if (li == 0 && T.has_stem) {
            cname = T.prefix + "conv", bname = T.prefix + "bn", cin_real = T.cin0_real;
        } else {
            const int lb = static_cast<int>(li) - (T.has_stem ? 1 : 0);
            const int blk = lb / 2, which = lb % 2 + 1;
            cname = T.prefix + "residual_blocks." + std::to_string(blk) + ".conv" + std::to_string(which);
            bname = T.prefix + "residual_blocks." + std::to_string(blk) + ".bn" + std::to_string(which);
            cin_real = Ch;
        }
        std::vector<float> bias;
        std::vector<float> w = fold_conv(e, cname, bname, Ch, cin_real, 3, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        __half* wd = reinterpret_cast<__half*>(host.data() + L.w_off);
        float* bd = reinterpret_cast<float*>(host.data() + L.b_off);
        for (int co = 0; co < Ch; ++co) {
            bd[co] = bias[co];
            for (int ci = 0; ci < cin_real; ++ci) {
                for (int tap = 0; tap < 9; ++tap) {
                    wd[(static_cast<size_t>(tap) * L.cout + co) * L.cin + ci] = __float2half_rn(w[(static_cast<size_t>(co) * cin_real + ci) * 9 + tap]);
                }
            }
        }
    }
    }
    const std::string hp = (e->cfg.muzero ? "prediction_network." : ""); // muzero_network.py:44-45
    if (e->atari) {
        int rc = pack_atari_weights(e, host);
        if (rc) { return rc; }
        rc = alloc_net(e);
        if (rc) { return rc; }
        CUDA_OK(cudaMemcpyAsync(e->d_blob, host.data(), host.size(), cudaMemcpyHostToDevice, e->stream));
        CUDA_OK(cudaStreamSynchronize(e->stream));
        e->tensors.clear();
        e->net_ready = true;
        return MZ_OK;
    }
    // heads, fp32
    auto put = [&](int idx, const std::vector<float>& v) { std::memcpy(host.data() + e->off_head[idx], v.data(), v.size() * sizeof(float)); };
    auto raw = [&](const std::string& name, size_t n, std::vector<float>& out) -> bool {
        auto it = e->tensors.find(name);
        if (it == e->tensors.end() || it->second.size() != n) {
            err = "missing or mis-sized tensor " + name;
            return false;
        }
        out = it->second;
        return true;
    };
    {
        std::vector<float> bias, w = fold_conv(e, hp + "policy.conv", hp + "policy.bn", e->pol_ch, Ch, 1, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        std::vector<float> wp(static_cast<size_t>(e->pol_ch) * cp, 0.0f);
        for (int o = 0; o < e->pol_ch; ++o) {
            for (int c = 0; c < Ch; ++c) { wp[static_cast<size_t>(o) * cp + c] = w[static_cast<size_t>(o) * Ch + c]; }
        }
        put(0, wp), put(1, bias);
        std::vector<float> t;
        if (!raw(hp + "policy.fc.weight", static_cast<size_t>(nd.action_size) * e->pol_ch * hw, t)) { return fail(MZ_ERR_ARG, err); }
        {
            const int nin = e->pol_ch * hw;
            std::vector<float> tt(t.size());
            for (int o = 0; o < nd.action_size; ++o) {
                for (int i = 0; i < nin; ++i) { tt[static_cast<size_t>(i) * nd.action_size + o] = t[static_cast<size_t>(o) * nin + i]; }
            }
            put(2, tt);
        }
        if (!raw(hp + "policy.fc.bias", nd.action_size, t)) { return fail(MZ_ERR_ARG, err); }
        put(3, t);
    }
    {
        std::vector<float> bias, w = fold_conv(e, hp + "value.conv", hp + "value.bn", 1, Ch, 1, bias, err);
        if (w.empty()) { return fail(MZ_ERR_ARG, err); }
        std::vector<float> wp(cp, 0.0f);
        for (int c = 0; c < Ch; ++c) { wp[c] = w[c]; }
        put(4, wp), put(5, bias);
        std::vector<float> t;
        if (!raw(hp + "value.fc1.weight", static_cast<size_t>(nd.num_value_hidden_channels) * hw, t)) { return fail(MZ_ERR_ARG, err); }
        {
            const int vh = nd.num_value_hidden_channels;
            std::vector<float> tt(t.size());
            for (int j = 0; j < vh; ++j) {
                for (int i = 0; i < hw; ++i) { tt[static_cast<size_t>(i) * vh + j] = t[static_cast<size_t>(j) * hw + i]; }
            }
            put(6, tt);
        }
        if (!raw(hp + "value.fc1.bias", nd.num_value_hidden_channels, t)) { return fail(MZ_ERR_ARG, err); }
        put(7, t);
        if (!raw(hp + "value.fc2.weight", nd.num_value_hidden_channels, t)) { return fail(MZ_ERR_ARG, err); }
        put(8, t);
        if (!raw(hp + "value.fc2.bias", 1, t)) { return fail(MZ_ERR_ARG, err); }
        put(9, t);
    }
    int rc = alloc_net(e);
    if (rc) { return rc; }
    CUDA_OK(cudaMemcpyAsync(e->d_blob, host.data(), host.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    e->tensors.clear();
    e->net_ready = true;
    return MZ_OK;
}

int mz_net_blob(mz_engine* e, void** device_ptr, int64_t* bytes)
{
    if (!e || !device_ptr || !bytes) { return fail(MZ_ERR_ARG, "null argument"); }
    if (!e->d_blob) { return fail(MZ_ERR_STATE, "network not allocated"); }
    *device_ptr = e->d_blob;
    *bytes = static_cast<int64_t>(e->blob.size);
    return MZ_OK;
}

namespace {

// outputs of the last forward() back to the host; MuZero: also the scaled hidden state (stored in slot 0 by the hooks)
int read_outputs(mz_engine* e, int32_t n, float* policy, float* logits, float* value, float* hidden_out)
{
    const mz_dims& d = e->d;
    if (policy) { CUDA_OK(cudaMemcpyAsync(policy, e->s.policy, sizeof(float) * n * d.A, cudaMemcpyDeviceToHost, e->stream)); }
    if (logits) { CUDA_OK(cudaMemcpyAsync(logits, e->s.logits, sizeof(float) * n * d.A, cudaMemcpyDeviceToHost, e->stream)); }
    if (value) { CUDA_OK(cudaMemcpyAsync(value, e->s.nn_value, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream)); }
    if (hidden_out) {
        const int ch = e->nd.num_hidden_channels;
        mznn::unpack_hidden_kernel<<<148, 256, 0, e->stream>>>(reinterpret_cast<const __half*>(e->s.hid), e->d_hidden_f32, n, ch, d.N, e->cpad, d.S + 1, 0);
        e->launches++;
        CUDA_OK(cudaMemcpyAsync(hidden_out, e->d_hidden_f32, sizeof(float) * n * ch * d.N * d.N, cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

} // namespace

int mz_eval_initial(mz_engine* e, const float* features, int32_t n, float* policy, float* logits, float* value, float* hidden_out)
{
    if (!e || !features || n < 1 || n > e->d.B) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (!e->net_ready) { return fail(MZ_ERR_STATE, "network not finalized"); }
    if (hidden_out && !e->cfg.muzero) { return fail(MZ_ERR_STATE, "hidden states exist only for a muzero network"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    if (e->atari) { // planes [n][32][96][96] -> space-to-depth rows of the first representation stage
        const size_t F = static_cast<size_t>(mzat::PLANES) * mzat::RES * mzat::RES;
        if (!e->d_planes_f32) {
            int rc0 = e->dalloc(&e->d_planes_f32, static_cast<size_t>(d.B) * F);
            if (rc0) { return rc0; }
        }
        CUDA_OK(cudaMemcpyAsync(e->d_planes_f32, features, n * F * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        mzat::pack_input_kernel<true><<<4 * e->num_sms, 256, 0, e->stream>>>(nullptr, nullptr, e->d_planes_f32, e->ast[0].in, n);
        e->launches++;
    } else {
    const size_t F = static_cast<size_t>(d.C) * d.N * d.N;
    CUDA_OK(cudaMemcpyAsync(e->d_feat_f32, features, n * F * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemsetAsync(e->s.nn_in, 0, static_cast<size_t>(e->rows_alloc) * MZ_NN_CPAD * sizeof(uint16_t), e->stream));
    mznn::pack_features_kernel<<<148, 256, 0, e->stream>>>(e->d_feat_f32, reinterpret_cast<__half*>(e->s.nn_in), n, d.C, d.N, d.slots, MZ_NN_CPAD);
    e->launches++;
    }
    if (e->cfg.muzero) { CUDA_OK(cudaMemsetAsync(e->s.eval_slot, 0, sizeof(int32_t) * d.B, e->stream)); } // the hooks keep their hidden states in slot 0
    int rc = forward(e, 0);
    if (rc) { return rc; }
    return read_outputs(e, n, policy, logits, value, hidden_out);
}

int mz_eval_batch(mz_engine* e, const float* features, int32_t n, float* policy, float* logits, float* value)
{
    return mz_eval_initial(e, features, n, policy, logits, value, nullptr);
}

int mz_eval_recurrent(mz_engine* e, const float* hidden, const int32_t* actions, int32_t n, float* policy, float* logits, float* value, float* hidden_out)
{
    if (!e || !hidden || !actions || n < 1 || n > e->d.B) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (!e->net_ready) { return fail(MZ_ERR_STATE, "network not finalized"); }
    if (!e->cfg.muzero) { return fail(MZ_ERR_STATE, "recurrent inference needs a muzero network"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    const int ch = e->nd.num_hidden_channels;
    CUDA_OK(cudaMemcpyAsync(e->d_hidden_f32, hidden, sizeof(float) * n * ch * d.N * d.N, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(e->d_actions, actions, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemsetAsync(e->s.dyn_in, 0, static_cast<size_t>(e->rows_alloc) * d.dyn_c * sizeof(uint16_t), e->stream));
    mznn::pack_hidden_kernel<<<148, 256, 0, e->stream>>>(e->d_hidden_f32, e->d_actions, reinterpret_cast<__half*>(e->s.dyn_in), n, ch, d.N, d.slots, d.dyn_c, d.act_col, d.act_planes);
    e->launches++;
    CUDA_OK(cudaMemsetAsync(e->s.eval_slot, 0, sizeof(int32_t) * d.B, e->stream));
    int rc = forward(e, 1);
    if (rc) { return rc; }
    return read_outputs(e, n, policy, logits, value, hidden_out);
}

int mz_eval_rewards(mz_engine* e, int32_t n, float* reward)
{
    if (!e || !reward || n < 1 || n > e->d.B) { return fail(MZ_ERR_ARG, "bad argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    CUDA_OK(cudaMemcpyAsync(reward, e->s.nn_reward, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return MZ_OK;
}

int mz_atari_observe(mz_engine* e, const int32_t* actions, const uint8_t* frames)
{
    if (!e || !actions || !frames) { return fail(MZ_ERR_ARG, "null argument"); }
    if (!e->atari) { return fail(MZ_ERR_STATE, "not an Atari engine"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const int B = e->d.B;
    CUDA_OK(cudaMemcpyAsync(e->d_actions, actions, sizeof(int32_t) * B, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(e->d_at_frames_in, frames, static_cast<size_t>(B) * MZ_ATARI_FRAME, cudaMemcpyHostToDevice, e->stream));
    k_atari_observe<<<B, 256, 0, e->stream>>>(e->d, e->s, e->d_actions, e->d_at_frames_in);
    e->launches++;
    CUDA_OK(cudaStreamSynchronize(e->stream)); // the caller's buffers are free again
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_get_root_rewards(mz_engine* e, float* reward, int32_t* bound_size, float* bound_lo, float* bound_hi)
{
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const int B = e->d.B;
    k_gather_root_rewards<<<B, 32, 0, e->stream>>>(e->d, e->s, e->d_root_reward, e->d_bound_size, e->d_bound_lo, e->d_bound_hi);
    e->launches++;
    if (reward) { CUDA_OK(cudaMemcpyAsync(reward, e->d_root_reward, sizeof(float) * B * e->d.A, cudaMemcpyDeviceToHost, e->stream)); }
    if (bound_size) { CUDA_OK(cudaMemcpyAsync(bound_size, e->d_bound_size, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream)); }
    if (bound_lo) { CUDA_OK(cudaMemcpyAsync(bound_lo, e->d_bound_lo, sizeof(float) * B, cudaMemcpyDeviceToHost, e->stream)); }
    if (bound_hi) { CUDA_OK(cudaMemcpyAsync(bound_hi, e->d_bound_hi, sizeof(float) * B, cudaMemcpyDeviceToHost, e->stream)); }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_replay_features(mz_engine* e, const int32_t* actions, int32_t max_len, const int32_t* positions, const uint8_t* rotations, int32_t n, float* features_out)
{
    if (!e || !actions || !positions || !features_out || n < 1 || n > e->d.B || max_len < 1) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (e->atari) { return fail(MZ_ERR_STATE, "Atari records carry their observations (OBS tag): there is nothing to replay on the device"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    int32_t *d_act = nullptr, *d_pos = nullptr;
    uint8_t* d_rot = nullptr;
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d_act), sizeof(int32_t) * n * max_len, e->stream));
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d_pos), sizeof(int32_t) * n, e->stream));
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&d_rot), n, e->stream));
    CUDA_OK(cudaMemcpyAsync(d_act, actions, sizeof(int32_t) * n * max_len, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(d_pos, positions, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream));
    if (rotations) { CUDA_OK(cudaMemcpyAsync(d_rot, rotations, n, cudaMemcpyHostToDevice, e->stream)); }
    k_replay_features<<<n, 32, 0, e->stream>>>(d, e->s, d_act, max_len, d_pos, rotations ? d_rot : nullptr, n);
    e->launches++;
    const size_t F = static_cast<size_t>(d.C) * d.N * d.N;
    mznn::unpack_features_kernel<<<148, 256, 0, e->stream>>>(reinterpret_cast<const __half*>(e->s.nn_in), e->d_feat_f32, n, d.C, d.N, d.slots, MZ_NN_CPAD);
    e->launches++;
    CUDA_OK(cudaMemcpyAsync(features_out, e->d_feat_f32, sizeof(float) * n * F, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaFreeAsync(d_act, e->stream));
    CUDA_OK(cudaFreeAsync(d_pos, e->stream));
    CUDA_OK(cudaFreeAsync(d_rot, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_search_leaf(mz_engine* e, int32_t* parent_slot, int32_t* leaf_action, int32_t* path_actions)
{
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    std::vector<int32_t> lp(static_cast<size_t>(d.B) * 2);
    CUDA_OK(cudaMemcpyAsync(lp.data(), e->s.leaf_parent, sizeof(int32_t) * d.B * 2, cudaMemcpyDeviceToHost, e->stream));
    if (path_actions) {
        k_path_actions<<<d.B, 64, 0, e->stream>>>(d, e->s, e->d_path_actions);
        e->launches++;
        CUDA_OK(cudaMemcpyAsync(path_actions, e->d_path_actions, sizeof(int32_t) * d.B * (d.S + 2), cudaMemcpyDeviceToHost, e->stream));
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    for (int g = 0; g < d.B; ++g) {
        if (parent_slot) { parent_slot[g] = lp[g * 2]; }
        if (leaf_action) { leaf_action[g] = lp[g * 2 + 1]; }
    }
    return MZ_OK;
}

int mz_gumbel_best_actions(mz_engine* e, int32_t* actions_out)
{
    if (!e || !actions_out) { return fail(MZ_ERR_ARG, "null argument"); }
    if (!e->cfg.use_gumbel) { return fail(MZ_ERR_STATE, "engine created without actor_use_gumbel"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    k_gumbel_best<<<e->d.B, 32, 0, e->stream>>>(e->d, e->s, e->d_actions);
    e->launches++;
    CUDA_OK(cudaMemcpyAsync(actions_out, e->d_actions, sizeof(int32_t) * e->d.B, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_reset_game(mz_engine* e, int32_t g)
{
    if (!e || g >= e->d.B) { return fail(MZ_ERR_ARG, "bad argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    // stream-ordered like everything else of the engine: several games ending on the same move cost one small launch each and no host round trip
    k_reset<<<(g >= 0 ? 1 : e->d.B), 32, 0, e->stream>>>(e->d, e->s, g);
    e->launches++;
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_play(mz_engine* e, const int32_t* actions, mz_play_result* results)
{
    NvtxRange nvtx_range("mz_play");
    if (!e || !actions || !results) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const int B = e->d.B;
    CUDA_OK(cudaMemcpyAsync(e->d_actions, actions, sizeof(int32_t) * B, cudaMemcpyHostToDevice, e->stream));
    k_play<<<B, 32, 0, e->stream>>>(e->d, e->s, e->d_actions, e->d_play_out, e->d_play_score);
    e->launches++;
    std::vector<int32_t> out(static_cast<size_t>(B) * 4);
    std::vector<float> score(B);
    CUDA_OK(cudaMemcpyAsync(out.data(), e->d_play_out, sizeof(int32_t) * B * 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaMemcpyAsync(score.data(), e->d_play_score, sizeof(float) * B, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    for (int g = 0; g < B; ++g) {
        results[g].applied = out[g * 4 + 0], results[g].terminal = out[g * 4 + 1], results[g].num_legal = out[g * 4 + 2], results[g].turn = out[g * 4 + 3];
        results[g].eval_score = score[g];
    }
    return MZ_OK;
}

int mz_play_max_count(mz_engine* e, int32_t auto_reset, int32_t* actions_out, mz_play_result* results)
{
    NvtxRange nvtx_range("mz_play_max_count");
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const int B = e->d.B;
    k_play_max_count<<<B, 32, 0, e->stream>>>(e->d, e->s, e->d_actions, e->d_play_out, e->d_play_score, auto_reset);
    e->launches++;
    if (!actions_out && !results) { return MZ_OK; } // asynchronous: nothing read back
    std::vector<int32_t> out(static_cast<size_t>(B) * 4);
    std::vector<float> score(B);
    if (actions_out) { CUDA_OK(cudaMemcpyAsync(actions_out, e->d_actions, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, e->stream)); }
    CUDA_OK(cudaMemcpyAsync(out.data(), e->d_play_out, sizeof(int32_t) * B * 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaMemcpyAsync(score.data(), e->d_play_score, sizeof(float) * B, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    if (results) {
        for (int g = 0; g < B; ++g) {
            results[g].applied = out[g * 4 + 0], results[g].terminal = out[g * 4 + 1], results[g].num_legal = out[g * 4 + 2], results[g].turn = out[g * 4 + 3];
            results[g].eval_score = score[g];
        }
    }
    return MZ_OK;
}

int mz_sync(mz_engine* e)
{
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_timer_begin(mz_engine* e)
{
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    CUDA_OK(cudaEventRecord(e->ev2, e->stream));
    return MZ_OK;
}

int mz_timer_end(mz_engine* e, float* device_ms)
{
    if (!e || !device_ms) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    CUDA_OK(cudaEventRecord(e->ev3, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventElapsedTime(device_ms, e->ev2, e->ev3));
    return MZ_OK;
}

int mz_get_roots(mz_engine* e, mz_root_info* info, int32_t* action, float* count, float* mean, float* policy, float* logit, float* noise, float* value)
{
    NvtxRange nvtx_range("mz_get_roots");
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const int B = e->d.B;
    const size_t BA = static_cast<size_t>(B) * e->d.A;
    k_gather_roots<<<B, 32, 0, e->stream>>>(e->d, e->s, e->d_root_info, e->d_root_action, e->d_root_f[0], e->d_root_f[1], e->d_root_f[2], e->d_root_f[3], e->d_root_f[4],
                                            e->d_root_f[5]);
    e->launches++;
    std::vector<float> hinfo(static_cast<size_t>(B) * 4);
    CUDA_OK(cudaMemcpyAsync(hinfo.data(), e->d_root_info, sizeof(float) * B * 4, cudaMemcpyDeviceToHost, e->stream));
    if (action) { CUDA_OK(cudaMemcpyAsync(action, e->d_root_action, sizeof(int32_t) * BA, cudaMemcpyDeviceToHost, e->stream)); }
    float* outs[6] = {count, mean, policy, logit, noise, value};
    for (int i = 0; i < 6; ++i) {
        if (outs[i]) { CUDA_OK(cudaMemcpyAsync(outs[i], e->d_root_f[i], sizeof(float) * BA, cudaMemcpyDeviceToHost, e->stream)); }
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    if (info) {
        for (int g = 0; g < B; ++g) {
            int32_t nc;
            std::memcpy(&nc, &hinfo[g * 4], 4);
            info[g].num_children = nc, info[g].count = hinfo[g * 4 + 1], info[g].mean = hinfo[g * 4 + 2], info[g].value = hinfo[g * 4 + 3];
        }
    }
    return MZ_OK;
}

int mz_search_select(mz_engine* e, const uint8_t* rotations, float* features_out, int32_t* path_len_out)
{
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    if (rotations) { CUDA_OK(cudaMemcpyAsync(e->d_rot_all, rotations, d.B, cudaMemcpyHostToDevice, e->stream)); }
    step(e, STEP_BEFORE, rotations ? e->d_rot_all : nullptr);
    if (features_out && e->atari) { // AtariEnv::getFeatures of every root's observation history (atari.cpp:106-116)
        const size_t F = static_cast<size_t>(mzat::PLANES) * mzat::RES * mzat::RES;
        if (!e->d_planes_f32) {
            int rc0 = e->dalloc(&e->d_planes_f32, static_cast<size_t>(d.B) * F);
            if (rc0) { return rc0; }
        }
        mzat::planes_f32_kernel<<<4 * e->num_sms, 256, 0, e->stream>>>(e->s.at_frames, e->s.at_meta, e->d_planes_f32, d.B);
        e->launches++;
        CUDA_OK(cudaMemcpyAsync(features_out, e->d_planes_f32, sizeof(float) * d.B * F, cudaMemcpyDeviceToHost, e->stream));
    } else if (features_out) {
        const size_t F = static_cast<size_t>(d.C) * d.N * d.N;
        mznn::unpack_features_kernel<<<148, 256, 0, e->stream>>>(reinterpret_cast<const __half*>(e->s.nn_in), e->d_feat_f32, d.B, d.C, d.N, d.slots, MZ_NN_CPAD);
        e->launches++;
        CUDA_OK(cudaMemcpyAsync(features_out, e->d_feat_f32, sizeof(float) * d.B * F, cudaMemcpyDeviceToHost, e->stream));
    }
    if (path_len_out) { CUDA_OK(cudaMemcpyAsync(path_len_out, e->s.path_len, sizeof(int32_t) * d.B, cudaMemcpyDeviceToHost, e->stream)); }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_search_apply(mz_engine* e, const float* policy, const float* logits, const float* value, const float* noise)
{
    return mz_search_apply_reward(e, policy, logits, value, nullptr, noise);
}

int mz_search_apply_reward(mz_engine* e, const float* policy, const float* logits, const float* value, const float* reward, const float* noise)
{
    if (!e || !policy || !logits || !value) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    const size_t BA = static_cast<size_t>(d.B) * d.A;
    if (reward) {
        CUDA_OK(cudaMemcpyAsync(e->s.nn_reward, reward, sizeof(float) * d.B, cudaMemcpyHostToDevice, e->stream));
    } else {
        CUDA_OK(cudaMemsetAsync(e->s.nn_reward, 0, sizeof(float) * d.B, e->stream));
    }
    CUDA_OK(cudaMemcpyAsync(e->s.policy, policy, sizeof(float) * BA, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(e->s.logits, logits, sizeof(float) * BA, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(e->s.nn_value, value, sizeof(float) * d.B, cudaMemcpyHostToDevice, e->stream));
    if (noise) { CUDA_OK(cudaMemcpyAsync(e->d_noise, noise, sizeof(float) * BA, cudaMemcpyHostToDevice, e->stream)); }
    const bool saved = e->noise_enabled;
    e->noise_enabled = (noise != nullptr);
    step(e, STEP_AFTER, nullptr);
    e->noise_enabled = saved;
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

int mz_search_set_inputs(mz_engine* e, const uint8_t* rotations, const float* noise)
{
    NvtxRange nvtx_range("mz_search_set_inputs");
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    if (rotations) { CUDA_OK(cudaMemcpyAsync(e->d_rot_all, rotations, static_cast<size_t>(d.B) * (d.S + 1), cudaMemcpyHostToDevice, e->stream)); }
    if (noise) { CUDA_OK(cudaMemcpyAsync(e->d_noise, noise, sizeof(float) * d.B * d.A, cudaMemcpyHostToDevice, e->stream)); }
    e->rot_enabled = (rotations != nullptr);
    e->noise_enabled = (noise != nullptr);
    return MZ_OK;
}

int mz_search_run(mz_engine* e, int32_t num_evals, float* device_ms)
{
    NvtxRange nvtx_range("mz_search_run");
    if (!e) { return fail(MZ_ERR_ARG, "null argument"); }
    if (!e->net_ready) { return fail(MZ_ERR_STATE, "network not finalized"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const mz_dims& d = e->d;
    if (num_evals <= 0) { num_evals = d.S + 1; }
    if (num_evals > d.S + 1) { return fail(MZ_ERR_ARG, "num_evals exceeds actor_num_simulation + 1"); }
    if (e->cfg.muzero && num_evals != d.S + 1) { return fail(MZ_ERR_ARG, "a muzero search runs whole (the first evaluation is the initial inference): num_evals must be 0 or S + 1"); }
    if (d.think_k) {
        // ZeroActor::think (zero_actor.cpp:36-49): batched steps until every tree holds S + 1 simulations. How many steps that takes depends on how many
        // selections of a step hit the same leaf, so the loop is driven from the host (the console path: one tree, a handful of steps per second of
        // thinking time) instead of being captured whole; a step evaluates at least one new leaf per unfinished tree, so S + 1 steps always suffice
        if (num_evals != d.S + 1) { return fail(MZ_ERR_ARG, "a think() search runs whole: num_evals must be 0 or S + 1"); }
        const int trees = e->think_trees;
        std::vector<float> counts(trees);
        CUDA_OK(cudaEventRecord(e->ev0, e->stream));
        e->think_steps = 0;
        for (int c = 0; c <= d.S; ++c) {
            int rc = step(e, STEP_BEFORE, e->rot_enabled ? e->d_rot_all + static_cast<size_t>(c) * d.B : nullptr);
            if (!rc) { rc = forward(e, (e->cfg.muzero && c > 0) ? 1 : 0, true); } // MuZero: the root's initial inference (a batch of one lane), recurrent below
            if (!rc) { rc = step(e, STEP_AFTER, nullptr); }
            if (rc) { return rc; }
            ++e->think_steps;
            CUDA_OK(cudaMemcpy2DAsync(counts.data(), sizeof(float), e->s.hot, sizeof(mz_hot) * d.NP, sizeof(float), trees, cudaMemcpyDeviceToHost, e->stream));
            CUDA_OK(cudaStreamSynchronize(e->stream));
            bool all_done = true;
            for (int g = 0; g < trees; ++g) { all_done = all_done && (counts[g] >= static_cast<float>(d.S + 1)); }
            if (all_done) { break; }
        }
        CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        CUDA_OK(cudaStreamSynchronize(e->stream));
        CUDA_OK(cudaGetLastError());
        if (device_ms) { CUDA_OK(cudaEventElapsedTime(device_ms, e->ev0, e->ev1)); }
        return MZ_OK;
    }
    const int key = num_evals * 4 + (e->noise_enabled ? 2 : 0) + (e->rot_enabled ? 1 : 0);
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        // capture the whole search: before | (NN, after+before) x (n-1) | NN, after
        cudaGraph_t graph = nullptr;
        const int64_t launches_before = e->launches, memsets_before = e->memsets;
        CUDA_OK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        int rc = MZ_OK;
        for (int c = 0; c < num_evals && !rc; ++c) {
            step(e, (c > 0 ? STEP_AFTER : 0) | STEP_BEFORE, e->rot_enabled ? e->d_rot_all + static_cast<size_t>(c) * d.B : nullptr);
            if (e->atari && c == 0) { // the root's planes: observation history -> space-to-depth rows of the first representation stage
                mzat::pack_input_kernel<false><<<4 * e->num_sms, 256, 0, e->stream>>>(e->s.at_frames, e->s.at_meta, nullptr, e->ast[0].in, d.B);
                e->launches++;
            }
            rc = forward(e, (e->cfg.muzero && c > 0) ? 1 : 0, true); // MuZero: initial inference for the root, recurrent below
        }
        step(e, STEP_AFTER, nullptr);
        cudaError_t cerr = cudaStreamEndCapture(e->stream, &graph);
        const int64_t captured = e->launches - launches_before; // kernel launches recorded into the graph: every launch site counts itself
        e->launches = launches_before;
        if (rc || cerr != cudaSuccess) {
            if (graph) { cudaGraphDestroy(graph); }
            return rc ? rc : fail(MZ_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(cerr));
        }
        size_t num_nodes = 0;
        cudaGraphGetNodes(graph, nullptr, &num_nodes);
        if (static_cast<int64_t>(num_nodes) != captured + (e->memsets - memsets_before)) { // kernel + memset nodes; a mismatch means a launch site does not count itself
            cudaGraphDestroy(graph);
            return fail(MZ_ERR_STATE, "captured graph has " + std::to_string(num_nodes) + " nodes but " + std::to_string(captured) + " launches were counted");
        }
        cudaGraphExec_t exec = nullptr;
        cerr = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (cerr != cudaSuccess) { return fail(MZ_ERR_CUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(cerr)); }
        it = e->graphs.emplace(key, SearchGraph{exec, captured}).first;
    }
    e->launches += it->second.kernels;
    if (!device_ms) { // asynchronous: the caller brackets several calls with mz_timer_begin / mz_timer_end or mz_sync
        CUDA_OK(cudaGraphLaunch(it->second.exec, e->stream));
        return MZ_OK;
    }
    CUDA_OK(cudaEventRecord(e->ev0, e->stream));
    CUDA_OK(cudaGraphLaunch(it->second.exec, e->stream));
    CUDA_OK(cudaEventRecord(e->ev1, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventElapsedTime(device_ms, e->ev0, e->ev1));
    return MZ_OK;
}

int mz_debug_tree_timing(mz_engine* e, uint64_t* out)
{
    // counters accumulated by every tree step since mz_create when MZ_DEBUG_TREE=1 was set (profiling runs only):
    // out [num_games][8] = cycles in {selection, transition, leaf analysis, features, expand+backup}, steps, max path, sum of paths
    if (!e || !out) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (!e->s.dbg) { return fail(MZ_ERR_STATE, "set MZ_DEBUG_TREE=1 before creating the engine"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    const size_t n = static_cast<size_t>(e->d.B) * 16;
    CUDA_OK(cudaMemcpyAsync(out, e->s.dbg, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaMemsetAsync(e->s.dbg, 0, sizeof(unsigned long long) * n, e->stream));
    return MZ_OK;
}

int mz_debug_tower_timing(mz_engine* e, uint64_t* out, int32_t max_ctas)
{
    if (!e || !out) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (!e->tw[0].params || !e->tw[0].params->dbg) { return fail(MZ_ERR_STATE, "set MZ_DEBUG_TOWER=1 before loading the network"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    int rc = forward(e);
    if (rc) { return rc; }
    const int n = (max_ctas < e->num_sms ? max_ctas : e->num_sms);
    CUDA_OK(cudaMemcpyAsync(out, e->tw[0].params->dbg, sizeof(unsigned long long) * 8 * n, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return n;
}

int mz_profile_kernels(mz_engine* e, int32_t iters, float* conv_ms, float* tree_ms, float* heads_ms)
{
    if (!e || iters < 1) { return fail(MZ_ERR_ARG, "bad argument"); }
    if (!e->net_ready) { return fail(MZ_ERR_STATE, "network not finalized"); }
    CUDA_OK(cudaSetDevice(e->cfg.device));
    float ms = 0.0f;
    if (conv_ms && e->conv_mode == 3) { // the whole tower is one launch
        const int which = e->num_towers - 1; // MuZero: the dynamics tower (S of the S + 1 evaluations of a move)
        for (int i = 0; i < 3; ++i) { launch_tower(e, which); }
        CUDA_OK(cudaEventRecord(e->ev0, e->stream));
        for (int i = 0; i < iters; ++i) { launch_tower(e, which); }
        CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        {   // leave the completion counters as a network forward expects them (zero; normally the heads kernel re-zeroes them)
            const NetTower& T = e->tw[which];
            CUDA_OK(cudaMemsetAsync(T.d_done, 0, sizeof(int) * T.done_count, e->stream));
        }
        CUDA_OK(cudaStreamSynchronize(e->stream));
        CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        *conv_ms = ms / iters;
    } else if (conv_ms) {
        const ConvLayer& L = e->tw[0].convs.back();
        for (int i = 0; i < 3; ++i) { conv(e, e->map_act[0], e->map_act_ext[0], L, e->act[1], e->act[2]); }
        CUDA_OK(cudaEventRecord(e->ev0, e->stream));
        for (int i = 0; i < iters; ++i) { conv(e, e->map_act[0], e->map_act_ext[0], L, e->act[1], e->act[2]); }
        CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        CUDA_OK(cudaStreamSynchronize(e->stream));
        CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        *conv_ms = ms / iters;
    }
    if (heads_ms) {
        auto heads = [&]() { return e->atari ? launch_atari_heads(e, e->act[0], true) : launch_heads(e, e->act[0]); };
        for (int i = 0; i < 3; ++i) { heads(); }
        CUDA_OK(cudaEventRecord(e->ev0, e->stream));
        for (int i = 0; i < iters; ++i) { heads(); }
        CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        CUDA_OK(cudaStreamSynchronize(e->stream));
        CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        *heads_ms = ms / iters;
    }
    if (tree_ms) {
        // selection + transition + leaf analysis on the trees as they stand (STEP_BEFORE does not modify them)
        for (int i = 0; i < 3; ++i) { step(e, STEP_BEFORE, nullptr); }
        CUDA_OK(cudaEventRecord(e->ev0, e->stream));
        for (int i = 0; i < iters; ++i) { step(e, STEP_BEFORE, nullptr); }
        CUDA_OK(cudaEventRecord(e->ev1, e->stream));
        CUDA_OK(cudaStreamSynchronize(e->stream));
        CUDA_OK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        *tree_ms = ms / iters;
    }
    CUDA_OK(cudaGetLastError());
    return MZ_OK;
}

} // extern "C"
