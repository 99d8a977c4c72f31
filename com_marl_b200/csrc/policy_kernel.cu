// policy_kernel.cu — fused CommCategoricalMLPPolicy forward (+ action sampling) for sm_100a, fp32.
//
// Reference formula (comm_categorical_mlp_policy.py:48-119, comm_base_net.py:80-108,
// attention_module.py:38-49, graph_conv_module.py:51-72, categorical_mlp_module.py:64-80):
//   h  = tanh(obs W1^T + b1)            D   -> 128      encoder._layers.0
//   E  = tanh(h   W2^T + b2)            128 -> 64       encoder._output_layers.0 (output_nonlinearity=tanh)
//   M  = softmax_j((E Wa^T) E^T)        per env, n x n  attention_layer ('general')
//   A_l = M * Range * chan_l ; A_l /= (sum_j A_l + 1e-12)
//   H_{l+1} = tanh(A_l (H_l Wg_l) + bg_l),  H_0 = E      gcn_layers.l   (Wg is (in,out))
//   X  = E + H_L                                         residual
//   logits = head(X): tanh 64->128, tanh 128->64, tanh 64->32, linear 32->5
//   probs  = softmax(logits) * avail / sum ; action ~ Categorical(probs) | argmax
//
// One CTA (256 threads) owns a tile of 64 agent rows made of WHOLE environments (n <= 64), so the
// per-env attention / graph-convolution stay inside the tile and every activation lives in shared
// memory from the observation load to the sampled action: HBM traffic is the obs read plus the
// probs / action (and optional logits / attention) writes.  All nine matrix products run through one
// register-tiled 64-row microkernel (4 rows x N/16 columns per thread, operands k-major so that both
// operand loads are 128-bit).  The per-env n x n products are evaluated as dense 64 x 64 tile products
// whose off-env entries are exact zeros.  Weights (<= 198 KB) are read through the read-only path and
// stay L1/L2 resident.  Exact fp32 FFMA: this is the variant that meets the 1e-5 logit tolerance; a
// tcgen05 variant needs a stated tolerance.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "commarl_b200.h"
#include "common.cuh"

namespace cm {

static constexpr int kTile = 64;        // agent rows per CTA tile
static constexpr int kPitch = 68;       // floats per k-major row (64 + 4: keeps float4 alignment, spreads banks)
static constexpr int kThreads = 256;
static constexpr int kH1 = 128, kE = 64, kC1 = 128, kC2 = 64, kC3 = 32;

struct Blob {   // float offsets into the weight blob (commarl_b200.h)
    int enc_w1, enc_b1, enc_w2, enc_b2, att_w, gcn_w, gcn_b, head_w1, head_b1, head_w2, head_b2, head_w3, head_b3,
        head_w4, head_b4, total;
};

__host__ __device__ inline Blob blob_layout(int D, int L)
{
    Blob o;
    int p = 0;
    o.enc_w1 = p; p += D * kH1;
    o.enc_b1 = p; p += kH1;
    o.enc_w2 = p; p += kH1 * kE;
    o.enc_b2 = p; p += kE;
    o.att_w = p; p += kE * kE;
    o.gcn_w = p; p += L * kE * kE;
    o.gcn_b = p; p += L * kE;
    o.head_w1 = p; p += kE * kC1;
    o.head_b1 = p; p += kC1;
    o.head_w2 = p; p += kC1 * kC2;
    o.head_b2 = p; p += kC2;
    o.head_w3 = p; p += kC2 * kC3;
    o.head_b3 = p; p += kC3;
    o.head_w4 = p; p += kC3 * CM_ACTIONS;
    o.head_b4 = p; p += CM_ACTIONS;
    o.total = p;
    return o;
}

struct PolicyArgs {
    cm_policy_desc d;
    cm_policy_io io;
    int envs_per_tile;
    int64_t n_tiles;
};

// ------------------------------------------------------------------------------------------------
// 64-row register-tiled product: acc[r][q] = sum_k At[k][4ty + r] * Bm[k][col(q)]
//   At : shared, k-major, pitch kPitch.      Bm : global weights (k-major, ld = N) or shared (pitch ldb)
//   thread (tx = tid & 15, ty = tid >> 4) owns rows 4ty..4ty+3 and columns
//   N=128: 4tx..4tx+3 and 64+4tx..64+4tx+3 | N=64: 4tx..4tx+3 | N=32: 2tx, 2tx+1
// ------------------------------------------------------------------------------------------------
template <int N, bool kGlobalB>
__device__ __forceinline__ void gemm64(const float *__restrict__ At, const float *__restrict__ Bm, int ldb, int K,
                                       float (&acc)[4][N / 16])
{
    constexpr int CN = N / 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < CN; ++q) acc[r][q] = 0.0f;
    const float *ap = At + 4 * ty;
    const float *bp = Bm + (CN == 2 ? 2 * tx : 4 * tx);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a4 = *reinterpret_cast<const float4 *>(ap + k * kPitch);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float b[CN];
        if constexpr (CN == 2) {
            float2 v;
            if constexpr (kGlobalB) v = __ldg(reinterpret_cast<const float2 *>(bp + (size_t)k * ldb));
            else v = *reinterpret_cast<const float2 *>(bp + k * ldb);
            b[0] = v.x; b[1] = v.y;
        } else {
#pragma unroll
            for (int h = 0; h < CN / 4; ++h) {
                float4 v;
                if constexpr (kGlobalB) v = __ldg(reinterpret_cast<const float4 *>(bp + (size_t)k * ldb + 64 * h));
                else v = *reinterpret_cast<const float4 *>(bp + k * ldb + 64 * h);
                b[4 * h + 0] = v.x; b[4 * h + 1] = v.y; b[4 * h + 2] = v.z; b[4 * h + 3] = v.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < CN; ++q) acc[r][q] = fmaf(a[r], b[q], acc[r][q]);
    }
}

template <int N>
__device__ __forceinline__ int col_of(int q)
{
    const int tx = threadIdx.x & 15;
    if constexpr (N == 32) return 2 * tx + q;
    else return (q < 4) ? 4 * tx + q : 64 + 4 * tx + (q - 4);
}

// store the thread's 4 x CN block k-major (transposed) for the next product: outT[col][4ty..4ty+3],
// optionally adding a bias (global) and applying tanh
template <int N, bool kBias, bool kTanh>
__device__ __forceinline__ void store_kmajor(float *__restrict__ outT, const float (&acc)[4][N / 16],
                                             const float *__restrict__ bias)
{
    const int ty = threadIdx.x >> 4;
#pragma unroll
    for (int q = 0; q < N / 16; ++q) {
        const int c = col_of<N>(q);
        const float bv = kBias ? __ldg(bias + c) : 0.0f;
        float4 v = make_float4(acc[0][q] + bv, acc[1][q] + bv, acc[2][q] + bv, acc[3][q] + bv);
        if (kTanh) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
        *reinterpret_cast<float4 *>(outT + c * kPitch + 4 * ty) = v;
    }
}

// store row-major out[row][col] (pitch kPitch) — the layout a product consumes as its B operand (N = 64)
__device__ __forceinline__ void store_rowmajor(float *__restrict__ out, const float (&acc)[4][4])
{
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4 *>(out + (4 * ty + r) * kPitch + 4 * tx) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) policy_small_kernel(const PolicyArgs A)
{
    extern __shared__ __align__(16) float sm[];
    float *bufA = sm;                              // 128 x pitch
    float *bufB = bufA + 128 * kPitch;             // 128 x pitch
    float *bufE = bufB + 128 * kPitch;             // 64 x pitch: E^T, kept for attention keys and the residual
    float *logit_s = bufE + kE * kPitch;           // 64 x 5
    const cm_policy_desc &d = A.d;
    const cm_policy_io &io = A.io;
    const int n = d.n_agents, D = d.obs_dim, L = d.n_layers, W = (n + 31) >> 5;
    const Blob o = blob_layout(D, L);
    const float *__restrict__ wts = io.weights;
    const int tid = threadIdx.x;

    for (int64_t tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        const int64_t env0 = tile * A.envs_per_tile;
        const int envs = (int)min((int64_t)A.envs_per_tile, io.n_envs - env0);
        const int rows = envs * n;                 // valid rows of this tile
        const int64_t row0 = env0 * n;             // first global agent row
        __syncthreads();
        // ---- 0. observations -> bufA as obs^T[d][r]; rows beyond `rows` are zero so that they stay finite ----
        {
            const float *src = io.obs + row0 * D;
            for (int e = tid; e < kTile * D; e += kThreads) {
                const int r = e / D, c = e - r * D;
                bufA[c * kPitch + r] = (r < rows) ? __ldg(src + e) : 0.0f;
            }
        }
        __syncthreads();
        // ---- 1. h = tanh(obs W1^T + b1) -> bufB[128] ----
        {
            float acc[4][8];
            gemm64<128, true>(bufA, wts + o.enc_w1, kH1, D, acc);
            store_kmajor<128, true, true>(bufB, acc, wts + o.enc_b1);
        }
        __syncthreads();
        // ---- 2. E = tanh(h W2^T + b2) -> bufE ----
        {
            float acc[4][4];
            gemm64<64, true>(bufB, wts + o.enc_w2, kE, kH1, acc);
            store_kmajor<64, true, true>(bufE, acc, wts + o.enc_b2);
        }
        __syncthreads();
        // ---- 3. Q = E Wa^T -> bufA[0..63] ----
        float *QT = bufA, *HW = bufA + 64 * kPitch, *MT = bufB, *HT = bufB + 64 * kPitch;
        {
            float acc[4][4];
            gemm64<64, true>(bufE, wts + o.att_w, kE, kE, acc);
            store_kmajor<64, false, false>(QT, acc, nullptr);
        }
        __syncthreads();
        // ---- 4. S = Q E^T (64 x 64 tile product; only the per-env diagonal blocks are used) -> MT[j][i] ----
        {
            float acc[4][4];
            gemm64<64, false>(QT, bufE, kPitch, kE, acc);
            store_kmajor<64, false, false>(MT, acc, nullptr);
        }
        __syncthreads();
        // ---- 5. row softmax over the agents of the same env; everything else in the row becomes 0 ----
        if (tid < kTile) {
            const int i = tid;
            if (i < rows) {
                const int j0 = (i / n) * n;
                float mx = -INFINITY;
                for (int j = j0; j < j0 + n; ++j) mx = fmaxf(mx, MT[j * kPitch + i]);
                float sum = 0.0f;
                for (int j = j0; j < j0 + n; ++j) { const float e = expf(MT[j * kPitch + i] - mx); MT[j * kPitch + i] = e; sum += e; }
                for (int j = 0; j < kTile; ++j) {
                    const bool in = (j >= j0) && (j < j0 + n);
                    MT[j * kPitch + i] = in ? MT[j * kPitch + i] / sum : 0.0f;
                }
            } else {
                for (int j = 0; j < kTile; ++j) MT[j * kPitch + i] = 0.0f;
            }
        }
        __syncthreads();
        if (io.attention) {   // agent_infos['attention_weights']: the UNMASKED softmax (comm_base_net.py:93)
            float *dst = io.attention + row0 * n;
            for (int e = tid; e < rows * n; e += kThreads) {
                const int r = e / n, jl = e - r * n;
                dst[e] = MT[((r / n) * n + jl) * kPitch + r];
            }
        }
        // ---- 6. graph convolutions ----
        for (int l = 0; l < L; ++l) {
            const float *Hin = (l == 0) ? bufE : HT;
            {
                float acc[4][4];
                gemm64<64, true>(Hin, wts + o.gcn_w + l * kE * kE, kE, kE, acc);   // H_l Wg_l  (Wg is (in,out))
                store_rowmajor(HW, acc);
            }
            // A_l = M * Range * chan_l, renormalised with eps = 1e-12 (comm_base_net.py:101-103) -> QT as A^T[j][i]
            if (tid < kTile) {
                const int i = tid;
                float *AT = QT;
                if (i < rows) {
                    const int el = i / n, il = i - el * n, j0 = el * n;
                    const int64_t env = env0 + el;
                    uint32_t m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu;
                    if (io.adj_bits) {
                        const uint32_t *p = io.adj_bits + (env * n + il) * W;
                        m0 &= __ldg(p);
                        if (W > 1) m1 &= __ldg(p + 1);
                    }
                    if (io.chan_bits) {
                        const uint32_t *p = io.chan_bits + ((env * L + l) * n + il) * W;
                        m0 &= __ldg(p);
                        if (W > 1) m1 &= __ldg(p + 1);
                    }
                    float sum = 0.0f;
                    for (int jl = 0; jl < n; ++jl) {
                        const uint32_t bit = ((jl < 32 ? m0 : m1) >> (jl & 31)) & 1u;
                        sum += bit ? MT[(j0 + jl) * kPitch + i] : 0.0f;
                    }
                    const float den = sum + 1e-12f;
                    for (int j = 0; j < kTile; ++j) {
                        const int jl = j - j0;
                        const bool in = (jl >= 0) && (jl < n) && ((((jl < 32 ? m0 : m1) >> (jl & 31)) & 1u) != 0);
                        AT[j * kPitch + i] = in ? MT[j * kPitch + i] / den : 0.0f;
                    }
                } else {
                    for (int j = 0; j < kTile; ++j) AT[j * kPitch + i] = 0.0f;
                }
            }
            __syncthreads();
            {
                float acc[4][4];
                gemm64<64, false>(QT, HW, kPitch, kTile, acc);                       // A_l (H_l Wg_l)
                store_kmajor<64, true, true>(HT, acc, wts + o.gcn_b + l * kE);       // tanh(. + b)
            }
            __syncthreads();
        }
        // ---- 7. residual X = E + H_L (comm_categorical_mlp_policy.py:74-77) ----
        if (d.residual) {
            for (int e = tid; e < kE * kTile; e += kThreads) {
                const int k = e >> 6, r = e & 63;
                HT[k * kPitch + r] += bufE[k * kPitch + r];
            }
            __syncthreads();
        }
        // ---- 8..10. categorical head 64 -> 128 -> 64 -> 32 (tanh) ----
        {
            float acc[4][8];
            gemm64<128, true>(HT, wts + o.head_w1, kC1, kE, acc);
            store_kmajor<128, true, true>(bufA, acc, wts + o.head_b1);
        }
        __syncthreads();
        {
            float acc[4][4];
            gemm64<64, true>(bufA, wts + o.head_w2, kC2, kC1, acc);
            store_kmajor<64, true, true>(bufB, acc, wts + o.head_b2);
        }
        __syncthreads();
        {
            float acc[4][2];
            gemm64<32, true>(bufB, wts + o.head_w3, kC3, kC2, acc);
            store_kmajor<32, true, true>(bufB + 64 * kPitch, acc, wts + o.head_b3);
        }
        __syncthreads();
        // ---- 11. logits = x W4^T + b4 ----
        for (int e = tid; e < kTile * CM_ACTIONS; e += kThreads) {
            const int r = e / CM_ACTIONS, a = e - r * CM_ACTIONS;
            const float *x = bufB + 64 * kPitch + r;
            float s = 0.0f;
#pragma unroll 8
            for (int k = 0; k < kC3; ++k) s = fmaf(x[k * kPitch], __ldg(wts + o.head_w4 + k * CM_ACTIONS + a), s);
            logit_s[e] = s + __ldg(wts + o.head_b4 + a);
        }
        __syncthreads();
        // ---- 12. softmax, availability mask, renormalise, sample ----
        if (tid < rows) {
            const int r = tid;
            const int64_t g = row0 + r;
            float lg[CM_ACTIONS], pr[CM_ACTIONS];
            float mx = -INFINITY;
#pragma unroll
            for (int a = 0; a < CM_ACTIONS; ++a) { lg[a] = logit_s[r * CM_ACTIONS + a]; mx = fmaxf(mx, lg[a]); }
            float sum = 0.0f;
#pragma unroll
            for (int a = 0; a < CM_ACTIONS; ++a) { pr[a] = expf(lg[a] - mx); sum += pr[a]; }
            const uint32_t av = io.avail_bits ? io.avail_bits[g] : 0x1Fu;
            float msum = 0.0f;
#pragma unroll
            for (int a = 0; a < CM_ACTIONS; ++a) { pr[a] = ((av >> a) & 1u) ? pr[a] / sum : 0.0f; msum += pr[a]; }
#pragma unroll
            for (int a = 0; a < CM_ACTIONS; ++a) pr[a] = pr[a] / msum;
            if (io.logits) for (int a = 0; a < CM_ACTIONS; ++a) io.logits[g * CM_ACTIONS + a] = lg[a];
            if (io.probs) for (int a = 0; a < CM_ACTIONS; ++a) io.probs[g * CM_ACTIONS + a] = pr[a];
            if (io.actions) {
                int act;
                if (d.greedy) {                    // np.argmax: first maximum (:111-112)
                    act = 0;
                    for (int a = 1; a < CM_ACTIONS; ++a) if (pr[a] > pr[act]) act = a;
                } else {                           // inverse CDF, sequential fp32 cumulative sum (stream spec)
                    const int el = r / n, il = r - el * n;
                    float u;
                    if (io.sample_u) u = io.sample_u[g];
                    else {
                        const int64_t env = env0 + el;
                        const uint4 blk = philox4x32_10(
                            make_uint4((uint32_t)(d.env_id0 + env), io.tick[env], kStreamAct | (io.episode[env] << 8), (uint32_t)(il >> 2)),
                            make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
                        const uint32_t w = (il & 3) == 0 ? blk.x : ((il & 3) == 1 ? blk.y : ((il & 3) == 2 ? blk.z : blk.w));
                        u = u24(w);
                    }
                    int last = 4;
                    for (int a = 0; a < CM_ACTIONS; ++a) if (pr[a] > 0.0f) last = a;
                    act = -1;
                    float c = 0.0f;
                    for (int a = 0; a < CM_ACTIONS; ++a) { c += pr[a]; if (act < 0 && u < c) act = a; }
                    if (act < 0) act = last;
                }
                io.actions[g] = (int8_t)act;
            }
        }
    }
}

static size_t small_smem_bytes() { return (size_t)(2 * 128 * kPitch + kE * kPitch + kTile * CM_ACTIONS) * sizeof(float); }

}  // namespace cm

extern "C" size_t cm_policy_blob_floats(int32_t obs_dim, int32_t n_layers)
{
    return (size_t)cm::blob_layout(obs_dim, n_layers).total;
}

extern "C" int cm_policy_forward(const cm_policy_desc *desc, const cm_policy_io *io, cm_stream_t stream)
{
    using namespace cm;
    if (!desc || !io || !io->weights || !io->obs) return CM_EINVAL;
    if (!io->probs && !io->actions && !io->logits && !io->attention) return CM_EINVAL;
    if (desc->n_agents < 1 || desc->obs_dim < 1 || desc->obs_dim > 128) return CM_EUNSUPPORTED;
    if (desc->n_layers < 1 || desc->n_layers > CM_MAX_LAYERS) return CM_EUNSUPPORTED;
    if (io->actions && !desc->greedy && !io->sample_u && (!io->tick || !io->episode)) return CM_EINVAL;
    if (io->n_envs < 0) return CM_EINVAL;
    if (io->n_envs == 0) return CM_OK;
    if (desc->n_agents > kTile) return CM_EUNSUPPORTED;   // large-team variant: see policy_large (round 2)
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ENODEVICE);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ECUDA);
    PolicyArgs A;
    A.d = *desc;
    A.io = *io;
    A.envs_per_tile = kTile / desc->n_agents;
    A.n_tiles = (io->n_envs + A.envs_per_tile - 1) / A.envs_per_tile;
    const size_t smem = small_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(policy_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, CM_ECUDA);
    int ctas_per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, policy_small_kernel, kThreads, smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int64_t cap = (int64_t)sms * ctas_per_sm;   // persistent: a whole number of CTAs per SM
    const int grid = (int)(A.n_tiles < cap ? A.n_tiles : cap);
    policy_small_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, CM_ECUDA);
    return CM_OK;
}
