// policy_attn_kernel.cu — attention + graph convolutions of CommCategoricalMLPPolicy for LARGE teams (64 < n <= 256).
//
// Middle stage of the large-team pipeline (policy_tc_kernel.cu holds the row-wise halves on the tensor cores):
//   1. policy_tc_kernel<mode Enc> : obs rows -> E, Q = E Wa^T, H_0 Wg_0          (rows of 64 floats in global scratch)
//   2. policy_attn_kernel (here)  : per env  M = softmax_j(Q E^T),  A_l = M * Range * chan_l / (sum + 1e-12),
//                                   H_{l+1} = tanh(A_l (H_l Wg_l) + bg_l),  X = E + H_L         (comm_base_net.py:80-108,
//                                   attention_module.py:38-49, graph_conv_module.py:51-72)
//   3. policy_tc_kernel<mode Head>: X rows -> logits / probs / actions
// The n x n pieces are exact fp32 on the CUDA cores.  One CTA of 8 warps owns an environment; the keys E and the values
// H_l Wg_l of the whole team sit in shared memory (row-major, <= 256 rows), the query rows are processed in blocks of
// 8 * RT rows, and EVERYTHING between the scores and the next layer's rows is private to a warp: warp w owns RT query
// rows, computes their scores against all keys with an RT x KT register tile (lane = key mod 32), takes the softmax and
// the masked sums with warp shuffles, parks the un-normalised masked attention rows in its own slice of shared memory
// and multiplies them with the values (lane = 4 output columns x half of the keys).  No CTA barrier inside a layer.
// The scores are recomputed per layer (64 n^2 FMAs, as many as the aggregation) instead of keeping n x n floats per env.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "commarl_b200.h"
#include "common.cuh"
#include "policy_layout.cuh"

namespace cm {

static constexpr int kAThreads = 256, kAWarps = 8;
static constexpr int kEPitch = 68;        // floats per key row: 272 bytes, so that 8 lanes x 16 bytes touch 32 distinct banks

struct AttnArgs {
    cm_policy_desc d;
    const float *weights;                 // fp32 blob: Wg_l (l >= 1) and the graph-convolution biases
    const uint32_t *adj_bits, *chan_bits;
    float *attention;                     // [B][n][n] or NULL
    const float *scr_e;                   // [B n][64]  E
    float *scr_q;                         // [B n][64]  in: Q rows, out: X rows (in place, row by row)
    float *scr_hw;                        // [B n][64]  in: H_0 Wg_0; rewritten in place with H_l Wg_l of the later layers
    int64_t n_envs;
};

template <int RT, int KT>
__global__ void __launch_bounds__(kAThreads, (RT * KT <= 30) ? 2 : 1) policy_attn_kernel(const AttnArgs A)
{
    constexpr int NK = 32 * KT, AP = NK + 4;
    extern __shared__ __align__(16) float sm[];
    float *Es = sm;                              // [NK][68]   keys (rows >= n are zero)
    float *HWs = Es + NK * kEPitch;              // [NK][64]   values of the current layer (rows >= n are zero)
    float *As = HWs + NK * 64;                   // [8][RT][AP] per-warp: query rows, then attention rows, then H rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *Aw = As + warp * RT * AP;
    const int n = A.d.n_agents, L = A.d.n_layers, W = (n + 31) >> 5;
    const Blob o = blob_layout(A.d.obs_dim, L);
    const int block_rows = kAWarps * RT, nb = (n + block_rows - 1) / block_rows;
    const int half = lane >> 4, cl = (lane & 15) << 2;
    const int n_chunks = (n + 3) >> 2;           // key chunks of 4 (attention values / value rows beyond n are zero)

    for (int64_t env = blockIdx.x; env < A.n_envs; env += gridDim.x) {
        const size_t r_env = (size_t)env * n;
        __syncthreads();                         // the previous env's keys / values are dead
        for (int e = tid; e < NK * 16; e += kAThreads) {
            const int j = e >> 4, c = (e & 15) << 2;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (j < n) v = __ldcg(reinterpret_cast<const float4 *>(A.scr_e + (r_env + j) * 64 + c));
            *reinterpret_cast<float4 *>(Es + j * kEPitch + c) = v;
        }
        for (int l = 0; l < L; ++l) {
            if (l) __syncthreads();              // every warp has written its H_l Wg_l rows and is done with the old values
            for (int e = tid; e < NK * 16; e += kAThreads) {
                const int j = e >> 4, c = (e & 15) << 2;
                float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (j < n) v = __ldcg(reinterpret_cast<const float4 *>(A.scr_hw + (r_env + j) * 64 + c));
                *reinterpret_cast<float4 *>(HWs + j * 64 + c) = v;
            }
            __syncthreads();
            const float *bias = A.weights + o.gcn_b + l * kE;
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + cl));
            for (int rb = 0; rb < nb; ++rb) {
                const int i0 = rb * block_rows + warp * RT;          // this warp's first query row
                if (i0 >= n) break;                                  // warp-uniform: nothing left for this warp
                // ---- query rows -> the warp's slice ----
                __syncwarp();
                for (int e = lane; e < RT * 16; e += 32) {
                    const int r = e >> 4, c = (e & 15) << 2;
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (i0 + r < n) v = __ldcg(reinterpret_cast<const float4 *>(A.scr_q + (r_env + i0 + r) * 64 + c));
                    *reinterpret_cast<float4 *>(Aw + r * AP + c) = v;
                }
                __syncwarp();
                // ---- scores: s[r][t] = Q[i0 + r] . E[lane + 32 t] ----
                float s[RT][KT];
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int t = 0; t < KT; ++t) s[r][t] = 0.0f;
#pragma unroll 2
                for (int k = 0; k < kE; k += 4) {
                    float4 q[RT];
#pragma unroll
                    for (int r = 0; r < RT; ++r) q[r] = *reinterpret_cast<const float4 *>(Aw + r * AP + k);
#pragma unroll
                    for (int t = 0; t < KT; ++t) {
                        const float4 ev = *reinterpret_cast<const float4 *>(Es + (lane + 32 * t) * kEPitch + k);
#pragma unroll
                        for (int r = 0; r < RT; ++r)
                            s[r][t] = fmaf(q[r].w, ev.w, fmaf(q[r].z, ev.z, fmaf(q[r].y, ev.y, fmaf(q[r].x, ev.x, s[r][t]))));
                    }
                }
                __syncwarp();                                        // every lane has read the query rows
                // ---- softmax over the keys (attention_module.py:44-49), mask, un-normalised attention rows ----
                float den[RT];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const int i = i0 + r;
                    float mx = -INFINITY;
#pragma unroll
                    for (int t = 0; t < KT; ++t)
                        if (lane + 32 * t < n) mx = fmaxf(mx, s[r][t]);
                    mx = warp_max(mx);
                    float sum = 0.0f;
#pragma unroll
                    for (int t = 0; t < KT; ++t) {
                        s[r][t] = lane + 32 * t < n ? __expf(s[r][t] - mx) : 0.0f;
                        sum += s[r][t];
                    }
                    sum = warp_sumf(sum);
                    uint32_t mw = 0u;                                // lane t holds word t of the row's neighbour mask
                    if (i < n && lane < W) {
                        mw = 0xFFFFFFFFu;
                        if (A.adj_bits) mw &= __ldg(A.adj_bits + (r_env + i) * W + lane);
                        if (A.chan_bits) mw &= __ldg(A.chan_bits + (((size_t)env * L + l) * n + i) * W + lane);
                    }
                    float dsum = 0.0f;
#pragma unroll
                    for (int t = 0; t < KT; ++t) {
                        const float p = s[r][t] / sum;
                        if (l == 0 && A.attention && i < n && lane + 32 * t < n)     // the UNMASKED softmax (comm_base_net.py:93)
                            A.attention[(r_env + i) * n + lane + 32 * t] = p;
                        const uint32_t word = __shfl_sync(0xFFFFFFFFu, mw, t);
                        const float a = ((word >> lane) & 1u) ? p : 0.0f;
                        dsum += a;
                        Aw[r * AP + lane + 32 * t] = a;
                    }
                    den[r] = warp_sumf(dsum);
                }
                __syncwarp();
                // ---- aggregation: out[r][cl .. cl+3] = sum_j a[r][j] HW[j][cl ..]; the two half-warps split the key chunks ----
                float acc[RT][4];
#pragma unroll
                for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
#pragma unroll 2
                for (int c = half; c < n_chunks; c += 2) {
                    const int j = c << 2;
                    const float4 h0 = *reinterpret_cast<const float4 *>(HWs + (j + 0) * 64 + cl);
                    const float4 h1 = *reinterpret_cast<const float4 *>(HWs + (j + 1) * 64 + cl);
                    const float4 h2 = *reinterpret_cast<const float4 *>(HWs + (j + 2) * 64 + cl);
                    const float4 h3 = *reinterpret_cast<const float4 *>(HWs + (j + 3) * 64 + cl);
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const float4 a = *reinterpret_cast<const float4 *>(Aw + r * AP + j);
                        acc[r][0] = fmaf(a.w, h3.x, fmaf(a.z, h2.x, fmaf(a.y, h1.x, fmaf(a.x, h0.x, acc[r][0]))));
                        acc[r][1] = fmaf(a.w, h3.y, fmaf(a.z, h2.y, fmaf(a.y, h1.y, fmaf(a.x, h0.y, acc[r][1]))));
                        acc[r][2] = fmaf(a.w, h3.z, fmaf(a.z, h2.z, fmaf(a.y, h1.z, fmaf(a.x, h0.z, acc[r][2]))));
                        acc[r][3] = fmaf(a.w, h3.w, fmaf(a.z, h2.w, fmaf(a.y, h1.w, fmaf(a.x, h0.w, acc[r][3]))));
                    }
                }
#pragma unroll
                for (int r = 0; r < RT; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] += __shfl_xor_sync(0xFFFFFFFFu, acc[r][c], 16);
                __syncwarp();                                        // every lane has read the attention rows
                // ---- H_{l+1} = tanh(out / (sum + 1e-12) + b)   (comm_base_net.py:103, graph_conv_module.py:66-72) ----
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float dn = den[r] + 1e-12f;
                    acc[r][0] = tanhf(acc[r][0] / dn + b4.x);
                    acc[r][1] = tanhf(acc[r][1] / dn + b4.y);
                    acc[r][2] = tanhf(acc[r][2] / dn + b4.z);
                    acc[r][3] = tanhf(acc[r][3] / dn + b4.w);
                }
                if (l + 1 < L) {
                    // ---- next layer's value rows: H_{l+1} Wg_{l+1} (lane = 2 output columns), written over this warp's rows ----
                    if (half == 0) {
#pragma unroll
                        for (int r = 0; r < RT; ++r)
                            *reinterpret_cast<float4 *>(Aw + r * AP + cl) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                    }
                    __syncwarp();
                    const float *wg = A.weights + o.gcn_w + (size_t)(l + 1) * kE * kE + 2 * lane;
                    float o2[RT][2];
#pragma unroll
                    for (int r = 0; r < RT; ++r) o2[r][0] = o2[r][1] = 0.0f;
#pragma unroll 2
                    for (int k = 0; k < kE; k += 4) {
                        const float2 w0 = __ldg(reinterpret_cast<const float2 *>(wg + (k + 0) * kE));
                        const float2 w1 = __ldg(reinterpret_cast<const float2 *>(wg + (k + 1) * kE));
                        const float2 w2 = __ldg(reinterpret_cast<const float2 *>(wg + (k + 2) * kE));
                        const float2 w3 = __ldg(reinterpret_cast<const float2 *>(wg + (k + 3) * kE));
#pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            const float4 hv = *reinterpret_cast<const float4 *>(Aw + r * AP + k);
                            o2[r][0] = fmaf(hv.w, w3.x, fmaf(hv.z, w2.x, fmaf(hv.y, w1.x, fmaf(hv.x, w0.x, o2[r][0]))));
                            o2[r][1] = fmaf(hv.w, w3.y, fmaf(hv.z, w2.y, fmaf(hv.y, w1.y, fmaf(hv.x, w0.y, o2[r][1]))));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RT; ++r)
                        if (i0 + r < n)
                            *reinterpret_cast<float2 *>(A.scr_hw + (r_env + i0 + r) * 64 + 2 * lane) = make_float2(o2[r][0], o2[r][1]);
                } else if (half == 0) {
                    // ---- X = E + H_L (comm_base_net.py:105-106) over the query rows, which are dead ----
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        if (i0 + r < n) {
                            float4 x = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                            if (A.d.residual) {
                                const float4 ev = *reinterpret_cast<const float4 *>(Es + (i0 + r) * kEPitch + cl);
                                x.x += ev.x; x.y += ev.y; x.z += ev.z; x.w += ev.w;
                            }
                            *reinterpret_cast<float4 *>(A.scr_q + (r_env + i0 + r) * 64 + cl) = x;
                        }
                    }
                }
            }
        }
    }
}

template <int RT, int KT>
static int launch_attn_t(const AttnArgs &A, cudaStream_t stream)
{
    constexpr int NK = 32 * KT;
    constexpr size_t smem = (size_t)(NK * kEPitch + NK * 64 + kAWarps * RT * (NK + 4)) * sizeof(float);
    static thread_local struct { int dev; int slots; } cache = {-1, 0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ENODEVICE);
    if (cache.dev != dev) {
        int sms = 0, ctas = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ECUDA);
        cudaError_t e = cudaFuncSetAttribute(policy_attn_kernel<RT, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, CM_ECUDA);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, policy_attn_kernel<RT, KT>, kAThreads, smem) != cudaSuccess)
            return set_cuda_error(cudaGetLastError(), CM_ECUDA);
        cache.dev = dev;
        cache.slots = sms * (ctas < 1 ? 1 : ctas);
    }
    const int grid = (int)(A.n_envs < cache.slots ? A.n_envs : cache.slots);
    policy_attn_kernel<RT, KT><<<grid, kAThreads, smem, stream>>>(A);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CM_OK : set_cuda_error(e, CM_ECUDA);
}

template <int RT>
static int launch_attn_r(int KT, const AttnArgs &A, cudaStream_t stream)
{
    switch (KT) {
    case 3: return launch_attn_t<RT, 3>(A, stream);
    case 4: return launch_attn_t<RT, 4>(A, stream);
    case 5: return launch_attn_t<RT, 5>(A, stream);
    case 6: return launch_attn_t<RT, 6>(A, stream);
    case 7: return launch_attn_t<RT, 7>(A, stream);
    case 8: return launch_attn_t<RT, 8>(A, stream);
    }
    return CM_EUNSUPPORTED;
}

int launch_policy_tc_encode(const cm_policy_desc *desc, const cm_policy_io *io, float *scr_e, float *scr_q, float *scr_hw, cudaStream_t stream);
int launch_policy_tc_head(const cm_policy_desc *desc, const cm_policy_io *io, const float *x_rows, cudaStream_t stream);

size_t tc_large_ws_floats(int n, int64_t n_envs) { return (size_t)3 * (size_t)n_envs * (size_t)n * kE; }

// Comm-DP forward for teams of 65..256 agents: encoder (tensor cores) -> attention / graph convolutions (here) -> head
// (tensor cores), three launches on one stream, rows handed over through the caller's workspace.
int launch_policy_tc_large(const cm_policy_desc *desc, const cm_policy_io *io, cudaStream_t stream)
{
    const int n = desc->n_agents;
    if (n <= 64 || n > CM_MAX_AGENTS) return CM_EUNSUPPORTED;
    const size_t rows = (size_t)io->n_envs * n;
    if (!io->workspace || io->workspace_bytes < tc_large_ws_floats(n, io->n_envs) * sizeof(float)) return CM_EINVAL;
    if ((reinterpret_cast<uintptr_t>(io->workspace) & 15u) != 0) return CM_EINVAL;
    float *scr_e = io->workspace, *scr_q = scr_e + rows * kE, *scr_hw = scr_q + rows * kE;
    int rc = launch_policy_tc_encode(desc, io, scr_e, scr_q, scr_hw, stream);
    if (rc) return rc;
    AttnArgs A;
    A.d = *desc;
    A.weights = io->weights;
    A.adj_bits = io->adj_bits;
    A.chan_bits = io->chan_bits;
    A.attention = io->attention;
    A.scr_e = scr_e; A.scr_q = scr_q; A.scr_hw = scr_hw;
    A.n_envs = io->n_envs;
    // query rows per warp: the team is cut into ceil(n / 64) blocks of 8 * RT rows; keys per lane: ceil(n / 32)
    const int nb = (n + 63) / 64, RT = (n + 8 * nb - 1) / (8 * nb), KT = (n + 31) / 32;
    switch (RT) {
    case 5: rc = launch_attn_r<5>(KT, A, stream); break;
    case 6: rc = launch_attn_r<6>(KT, A, stream); break;
    case 7: rc = launch_attn_r<7>(KT, A, stream); break;
    case 8: rc = launch_attn_r<8>(KT, A, stream); break;
    default: rc = CM_EUNSUPPORTED;
    }
    if (rc) return rc;
    if (!io->probs && !io->actions && !io->logits) return CM_OK;
    return launch_policy_tc_head(desc, io, scr_q, stream);
}

}  // namespace cm
