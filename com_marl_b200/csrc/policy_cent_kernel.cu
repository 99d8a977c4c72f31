// policy_cent_kernel.cu — CENT policy (CentralizedCategoricalMLPPolicy, centralized_categorical_mlp_policy.py:11-97):
// ONE multi-layer perceptron over the concatenated observation of the whole team,
//   logits[B][5n] = W4 act(W3 act(W2 act(W1 obs[B][n*D] + b1) + b2) + b3) + b4        (hidden sizes 128, 64, 32; act = tanh | relu)
// reshaped to [B][n][5], softmax per agent, availability mask, renormalisation, sampling (:73-96, :98-117).
//
// Rows of the product are ENVS here (not agents): a CTA of 256 threads owns a tile of 64 envs and walks the four layers
// with everything but the first layer's input in shared memory.  The first layer is the only product whose K grows with
// the team (K = n*D: 84 at C1, 1 696 at C3, 10 600 at C5): observations and W1 are streamed through shared memory in chunks
// of 32 k with the next chunk's global loads in flight (registers) while the current one is multiplied; the observation
// row block is read exactly once (the kernel's only HBM stream), W1 comes from L2.  Thread (ty, tx) of the 16 x 16 grid
// owns rows 4ty..4ty+3 and the columns {4tx..4tx+3} + 64j: every shared-memory operand read is a conflict-free 128-bit
// load, 3 loads per 32 FMAs.  The output layer is walked in passes of 16 agents (80 columns): a thread computes the five
// logits of ONE agent for its four envs, so softmax / mask / sampling finish in registers (common.cuh::categorical_finish,
// the same tail and random-stream specification as the Comm-DP / Obs-DP kernels).  Exact fp32 (FFMA, tanhf).
#include <cuda_runtime.h>

#include "common.cuh"
#include "policy_layout.cuh"

namespace cm {

static constexpr int kCThreads = 256, kCRows = 64, kCK = 32;
static constexpr int kPH1 = kC1 + 4, kPH2 = kC2 + 4, kPH3 = kC3 + 4, kPA = kCK + 4;      // pitches: rows stay 16-byte aligned
static constexpr int kCAgentsPerPass = 16, kCOutCols = kCAgentsPerPass * CM_ACTIONS;     // 80
static constexpr int kCentSmemFloats = kCRows * (kPH1 + kPH2 + kPH3 + kPA) + kCK * kC1;

struct CentArgs {
    cm_policy_desc d;
    cm_policy_io io;
    int K, NO, relu;
    CentBlob o;
};

template <bool RELU>
__device__ __forceinline__ float cent_act(float x) { return RELU ? fmaxf(x, 0.0f) : tanhf(x); }

// acc[4][4*J] += A[rows 4ty..][k0..k0+kc) * Bs[k][cols], Bs pitch = 64*J, kc a multiple of 4 (zero padded)
template <int J>
__device__ __forceinline__ void cent_mac(const float *__restrict__ As, int lda, const float *__restrict__ Bs, int kc, int ty, int tx,
                                         float (&acc)[4][4 * J])
{
#pragma unroll 2
    for (int k = 0; k < kc; k += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(As + (ty * 4 + i) * lda + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 b[J];
#pragma unroll
            for (int j = 0; j < J; ++j) b[j] = *reinterpret_cast<const float4 *>(Bs + (k + kk) * (64 * J) + j * 64 + tx * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    acc[i][4 * j + 0] = fmaf(av, b[j].x, acc[i][4 * j + 0]);
                    acc[i][4 * j + 1] = fmaf(av, b[j].y, acc[i][4 * j + 1]);
                    acc[i][4 * j + 2] = fmaf(av, b[j].z, acc[i][4 * j + 2]);
                    acc[i][4 * j + 3] = fmaf(av, b[j].w, acc[i][4 * j + 3]);
                }
            }
        }
    }
}

// out[row][col] = act(acc + bias[col]) for the thread's 4 x 4J block
template <int J, bool RELU>
__device__ __forceinline__ void cent_store(float *__restrict__ out, int ldo, const float (&acc)[4][4 * J], const float *__restrict__ bias,
                                           int ty, int tx)
{
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + j * 64 + tx * 4));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v;
            v.x = cent_act<RELU>(acc[i][4 * j + 0] + bv.x);
            v.y = cent_act<RELU>(acc[i][4 * j + 1] + bv.y);
            v.z = cent_act<RELU>(acc[i][4 * j + 2] + bv.z);
            v.w = cent_act<RELU>(acc[i][4 * j + 3] + bv.w);
            *reinterpret_cast<float4 *>(out + (ty * 4 + i) * ldo + j * 64 + tx * 4) = v;
        }
    }
}

// hidden layer with its input already in shared memory: out = act(in[64][K] W[K][64*J] + b), W streamed in 32-row chunks.
// (32-wide layer: J = 1 with the upper half of the 64 columns padded by zeros — W has only N = 32 real columns.)
template <int J, bool RELU>
__device__ __forceinline__ void cent_hidden(const float *__restrict__ in, int ldi, int K, const float *__restrict__ W, int N,
                                            const float *__restrict__ bias, float *__restrict__ out, int ldo, float *__restrict__ Bch,
                                            int tid, int ty, int tx)
{
    float acc[4][4 * J];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4 * J; ++c) acc[i][c] = 0.0f;
    const int NP = 64 * J;                                       // padded width of the chunk
    for (int k0 = 0; k0 < K; k0 += kCK) {
        __syncthreads();                                         // previous chunk consumed (and `in` complete on the first pass)
        for (int e = tid; e < kCK * NP / 4; e += kCThreads) {
            const int kk = e / (NP / 4), c4 = e - kk * (NP / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 * 4 < N) v = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(k0 + kk) * N + c4 * 4));
            *reinterpret_cast<float4 *>(Bch + kk * NP + c4 * 4) = v;
        }
        __syncthreads();
        cent_mac<J>(in + k0, ldi, Bch, kCK, ty, tx, acc);
    }
    if (tx * 4 < N || J > 1) {
        if (J == 1 && N < 64) {                                  // 32-wide layer: only tx < 8 hold real columns
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + tx * 4));
                float4 v;
                v.x = cent_act<RELU>(acc[i][0] + bv.x); v.y = cent_act<RELU>(acc[i][1] + bv.y);
                v.z = cent_act<RELU>(acc[i][2] + bv.z); v.w = cent_act<RELU>(acc[i][3] + bv.w);
                *reinterpret_cast<float4 *>(out + (ty * 4 + i) * ldo + tx * 4) = v;
            }
        } else {
            cent_store<J, RELU>(out, ldo, acc, bias, ty, tx);
        }
    }
}

template <bool RELU>
__global__ void __launch_bounds__(kCThreads, 2) policy_cent_kernel(const CentArgs A)
{
    extern __shared__ __align__(16) float smem[];
    float *h1 = smem;                                  // [64][132]
    float *h2 = h1 + kCRows * kPH1;                    // [64][68]
    float *h3 = h2 + kCRows * kPH2;                    // [64][36]
    float *Ach = h3 + kCRows * kPH3;                   // [64][36]   streamed observation chunk
    float *Bch = Ach + kCRows * kPA;                   // [32][128]  streamed weight chunk (also [32][80] of the output layer)
    const cm_policy_io &io = A.io;
    const float *__restrict__ wts = io.weights;
    const CentBlob &o = A.o;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
    const int K = A.K, n = A.d.n_agents;
    const int64_t row0 = (int64_t)blockIdx.x * kCRows;
    const int rows = (int)min((int64_t)kCRows, io.n_envs - row0);

    // ---- layer 1: h1 = act(obs W1 + b1), K streamed ----
    {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = 0.0f;
        float ra[8];
        float4 rb[4];
        const float *__restrict__ obs = io.obs + row0 * K;
        auto fetch = [&](int k0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {                        // lane = k, warp + 8j = row: 128-byte row segments
                const int r = warp + 8 * j;
                ra[j] = (r < rows && k0 + lane < K) ? __ldg(obs + (size_t)r * K + k0 + lane) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = tid + kCThreads * j, kk = e >> 5, c4 = e & 31;
                rb[j] = (k0 + kk < K) ? __ldg(reinterpret_cast<const float4 *>(wts + o.w1 + (size_t)(k0 + kk) * kC1 + c4 * 4))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        fetch(0);
        for (int k0 = 0; k0 < K; k0 += kCK) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) Ach[(warp + 8 * j) * kPA + lane] = ra[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = tid + kCThreads * j, kk = e >> 5, c4 = e & 31;
                *reinterpret_cast<float4 *>(Bch + kk * kC1 + c4 * 4) = rb[j];
            }
            __syncthreads();
            if (k0 + kCK < K) fetch(k0 + kCK);
            cent_mac<2>(Ach, kPA, Bch, kCK, ty, tx, acc);
        }
        cent_store<2, RELU>(h1, kPH1, acc, wts + o.b1, ty, tx);
    }
    // ---- layers 2, 3 ----
    cent_hidden<1, RELU>(h1, kPH1, kC1, wts + o.w2, kC2, wts + o.b2, h2, kPH2, Bch, tid, ty, tx);
    cent_hidden<1, RELU>(h2, kPH2, kC2, wts + o.w3, kC3, wts + o.b3, h3, kPH3, Bch, tid, ty, tx);
    // ---- output layer in passes of 16 agents; thread = one agent x four envs ----
    const int NO = A.NO;
    for (int a0 = 0; a0 < n; a0 += kCAgentsPerPass) {
        __syncthreads();                                         // h3 complete / previous pass consumed
        for (int e = tid; e < kC3 * kCOutCols; e += kCThreads) {
            const int kk = e / kCOutCols, c = e - kk * kCOutCols, col = a0 * CM_ACTIONS + c;
            Bch[e] = col < NO ? __ldg(wts + o.w4 + (size_t)kk * NO + col) : 0.0f;
        }
        __syncthreads();
        const int il = a0 + tx;
        float lg[4][CM_ACTIONS];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int a = 0; a < CM_ACTIONS; ++a) lg[i][a] = 0.0f;
#pragma unroll 2
        for (int k = 0; k < kC3; k += 4) {
            float4 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4 *>(h3 + (ty * 4 + i) * kPH3 + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float w[CM_ACTIONS];
#pragma unroll
                for (int a = 0; a < CM_ACTIONS; ++a) w[a] = Bch[(k + kk) * kCOutCols + tx * CM_ACTIONS + a];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xv = kk == 0 ? x[i].x : (kk == 1 ? x[i].y : (kk == 2 ? x[i].z : x[i].w));
#pragma unroll
                    for (int a = 0; a < CM_ACTIONS; ++a) lg[i][a] = fmaf(xv, w[a], lg[i][a]);
                }
            }
        }
        if (il < n) {
            float b4[CM_ACTIONS];
#pragma unroll
            for (int a = 0; a < CM_ACTIONS; ++a) b4[a] = __ldg(wts + o.b4 + il * CM_ACTIONS + a);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i;
                if (r < rows) {
                    float l5[CM_ACTIONS];
#pragma unroll
                    for (int a = 0; a < CM_ACTIONS; ++a) l5[a] = lg[i][a] + b4[a];
                    const int64_t env = row0 + r;
                    categorical_finish(A.d, io, l5, env * n + il, env, il);
                }
            }
        }
    }
}

int launch_policy_cent(const cm_policy_desc *desc, const cm_policy_io *io, cudaStream_t stream)
{
    if (desc->math != 0) return CM_EUNSUPPORTED;                 // exact fp32 only
    if (io->attention) return CM_EINVAL;                         // no communication, no attention weights
    if (io->n_envs == 0) return CM_OK;
    if (io->n_envs * desc->n_agents > (int64_t)1 << 30) return CM_EUNSUPPORTED;
    CentArgs A;
    A.d = *desc;
    A.io = *io;
    A.K = desc->n_agents * desc->obs_dim;
    A.NO = desc->n_agents * CM_ACTIONS;
    A.relu = desc->flags & CM_POLICY_FLAG_RELU;
    A.o = cent_blob_layout(desc->n_agents, desc->obs_dim);
    const size_t smem = kCentSmemFloats * sizeof(float);
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_cuda_error(cudaGetLastError(), CM_ENODEVICE);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(policy_cent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_cent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, CM_ECUDA);
        attr_set[dev] = true;
    }
    const unsigned grid = (unsigned)((io->n_envs + kCRows - 1) / kCRows);
    if (A.relu) policy_cent_kernel<true><<<grid, kCThreads, smem, stream>>>(A);
    else policy_cent_kernel<false><<<grid, kCThreads, smem, stream>>>(A);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CM_OK : set_cuda_error(e, CM_ECUDA);
}

}  // namespace cm

extern "C" size_t cm_policy_cent_blob_floats(int32_t n_agents, int32_t obs_dim)
{
    if (n_agents < 1 || obs_dim < 1) return 0;
    return (size_t)cm::cent_blob_layout(n_agents, obs_dim).total;
}
