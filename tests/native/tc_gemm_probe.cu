// tc_gemm_probe.cu — test-only probe of the tcgen05 building block used by the policy kernel:
// C[128][N] = A[128][K] * B[N][K]^T with error-compensated TF32 (passes = 3: hi*hi + lo*hi + hi*lo).
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

using namespace cm::tc;

__global__ void __launch_bounds__(128) tc_gemm_probe_kernel(const float *A, const float *B, float *C, int N, int K, int passes, int *status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    float *Ahi = reinterpret_cast<float *>(smem);
    float *Alo = Ahi + 128 * K;
    float *Bhi = Alo + 128 * K;
    float *Blo = Bhi + N * K;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    // operands into the canonical layout, split hi / lo
    for (int kc = 0; kc < K / 4; ++kc) {
        const float4 a = *reinterpret_cast<const float4 *>(A + (size_t)tid * K + 4 * kc);
        const float4 h = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
        const uint32_t off = canon_off(tid, 4 * kc, K);
        *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(Ahi) + off) = h;
        *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(Alo) + off) = make_float4(tf32_lo(a.x, h.x), tf32_lo(a.y, h.y), tf32_lo(a.z, h.z), tf32_lo(a.w, h.w));
    }
    for (int r = tid; r < N; r += 128)
        for (int kc = 0; kc < K / 4; ++kc) {
            const float4 b = *reinterpret_cast<const float4 *>(B + (size_t)r * K + 4 * kc);
            const float4 h = make_float4(tf32_hi(b.x), tf32_hi(b.y), tf32_hi(b.z), tf32_hi(b.w));
            const uint32_t off = canon_off(r, 4 * kc, K);
            *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(Bhi) + off) = h;
            *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(Blo) + off) = make_float4(tf32_lo(b.x, h.x), tf32_lo(b.y, h.y), tf32_lo(b.z, h.z), tf32_lo(b.w, h.w));
        }
    fence_proxy_async();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        uint32_t acc = 0;
        for (int p = 0; p < passes; ++p) {
            const float *a = (p == 1) ? Alo : Ahi;
            const float *b = (p == 2) ? Blo : Bhi;
            for (int j = 0; j < K / 8; ++j) {
                mma_tf32(tmem_base, make_smem_desc(smem_u32(a), K, j), make_smem_desc(smem_u32(b), K, j), idesc, acc);
                acc = 1;
            }
        }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    if (!ok && tid == 0) atomicExch(status, 1);
    fence_after_thread_sync();
    if (ok) {
        for (int c = 0; c < N; c += 8) {
            float v[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) C[(size_t)tid * N + c + i] = v[i];
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// fp16 split with a scaled remainder and a stacked B operand:
//   x = x_hi + 2^-12 x_lo',  x_hi = fp16(x), x_lo' = fp16((x - x_hi) * 4096)
//   pass 1: A_hi x [B_hi ; B_lo'] (N' = 2N)  -> D[:, 0:N] = hi*hi, D[:, N:2N] = hi*lo'
//   pass 2: A_lo' x B_hi (N' = N)            -> accumulates into D[:, N:2N]
//   C = D[:, c] + 2^-12 D[:, N + c]
#include <cuda_fp16.h>
__global__ void __launch_bounds__(128) tc_gemm_probe16_kernel(const float *A, const float *B, float *C, int N, int K, int *status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *Ahi = smem, *Alo = Ahi + 128 * K * 2, *Bst = Alo + 128 * K * 2;   // Bst: 2N rows x K
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int k = 0; k < K; ++k) {
        const float a = A[(size_t)tid * K + k];
        const __half h = __float2half_rn(a);
        const __half l = __float2half_rn((a - __half2float(h)) * 4096.0f);
        *reinterpret_cast<__half *>(Ahi + canon_off16(tid, k, K)) = h;
        *reinterpret_cast<__half *>(Alo + canon_off16(tid, k, K)) = l;
    }
    for (int r = tid; r < N; r += 128)
        for (int k = 0; k < K; ++k) {
            const float b = B[(size_t)r * K + k];
            const __half h = __float2half_rn(b);
            const __half l = __float2half_rn((b - __half2float(h)) * 4096.0f);
            *reinterpret_cast<__half *>(Bst + canon_off16(r, k, K)) = h;
            *reinterpret_cast<__half *>(Bst + canon_off16(N + r, k, K)) = l;
        }
    fence_proxy_async();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        for (int j = 0; j < K / 16; ++j)
            mma_f16(tmem_base, make_smem_desc16(smem_u32(Ahi), K, j), make_smem_desc16(smem_u32(Bst), K, j), make_idesc_f16(128, 2 * N), j > 0);
        for (int j = 0; j < K / 16; ++j)
            mma_f16(tmem_base + N, make_smem_desc16(smem_u32(Alo), K, j), make_smem_desc16(smem_u32(Bst), K, j), make_idesc_f16(128, N), 1);
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    if (!ok && tid == 0) atomicExch(status, 1);
    fence_after_thread_sync();
    if (ok) {
        for (int c = 0; c < N; c += 8) {
            float v[8], w[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N + c), w);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) C[(size_t)tid * N + c + i] = v[i] + w[i] * (1.0f / 4096.0f);
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

extern "C" int tc_gemm_probe16(const float *A, const float *B, float *C, int N, int K, int *status, void *stream)
{
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * 2;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_probe16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc_gemm_probe16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K, status);
    return (int)cudaGetLastError();
}

extern "C" int tc_gemm_probe(const float *A, const float *B, float *C, int N, int K, int passes, int *status, void *stream)
{
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc_gemm_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K, passes, status);
    return (int)cudaGetLastError();
}

// ---- TS-mode probe: A operand from tensor memory (tcgen05.mma [d], [a_tmem], b_desc) vs from shared memory, fp16 single
// pass; also times `reps` repetitions of the K/16 instruction series (cycles from first issue to completion).
// Assumed A layout in TMEM (checked numerically by the test): lane = row, 32-bit column c = elements (2c, 2c + 1). ----
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128) tc_ts_probe_kernel(const float *A, const float *B, float *C, int N, int K, int mode, int reps,
                                                          long long *cycles, int *status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *As = smem, *Bs = As + 128 * K * 2;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s, lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_col = 256;      // A operand columns [256, 256 + K/2)
    for (int k = 0; k < K; ++k) *reinterpret_cast<__half *>(As + canon_off16(tid, k, K)) = __float2half_rn(A[(size_t)tid * K + k]);
    for (int k = 0; k < K; k += 16) {          // 16 halves = 8 packed columns per store
        float pk[8];
        for (int i = 0; i < 8; ++i) {
            const __half2 h2 = __floats2half2_rn(A[(size_t)tid * K + k + 2 * i], A[(size_t)tid * K + k + 2 * i + 1]);
            pk[i] = __uint_as_float(*reinterpret_cast<const uint32_t *>(&h2));
        }
        tmem_st<8>(lane_addr + a_col + (uint32_t)(k / 2), pk);
    }
    tmem_st_wait();
    for (int r = tid; r < N; r += 128)
        for (int k = 0; k < K; ++k) *reinterpret_cast<__half *>(Bs + canon_off16(r, k, K)) = __float2half_rn(B[(size_t)r * K + k]);
    fence_proxy_async();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        t0 = clock64();
        if (mode < 2) {
        for (int r = 0; r < reps; ++r)
            for (int j = 0; j < K / 16; ++j) {
                if (elect_one()) {
                    if (mode == 0) mma_f16(tmem_base, make_smem_desc16(smem_u32(As), K, j), make_smem_desc16(smem_u32(Bs), K, j), idesc, (r | j) != 0);
                    else mma_f16_ts(tmem_base, tmem_base + a_col + (uint32_t)(8 * j), make_smem_desc16(smem_u32(Bs), K, j), idesc, (r | j) != 0);
                }
                __syncwarp();
            }
        } else if (mode == 2) {        // two independent accumulators, alternating (rep r -> D block r & 1 at column 0 / 256... N <= 128)
            for (int r = 0; r < reps; ++r)
                for (int j = 0; j < K / 16; ++j) {
                    if (elect_one())
                        mma_f16(tmem_base + (uint32_t)((j & 1) * 128), make_smem_desc16(smem_u32(As), K, j), make_smem_desc16(smem_u32(Bs), K, j), idesc, (r | (j >> 1)) != 0);
                    __syncwarp();
                }
        } else if (mode == 3) {        // one thread issues, no elect / syncwarp in the loop
            if (threadIdx.x == 0) {
                uint64_t da = make_smem_desc16(smem_u32(As), K, 0), db = make_smem_desc16(smem_u32(Bs), K, 0);
                for (int r = 0; r < reps; ++r)
#pragma unroll 4
                    for (int j = 0; j < K / 16; ++j) mma_f16(tmem_base, da + 16 * j, db + 16 * j, idesc, (r | j) != 0);
            }
            __syncwarp();
        } else if (mode == 5) {        // all lanes run the loop (uniform operands), the instruction itself is predicated on the leader
            const uint32_t leader = elect_one() ? 1u : 0u;
            uint64_t da = make_smem_desc16(smem_u32(As), K, 0), db = make_smem_desc16(smem_u32(Bs), K, 0);
            for (int r = 0; r < reps; ++r)
#pragma unroll 4
                for (int j = 0; j < K / 16; ++j) {
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p, q;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "setp.ne.b32 q, %5, 0;\n\t"
                        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                        "}\n" ::"r"(tmem_base), "l"(da + 16 * j), "l"(db + 16 * j), "r"(idesc), "r"((uint32_t)((r | j) != 0)), "r"(leader)
                        : "memory");
                }
            __syncwarp();
        } else if (mode == 4) {        // four independent accumulators
            for (int r = 0; r < reps; ++r)
                for (int j = 0; j < K / 16; ++j) {
                    if (elect_one())
                        mma_f16(tmem_base + (uint32_t)((j & 3) * 64), make_smem_desc16(smem_u32(As), K, j), make_smem_desc16(smem_u32(Bs), K, j), idesc, r != 0);
                    __syncwarp();
                }
        }
        if (elect_one()) mma_commit(&bar);
        __syncwarp();
    }
    const bool ok = mbar_wait(&bar, 0);
    t1 = clock64();
    if (tid == 0) { cycles[0] = t1 - t0; }
    if (!ok && tid == 0) atomicExch(status, 1);
    fence_after_thread_sync();
    if (ok) {
        for (int c = 0; c < N; c += 8) {
            float v[8];
            tmem_ld8(lane_addr + (uint32_t)c, v);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) C[(size_t)tid * N + c + i] = v[i];
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

extern "C" int tc_ts_probe(const float *A, const float *B, float *C, int N, int K, int mode, int reps, long long *cycles, int *status, void *stream)
{
    const size_t smem = (size_t)(128 * K + N * K) * 2;
    cudaError_t e = cudaFuncSetAttribute(tc_ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc_ts_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, C, N, K, mode, reps, cycles, status);
    return (int)cudaGetLastError();
}
