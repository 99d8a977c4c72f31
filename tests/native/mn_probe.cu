// mn_probe.cu — test-only probe of two tcgen05 operand forms the fused policy kernel's attention uses:
//   (1) an MN-major B operand (memory [K][N], N contiguous) without swizzle: which of LBO / SBO is the K-group stride;
//   (2) an MN-major A operand with M = 64 and the lane layout of an M = 64 accumulator in tensor memory.
// D is dumped for all 128 lanes so the host can find where each row landed.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

using namespace cm::tc;

__device__ __forceinline__ uint64_t desc_raw(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// X: [R rows][C cols] fp32 (row-major) -> fp16 "row-chunk" layout: byte(r, c) = (r >> 3) * (C / 8) * 128 + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2
// (this is canon_off16(r, c, C): K-major when the columns are K, MN-major when the columns are M / N)
__device__ void store_chunks(unsigned char *dst, const float *X, int R, int C, int tid, int nthreads)
{
    for (int e = tid; e < R * C; e += nthreads) {
        const int r = e / C, c = e - r * C;
        *reinterpret_cast<__half *>(dst + canon_off16(r, c, C)) = __float2half_rn(X[e]);
    }
}

// mode 0: D[128][N] = A[128][K] * B[K][N], A K-major, B MN-major (memory [K][N]); variant selects the LBO / SBO roles
// mode 1: D[64][N]  = A^T: A given as [K][64] (MN-major A), B given as [N][K] (K-major)
__global__ void __launch_bounds__(128) mn_probe_kernel(const float *A, const float *B, float *D, int M, int N, int K, int mode, int variant,
                                                       int *status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *As = smem, *Bs = smem + 32768;
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (mode == 0) {
        store_chunks(As, A, 128, K, tid, 128);      // [row][k]: K-major
        store_chunks(Bs, B, K, N, tid, 128);        // [k][n]:  MN-major
    } else {
        store_chunks(As, A, K, 64, tid, 128);       // [k][m]:  MN-major A
        store_chunks(Bs, B, N, K, tid, 128);        // [n][k]:  K-major
    }
    fence_proxy_async();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s, lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    {   // zero the dump area
        float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 128; c += 8) tmem_st<8>(lane_addr + (uint32_t)c, z);
        tmem_st_wait();
    }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    if (tid == 0) {
        uint32_t idesc = make_idesc_f16(M, N);
        if (mode == 0) {
            idesc |= 1u << 16;                       // b_major = MN
            const uint32_t kgroup = (uint32_t)(N / 8) * 128u, mngroup = 128u;      // strides of the B buffer
            for (int j = 0; j < K / 16; ++j) {
                const uint64_t da = make_smem_desc16(smem_u32(As), K, j);
                const uint32_t b0 = smem_u32(Bs) + (uint32_t)j * 2u * kgroup;
                const uint64_t db = variant == 0 ? desc_raw(b0, kgroup, mngroup) : desc_raw(b0, mngroup, kgroup);
                mma_f16(tmem_base, da, db, idesc, j > 0);
            }
        } else {
            idesc |= 1u << 15;                       // a_major = MN
            const uint32_t kgroup = (uint32_t)(64 / 8) * 128u, mngroup = 128u;     // strides of the A buffer ([k][64])
            for (int j = 0; j < K / 16; ++j) {
                const uint32_t a0 = smem_u32(As) + (uint32_t)j * 2u * kgroup;
                const uint64_t da = variant == 0 ? desc_raw(a0, kgroup, mngroup) : desc_raw(a0, mngroup, kgroup);
                const uint64_t db = make_smem_desc16(smem_u32(Bs), K, j);
                mma_f16(tmem_base, da, db, idesc, j > 0);
            }
        }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    if (!ok && tid == 0) atomicExch(status, 1);
    fence_after_thread_sync();
    if (ok) {
        for (int c = 0; c < 128; c += 8) {          // dump [128 lanes][128 columns]
            float v[8];
            tmem_ld8(lane_addr + (uint32_t)c, v);
            tmem_ld_wait();
            for (int i = 0; i < 8; ++i) D[(size_t)tid * 128 + c + i] = v[i];
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 128);
}

extern "C" int mn_probe(const float *A, const float *B, float *D, int M, int N, int K, int mode, int variant, int *status, void *stream)
{
    const size_t smem = 65536;
    cudaError_t e = cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    mn_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, M, N, K, mode, variant, status);
    return (int)cudaGetLastError();
}
