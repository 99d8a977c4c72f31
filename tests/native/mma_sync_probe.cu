// mma_sync_probe.cu — issue rate of the legacy warp-level tensor path (mma.sync, SASS HMMA) on sm_100a:
// m16n8k16 f16 -> f32 and m16n8k8 tf32 -> f32, NACC independent accumulators per warp, W warps per CTA, one CTA per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int NACC, bool TF32>
__global__ void probe(float *out, long long *cycles, int iters)
{
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (TF32)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NACC, bool TF32>
static void run(int warps)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    probe<NACC, TF32><<<148, warps * 32>>>(out, cyc, iters);
    probe<NACC, TF32><<<148, warps * 32>>>(out, cyc, iters);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double mmas = (double)iters * NACC * warps, macs = mmas * 16 * 8 * (TF32 ? 8 : 16);
    printf("%s NACC=%d warps/SM=%2d : %.2f cycles per MMA per SMSP, %.0f MAC/clk/SM\n", TF32 ? "tf32 m16n8k8 " : "f16  m16n8k16", NACC, warps,
           avg / (mmas / 4), macs / avg);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<8, false>(4); run<8, false>(8); run<8, false>(16);
    run<8, true>(4); run<8, true>(8); run<8, true>(16);
    run<2, false>(8); run<2, true>(8);
    return 0;
}
