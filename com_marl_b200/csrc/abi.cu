// abi.cu — version / error plumbing of the C ABI and the dense <-> bit-row mask converters.
#include <cuda_runtime.h>
#include <stdint.h>

#include "commarl_b200.h"
#include "common.cuh"

namespace cm {

static thread_local int g_last_cuda_error = 0;

int set_cuda_error(cudaError_t e, int rc)
{
    g_last_cuda_error = (int)e;
    return rc;
}

// dense float32 mask rows (the reference's dist_adj / channels arrays) -> one u32 word per 32 columns.
// One warp per (row, word): lanes read 32 consecutive floats (coalesced), __ballot_sync packs them.
__global__ void mask_pack_kernel(const float *__restrict__ dense, uint32_t *__restrict__ bits, int64_t rows, int n, int W)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < rows * W; item += nwarps) {
        const int64_t row = item / W;
        const int w = (int)(item - row * W);
        const int j = w * 32 + lane;
        const float v = (j < n) ? __ldg(dense + row * n + j) : 0.0f;
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, v != 0.0f);
        if (lane == 0) bits[item] = word;
    }
}

__global__ void mask_unpack_kernel(const uint32_t *__restrict__ bits, float *__restrict__ dense, int64_t rows, int n, int W)
{
    const int64_t total = rows * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / n;
        const int j = (int)(e - row * n);
        dense[e] = (float)((__ldg(bits + row * W + (j >> 5)) >> (j & 31)) & 1u);
    }
}

}  // namespace cm

extern "C" int cm_abi_version(void) { return CM_ABI_VERSION; }

extern "C" const char *cm_strerror(int err)
{
    switch (err) {
    case CM_OK: return "ok";
    case CM_EINVAL: return "invalid argument (null pointer or bad size)";
    case CM_EUNSUPPORTED: return "unsupported configuration";
    case CM_ECUDA: return "CUDA runtime error (see cm_last_cuda_error)";
    case CM_ENODEVICE: return "no CUDA device";
    case CM_EACTION: return "Action Not found!";
    default: return "unknown error";
    }
}

extern "C" int cm_last_cuda_error(void) { return cm::g_last_cuda_error; }

extern "C" int cm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int grid_for(int64_t work_items, int threads)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sms * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

extern "C" int cm_mask_pack(const float *dense, uint32_t *bits, int64_t rows, int32_t n, cm_stream_t stream)
{
    if (!dense || !bits || rows < 0 || n < 1) return CM_EINVAL;
    if (rows == 0) return CM_OK;
    if (cm_device_count() < 1) return CM_ENODEVICE;
    const int W = (n + 31) / 32;
    cm::mask_pack_kernel<<<grid_for(rows * W * 32, 256), 256, 0, (cudaStream_t)stream>>>(dense, bits, rows, n, W);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CM_OK : cm::set_cuda_error(e, CM_ECUDA);
}

extern "C" int cm_mask_unpack(const uint32_t *bits, float *dense, int64_t rows, int32_t n, cm_stream_t stream)
{
    if (!dense || !bits || rows < 0 || n < 1) return CM_EINVAL;
    if (rows == 0) return CM_OK;
    if (cm_device_count() < 1) return CM_ENODEVICE;
    const int W = (n + 31) / 32;
    cm::mask_unpack_kernel<<<grid_for(rows * n, 256), 256, 0, (cudaStream_t)stream>>>(bits, dense, rows, n, W);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CM_OK : cm::set_cuda_error(e, CM_ECUDA);
}
