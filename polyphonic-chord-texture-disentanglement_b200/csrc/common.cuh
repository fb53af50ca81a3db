// Shared helpers for libpolydis_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PD_API extern "C" __attribute__((visibility("default")))

// Every entry point returns 0 on success or a cudaError_t / negative argument-error code.
#define PD_BAD_ARG (-22)

static inline int pd_launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

static inline unsigned pd_blocks(long n, int per) { return (unsigned)((n + per - 1) / per); }

// SM count of the CURRENT device (cached per device; grids are sized from it, nothing is compiled in)
static inline int pd_num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = cache[dev & 63];
    if (c == 0 && (cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c <= 0)) c = 148;
    return c;
}
#define PD_NUM_SMS pd_num_sms()

// cudaFuncSetAttribute is per DEVICE: true the first time a kernel's launcher runs on the current device
static inline bool pd_first_use_on_device(unsigned long long& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

__device__ __forceinline__ float pd_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU forms for the TF32-mode recurrent kernels, which are instruction-bound on the gate math: ex2.approx +
// rcp.approx, absolute error ~1e-6 (three orders below the TF32 operand rounding of the matvec next to them);
// saturate correctly (exp -> inf gives 1/inf = 0).  The fp32 / tf32x3 parity paths keep expf / tanhf.
__device__ __forceinline__ float pd_sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float pd_tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
