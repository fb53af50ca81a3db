// Fused optimizer tail: global-norm gradient clip + Adam + exponential LR decay with a
// floor, over flat fp32 parameter / gradient / moment buffers -- what TrainingInterface.train does after
// backward (amc_dl/torch_plus/module.py:142-143 clip_grad_norm_, scheduler.py:69-74 Adam.step + LR step,
// example.py:4-12 MinExponentialLR).  Two HBM-bound passes per step: sum of squares (4 B/param read) and the
// update (16 B read + 12 B written per param); step count and norm stay on the device (graph-capturable).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, float* out) {
    float s = 0.0f;
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long)gridDim.x * blockDim.x * 4) {
        if (i + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(g + i);
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        } else {
            for (long k = i; k < n; ++k) s += g[k] * g[k];
        }
    }
    s = warp_sum(s);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int k = 0; k < 8; ++k) a += sh[k];
        atomicAdd(out, a);
    }
}

__global__ void counter_inc_kernel(int* c) { c[0] += 1; }

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long n,
                                                        const float* __restrict__ sumsq, const int* __restrict__ step,
                                                        float lr0, float gamma, float lr_min, float b1, float b2,
                                                        float eps, float clip) {
    const int t = step[0];
    float coef = 1.0f;
    if (clip > 0.0f) coef = fminf(1.0f, clip / (sqrtf(sumsq[0]) + 1e-6f));
    float lr = lr0;
    if (gamma > 0.0f) lr = fmaxf(lr0 * powf(gamma, (float)(t - 1)), lr_min);
    const float bc1 = 1.0f - powf(b1, (float)t), bc2s = sqrtf(1.0f - powf(b2, (float)t));
    const float step_size = lr / bc1;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * coef;
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
}

}  // namespace

// out[0] += sum g^2 (caller zeroes out once per step, then calls this for every gradient buffer)
PD_API int pd_sumsq_f32(const float* g, long n, float* out, void* stream) {
    if (n <= 0) return 0;
    long blocks = (n / 4 + 255) / 256;
    if (blocks > 8L * PD_NUM_SMS) blocks = 8L * PD_NUM_SMS;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
    return pd_launch_status();
}

PD_API int pd_counter_inc(int* counter, void* stream) {
    counter_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    return pd_launch_status();
}

// clip (global norm from sumsq[0]; clip <= 0 disables) + Adam; lr = max(lr0 * gamma^(step-1), lr_min) when gamma > 0.
PD_API int pd_adam_clip_step(float* p, const float* g, float* m, float* v, long n, const float* sumsq, const int* step,
                             float lr0, float gamma, float lr_min, float b1, float b2, float eps, float clip,
                             void* stream) {
    if (n <= 0) return 0;
    adam_clip_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sumsq, step, lr0, gamma, lr_min,
                                                                         b1, b2, eps, clip);
    return pd_launch_status();
}

// ---- library-wide switch: programmatic dependent launch for the opted-in kernels (common.cuh) -----------------------------
int g_pd_pdl = 0;
PD_API int pd_set_pdl(int on) {
    g_pd_pdl = on ? 1 : 0;
    return 0;
}
