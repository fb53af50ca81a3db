// Texture-encoder front end: Conv2d(1->C, k=(4,12), stride=(4,1)) + ReLU + MaxPool2d((1,4),(1,4)) over the
// (B,32,128) piano-roll, fused.  Replaces ptvae.py:95-99 / :114.  The output is written channel-major
// (B,C,8,29) and the caller REINTERPRETS it as (B,8,29*C) exactly like the reference's `.view(bs,8,-1)`
// (a memory reinterpretation, not a transpose -- trained weights depend on it).
//
// HBM-bound: 16 KB in, C*8*29*4 B out per sample.  One CTA per (sample, 4-row band): the band (4x128
// floats) and the C*48 filter taps are staged in shared memory, each thread produces one pooled
// output (4 conv positions x 48 taps).  The backward pass recomputes the conv to find the pooled
// argmax and reduces the filter gradient in shared memory (one atomicAdd per tap per CTA).
#include "common.cuh"

namespace {

constexpr int KH = 4, KW = 12, W_IN = 128, W_POOL = 29, MAXC = 16;

__global__ void __launch_bounds__(320) texture_fwd_kernel(const float* __restrict__ pr, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int C, float* out) {
    __shared__ float band[KH][W_IN];
    __shared__ float ws[MAXC * KH * KW];
    __shared__ float bs[MAXC];
    const int b = blockIdx.x >> 3, i = blockIdx.x & 7;
    const float* src = pr + ((long)b * 32 + i * 4) * W_IN;
    for (int k = threadIdx.x; k < KH * W_IN; k += blockDim.x) band[k / W_IN][k % W_IN] = src[k];
    for (int k = threadIdx.x; k < C * KH * KW; k += blockDim.x) ws[k] = w[k];
    if (threadIdx.x < C) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    for (int o = threadIdx.x; o < C * W_POOL; o += blockDim.x) {
        const int ch = o / W_POOL, wp = o % W_POOL;
        const float* f = ws + ch * KH * KW;
        float best = 0.0f;   // relu floor
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s = bs[ch];
            const int w0 = wp * 4 + q;
#pragma unroll
            for (int dr = 0; dr < KH; ++dr)
#pragma unroll
                for (int dc = 0; dc < KW; ++dc) s = fmaf(f[dr * KW + dc], band[dr][w0 + dc], s);
            best = fmaxf(best, s);
        }
        out[(((long)b * C + ch) * 8 + i) * W_POOL + wp] = best;
    }
}

__global__ void __launch_bounds__(320) texture_bwd_kernel(const float* __restrict__ pr, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int C,
                                                          const float* __restrict__ gout, float* dw, float* dbias) {
    __shared__ float band[KH][W_IN];
    __shared__ float ws[MAXC * KH * KW];
    __shared__ float bs[MAXC];
    __shared__ float dws[MAXC * KH * KW];
    __shared__ float dbs[MAXC];
    const int b = blockIdx.x >> 3, i = blockIdx.x & 7;
    const float* src = pr + ((long)b * 32 + i * 4) * W_IN;
    for (int k = threadIdx.x; k < KH * W_IN; k += blockDim.x) band[k / W_IN][k % W_IN] = src[k];
    for (int k = threadIdx.x; k < C * KH * KW; k += blockDim.x) { ws[k] = w[k]; dws[k] = 0.0f; }
    if (threadIdx.x < C) { bs[threadIdx.x] = bias[threadIdx.x]; dbs[threadIdx.x] = 0.0f; }
    __syncthreads();
    for (int o = threadIdx.x; o < C * W_POOL; o += blockDim.x) {
        const int ch = o / W_POOL, wp = o % W_POOL;
        const float g = gout[(((long)b * C + ch) * 8 + i) * W_POOL + wp];
        if (g == 0.0f) continue;
        const float* f = ws + ch * KH * KW;
        float best = 0.0f;
        int bq = -1;          // -1: relu inactive everywhere -> no gradient
        for (int q = 0; q < 4; ++q) {
            float s = bs[ch];
            const int w0 = wp * 4 + q;
#pragma unroll
            for (int dr = 0; dr < KH; ++dr)
#pragma unroll
                for (int dc = 0; dc < KW; ++dc) s = fmaf(f[dr * KW + dc], band[dr][w0 + dc], s);
            if (s > best) { best = s; bq = q; }   // first maximum wins (max_pool2d backward convention)
        }
        if (bq < 0) continue;
        const int w0 = wp * 4 + bq;
        atomicAdd(&dbs[ch], g);
        for (int dr = 0; dr < KH; ++dr)
            for (int dc = 0; dc < KW; ++dc) {
                float v = band[dr][w0 + dc];
                if (v != 0.0f) atomicAdd(&dws[ch * KH * KW + dr * KW + dc], g * v);
            }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C * KH * KW; k += blockDim.x)
        if (dws[k] != 0.0f) atomicAdd(dw + k, dws[k]);
    if (threadIdx.x < C && dbs[threadIdx.x] != 0.0f) atomicAdd(dbias + threadIdx.x, dbs[threadIdx.x]);
}

}  // namespace

PD_API int pd_texture_frontend_fwd(const float* pr_mat, const float* w, const float* bias, int B, int C,
                                   float* out, void* stream) {
    if (B <= 0) return 0;
    if (C < 1 || C > MAXC) return PD_BAD_ARG;
    texture_fwd_kernel<<<B * 8, 320, 0, (cudaStream_t)stream>>>(pr_mat, w, bias, C, out);
    return pd_launch_status();
}

// dw (C*48) and dbias (C) are ACCUMULATED into (caller zeroes them).
PD_API int pd_texture_frontend_bwd(const float* pr_mat, const float* w, const float* bias, int B, int C,
                                   const float* gout, float* dw, float* dbias, void* stream) {
    if (B <= 0) return 0;
    if (C < 1 || C > MAXC) return PD_BAD_ARG;
    texture_bwd_kernel<<<B * 8, 320, 0, (cudaStream_t)stream>>>(pr_mat, w, bias, C, gout, dw, dbias);
    return pd_launch_status();
}
