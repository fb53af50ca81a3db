// Texture-encoder front end: Conv2d(1->C, k=(4,12), stride=(4,1)) + ReLU + MaxPool2d((1,4),(1,4)) over the
// (B,32,128) piano-roll, fused.  Replaces ptvae.py:95-99 / :114.  The output is written channel-major
// (B,C,8,29) and the caller REINTERPRETS it as (B,8,29*C) exactly like the reference's `.view(bs,8,-1)`
// (a memory reinterpretation, not a transpose -- trained weights depend on it).
//
// HBM-bound: 16 KB in, C*8*29*4 B out per sample.  One CTA per (sample, 4-row band): the band (4x128
// floats) and the C*48 filter taps are staged in shared memory, each thread produces one pooled
// output (4 conv positions x 48 taps).  The backward pass recomputes the conv to find the pooled
// argmax and accumulates every filter-gradient element in a register of its owner thread.
#include "common.cuh"

namespace {

constexpr int KH = 4, KW = 12, W_IN = 128, W_POOL = 29, MAXC = 16;

__global__ void __launch_bounds__(320) texture_fwd_kernel(const float* __restrict__ pr, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int C, float* out,
                                                          signed char* amax) {
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
        int bq = -1;         // pooled arg-max position; -1: relu inactive everywhere (no gradient flows)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s = bs[ch];
            const int w0 = wp * 4 + q;
#pragma unroll
            for (int dr = 0; dr < KH; ++dr)
#pragma unroll
                for (int dc = 0; dc < KW; ++dc) s = fmaf(f[dr * KW + dc], band[dr][w0 + dc], s);
            if (s > best) { best = s; bq = q; }          // first maximum wins (max_pool2d backward convention)
        }
        out[(((long)b * C + ch) * 8 + i) * W_POOL + wp] = best;
        if (amax) amax[(((long)b * C + ch) * 8 + i) * W_POOL + wp] = (signed char)bq;
    }
}

// Backward: only the filter / bias gradients exist (the piano-roll needs none).  Persistent CTAs walk the (sample, band)
// items; per item the C*29 pooled outputs recompute their conv window to find the arg-max position (phase 1), then
// thread (channel, tap) -- the OWNER of one filter-gradient element, accumulated in a register across all items of the
// CTA -- adds gout * input over the channel's 29 outputs (phase 2).  No shared-memory atomics (the previous one-CTA-per-
// band kernel spent its 204 us in 29-way conflicting shared atomics on 48 addresses per channel and 2 M global atomics);
// one global atomicAdd per filter element per CTA.
constexpr int BWD_THREADS = 512;
__global__ void __launch_bounds__(BWD_THREADS) texture_bwd_kernel(const float* __restrict__ pr, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, int C, long n_items,
                                                                   const float* __restrict__ gout, float* dw, float* dbias,
                                                                   const signed char* __restrict__ amax) {
    __shared__ float band[KH][W_IN];
    __shared__ float ws[MAXC * KH * KW];
    __shared__ float bs[MAXC];
    __shared__ float gq[MAXC * W_POOL];            // gout of the pooled output, 0 where no gradient flows
    __shared__ int qpos[MAXC * W_POOL];            // first column of its arg-max conv window
    const int t = threadIdx.x;
    if (amax == nullptr) {
        for (int k = t; k < C * KH * KW; k += BWD_THREADS) ws[k] = w[k];
        if (t < C) bs[t] = bias[t];
    }
    const int ch_own = t / (KH * KW), tap = t % (KH * KW), dr_own = tap / KW, dc_own = tap % KW;
    const bool owner = t < C * KH * KW;
    float acc = 0.0f, dbacc = 0.0f;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const long b = item >> 3;
        const int i = (int)(item & 7);
        __syncthreads();                           // previous item's phase 2 is done with band / gq / qpos
        const float* src = pr + (b * 32 + i * 4) * W_IN;
        for (int k = t; k < KH * W_IN; k += BWD_THREADS) band[k / W_IN][k % W_IN] = src[k];
        __syncthreads();
        for (int o = t; o < C * W_POOL; o += BWD_THREADS) {
            const int ch = o / W_POOL, wp = o % W_POOL;
            const float g = gout[((b * C + ch) * 8 + i) * W_POOL + wp];
            float best = 0.0f;
            int bq = -1;                           // -1: relu inactive everywhere -> no gradient
            if (amax != nullptr) {                 // the forward saved the arg-max: no recomputation of the conv windows
                bq = amax[((b * C + ch) * 8 + i) * W_POOL + wp];
            } else if (g != 0.0f) {
                const float* f = ws + ch * KH * KW;
                for (int q = 0; q < 4; ++q) {
                    float s = bs[ch];
                    const int w0 = wp * 4 + q;
#pragma unroll
                    for (int dr = 0; dr < KH; ++dr)
#pragma unroll
                        for (int dc = 0; dc < KW; ++dc) s = fmaf(f[dr * KW + dc], band[dr][w0 + dc], s);
                    if (s > best) { best = s; bq = q; }   // first maximum wins (max_pool2d backward convention)
                }
            }
            gq[o] = bq < 0 ? 0.0f : g;
            qpos[o] = wp * 4 + (bq < 0 ? 0 : bq);
        }
        __syncthreads();
        if (owner) {
            const float* gr = gq + ch_own * W_POOL;
            const int* qp = qpos + ch_own * W_POOL;
            const float* row = band[dr_own] + dc_own;
#pragma unroll 1
            for (int wp = 0; wp < W_POOL; ++wp) acc = fmaf(gr[wp], row[qp[wp]], acc);
        }
        if (t < C) {
            const float* gr = gq + t * W_POOL;
            for (int wp = 0; wp < W_POOL; ++wp) dbacc += gr[wp];
        }
    }
    if (owner && acc != 0.0f) atomicAdd(dw + t, acc);
    if (t < C && dbacc != 0.0f) atomicAdd(dbias + t, dbacc);
}

}  // namespace

PD_API int pd_texture_frontend_fwd(const float* pr_mat, const float* w, const float* bias, int B, int C,
                                   float* out, void* stream) {
    if (B <= 0) return 0;
    if (C < 1 || C > MAXC) return PD_BAD_ARG;
    texture_fwd_kernel<<<B * 8, 320, 0, (cudaStream_t)stream>>>(pr_mat, w, bias, C, out, nullptr);
    return pd_launch_status();
}

// Training form: also writes amax (B,C,8,29) int8, the pooled arg-max position of every output (-1: ReLU inactive), which
// pd_texture_frontend_bwd_ix reads instead of recomputing the convolution (the recomputation, not the reduction, was the
// cost of the backward: 182 of 204 us)
PD_API int pd_texture_frontend_fwd_ix(const float* pr_mat, const float* w, const float* bias, int B, int C, float* out,
                                      signed char* amax, void* stream) {
    if (B <= 0) return 0;
    if (C < 1 || C > MAXC || amax == nullptr) return PD_BAD_ARG;
    texture_fwd_kernel<<<B * 8, 320, 0, (cudaStream_t)stream>>>(pr_mat, w, bias, C, out, amax);
    return pd_launch_status();
}

// dw (C*48) and dbias (C) are ACCUMULATED into (caller zeroes them).
static int texture_bwd_impl(const float* pr_mat, const float* w, const float* bias, int B, int C,
                            const float* gout, float* dw, float* dbias, const signed char* amax, void* stream) {
    if (B <= 0) return 0;
    if (C < 1 || C > MAXC) return PD_BAD_ARG;
    if (C * KH * KW > BWD_THREADS) return PD_BAD_ARG;
    const long n_items = (long)B * 8;
    const long want = 2L * PD_NUM_SMS;
    texture_bwd_kernel<<<(unsigned)(n_items < want ? n_items : want), BWD_THREADS, 0, (cudaStream_t)stream>>>(pr_mat, w, bias, C,
                                                                                                          n_items, gout, dw, dbias, amax);
    return pd_launch_status();
}

PD_API int pd_texture_frontend_bwd(const float* pr_mat, const float* w, const float* bias, int B, int C,
                                   const float* gout, float* dw, float* dbias, void* stream) {
    return texture_bwd_impl(pr_mat, w, bias, B, C, gout, dw, dbias, nullptr, stream);
}

PD_API int pd_texture_frontend_bwd_ix(const float* pr_mat, const signed char* amax, int B, int C, const float* gout, float* dw,
                                      float* dbias, void* stream) {
    if (amax == nullptr) return PD_BAD_ARG;
    return texture_bwd_impl(pr_mat, nullptr, nullptr, B, C, gout, dw, dbias, amax, stream);
}
