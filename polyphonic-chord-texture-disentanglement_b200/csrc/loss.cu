// Loss-side kernels: masked mean cross-entropy (fwd + bwd), reparameterisation, KL to N(0,I), exp.
// Replaces nn.CrossEntropyLoss(ignore_index) at ptvae.py:498-511 / model.py:70-83, Normal.rsample and
// kl_divergence(...).mean() at train_utils.py:33-49.  HBM-bound elementwise / row reductions.
#include "common.cuh"

namespace {

// acc[0] += sum over valid rows of (logsumexp(row) - row[target]); acc[1] += number of valid rows.
// Wide rows (C > 32): one warp per row.  Narrow rows: one thread per row.
__global__ void __launch_bounds__(256) ce_fwd_wide_kernel(const float* __restrict__ logits, long ldl,
                                                          const int* __restrict__ tgt, long R, int C, int ignore,
                                                          float* acc) {
    long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    float loss = 0.0f, cnt = 0.0f;
    if (r < R) {
        int t = tgt[r];
        if (t != ignore) {
            const float* p = logits + r * ldl;
            float mx = -INFINITY;
            for (int i = lane; i < C; i += 32) mx = fmaxf(mx, p[i]);
            mx = warp_max(mx);
            float s = 0.0f;
            for (int i = lane; i < C; i += 32) s += expf(p[i] - mx);
            s = warp_sum(s);
            loss = logf(s) + mx - p[t];
            cnt = 1.0f;
        }
    }
    // block reduce (lane 0 of each warp holds the row value)
    __shared__ float sl[8], sc[8];
    if (lane == 0) { sl[threadIdx.x >> 5] = loss; sc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += sl[i]; b += sc[i]; }
        if (b != 0.0f) { atomicAdd(acc, a); atomicAdd(acc + 1, b); }
    }
}

__global__ void __launch_bounds__(256) ce_fwd_narrow_kernel(const float* __restrict__ logits, long ldl,
                                                            const int* __restrict__ tgt, long R, int C, int ignore,
                                                            float* acc) {
    long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.0f, cnt = 0.0f;
    if (r < R) {
        int t = tgt[r];
        if (t != ignore) {
            const float* p = logits + r * ldl;
            float mx = p[0];
            for (int i = 1; i < C; ++i) mx = fmaxf(mx, p[i]);
            float s = 0.0f;
            for (int i = 0; i < C; ++i) s += expf(p[i] - mx);
            loss = logf(s) + mx - p[t];
            cnt = 1.0f;
        }
    }
    loss = warp_sum(loss); cnt = warp_sum(cnt);
    __shared__ float sl[8], sc[8];
    const int lane = threadIdx.x & 31;
    if (lane == 0) { sl[threadIdx.x >> 5] = loss; sc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += sl[i]; b += sc[i]; }
        if (b != 0.0f) { atomicAdd(acc, a); atomicAdd(acc + 1, b); }
    }
}

// dlogits[r, i] = (softmax(row)[i] - [i == t]) * gout / count   (0 for ignored rows)
__global__ void __launch_bounds__(256) ce_bwd_wide_kernel(const float* __restrict__ logits, long ldl,
                                                          const int* __restrict__ tgt, long R, int C, int ignore,
                                                          const float* __restrict__ acc,
                                                          const float* __restrict__ gout, float* dl, long lddl) {
    long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const int lane = threadIdx.x & 31;
    int t = tgt[r];
    float* d = dl + r * lddl;
    if (t == ignore) { for (int i = lane; i < C; i += 32) d[i] = 0.0f; return; }
    const float scale = gout[0] / acc[1];
    const float* p = logits + r * ldl;
    float mx = -INFINITY;
    for (int i = lane; i < C; i += 32) mx = fmaxf(mx, p[i]);
    mx = warp_max(mx);
    float s = 0.0f;
    for (int i = lane; i < C; i += 32) s += expf(p[i] - mx);
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int i = lane; i < C; i += 32) d[i] = (expf(p[i] - mx) * inv - (i == t ? 1.0f : 0.0f)) * scale;
}

__global__ void __launch_bounds__(256) ce_bwd_narrow_kernel(const float* __restrict__ logits, long ldl,
                                                            const int* __restrict__ tgt, long R, int C, int ignore,
                                                            const float* __restrict__ acc,
                                                            const float* __restrict__ gout, float* dl, long lddl) {
    long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int t = tgt[r];
    float* d = dl + r * lddl;
    if (t == ignore) { for (int i = 0; i < C; ++i) d[i] = 0.0f; return; }
    const float scale = gout[0] / acc[1];
    const float* p = logits + r * ldl;
    float mx = p[0];
    for (int i = 1; i < C; ++i) mx = fmaxf(mx, p[i]);
    float s = 0.0f;
    for (int i = 0; i < C; ++i) s += expf(p[i] - mx);
    const float inv = 1.0f / s;
    for (int i = 0; i < C; ++i) d[i] = (expf(p[i] - mx) * inv - (i == t ? 1.0f : 0.0f)) * scale;
}

__global__ void ce_finish_kernel(const float* acc, float* loss) { loss[0] = acc[0] / acc[1]; }

// ---- posterior helpers -------------------------------------------------------------------------
__global__ void exp_fwd_kernel(const float* __restrict__ x, long n, float* y) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = expf(x[i]);
}
// dx = dy * y
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, float* out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, float* out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
// lo = x - rn_tf32(x): the part of an fp32 operand a TF32 tensor-core multiply drops (error-compensated
// "3xTF32" GEMM: A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi, the hi parts being what TMA's TFLOAT32 load yields).
// hi is written explicitly (fp32 with the 13 low mantissa bits cleared) so the result does not depend on how
// the TMA / tensor-core path rounds an fp32 operand to TF32.
__global__ void tf32_split_kernel(const float* __restrict__ x, long ldx, long rows, int cols, float* __restrict__ hi_out,
                                  float* __restrict__ lo, long ldo) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    long r = i / cols;
    int c = (int)(i % cols);
    float v = x[r * ldx + c];
    uint32_t hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    hi_out[r * ldo + c] = __uint_as_float(hi);
    lo[r * ldo + c] = v - __uint_as_float(hi);
}
// Operand of the single-launch 3xTF32 GEMM: out row = three K-segments of width kp (cols padded to a multiple of 4,
// padding zero): order 0 (A side) [hi | hi | lo], order 1 (B side) [hi | lo | hi], so that ONE GEMM over K = 3*kp
// accumulates A_hi.B_hi + A_hi.B_lo + A_lo.B_hi in TMEM instead of three launches meeting in an L2-reduction epilogue.
__global__ void tf32_split3_kernel(const float* __restrict__ x, long ldx, long rows, int cols, int kp, float* __restrict__ out,
                                   long ldo, int order) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * kp) return;
    long r = i / kp;
    int c = (int)(i % kp);
    float v = c < cols ? x[r * ldx + c] : 0.0f;
    uint32_t hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float hi = __uint_as_float(hb), lo = v - hi;
    float* o = out + r * ldo + c;
    o[0] = hi;
    o[kp] = order ? lo : hi;
    o[2 * kp] = order ? hi : lo;
}
// z[b, j] = mu + std * eps, written with row stride ldz (into its half of dec_z)
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sd,
                                   const float* __restrict__ eps, int B, int D, float* z, long ldz) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    int b = (int)(i / D), j = (int)(i % D);
    z[(long)b * ldz + j] = eps ? fmaf(sd[i], eps[i], mu[i]) : mu[i];
}
// dmu = dz ; dstd = dz * eps  (dz read with row stride lddz)
__global__ void reparam_bwd_kernel(const float* __restrict__ dz, long lddz, const float* __restrict__ eps, int B,
                                   int D, float* dmu, float* dsd) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    int b = (int)(i / D), j = (int)(i % D);
    float g = dz[(long)b * lddz + j];
    dmu[i] = g;
    dsd[i] = eps ? g * eps[i] : 0.0f;
}
// out += sum_i (-log sd + (sd^2 + mu^2)/2 - 1/2) / n
__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sd,
                                                     long n, float inv_n, float* out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.0f;
    if (i < n) { float s = sd[i], m = mu[i]; v = (-logf(s) + 0.5f * (s * s + m * m) - 0.5f) * inv_n; }
    v = warp_sum(v);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int k = 0; k < 8; ++k) a += sh[k];
        atomicAdd(out, a);
    }
}
// dmu = g * mu / n ; dsd = g * (sd - 1/sd) / n
__global__ void kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ sd, long n, float inv_n,
                              const float* __restrict__ gout, float* dmu, float* dsd) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = gout[0] * inv_n, s = sd[i];
    dmu[i] = g * mu[i];
    dsd[i] = g * (s - 1.0f / s);
}

// Scheduled sampling with a DEVICE-resident teacher-forcing plan (one CUDA graph for every ratio): the next input is
// the ground-truth row if *flag != 0, else the predicted row (ptvae.py:420-424, :476-486, :84-86); one decision per
// step for the whole batch, like the reference's single random.random() per step.
__global__ void select_rows_kernel(const float* __restrict__ a, long lda, const float* __restrict__ b, long ldb,
                                   const int* __restrict__ flag, float* __restrict__ out, long ldo, long rows, int cols) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long r = i / cols;
    const int c = (int)(i % cols);
    out[r * ldo + c] = (*flag != 0) ? a[r * lda + c] : b[r * ldb + c];
}
// gradient routing of the select: the branch that was not taken gets zeros
__global__ void select_rows_bwd_kernel(const float* __restrict__ dout, long ldd, const int* __restrict__ flag,
                                       float* __restrict__ da, long ldda, float* __restrict__ db, long lddb, long rows,
                                       int cols) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long r = i / cols;
    const int c = (int)(i % cols);
    const float g = dout[r * ldd + c];
    const bool take_a = *flag != 0;
    if (da) da[r * ldda + c] = take_a ? g : 0.0f;
    if (db) db[r * lddb + c] = take_a ? 0.0f : g;
}

}  // namespace

// loss[0] = mean over rows with target != ignore of CE(logits[r], target[r]); acc2 = {sum, count} scratch
// (kept for the backward pass).
PD_API int pd_ce_fwd(const float* logits, long ldl, const int* targets, long R, int C, int ignore, float* acc2,
                     float* loss, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(acc2, 0, 2 * sizeof(float), st);
    if (R > 0) {
        if (C > 32) ce_fwd_wide_kernel<<<pd_blocks(R * 32, 256), 256, 0, st>>>(logits, ldl, targets, R, C, ignore, acc2);
        else ce_fwd_narrow_kernel<<<pd_blocks(R, 256), 256, 0, st>>>(logits, ldl, targets, R, C, ignore, acc2);
    }
    ce_finish_kernel<<<1, 1, 0, st>>>(acc2, loss);
    return pd_launch_status();
}

PD_API int pd_ce_bwd(const float* logits, long ldl, const int* targets, long R, int C, int ignore,
                     const float* acc2, const float* gout, float* dlogits, long lddl, void* stream) {
    if (R <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (C > 32)
        ce_bwd_wide_kernel<<<pd_blocks(R * 32, 256), 256, 0, st>>>(logits, ldl, targets, R, C, ignore, acc2, gout,
                                                                   dlogits, lddl);
    else
        ce_bwd_narrow_kernel<<<pd_blocks(R, 256), 256, 0, st>>>(logits, ldl, targets, R, C, ignore, acc2, gout,
                                                                dlogits, lddl);
    return pd_launch_status();
}

PD_API int pd_exp_fwd(const float* x, long n, float* y, void* stream) {
    if (n <= 0) return 0;
    exp_fwd_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, y);
    return pd_launch_status();
}

PD_API int pd_mul_f32(const float* a, const float* b, long n, float* out, void* stream) {
    if (n <= 0) return 0;
    mul_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    return pd_launch_status();
}

PD_API int pd_add_f32(const float* a, const float* b, long n, float* out, void* stream) {
    if (n <= 0) return 0;
    add_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    return pd_launch_status();
}

// hi = round_to_tf32(x), lo = x - hi for x (rows, cols; row stride ldx); outputs with row stride ldo
PD_API int pd_tf32_split(const float* x, long ldx, long rows, int cols, float* hi, float* lo, long ldo, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    tf32_split_kernel<<<pd_blocks(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, hi, lo, ldo);
    return pd_launch_status();
}

// out (rows, 3*kp; row stride ldo), kp = cols rounded up to a multiple of 4: [hi | hi | lo] (order 0) or [hi | lo | hi]
// (order 1) of x (rows, cols; row stride ldx)
PD_API int pd_tf32_split3(const float* x, long ldx, long rows, int cols, float* out, long ldo, int order, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    const int kp = (cols + 3) / 4 * 4;
    if (ldo < 3L * kp) return PD_BAD_ARG;
    tf32_split3_kernel<<<pd_blocks(rows * kp, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, kp, out, ldo, order);
    return pd_launch_status();
}

PD_API int pd_reparam_fwd(const float* mu, const float* sd, const float* eps, int B, int D, float* z, long ldz,
                          void* stream) {
    long n = (long)B * D;
    if (n <= 0) return 0;
    reparam_fwd_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(mu, sd, eps, B, D, z, ldz);
    return pd_launch_status();
}

PD_API int pd_reparam_bwd(const float* dz, long lddz, const float* eps, int B, int D, float* dmu, float* dsd,
                          void* stream) {
    long n = (long)B * D;
    if (n <= 0) return 0;
    reparam_bwd_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(dz, lddz, eps, B, D, dmu, dsd);
    return pd_launch_status();
}

PD_API int pd_kl_fwd(const float* mu, const float* sd, long n, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(float), st);
    if (n > 0) kl_fwd_kernel<<<pd_blocks(n, 256), 256, 0, st>>>(mu, sd, n, 1.0f / (float)n, out);
    return pd_launch_status();
}

PD_API int pd_kl_bwd(const float* mu, const float* sd, long n, const float* gout, float* dmu, float* dsd,
                     void* stream) {
    if (n <= 0) return 0;
    kl_bwd_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(mu, sd, n, 1.0f / (float)n, gout, dmu, dsd);
    return pd_launch_status();
}

// out (rows, cols) = *flag ? a : b  (row strides in floats; flag: one int32 on the device)
PD_API int pd_select_rows(const float* a, long lda, const float* b, long ldb, const int* flag, float* out, long ldo,
                          long rows, int cols, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    select_rows_kernel<<<pd_blocks(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, flag, out, ldo, rows, cols);
    return pd_launch_status();
}

// da = *flag ? dout : 0, db = *flag ? 0 : dout  (either may be NULL)
PD_API int pd_select_rows_bwd(const float* dout, long ldd, const int* flag, float* da, long ldda, float* db, long lddb,
                              long rows, int cols, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    select_rows_bwd_kernel<<<pd_blocks(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(dout, ldd, flag, da, ldda, db, lddb,
                                                                                         rows, cols);
    return pd_launch_status();
}
