// Generic fp32 GEMM on the CUDA cores (FFMA), any shape / any of the NT, NN, TN layouts.
//
//   C[m, n] (+)= sum_k A(m,k) * B(k,n) (+ bias[n])
//   A(m,k) = A[m*sam + k*sak]   with sak == 1 (row-major A) or sam == 1 (A given transposed)
//   B(k,n) = B[k*sbk + n*sbn]   with sbk == 1 (nn.Linear weight [N,K]) or sbn == 1 ([K,N])
//
// Role in the path: (1) the fp32-faithful arithmetic the greedy decoder needs for token parity
// (bf16/tf32 operands flip near-tied argmaxes), (2) shapes TMA cannot address
// (K = 290, N = 130, ...), (3) the on-GPU cross-check for the tcgen05 kernel in gemm_tc.cu.
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread (as two 4-wide halves so every
// shared-memory read is a conflict-free LDS.128), global->register prefetch of the next k-slab while
// the current one is multiplied, optional split-K (atomicAdd epilogue) when the tile grid alone
// cannot fill 148 SMs (the weight-gradient GEMMs: small MxN, K = rows of the batch).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

struct GemmArgs {
    const float* A; long sam, sak;
    const float* B; long sbk, sbn;
    float* C; long ldc;
    const float* bias;
    int M, N, K;
    int accumulate;   // C += result
    int kchunk;       // K range per blockIdx.z (multiple of BK)
    int atomic;       // split-K: atomicAdd epilogue
};

template <bool KCONTIG>
__device__ __forceinline__ void load_slab(const float* __restrict__ P, long s_outer, long s_k, int outer0,
                                          int k0, int n_outer, int k_end, float (&r)[8]) {
    // tile is 128 (outer: m or n) x 16 (k).  KCONTIG: k fastest in memory.
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int idx = i * NT + tid;
        int o, k;
        if (KCONTIG) { o = idx >> 4; k = idx & 15; } else { k = idx >> 7; o = idx & 127; }
        int go = outer0 + o, gk = k0 + k;
        r[i] = (go < n_outer && gk < k_end) ? __ldg(P + (long)go * s_outer + (long)gk * s_k) : 0.0f;
    }
}

template <bool KCONTIG>
__device__ __forceinline__ void store_slab(float (*S)[BM + PAD], const float (&r)[8]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int idx = i * NT + tid;
        int o, k;
        if (KCONTIG) { o = idx >> 4; k = idx & 15; } else { k = idx >> 7; o = idx & 127; }
        S[k][o] = r[i];
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(NT, 2) gemm_f32_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * g.kchunk;
    const int kend = min(g.K, kbeg + g.kchunk);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    float ra[8], rb[8];
    load_slab<A_KC>(g.A, g.sam, g.sak, m0, kbeg, g.M, kend, ra);
    load_slab<B_KC>(g.B, g.sbn, g.sbk, n0, kbeg, g.N, kend, rb);

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        store_slab<A_KC>(As, ra);
        store_slab<B_KC>(Bs, rb);
        __syncthreads();
        if (k0 + BK < kend) {
            load_slab<A_KC>(g.A, g.sam, g.sak, m0, k0 + BK, g.M, kend, ra);
            load_slab<B_KC>(g.B, g.sbn, g.sbk, n0, k0 + BK, g.N, kend, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool add_bias = g.bias != nullptr && blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= g.N) continue;
            float v = acc[i][j] + (add_bias ? __ldg(g.bias + n) : 0.0f);
            float* c = g.C + (long)m * g.ldc + n;
            if (g.atomic || g.accumulate) atomicAdd(c, v);   // RED at L2: no read round trip through the SM
            else *c = v;
        }
    }
}

__global__ void zero_2d_kernel(float* C, long ldc, int M, int N) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long)M * N) C[(i / N) * ldc + (i % N)] = 0.0f;
}

// out[n] (+)= sum_m X[m*ldx + n]; one warp-wide column strip per block.x, M split over block.y.
__global__ void colsum_kernel(const float* __restrict__ X, long ldx, int M, int N, float* out, int rows_per_blk) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int mb = blockIdx.y * rows_per_blk, me = min(M, mb + rows_per_blk);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int m = mb;
    for (; m + 3 < me; m += 4) {
        s0 += __ldg(X + (long)m * ldx + n);
        s1 += __ldg(X + (long)(m + 1) * ldx + n);
        s2 += __ldg(X + (long)(m + 2) * ldx + n);
        s3 += __ldg(X + (long)(m + 3) * ldx + n);
    }
    for (; m < me; ++m) s0 += __ldg(X + (long)m * ldx + n);
    atomicAdd(out + n, (s0 + s1) + (s2 + s3));
}

// Vector variant for 16-byte aligned inputs: a block sums one 128-column strip over its row range with 8 warps
// striding the rows (lane = 4 columns, float4 loads, 4 rows in flight per thread), reduces the warps through
// shared memory and issues one atomicAdd per column.
// SEQ: the rows are the (sequence, step) pairs of an (R, T, N) buffer and only the steps below a sequence's length are
// read (the rest are the zeros a masked recurrence's backward left there: 3/4 of the summariser's rows).  In SEQ mode M
// counts SEQUENCES, rows_per_blk sequences per block, and a warp walks the live steps of every 8th sequence (no per-row
// division; consecutive steps of a sequence are consecutive rows).
template <bool SEQ>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const float* __restrict__ X, long ldx, int M, int N, float* out,
                                                         int rows_per_blk, const int* __restrict__ lengths, int T) {
    __shared__ float4 part[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 128 + lane * 4;
    const int mb = blockIdx.y * rows_per_blk, me = min(M, mb + rows_per_blk);
    float4 s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        if (SEQ) {
            for (int q = mb + warp; q < me; q += 8) {
                const int len = min(__ldg(lengths + q), T);
                const float* row = X + (long)q * T * ldx + n;
                int t = 0;
                for (; t + 3 < len; t += 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(row + (long)(t + i) * ldx));
                        s[i].x += v.x; s[i].y += v.y; s[i].z += v.z; s[i].w += v.w;
                    }
                }
                for (; t < len; ++t) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(row + (long)t * ldx));
                    s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
                }
            }
        } else {
            int m = mb + warp;
            for (; m + 24 < me; m += 32) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(X + (long)(m + 8 * i) * ldx + n));
                    s[i].x += v.x; s[i].y += v.y; s[i].z += v.z; s[i].w += v.w;
                }
            }
            for (; m < me; m += 8) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(X + (long)m * ldx + n));
                s[0].x += v.x; s[0].y += v.y; s[0].z += v.z; s[0].w += v.w;
            }
        }
    }
    part[warp][lane] = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                                   (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
    __syncthreads();
    if (threadIdx.x < 128) {
        const int c = threadIdx.x, col = blockIdx.x * 128 + c;
        if (col < N) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += reinterpret_cast<const float*>(&part[w][0])[c];
            atomicAdd(out + col, t);
        }
    }
}

// out[r*ldo + c] = sum_t X[r*ldr + t*ldt + c]   (float4 lanes; C % 4 == 0)
__global__ void __launch_bounds__(256) sum_steps_kernel(const float* __restrict__ X, long ldr, long ldt, int T, float* out,
                                                        long ldo, long R, int C4) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * C4) return;
    const long r = idx / C4;
    const int c = (int)(idx % C4) * 4;
    const float* p = X + r * ldr + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    int t = 0;
    for (; t + 1 < T; t += 2) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(p + (long)t * ldt));
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + (long)(t + 1) * ldt));
        a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
        b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
    if (t < T) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(p + (long)t * ldt));
        a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
    }
    *reinterpret_cast<float4*>(out + r * ldo + c) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

}  // namespace

PD_API int pd_gemm_f32(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C,
                       long ldc, const float* bias, int M, int N, int K, int accumulate, void* stream) {
    if (M <= 0 || N <= 0) return 0;
    if (K < 0 || (sak != 1 && sam != 1) || (sbk != 1 && sbn != 1)) return PD_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    GemmArgs g{A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, 0, 0};
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, 1);   // M on x: no 65535 limit
    long tiles = (long)grid.x * grid.y;
    int ksteps = (K + BK - 1) / BK;
    int split = 1;
    if (tiles < PD_NUM_SMS && ksteps >= 16) {
        split = (int)((2 * PD_NUM_SMS + tiles - 1) / tiles);
        if (split > ksteps / 8) split = ksteps / 8;
        if (split < 1) split = 1;
    }
    int steps_per = (ksteps + split - 1) / split;
    if (steps_per < 1) steps_per = 1;          // K == 0: one empty slab, C = bias
    split = (ksteps + steps_per - 1) / steps_per;
    if (split < 1) split = 1;
    g.kchunk = steps_per * BK;
    g.atomic = split > 1;
    grid.z = split;
    if (g.atomic && !accumulate) {
        long n = (long)M * N;
        zero_2d_kernel<<<pd_blocks(n, 256), 256, 0, st>>>(C, ldc, M, N);
    }
    const bool akc = (sak == 1), bkc = (sbk == 1);
    if (akc && bkc) gemm_f32_kernel<true, true><<<grid, NT, 0, st>>>(g);
    else if (akc && !bkc) gemm_f32_kernel<true, false><<<grid, NT, 0, st>>>(g);
    else if (!akc && bkc) gemm_f32_kernel<false, true><<<grid, NT, 0, st>>>(g);
    else gemm_f32_kernel<false, false><<<grid, NT, 0, st>>>(g);
    return pd_launch_status();
}

PD_API int pd_colsum_f32(const float* X, long ldx, int M, int N, float* out, int accumulate, void* stream) {
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, st);
    if (M <= 0) return pd_launch_status();
    int nbx = (N + 127) / 128;
    if ((((uintptr_t)X) & 15) == 0 && (ldx & 3) == 0 && (N & 3) == 0 && M >= 256) {
        int want = (6 * PD_NUM_SMS + nbx - 1) / nbx;
        int rows = (M + want - 1) / want;
        rows = ((rows < 64 ? 64 : rows) + 7) / 8 * 8;
        dim3 gridv(nbx, (M + rows - 1) / rows);
        colsum_vec_kernel<false><<<gridv, 256, 0, st>>>(X, ldx, M, N, out, rows, nullptr, 1);
        return pd_launch_status();
    }
    int want_y = (4 * PD_NUM_SMS + nbx - 1) / nbx;
    int rows_per = (M + want_y - 1) / want_y;
    if (rows_per < 32) rows_per = 32;
    dim3 grid(nbx, (M + rows_per - 1) / rows_per);
    colsum_kernel<<<grid, 128, 0, st>>>(X, ldx, M, N, out, rows_per);
    return pd_launch_status();
}

// out[n] (+)= sum over the LIVE rows of X (R*T rows of N columns, row stride ldx): row m = (sequence m / T, step m % T) is
// live iff m % T < lengths[m / T].  The bias gradients of a length-masked recurrence (the note-summary bi-GRU): the dead
// rows hold zeros, so the result equals pd_colsum_f32's without reading them.  16-byte aligned X, ldx % 4 == 0, N % 4 == 0.
PD_API int pd_colsum_seq_f32(const float* X, long ldx, int R, int T, int N, const int* lengths, float* out, int accumulate,
                             void* stream) {
    if (N <= 0) return 0;
    if ((((uintptr_t)X) & 15) || (ldx & 3) || (N & 3) || T <= 0 || lengths == nullptr || (long)R * T > 2147483647L) return PD_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, st);
    if (R <= 0) return pd_launch_status();
    const int nbx = (N + 127) / 128;
    int want = (6 * PD_NUM_SMS + nbx - 1) / nbx;                 // blocks per column strip
    int seqs = (R + want - 1) / want;                           // sequences per block
    seqs = ((seqs < 8 ? 8 : seqs) + 7) / 8 * 8;
    dim3 gridv(nbx, (R + seqs - 1) / seqs);
    colsum_vec_kernel<true><<<gridv, 256, 0, st>>>(X, ldx, R, N, out, seqs, lengths, T);
    return pd_launch_status();
}

// out (R,C; row stride ldo) = sum over the T steps of X (R,T,C; strides ldr, ldt in floats): the gradient of a
// projection that is broadcast over a GRU's steps.  C % 4 == 0, strides % 4 == 0, 16-byte aligned bases.
PD_API int pd_sum_steps_f32(const float* X, long ldr, long ldt, int T, float* out, long ldo, long R, int C, void* stream) {
    if (R <= 0 || C <= 0) return 0;
    if ((C & 3) || (ldr & 3) || (ldt & 3) || (ldo & 3) || (((uintptr_t)X | (uintptr_t)out) & 15) || T < 0) return PD_BAD_ARG;
    const long n = R * (C / 4);
    sum_steps_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(X, ldr, ldt, T, out, ldo, R, C / 4);
    return pd_launch_status();
}
