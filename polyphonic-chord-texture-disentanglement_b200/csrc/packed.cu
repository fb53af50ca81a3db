// Packed note level of the teacher-forced PianoTree decoder (ops.py "packed notes").
//
// The reference runs all 15 note-GRU steps, the pitch / duration heads and the 5-step duration GRU for every one of the
// 32 x B (segment, time step) rows, although a step with k notes only has k + 1 target tokens: the loss ignores the other
// slots (ptvae.py:498-511, ignore_index = PAD) and no gradient leaves them.  With 4.7 tokens per step on average, 3/4 of
// that work is dead in loss mode.  Here the rows are SORTED by note count (the reference sorts too -- inside
// pack_padded_sequence for the note-summary GRU, ptvae.py:446-453) and every note-level buffer is SLOT-MAJOR
// (slot n, sorted row r), so the live rows of each slot are a prefix whose length is device data; kernels are launched for
// the full extent (one captured CUDA graph serves every batch) and skip dead tiles.
//
// This file: the sort itself, the token / target re-layout, row gathers between the (b,t) order of the time level and
// the sorted order of the note level, and the row-predicated reductions.
#include "common.cuh"

namespace {

constexpr int MAXLEN = 16;              // note slots per step; lengths are in [0, 16]
constexpr int NB = MAXLEN + 1;          // length buckets
constexpr int ORDER_THREADS = 512;

// Stable counting sort of the rows by length, DESCENDING (single CTA: R is 32 x batch <= 2^17).
// perm[i] = original row of sorted position i, inv[perm[i]] = i.  table (64 ints):
//   [0,17)  c[t]   = number of rows with length > t
//   [17,34) cp[t]  = min(R, c[t] rounded up to 128)     rows processed for slot t (whole 128-row tiles)
//   [34,51) cp6[t] = 6 * cp[t]                           the same in units of the duration decoder's 6 rows per note
__global__ void __launch_bounds__(ORDER_THREADS) pack_order_kernel(const int* __restrict__ lengths, int R, int* perm, int* inv,
                                                                   int* table) {
    __shared__ int hist[NB][ORDER_THREADS];      // rows of bucket b in thread t's chunk, then their start offsets
    __shared__ int base[NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // chunks are multiples of 4 rows, so a thread reads its rows as 16-byte vectors (it sits on the forward critical path
    // of the packed step: 60 us with scalar reads and a thread-serial scan, ~10 us now)
    const int chunk = ((R + ORDER_THREADS - 1) / ORDER_THREADS + 3) & ~3;
    const int r0 = min(R, tid * chunk), r1 = min(R, r0 + chunk);
    const bool vec = ((reinterpret_cast<uintptr_t>(lengths) & 15) == 0);
    int cnt[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) cnt[b] = 0;
    auto clampL = [](int L) { return L < 0 ? 0 : (L > MAXLEN ? MAXLEN : L); };
    auto count = [&](int L) {
        L = clampL(L);
#pragma unroll
        for (int b = 0; b < NB; ++b) cnt[b] += (L == b);
    };
    int r = r0;
    if (vec)
        for (; r + 3 < r1; r += 4) {
            const int4 v = *reinterpret_cast<const int4*>(lengths + r);
            count(v.x); count(v.y); count(v.z); count(v.w);
        }
    for (; r < r1; ++r) count(lengths[r]);
#pragma unroll
    for (int b = 0; b < NB; ++b) hist[b][tid] = cnt[b];
    __syncthreads();
    // exclusive scan of every bucket over the threads' chunks: one warp per bucket (warp 0 also takes the 17th), a lane
    // sums ORDER_THREADS / 32 consecutive entries, the warp scans the lane sums with shuffles
    constexpr int PER = ORDER_THREADS / 32;
    for (int b = warp; b < NB; b += ORDER_THREADS / 32) {
        int v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = hist[b][lane * PER + i]; sum += v[i]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        int run = incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) { hist[b][lane * PER + i] = run; run += v[i]; }
        if (lane == 31) base[b] = incl;          // bucket total
    }
    __syncthreads();
    if (tid == 0) {
        int tot[NB];
        for (int b = 0; b < NB; ++b) tot[b] = base[b];
        int run = 0;
        for (int b = MAXLEN; b >= 0; --b) { base[b] = run; run += tot[b]; }     // longest first
        int above = 0;                                                             // rows with length > t
        for (int t = MAXLEN; t >= 0; --t) {
            table[t] = above;
            const int p = (above + 127) / 128 * 128;
            table[NB + t] = p < R ? p : R;
            table[2 * NB + t] = 6 * (p < R ? p : R);
            above += tot[t];
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) cnt[b] = base[b] + hist[b][tid];
    auto place = [&](int row, int L) {
        L = clampL(L);
        int pos = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b)
            if (L == b) pos = cnt[b]++;
        perm[pos] = row;
        inv[row] = pos;
    };
    r = r0;
    if (vec)
        for (; r + 3 < r1; r += 4) {
            const int4 v = *reinterpret_cast<const int4*>(lengths + r);
            place(r, v.x); place(r + 1, v.y); place(r + 2, v.z); place(r + 3, v.w);
        }
    for (; r < r1; ++r) place(r, lengths[r]);
}

// tokens / targets of the sorted rows in slot-major layout: tok_s (16,R,6), pitch targets (15,R), duration targets
// (15,R,5), lengths of the sorted rows.  One thread per (slot, sorted row).
__global__ void pack_grid_kernel(const int* __restrict__ tok, const int* __restrict__ lengths, const int* __restrict__ perm,
                                 int R, int* tok_s, int* pt_s, int* dt_s, int* len_s) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)MAXLEN * R) return;
    const int n = (int)(idx / R), i = (int)(idx % R);
    const int src = perm[i];
    const int* t = tok + ((long)src * MAXLEN + n) * 6;
    int v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = t[k];
    int* o = tok_s + idx * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = v[k];
    if (n >= 1) {                                 // the target of note slot n - 1 is token n (ptvae.py:500-501)
        const long q = (long)(n - 1) * R + i;
        pt_s[q] = v[0];
#pragma unroll
        for (int k = 0; k < 5; ++k) dt_s[q * 5 + k] = v[1 + k];
    }
    if (n == 0) len_s[i] = lengths[src];
}

// dst[i, :] = src[idx[i], :]   (C % 4 == 0; float4 lanes)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, long lds, const int* __restrict__ idx,
                                                          long R, int C4, float* dst, long ldd) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * C4) return;
    const long i = t / C4;
    const int c = (int)(t % C4) * 4;
    *reinterpret_cast<float4*>(dst + i * ldd + c) = __ldg(reinterpret_cast<const float4*>(src + (long)idx[i] * lds + c));
}

// out[r, c] = sum over the slots t with r < cp[t] of X[t, r, c]   (X slot-major (T, R, C); cp non-increasing in t is NOT
// assumed).  The gradient of a projection that is broadcast over the note slots, from the live rows only.
__global__ void __launch_bounds__(256) sum_slots_rows_kernel(const float* __restrict__ X, long ldt, long ldr, int T,
                                                             const int* __restrict__ cp, float* out, long ldo, long R, int C4) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * C4) return;
    const long r = idx / C4;
    const int c = (int)(idx % C4) * 4;
    const float* p = X + r * ldr + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
        if (r >= cp[t]) continue;
        const float4 u = __ldg(reinterpret_cast<const float4*>(p + (long)t * ldt));
        a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
    }
    *reinterpret_cast<float4*>(out + r * ldo + c) = a;
}

// out[n] += sum over the LIVE rows m of X[m, n]   (rows slot-major; 32-row groups are live or dead as a whole:
// slot_rows % 32 == 0 and cp % 32 == 0 or cp == slot_rows).  256 threads = 8 warps x (32 lanes x float4 = 128 columns).
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ X, long ldx, long M, int N, float* out,
                                                          int groups_per_blk, PdRows live) {
    __shared__ float4 part[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 128 + lane * 4;
    const long g0 = (long)blockIdx.y * groups_per_blk, g1 = min((M + 31) / 32, g0 + groups_per_blk);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        for (long g = g0; g < g1; ++g) {
            if (!pd_rows_live(live, g * 32, 32)) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long m = g * 32 + warp + 8 * i;
                if (m < M) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(X + m * ldx + n));
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
            }
        }
    }
    part[warp][lane] = s;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int c = threadIdx.x, col = blockIdx.x * 128 + c;
        if (col < N) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += reinterpret_cast<const float*>(&part[w][0])[c];
            if (t != 0.0f) atomicAdd(out + col, t);
        }
    }
}

}  // namespace

// lengths (R) int32 in [0,16] -> perm / inv (R) and the 64-int live-row table (layout above).  R <= 131072.
PD_API int pd_pack_order(const int* lengths, int R, int* perm, int* inv, int* table, void* stream) {
    if (R <= 0) return 0;
    if (R > (1 << 17)) return PD_BAD_ARG;
    pack_order_kernel<<<1, ORDER_THREADS, 0, (cudaStream_t)stream>>>(lengths, R, perm, inv, table);
    return pd_launch_status();
}

// tok (R,16,6) int32 in row order, lengths (R), perm (R) -> slot-major sorted tok_s (16,R,6), pt_s (15,R), dt_s (15,R,5),
// len_s (R)
PD_API int pd_pack_grid(const int* tok, const int* lengths, const int* perm, int R, int* tok_s, int* pt_s, int* dt_s,
                        int* len_s, void* stream) {
    if (R <= 0) return 0;
    pack_grid_kernel<<<pd_blocks((long)MAXLEN * R, 256), 256, 0, (cudaStream_t)stream>>>(tok, lengths, perm, R, tok_s, pt_s, dt_s,
                                                                                       len_s);
    return pd_launch_status();
}

// dst (R,C; row stride ldd) = src[idx] (row stride lds); C % 4 == 0, strides % 4 == 0, 16-byte aligned bases
PD_API int pd_gather_rows_f32(const float* src, long lds, const int* idx, long R, int C, float* dst, long ldd, void* stream) {
    if (R <= 0 || C <= 0) return 0;
    if ((C & 3) || (lds & 3) || (ldd & 3) || (((uintptr_t)src | (uintptr_t)dst) & 15)) return PD_BAD_ARG;
    gather_rows_kernel<<<pd_blocks(R * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(src, lds, idx, R, C / 4, dst, ldd);
    return pd_launch_status();
}

// out (R,C; row stride ldo) = sum over slots t < T with r < cp[t] of X (T,R,C; slot stride ldt, row stride ldr)
PD_API int pd_sum_slots_rows_f32(const float* X, long ldt, long ldr, int T, const int* cp, float* out, long ldo, long R, int C,
                                 void* stream) {
    if (R <= 0 || C <= 0) return 0;
    if ((C & 3) || (ldr & 3) || (ldt & 3) || (ldo & 3) || (((uintptr_t)X | (uintptr_t)out) & 15) || T < 0 || cp == nullptr)
        return PD_BAD_ARG;
    sum_slots_rows_kernel<<<pd_blocks(R * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(X, ldt, ldr, T, cp, out, ldo, R, C / 4);
    return pd_launch_status();
}

// out (N) (+)= column sums of the LIVE rows of X (M,N; row stride ldx), rows slot-major with the live-row table cp
PD_API int pd_colsum_rows_f32(const float* X, long ldx, long M, int N, float* out, int accumulate, const int* cp, int slot_rows,
                              void* stream) {
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // (N need not be a multiple of 4: the float4 lanes may read into the row padding, ldx >= N rounded up to 4)
    if ((((uintptr_t)X) & 15) || (ldx & 3) || ((N + 3) / 4 * 4 > ldx) || cp == nullptr || slot_rows <= 0 || (slot_rows & 31))
        return PD_BAD_ARG;
    if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, st);
    if (M <= 0) return pd_launch_status();
    const int nbx = (N + 127) / 128;
    const long groups = (M + 31) / 32;
    long want = (8L * PD_NUM_SMS + nbx - 1) / nbx;
    long per = (groups + want - 1) / want;
    if (per < 1) per = 1;
    dim3 grid(nbx, (unsigned)((groups + per - 1) / per));
    colsum_rows_kernel<<<grid, 256, 0, st>>>(X, ldx, M, N, out, (int)per,
                                             PdRows{cp, slot_rows, (int)((M + slot_rows - 1) / slot_rows)});
    return pd_launch_status();
}
