// PianoTree grid handling: token preparation, note embedding as a gather (fwd) / column-owned
// reduction (bwd), greedy token picking and duration feedback tokens.  All HBM-bound integer/byte
// work: coalesced 512-byte rows, one warp per note, no dense multi-hot tensors.
#include "common.cuh"

namespace {

constexpr int NOTE_SLOTS = 16, TOK_W = 6, EMB = 128, P_RANGE = 130, NOTE_SIZE = 135;
constexpr int P_EOS = 129, P_PAD = 130;

// x (B,32,16,6) int64 -> tok int32 (same layout), lengths (B*32), pitch targets (B*32,15),
// dur targets (B*32,15,5).  One thread per (step, note slot): the 16 threads of a step read 16 x 48 contiguous bytes
// (a thread per step walked 768 bytes alone: 180 us for batch 512).             ptvae.py:292-297, :498-511
__global__ void __launch_bounds__(256) grid_prepare_kernel(const long long* __restrict__ x, long steps, int* tok, int* lengths,
                                                           int* pitch_tgt, int* dur_tgt) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long s = idx / NOTE_SLOTS;
    const int n = (int)(idx % NOTE_SLOTS);
    const bool ok = s < steps;
    int v[TOK_W];
#pragma unroll
    for (int k = 0; k < TOK_W; ++k) v[k] = ok ? (int)x[idx * TOK_W + k] : 0;
    if (ok) {
#pragma unroll
        for (int k = 0; k < TOK_W; ++k) tok[idx * TOK_W + k] = v[k];
        if (n >= 1) {
            if (pitch_tgt) pitch_tgt[s * 15 + (n - 1)] = v[0];
            if (dur_tgt) {
#pragma unroll
                for (int k = 1; k < TOK_W; ++k) dur_tgt[(s * 15 + (n - 1)) * 5 + (k - 1)] = v[k];
            }
        }
    }
    // PAD count of the step: the 16 slots of a step are 16 consecutive lanes of one warp
    const unsigned pad = __ballot_sync(0xffffffffu, ok && v[0] == P_PAD);
    if (ok && n == 0 && lengths) lengths[s] = NOTE_SLOTS - __popc((pad >> (threadIdx.x & 16)) & 0xffffu);
}

// emb[r, :] = bias + (p < 130 ? WT[p] : 0) + sum_k d_k * WT[130+k]   (== Linear(135->128) on the multi-hot,
// ptvae.py:299-313,:333,:531-535).  One warp per note, lane owns 4 consecutive outputs.
__global__ void __launch_bounds__(256) note_embed_fwd_kernel(const int* __restrict__ tok, long R,
                                                             const float* __restrict__ WT,
                                                             const float* __restrict__ bias, float* out, long ldo) {
    long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const int lane = threadIdx.x & 31;
    const int* t = tok + r * TOK_W;
    float4 acc = *reinterpret_cast<const float4*>(bias + lane * 4);
    int p = t[0];
    if (p >= 0 && p < P_RANGE) {
        float4 w = *reinterpret_cast<const float4*>(WT + (long)p * EMB + lane * 4);
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float d = (float)t[1 + k];
        if (d != 0.0f) {
            float4 w = *reinterpret_cast<const float4*>(WT + (long)(P_RANGE + k) * EMB + lane * 4);
            acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y);
            acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
        }
    }
    *reinterpret_cast<float4*>(out + r * ldo + lane * 4) = acc;
}

// dWT[p, j] += sum_r [tok_r.pitch == p] g[r, j] ; dWT[130+k, j] += sum_r d_k g[r, j] ; db[j] += sum_r g[r, j].
// Thread j owns column j of a CTA-private (135+1) x 128 accumulator in shared memory (no intra-CTA
// atomics), rows are streamed coalesced, one global atomicAdd per accumulator cell per CTA.
__global__ void __launch_bounds__(EMB) note_embed_bwd_kernel(const int* __restrict__ tok, long R,
                                                             const float* __restrict__ g, long ldg, float* dWT,
                                                             float* dbias, long rows_per_cta, PdRows live) {
    extern __shared__ float acc[];   // (NOTE_SIZE + 1) * EMB
    const int j = threadIdx.x;
    for (int i = 0; i <= NOTE_SIZE; ++i) acc[i * EMB + j] = 0.0f;
    // packed note level: rows are slot-major and only a prefix of each slot is live (dead rows carry no gradient and may
    // hold anything): the CTAs share the LIVE rows evenly and walk them in order
    PdLiveBlocks lb{live.cp, live.slot_rows, live.n_slots, 1, 0, 0, 0};
    long r0 = (long)blockIdx.x * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    if (live.cp) {
        const long total = lb.total();
        const long per = (total + gridDim.x - 1) / gridDim.x;
        r0 = min(total, (long)blockIdx.x * per);
        r1 = min(total, r0 + per);
        if (r0 < r1) lb.seek((int)r0);
    }
    for (long i = r0; i < r1; ++i) {
        long r = i;
        if (live.cp) { r = lb.row0(); lb.next(); }
        const int* t = tok + r * TOK_W;
        float v = g[r * ldg + j];
        int p = t[0];
        if (p >= 0 && p < P_RANGE) acc[p * EMB + j] += v;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            float d = (float)t[1 + k];
            if (d != 0.0f) acc[(P_RANGE + k) * EMB + j] += d * v;
        }
        acc[NOTE_SIZE * EMB + j] += v;
    }
    for (int i = 0; i < NOTE_SIZE; ++i) {
        float v = acc[i * EMB + j];
        if (v != 0.0f) atomicAdd(dWT + i * EMB + j, v);
    }
    atomicAdd(dbias + j, acc[NOTE_SIZE * EMB + j]);
}

// Greedy pick for one note slot n (1..15): argmax pitch (first maximum, like torch.max on CPU),
// argmax of each duration bit, token row for the embedding gather, EOS length bookkeeping
// (first n whose pitch is EOS; 15 if none).                           ptvae.py:408-416,:425
// one warp per row: returns (all lanes) the picked pitch; lane k < 5 returns its duration bit in `bit`
__device__ __forceinline__ int greedy_pick_row(const float* __restrict__ p, const float* __restrict__ d, int lane, int& bit) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < P_RANGE; i += 32) {
        float v = p[i];
        if (v > best || (v != v && best == best)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    bit = 0;
    if (lane < 5) bit = (d[lane * 2 + 1] > d[lane * 2]) ? 1 : 0;
    return bi;
}

__global__ void __launch_bounds__(256) greedy_pick_kernel(const float* __restrict__ pitch, long ldp,
                                                          const float* __restrict__ dur, long ldd, long R, int n,
                                                          int* tok, long ldtok, int* lens) {
    long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const int lane = threadIdx.x & 31;
    int bit;
    const int bi = greedy_pick_row(pitch + r * ldp, dur + r * ldd, lane, bit);
    if (lane == 0) {
        int* t = tok + r * ldtok;
        t[0] = bi;
        if (lens) {
            int L = lens[r];
            if (L == 0 && bi == P_EOS) L = n;
            if (n == NOTE_SLOTS - 1 && L == 0) L = NOTE_SLOTS - 1;
            lens[r] = L;
        }
    }
    if (lane < 5) tok[r * ldtok + 1 + lane] = bit;
}

// The same pick followed, in the same warp, by the embedding of the picked token (note_embed_fwd_kernel's arithmetic: bias
// first, then the pitch row, then the duration rows in bit order): the two tail kernels of a greedy note slot as one launch.
__global__ void __launch_bounds__(256) greedy_pick_embed_kernel(const float* __restrict__ pitch, long ldp,
                                                                const float* __restrict__ dur, long ldd, long R, int n,
                                                                int* tok, long ldtok, int* lens,
                                                                const float* __restrict__ WT, const float* __restrict__ bias,
                                                                float* emb, long lde) {
    long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const int lane = threadIdx.x & 31;
    int bit;
    const int bi = greedy_pick_row(pitch + r * ldp, dur + r * ldd, lane, bit);
    if (lane == 0) {
        int* t = tok + r * ldtok;
        t[0] = bi;
        if (lens) {
            int L = lens[r];
            if (L == 0 && bi == P_EOS) L = n;
            if (n == NOTE_SLOTS - 1 && L == 0) L = NOTE_SLOTS - 1;
            lens[r] = L;
        }
    }
    if (lane < 5) tok[r * ldtok + 1 + lane] = bit;
    const unsigned bits = __ballot_sync(0xffffffffu, bit != 0);
    float4 acc = *reinterpret_cast<const float4*>(bias + lane * 4);
    if (bi >= 0 && bi < P_RANGE) {
        float4 w = *reinterpret_cast<const float4*>(WT + (long)bi * EMB + lane * 4);
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (bits & (1u << k)) {
            float4 w = *reinterpret_cast<const float4*>(WT + (long)(P_RANGE + k) * EMB + lane * 4);
            acc.x = fmaf(1.0f, w.x, acc.x); acc.y = fmaf(1.0f, w.y, acc.y);
            acc.z = fmaf(1.0f, w.z, acc.z); acc.w = fmaf(1.0f, w.w, acc.w);
        }
    }
    *reinterpret_cast<float4*>(emb + r * lde + lane * 4) = acc;
}

// Duration feedback token: 5-wide vector with a single 1 at index == argmax bit (ptvae.py:322-326).
__global__ void dur_token_kernel(const float* __restrict__ logit, long ldl, long R, float* tok) {
    long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* l = logit + r * ldl;
    int one = (l[1] > l[0]) ? 1 : 0;
    float* t = tok + r * 5;
    t[0] = one ? 0.f : 1.f; t[1] = one ? 1.f : 0.f; t[2] = 0.f; t[3] = 0.f; t[4] = 0.f;
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* out) {
    __shared__ float tile[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (x < cols && y0 + i < rows) tile[i][threadIdx.x] = in[(long)(y0 + i) * cols + x];
    __syncthreads();
    int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (ox < rows && oy0 + i < cols) out[(long)(oy0 + i) * rows + ox] = tile[threadIdx.x][i];
}

// Chord decoder feedback (ptvae.py:73-78): the reference's advanced-index assignment makes every
// sample's root/bass one-hot the UNION over the batch of all samples' argmaxes.  Pass 1 marks the
// union flags, pass 2 writes the (B,36) tokens [union root | per-sample chroma argmax | union bass].
__global__ void chord_union_kernel(const float* __restrict__ root, long ldr, const float* __restrict__ bass,
                                   long ldb, int B, float* flags) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* r = root + (long)b * ldr;
    const float* s = bass + (long)b * ldb;
    int ri = 0, si = 0;
    for (int i = 1; i < 12; ++i) { if (r[i] > r[ri]) ri = i; if (s[i] > s[si]) si = i; }
    flags[ri] = 1.0f;
    flags[12 + si] = 1.0f;
}
__global__ void chord_token_kernel(const float* __restrict__ chroma, long ldc, const float* __restrict__ flags,
                                   int B, float* tok, long ldt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 36) return;
    int b = i / 36, k = i % 36;
    float v;
    if (k < 12) v = flags[k];
    else if (k < 24) { const float* c = chroma + (long)b * ldc + (k - 12) * 2; v = (c[1] > c[0]) ? 1.f : 0.f; }
    else v = flags[12 + (k - 24)];
    tok[(long)b * ldt + k] = v;
}

// c (B,8,36) -> CE targets: argmax root / chroma bit as class / argmax bass.   model.py:72-74
__global__ void chord_targets_kernel(const float* __restrict__ c, int rows, int* root, int* chroma, int* bass) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* p = c + (long)r * 36;
    int ri = 0, bi = 0;
    for (int i = 1; i < 12; ++i) { if (p[i] > p[ri]) ri = i; if (p[24 + i] > p[24 + bi]) bi = i; }
    root[r] = ri; bass[r] = bi;
    for (int i = 0; i < 12; ++i) chroma[r * 12 + i] = (int)p[12 + i];
}

// ---- data formats either side of the path ---------------------------------
// pr_mat (n_steps,128) durations-at-onset -> PianoTree grid x (n_steps,16,6) int64: the reference's
// converter.target_to_3dtarget as called in dataset.py:98-104.  One warp per time step: each lane owns 4
// pitches, a ballot + popc prefix gives every onset its slot (pitches ascending).  Steps with more than 14
// onsets do not fit the grid (the reference indexes out of bounds); they raise *overflow and are clipped.
__global__ void __launch_bounds__(256) prmat_to_grid_kernel(const float* __restrict__ pr, long n_steps,
                                                            long long* __restrict__ x, int* overflow) {
    long s = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_steps) return;
    const int lane = threadIdx.x & 31;
    float4 v = *reinterpret_cast<const float4*>(pr + s * 128 + lane * 4);
    const float d[4] = {v.x, v.y, v.z, v.w};
    long long* xs = x + s * NOTE_SLOTS * TOK_W;
    for (int i = lane; i < NOTE_SLOTS * TOK_W; i += 32) xs[i] = (i % TOK_W == 0) ? (i == 0 ? 128 : P_PAD) : 2;
    __syncwarp();
    int mine = (d[0] != 0.f) + (d[1] != 0.f) + (d[2] != 0.f) + (d[3] != 0.f);
    int pre = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, pre, o);
        if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    int slot = 1 + pre - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (d[k] != 0.f) {
            if (slot <= NOTE_SLOTS - 2) {
                int dur = (int)d[k] - 1;
                xs[slot * TOK_W] = lane * 4 + k;
#pragma unroll
                for (int b = 0; b < 5; ++b) xs[slot * TOK_W + 1 + b] = (dur >> (4 - b)) & 1;
            }
            ++slot;
        }
    }
    if (lane == 0) {
        if (total > NOTE_SLOTS - 2) atomicExch(overflow, 1);
        xs[(min(total, NOTE_SLOTS - 2) + 1) * TOK_W] = P_EOS;
    }
}

// decoded tokens (n_steps,15,6) int32 [pitch, 5 duration bits] -> pr_mat (n_steps,128): the loop of
// ptvae.py:558-575 (first 10 notes of a step, stop at EOS, duration = bits+1 clipped at the segment end).
// step index inside its 32-step segment = s % 32.
__global__ void grid_to_prmat_kernel(const int* __restrict__ tok, long n_steps, float* __restrict__ pr) {
    long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_steps) return;
    float* row = pr + s * 128;
    for (int i = 0; i < 128; ++i) row[i] = 0.0f;
    const int t = (int)(s % 32);
    const int* ts = tok + s * 15 * TOK_W;
    for (int n = 0; n < 10; ++n) {
        int p = ts[n * TOK_W];
        if (p == P_EOS) break;
        int dur = 0;
        for (int b = 0; b < 5; ++b) dur = dur * 2 + ts[n * TOK_W + 1 + b];
        dur += 1;
        if (p >= 0 && p < 128) row[p] = (float)min(dur, 32 - t);
    }
}

// decoded tokens int32 (R,6) [pitch 0..129, 5 duration bits] -> compact uint8 (R,2) [pitch, bits b0..b4 as 0b000b0b1b2b3b4]:
// what leaves the device after a decode (2 bytes per note instead of the 48 of the reference's int64 est_x,
// ptvae.py:537-544).  One thread per note; a warp reads 768 contiguous bytes and writes 64.
__global__ void pack_tokens_kernel(const int* __restrict__ tok, long R, uint8_t* __restrict__ out) {
    long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int* t = tok + r * TOK_W;
    int bits = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) bits = (bits << 1) | (t[1 + b] & 1);
    out[r * 2] = (uint8_t)t[0];
    out[r * 2 + 1] = (uint8_t)bits;
}

// ---- batch augmentation on device (dataset.py:67-120): transpose a segment by `shift` semitones -----------------
// pr_mat (B,32,128): np.roll along the pitch axis (converter.py:65-68; wraps around like np.roll).  The roll commutes
// with piano_roll_to_target (converter.py:87-113), which works column by column, so rolling pr_mat equals rolling
// the raw piano-roll first.  One thread per output element, coalesced along pitch.
__global__ void roll_prmat_kernel(const float* __restrict__ in, const int* __restrict__ shift, long B, float* __restrict__ out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 32 * 128) return;
    const int p = (int)(i & 127);
    const long row = i >> 7;
    int s = shift[row / 32] % 128;
    if (s < 0) s += 128;
    out[i] = in[row * 128 + ((p - s + 128) & 127)];
}

// chord (rows,14) [root, 12 chroma bits, bass] -> (rows,36) [root one-hot | rolled chroma | bass one-hot] transposed by
// the segment's shift (converter.py:150-164 expand_chord); rows_per_seg chord rows share one shift entry.
__global__ void expand_chord_kernel(const float* __restrict__ ch, const int* __restrict__ shift, long rows, int rows_per_seg,
                                    float* __restrict__ out) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 36) return;
    const long r = i / 36;
    const int j = (int)(i % 36);
    int s = shift[r / rows_per_seg] % 12;
    if (s < 0) s += 12;
    const float* c = ch + r * 14;
    float v;
    if (j < 12) v = (j == (((int)c[0] + s) % 12)) ? 1.0f : 0.0f;
    else if (j < 24) v = c[1 + ((j - 12 - s + 12) % 12)];
    else v = (j - 24 == (((int)c[13] + s) % 12)) ? 1.0f : 0.0f;
    out[i] = v;
}

// ---- latent-space interpolation on device (model.py:218-242 interp_path): spherical interpolation of the direction,
// log-linear interpolation of the norm, `count` points per (z1, z2) pair.  The reference runs this in float64 numpy on
// the host; the kernel keeps float64 (the data is B x count x D, a few KB) so results agree to fp32 rounding.
// One CTA per pair.
__global__ void slerp_path_kernel(const float* __restrict__ z1, const float* __restrict__ z2, int D, int count,
                                  float* __restrict__ out) {
    __shared__ double red[3][32];
    __shared__ double s_n1, s_n2, s_omega;
    const int b = blockIdx.x;
    const float* a = z1 + (long)b * D;
    const float* c = z2 + (long)b * D;
    double n1 = 0.0, n2 = 0.0, dot = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double x = a[i], y = c[i];
        n1 += x * x; n2 += y * y; dot += x * y;
    }
    for (int o = 16; o > 0; o >>= 1) {
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) { red[0][w] = n1; red[1][w] = n2; red[2][w] = dot; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A = 0, Bq = 0, Cq = 0;
        for (int i = 0; i < nw; ++i) { A += red[0][i]; Bq += red[1][i]; Cq += red[2][i]; }
        A = sqrt(A); Bq = sqrt(Bq);
        double cosv = Cq / (A * Bq);
        cosv = fmin(1.0, fmax(-1.0, cosv));
        s_n1 = A; s_n2 = Bq; s_omega = acos(cosv);
    }
    __syncthreads();
    const double N1 = s_n1, N2 = s_n2, om = s_omega, so = sin(om);
    const double l1 = log(N1), l2 = log(N2);
    for (int k = 0; k < count; ++k) {
        const double t = count > 1 ? (double)k / (double)(count - 1) : 0.0;
        const double w1 = sin((1.0 - t) * om) / so, w2 = sin(t * om) / so;
        const double len = exp(l1 + (l2 - l1) * t);
        for (int i = threadIdx.x; i < D; i += blockDim.x)
            out[((long)b * count + k) * D + i] = (float)((w1 * ((double)a[i] / N1) + w2 * ((double)c[i] / N2)) * len);
    }
}

}  // namespace

PD_API int pd_prmat_to_grid(const float* pr_mat, long n_steps, long long* x, int* overflow, void* stream) {
    if (n_steps <= 0) return 0;
    if (((uintptr_t)pr_mat & 15)) return PD_BAD_ARG;
    prmat_to_grid_kernel<<<pd_blocks(n_steps * 32, 256), 256, 0, (cudaStream_t)stream>>>(pr_mat, n_steps, x, overflow);
    return pd_launch_status();
}

PD_API int pd_grid_to_prmat(const int* tok, long n_steps, float* pr_mat, void* stream) {
    if (n_steps <= 0) return 0;
    grid_to_prmat_kernel<<<pd_blocks(n_steps, 128), 128, 0, (cudaStream_t)stream>>>(tok, n_steps, pr_mat);
    return pd_launch_status();
}

PD_API int pd_grid_prepare(const long long* x, long n_steps, int* tok, int* lengths, int* pitch_tgt,
                           int* dur_tgt, void* stream) {
    if (n_steps <= 0) return 0;
    grid_prepare_kernel<<<pd_blocks(n_steps * NOTE_SLOTS, 256), 256, 0, (cudaStream_t)stream>>>(x, n_steps, tok, lengths,
                                                                                                pitch_tgt, dur_tgt);
    return pd_launch_status();
}

PD_API int pd_note_embed_fwd(const int* tok, long R, const float* WT, const float* bias, float* out, long ldo,
                             void* stream) {
    if (R <= 0) return 0;
    if ((ldo & 3) || ((uintptr_t)out & 15) || ((uintptr_t)WT & 15) || ((uintptr_t)bias & 15)) return PD_BAD_ARG;
    note_embed_fwd_kernel<<<pd_blocks(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(tok, R, WT, bias, out, ldo);
    return pd_launch_status();
}

static int note_embed_bwd_impl(const int* tok, long R, const float* g, long ldg, float* dWT, float* dbias, PdRows live,
                               void* stream) {
    if (R <= 0) return 0;
    static unsigned long long attr_set = 0;
    const int smem = (NOTE_SIZE + 1) * EMB * (int)sizeof(float);
    if (pd_first_use_on_device(attr_set)) {
        cudaFuncSetAttribute(note_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    long ctas = 3 * PD_NUM_SMS;
    long rows_per = (R + ctas - 1) / ctas;
    if (rows_per < 64) rows_per = 64;
    ctas = (R + rows_per - 1) / rows_per;
    note_embed_bwd_kernel<<<(unsigned)ctas, EMB, smem, (cudaStream_t)stream>>>(tok, R, g, ldg, dWT, dbias, rows_per, live);
    return pd_launch_status();
}

PD_API int pd_note_embed_bwd(const int* tok, long R, const float* g, long ldg, float* dWT, float* dbias,
                             void* stream) {
    return note_embed_bwd_impl(tok, R, g, ldg, dWT, dbias, PdRows{nullptr, 0, 0}, stream);
}

// Packed note level: tok / g rows are slot-major (R = n_slots * slot_rows); only rows r % slot_rows < cp[r / slot_rows]
// contribute (common.cuh PdRows)
PD_API int pd_note_embed_bwd_rows(const int* tok, long R, const float* g, long ldg, float* dWT, float* dbias, const int* cp,
                                  int slot_rows, void* stream) {
    if (cp == nullptr || slot_rows <= 0) return PD_BAD_ARG;
    return note_embed_bwd_impl(tok, R, g, ldg, dWT, dbias, PdRows{cp, slot_rows, (int)((R + slot_rows - 1) / slot_rows)}, stream);
}

PD_API int pd_greedy_pick(const float* pitch, long ldp, const float* dur, long ldd, long R, int n, int* tok,
                          long ldtok, int* lens, void* stream) {
    if (R <= 0) return 0;
    greedy_pick_kernel<<<pd_blocks(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(pitch, ldp, dur, ldd, R, n, tok,
                                                                                 ldtok, lens);
    return pd_launch_status();
}

// pd_greedy_pick + pd_note_embed_fwd of the picked tokens in one launch: emb (R,128; row stride lde) = embedding of tok
PD_API int pd_greedy_pick_embed(const float* pitch, long ldp, const float* dur, long ldd, long R, int n, int* tok,
                                long ldtok, int* lens, const float* WT, const float* bias, float* emb, long lde, void* stream) {
    if (R <= 0) return 0;
    if ((lde & 3) || ((uintptr_t)emb & 15) || ((uintptr_t)WT & 15) || ((uintptr_t)bias & 15)) return PD_BAD_ARG;
    greedy_pick_embed_kernel<<<pd_blocks(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(pitch, ldp, dur, ldd, R, n, tok, ldtok,
                                                                                       lens, WT, bias, emb, lde);
    return pd_launch_status();
}

PD_API int pd_dur_token(const float* logit, long ldl, long R, float* tok, void* stream) {
    if (R <= 0) return 0;
    dur_token_kernel<<<pd_blocks(R, 256), 256, 0, (cudaStream_t)stream>>>(logit, ldl, R, tok);
    return pd_launch_status();
}

PD_API int pd_transpose_f32(const float* in, int rows, int cols, float* out, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), blk(32, 8);
    transpose_kernel<<<grid, blk, 0, (cudaStream_t)stream>>>(in, rows, cols, out);
    return pd_launch_status();
}

PD_API int pd_chord_feedback(const float* root, long ldr, const float* chroma, long ldc, const float* bass,
                             long ldb, int B, float* flags24, float* tok, long ldt, void* stream) {
    if (B <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(flags24, 0, 24 * sizeof(float), st);
    chord_union_kernel<<<pd_blocks(B, 128), 128, 0, st>>>(root, ldr, bass, ldb, B, flags24);
    chord_token_kernel<<<pd_blocks((long)B * 36, 128), 128, 0, st>>>(chroma, ldc, flags24, B, tok, ldt);
    return pd_launch_status();
}

PD_API int pd_chord_targets(const float* c, int rows, int* root, int* chroma, int* bass, void* stream) {
    if (rows <= 0) return 0;
    chord_targets_kernel<<<pd_blocks(rows, 128), 128, 0, (cudaStream_t)stream>>>(c, rows, root, chroma, bass);
    return pd_launch_status();
}

PD_API int pd_roll_prmat(const float* pr_in, const int* shift, long B, float* pr_out, void* stream) {
    if (B <= 0) return 0;
    if (pr_in == pr_out) return PD_BAD_ARG;
    roll_prmat_kernel<<<pd_blocks(B * 32 * 128, 256), 256, 0, (cudaStream_t)stream>>>(pr_in, shift, B, pr_out);
    return pd_launch_status();
}

PD_API int pd_expand_chord(const float* chord14, const int* shift, long rows, int rows_per_seg, float* c36, void* stream) {
    if (rows <= 0) return 0;
    if (rows_per_seg <= 0) return PD_BAD_ARG;
    expand_chord_kernel<<<pd_blocks(rows * 36, 256), 256, 0, (cudaStream_t)stream>>>(chord14, shift, rows, rows_per_seg, c36);
    return pd_launch_status();
}

PD_API int pd_slerp_path(const float* z1, const float* z2, int B, int D, int count, float* out, void* stream) {
    if (B <= 0 || count <= 0) return 0;
    if (D <= 0) return PD_BAD_ARG;
    slerp_path_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(z1, z2, D, count, out);
    return pd_launch_status();
}

PD_API int pd_pack_tokens(const int* tok, long R, unsigned char* out, void* stream) {
    if (R <= 0) return 0;
    pack_tokens_kernel<<<pd_blocks(R, 256), 256, 0, (cudaStream_t)stream>>>(tok, R, out);
    return pd_launch_status();
}
