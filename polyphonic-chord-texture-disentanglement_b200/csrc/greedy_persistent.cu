// Persistent greedy PianoTree decode for SMALL batches (<= 16 segments): the whole 32 x 15 x 5 loop nest of
// ptvae.py:430-491 (inference branch) in ONE cooperative launch.
//
// Why: a 16-segment decode (BASELINE configs[4]: a 256-bar arrangement = 128 segments over 8 GPUs) is ~5,400
// DEPENDENT kernel launches on the step-wise path -- 61 ms of launch / drain latency for microseconds of arithmetic.
// Here every SM keeps one CTA resident for the whole decode and the dependency chain advances through grid-wide
// barriers in L2 (release/acquire on one counter) instead of kernel boundaries:
//
//   per time step t (32):   P1  time-GRU cell (1024 units; W_hh / W_tok streamed from L2, one unit per warp)
//                           P2  note-level initial state (512) + summary projection of the note GRU (1536)
//     per note slot n (15): C   note-GRU cell: 4 hidden units per CTA, [W_hh | W_tok] rows RESIDENT IN SHARED MEMORY
//                               for the whole decode (30 KB per CTA, 128 CTAs)
//                           D   pitch head (130) + folded duration-hidden projection (64): rows resident in smem
//                           E   per segment (CTA b): argmax pitch, 5-step duration GRU with greedy bit feedback
//                               (W_hh rows in registers), token, EOS length, note-embedding gather
//                           P4a x-projections of the predicted notes for the summary bi-GRU (all CTAs)
//                           P4b variable-length bi-GRU(128) summary -> next time-step token (16 CTAs, W_hh in smem)
//
// Arithmetic is fp32 FFMA throughout (the fp32-faithful mode greedy token parity needs).  All data
// produced by one phase and consumed by another crosses CTAs through global memory: written with plain stores before
// the barrier's release, read with ld.global.cg (L2) after its acquire.
//
// Replaces, for small batches: aten::gru x 2,912 + Linear + argmax + one-hot per decode (ptvae.py:336-428,:460-486).
#include <cooperative_groups.h>
#include "common.cuh"

namespace {

constexpr int MAXB = 16;                 // segments per launch
constexpr int NT = 256, NWARP = 8;
constexpr int HT = 1024, HN = 512, E = 128, ZIN = 256, HE = 128, HD = 64;
constexpr int NHEAD = 194;               // 130 pitch logits + 64 duration-hidden units
constexpr int NSLOT = 16, T_STEPS = 32, P_EOS = 129, P_RANGE = 130;
constexpr int N_NOTE_CTAS = 128, UNITS_PER_CTA = 4;       // 128 x 4 = 512 note-GRU units
constexpr int N_SUM_CTAS = 16;                             // 2 directions x 8 row groups
constexpr int WE_LD = 132;                                 // padded row stride of the summary W_hh in smem

struct GreedyParams {
    int B;
    const float* h_time0;      // (B,1024)  z2dec_hid(z)
    const float* gi_z;         // (B,3072)  W_ih[:,256:] z_in + b_ih of the time GRU
    const float* wt_tok; long ld_wt;      // (3072,256), row stride ld_wt (= 512)
    const float* wt_hh;        // (3072,1024)
    const float* bt_hh;        // (3072)
    const float* init_tok;     // (256) dec_init_input
    const float* w_t2n; const float* b_t2n;           // (512,1024), (512)
    const float* wn_sum; long ld_wn; const float* bn_ih;   // (1536,1024) row stride ld_wn (= 1152); (1536)
    const float* wn_tok;       // (1536,128) row stride ld_wn
    const float* wn_hh; const float* bn_hh;           // (1536,512), (1536)
    const float* w_heads; const float* b_heads;       // (194,512), (194)
    const float* d_wih; const float* d_bih; const float* d_whh; const float* d_bhh;   // (192,5) (192) (192,64) (192)
    const float* d_sos; const float* d_wout; const float* d_bout;                     // (5) (2,64) (2)
    const float* emb_wt; const float* emb_b;          // (135,128) = note_embedding.weight^T, (128)
    const float* we_ih[2]; const float* we_hh[2]; const float* be_ih[2]; const float* be_hh[2];   // (384,128) ... per direction
    int* tokens;               // (32,15,B,6) int32
    int* lens_out;             // (32,B) or NULL
    float* ws;                 // workspace (see WS_* offsets)
    unsigned* bar;             // [0] barrier counter (zeroed by the entry point), [1] abort flag
};

// workspace layout (floats)
constexpr long WS_HTIME = 0;                                   // [2][MAXB][1024]
constexpr long WS_TOKT = WS_HTIME + 2L * MAXB * HT;            // [MAXB][256]
constexpr long WS_HN = WS_TOKT + (long)MAXB * ZIN;             // [2][MAXB][512]
constexpr long WS_GIS = WS_HN + 2L * MAXB * HN;                // [MAXB][1536]
constexpr long WS_HEADS = WS_GIS + (long)MAXB * 3 * HN;        // [MAXB][196]
constexpr long WS_PRED = WS_HEADS + (long)MAXB * 196;          // [MAXB][16][128]
constexpr long WS_GIE = WS_PRED + (long)MAXB * NSLOT * E;      // [2][MAXB][16][384]
constexpr long WS_LENS = WS_GIE + 2L * MAXB * NSLOT * 3 * HE;  // [MAXB] (int)
constexpr long WS_FLOATS = WS_LENS + MAXB;

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- grid barrier: monotonically increasing arrival counter in L2 --------------------------------------------------
struct GridBar {
    unsigned* ctr;
    unsigned* abort_flag;
    unsigned nblocks, epoch;
    bool dead;
};

__device__ __forceinline__ void grid_sync(GridBar& gb) {
    __syncthreads();
    if (threadIdx.x == 0 && !gb.dead) {
        const unsigned target = (++gb.epoch) * gb.nblocks;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
        unsigned v;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
            if (v < target && clock64() - t0 > 4000000000LL) {      // ~2 s: a peer never arrived -- give up, flag it
                atomicExch(gb.abort_flag, 1u);
                gb.dead = true;
                break;
            }
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// sum over the warp of N per-lane values; lane l ends up with the total of value (l % N).  N in {8, 16}.
template <int N>
__device__ __forceinline__ float reduce_scatter(float (&v)[N], int lane) {
#pragma unroll
    for (int half = N / 2; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    float r = v[0];
#pragma unroll
    for (int off = N; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// acc[b] += w[0..K) . act[b][0..K) for MAXB rows of a shared-memory activation matrix; lanes split K in float4s
template <bool WSMEM>
__device__ __forceinline__ void dot_rows(const float* __restrict__ w, int K, const float* act, int lda, int lane,
                                         float (&acc)[MAXB]) {
    for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = WSMEM ? *reinterpret_cast<const float4*>(w + k) : __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
            const float4 a = *reinterpret_cast<const float4*>(act + b * lda + k);
            acc[b] = fmaf(wv.x, a.x, fmaf(wv.y, a.y, fmaf(wv.z, a.z, fmaf(wv.w, a.w, acc[b]))));
        }
    }
}

// cooperative copy of `rows` x `cols` floats (global, row stride ld, produced by other CTAs) into shared memory
__device__ __forceinline__ void stage(float* dst, int ldd, const float* src, long lds, int rows, int cols) {
    const int c4 = cols >> 2;
    for (int i = threadIdx.x; i < rows * c4; i += NT) {
        const int r = i / c4, c = (i % c4) * 4;
        *reinterpret_cast<float4*>(dst + r * ldd + c) = ldcg4(src + r * lds + c);
    }
}

__global__ void __launch_bounds__(NT, 1) greedy_small_kernel(GreedyParams p) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int B = p.B;
    const bool is_sum = cta >= G - N_SUM_CTAS;                 // summary bi-GRU CTAs (the last 16)
    const bool is_note = cta < N_NOTE_CTAS;                    // hold 4 note-GRU units each
    const int n_work = G - N_SUM_CTAS;                         // CTAs that take part in the distributed mat-vec phases
    const int gw = cta * NWARP + warp, n_gw = n_work * NWARP;  // worker-warp index (valid when !is_sum)
    GridBar gb{p.bar, p.bar + 1, (unsigned)G, 0u, false};

    float* ws = p.ws;
    float* h_time = ws + WS_HTIME;
    float* tok_time = ws + WS_TOKT;
    float* h_n = ws + WS_HN;
    float* gi_s = ws + WS_GIS;
    float* heads = ws + WS_HEADS;
    float* pred = ws + WS_PRED;
    float* gi_e = ws + WS_GIE;
    int* lens = reinterpret_cast<int*>(ws + WS_LENS);

    // ---- shared memory carve-up (per role) ----------------------------------------------------------------------
    // worker CTAs: [note weights 4 x 3 x 640][head rows 2 x 512][activation staging 16 x 1280 (P1) / smaller later]
    // summary CTAs: [W_hh of one direction 384 x WE_LD][h 2 x 128][gh 2 x 384]
    float* wnote = sm;                                          // 12 rows x 640
    float* whead = wnote + UNITS_PER_CTA * 3 * (HN + E);        // 2 rows x 512
    float* act = whead + 2 * HN;                                // up to MAXB x 1280
    float* we_s = sm;                                           // summary role
    float* he_s = we_s + 3 * HE * WE_LD;                        // [2][128]
    float* ghe_s = he_s + 2 * HE;                               // [2][384]
    __shared__ float dur_gi[3][192];
    __shared__ __align__(16) float dur_h[HD];
    __shared__ float dur_gh[192];
    __shared__ __align__(16) float head_row[196];
    __shared__ float dur_lg[2];
    __shared__ int s_tok[6];

    // ---- one-time loads ---------------------------------------------------------------------------------------------
    if (is_sum) {
        const int dir = (cta - (G - N_SUM_CTAS)) / 8;
        for (int i = tid; i < 3 * HE * (HE / 4); i += NT) {
            const int r = i / (HE / 4), c = (i % (HE / 4)) * 4;
            *reinterpret_cast<float4*>(we_s + r * WE_LD + c) = __ldg(reinterpret_cast<const float4*>(p.we_hh[dir] + r * HE + c));
        }
    } else {
        if (is_note) {
            for (int i = tid; i < UNITS_PER_CTA * 3 * ((HN + E) / 4); i += NT) {
                const int row = i / ((HN + E) / 4), c = (i % ((HN + E) / 4)) * 4;
                const int u = cta * UNITS_PER_CTA + row / 3, gate = row % 3;
                const long wr = (long)gate * HN + u;
                const float4 v = c < HN ? __ldg(reinterpret_cast<const float4*>(p.wn_hh + wr * HN + c))
                                        : __ldg(reinterpret_cast<const float4*>(p.wn_tok + wr * p.ld_wn + (c - HN)));
                *reinterpret_cast<float4*>(wnote + row * (HN + E) + c) = v;
            }
        }
        for (int s = 0; s < 2; ++s) {
            const int j = cta + s * n_work;                    // head rows owned by this CTA
            if (j < NHEAD)
                for (int i = tid; i < HN / 4; i += NT)
                    *reinterpret_cast<float4*>(whead + s * HN + i * 4) = __ldg(reinterpret_cast<const float4*>(p.w_heads + (long)j * HN + i * 4));
        }
    }
    // duration GRU: thread j < 192 keeps row j of W_hh (192 x 64) in registers (row CTAs only use it)
    float dw[HD];
    float dbh = 0.f;
    if (cta < B && tid < 192) {
#pragma unroll
        for (int k = 0; k < HD; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.d_whh + tid * HD + k));
            dw[k] = v.x; dw[k + 1] = v.y; dw[k + 2] = v.z; dw[k + 3] = v.w;
        }
        dbh = p.d_bhh[tid];
        float bi = p.d_bih[tid], s0 = bi;
#pragma unroll
        for (int c = 0; c < 5; ++c) s0 = fmaf(p.d_wih[tid * 5 + c], p.d_sos[c], s0);
        dur_gi[0][tid] = s0;                                   // step 0: W_ih sos + b_ih
        dur_gi[1][tid] = p.d_wih[tid * 5 + 0] + bi;            // fed-back bit 0 -> one-hot at index 0
        dur_gi[2][tid] = p.d_wih[tid * 5 + 1] + bi;            // fed-back bit 1 -> one-hot at index 1 (ptvae.py:322-326)
    } else {
#pragma unroll
        for (int k = 0; k < HD; ++k) dw[k] = 0.f;
    }
    // initial state: h_time[0] = z2dec_hid(z), tok_time = dec_init_input (CTA 0 writes, everybody reads after the barrier)
    if (cta == 0) {
        for (int i = tid; i < B * HT; i += NT) h_time[i] = p.h_time0[i];
        for (int i = tid; i < B * ZIN; i += NT) tok_time[i] = p.init_tok[i % ZIN];
    }
    grid_sync(gb);

    int cur_t = 0;                                              // h_time ping-pong index
    for (int t = 0; t < T_STEPS; ++t) {
        // ================= P1: time-GRU cell ==========================================================================
        if (!is_sum) {
            constexpr int LDA = HT + ZIN;
            stage(act, LDA, h_time + (long)cur_t * MAXB * HT, HT, B, HT);
            stage(act + HT, LDA, tok_time, ZIN, B, ZIN);
            if (B < MAXB)
                for (int i = tid; i < (MAXB - B) * LDA; i += NT) act[B * LDA + i] = 0.f;
            __syncthreads();
            for (int u = gw; u < HT; u += n_gw) {
                float ar[MAXB], az[MAXB], anh[MAXB], ani[MAXB];
#pragma unroll
                for (int b = 0; b < MAXB; ++b) ar[b] = az[b] = anh[b] = ani[b] = 0.f;
                dot_rows<false>(p.wt_hh + (long)u * HT, HT, act, LDA, lane, ar);
                dot_rows<false>(p.wt_tok + (long)u * p.ld_wt, ZIN, act + HT, LDA, lane, ar);
                dot_rows<false>(p.wt_hh + (long)(HT + u) * HT, HT, act, LDA, lane, az);
                dot_rows<false>(p.wt_tok + (long)(HT + u) * p.ld_wt, ZIN, act + HT, LDA, lane, az);
                dot_rows<false>(p.wt_hh + (long)(2 * HT + u) * HT, HT, act, LDA, lane, anh);
                dot_rows<false>(p.wt_tok + (long)(2 * HT + u) * p.ld_wt, ZIN, act + HT, LDA, lane, ani);
                const float sr = reduce_scatter<MAXB>(ar, lane), sz = reduce_scatter<MAXB>(az, lane);
                const float snh = reduce_scatter<MAXB>(anh, lane), sni = reduce_scatter<MAXB>(ani, lane);
                if (lane < B) {
                    const float* gz = p.gi_z + (long)lane * 3 * HT;
                    const float r = pd_sigmoid(sr + gz[u] + p.bt_hh[u]);
                    const float z = pd_sigmoid(sz + gz[HT + u] + p.bt_hh[HT + u]);
                    const float n = tanhf(sni + gz[2 * HT + u] + r * (snh + p.bt_hh[2 * HT + u]));
                    h_time[(long)(cur_t ^ 1) * MAXB * HT + lane * HT + u] = (1.0f - z) * n + z * act[lane * LDA + u];
                }
            }
        }
        cur_t ^= 1;
        grid_sync(gb);
        // ================= P2: note-level initial state + summary projection =========================================
        if (!is_sum) {
            stage(act, HT, h_time + (long)cur_t * MAXB * HT, HT, B, HT);
            if (B < MAXB)
                for (int i = tid; i < (MAXB - B) * HT; i += NT) act[B * HT + i] = 0.f;
            __syncthreads();
            for (int j = gw; j < HN + 3 * HN; j += n_gw) {
                float a[MAXB];
#pragma unroll
                for (int b = 0; b < MAXB; ++b) a[b] = 0.f;
                const bool init = j < HN;
                const float* w = init ? p.w_t2n + (long)j * HT : p.wn_sum + (long)(j - HN) * p.ld_wn;
                dot_rows<false>(w, HT, act, HT, lane, a);
                const float s = reduce_scatter<MAXB>(a, lane);
                if (lane < B) {
                    if (init) h_n[lane * HN + j] = s + p.b_t2n[j];                       // ping-pong half 0
                    else gi_s[lane * 3 * HN + (j - HN)] = s + p.bn_ih[j - HN];
                }
            }
            if (cta == n_work - 1) {                            // slot 0 of every step is the SOS note; lengths restart
                for (int i = tid; i < B * E; i += NT) {
                    const int b = i / E, c = i % E;
                    float v = p.emb_b[c] + p.emb_wt[128 * E + c];                       // pitch SOS = 128
#pragma unroll
                    for (int k = 0; k < 5; ++k) v = fmaf(2.0f, p.emb_wt[(P_RANGE + k) * E + c], v);   // duration pad value 2
                    pred[(long)b * NSLOT * E + c] = v;
                }
                if (tid < B) lens[tid] = 0;
            }
        }
        grid_sync(gb);
        int cur_n = 0;
        for (int n = 1; n < NSLOT; ++n) {
            // ============= C: note-GRU cell (weights resident in smem) ================================================
            if (is_note) {
                constexpr int LDA = HN + E;
                stage(act, LDA, h_n + (long)cur_n * MAXB * HN, HN, B, HN);
                stage(act + HN, LDA, pred + (long)(n - 1) * E, (long)NSLOT * E, B, E);
                if (B < MAXB)
                    for (int i = tid; i < (MAXB - B) * LDA; i += NT) act[B * LDA + i] = 0.f;
                __syncthreads();
                const int ul = warp & 3, half = warp >> 2;     // unit within the CTA, row half (8 rows each)
                const int u = cta * UNITS_PER_CTA + ul;
                const float* wr = wnote + (ul * 3 + 0) * LDA;
                const float* wz = wnote + (ul * 3 + 1) * LDA;
                const float* wn = wnote + (ul * 3 + 2) * LDA;
                const float* a0 = act + half * 8 * LDA;
                float ar[8], az[8], anh[8], ani[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) ar[b] = az[b] = anh[b] = ani[b] = 0.f;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int k = lane * 4 + 128 * i;
                    const float4 vr = *reinterpret_cast<const float4*>(wr + k);
                    const float4 vz = *reinterpret_cast<const float4*>(wz + k);
                    const float4 vn = *reinterpret_cast<const float4*>(wn + k);
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const float4 a = *reinterpret_cast<const float4*>(a0 + b * LDA + k);
                        ar[b] = fmaf(vr.x, a.x, fmaf(vr.y, a.y, fmaf(vr.z, a.z, fmaf(vr.w, a.w, ar[b]))));
                        az[b] = fmaf(vz.x, a.x, fmaf(vz.y, a.y, fmaf(vz.z, a.z, fmaf(vz.w, a.w, az[b]))));
                        const float dn = fmaf(vn.x, a.x, fmaf(vn.y, a.y, fmaf(vn.z, a.z, vn.w * a.w)));
                        if (i < 4) anh[b] += dn; else ani[b] += dn;
                    }
                }
                const float sr = reduce_scatter<8>(ar, lane), sz = reduce_scatter<8>(az, lane);
                const float snh = reduce_scatter<8>(anh, lane), sni = reduce_scatter<8>(ani, lane);
                const int b = half * 8 + (lane & 7);
                if (lane < 8 && b < B) {
                    const float* gs = gi_s + (long)b * 3 * HN;
                    const float r = pd_sigmoid(sr + __ldcg(gs + u) + p.bn_hh[u]);
                    const float z = pd_sigmoid(sz + __ldcg(gs + HN + u) + p.bn_hh[HN + u]);
                    const float nn = tanhf(sni + __ldcg(gs + 2 * HN + u) + r * (snh + p.bn_hh[2 * HN + u]));
                    h_n[(long)(cur_n ^ 1) * MAXB * HN + b * HN + u] = (1.0f - z) * nn + z * act[b * LDA + u];
                }
            }
            cur_n ^= 1;
            grid_sync(gb);
            // ============= D: pitch head + duration-hidden projection (rows resident in smem) =========================
            if (!is_sum) {
                const bool has0 = cta < NHEAD, has1 = cta + n_work < NHEAD;
                if (has0) {
                    stage(act, HN, h_n + (long)cur_n * MAXB * HN, HN, B, HN);
                    if (B < MAXB)
                        for (int i = tid; i < (MAXB - B) * HN; i += NT) act[B * HN + i] = 0.f;
                    __syncthreads();
                    if (warp < 2 && (warp == 0 || has1)) {
                        const int j = cta + warp * n_work;
                        float a[MAXB];
#pragma unroll
                        for (int b = 0; b < MAXB; ++b) a[b] = 0.f;
                        dot_rows<true>(whead + warp * HN, HN, act, HN, lane, a);
                        const float s = reduce_scatter<MAXB>(a, lane);
                        if (lane < B) heads[lane * 196 + j] = s + p.b_heads[j];
                    }
                }
            }
            grid_sync(gb);
            // ============= E: per segment -- greedy pitch, duration GRU, token, length, embedding ======================
            if (cta < B) {
                const int b = cta;
                if (tid < 196 / 4) *reinterpret_cast<float4*>(head_row + tid * 4) = ldcg4(heads + b * 196 + tid * 4);
                __syncthreads();
                if (tid < HD) dur_h[tid] = head_row[P_RANGE + tid];
                if (warp == 7) {                                // argmax over the 130 pitch logits (first maximum)
                    float best = -INFINITY;
                    int bi = 0x7fffffff;
                    for (int i = lane; i < P_RANGE; i += 32) {
                        const float v = head_row[i];
                        if (v > best || (v != v && best == best)) { best = v; bi = i; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    if (lane == 0) s_tok[0] = bi;
                }
                __syncthreads();
                int table = 0;                                  // 0: dur_sos_token, 1 / 2: fed-back bit 0 / 1
                for (int k = 0; k < 5; ++k) {
                    if (tid < 192) {
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                        for (int q = 0; q < HD; q += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(dur_h + q);
                            a0 = fmaf(dw[q], v.x, a0); a1 = fmaf(dw[q + 1], v.y, a1);
                            a2 = fmaf(dw[q + 2], v.z, a2); a3 = fmaf(dw[q + 3], v.w, a3);
                        }
                        dur_gh[tid] = dbh + ((a0 + a1) + (a2 + a3));
                    }
                    __syncthreads();
                    if (tid < HD) {
                        const float* gi = dur_gi[table];
                        const float r = pd_sigmoid(gi[tid] + dur_gh[tid]);
                        const float z = pd_sigmoid(gi[HD + tid] + dur_gh[HD + tid]);
                        const float nn = tanhf(gi[2 * HD + tid] + r * dur_gh[2 * HD + tid]);
                        dur_h[tid] = (1.0f - z) * nn + z * dur_h[tid];
                    }
                    __syncthreads();
                    if (warp < 2) {                             // duration head: 2 logits
                        float a = fmaf(dur_h[lane], p.d_wout[warp * HD + lane], dur_h[lane + 32] * p.d_wout[warp * HD + lane + 32]);
                        a = warp_sum(a);
                        if (lane == 0) dur_lg[warp] = a + p.d_bout[warp];
                    }
                    __syncthreads();
                    const int bit = dur_lg[1] > dur_lg[0] ? 1 : 0;
                    if (tid == 0) s_tok[1 + k] = bit;
                    table = 1 + bit;
                    __syncthreads();
                }
                int* tk = p.tokens + (((long)t * (NSLOT - 1) + (n - 1)) * B + b) * 6;
                if (tid < 6) tk[tid] = s_tok[tid];
                if (tid == 0) {
                    int L = __ldcg(lens + b);
                    if (L == 0 && s_tok[0] == P_EOS) L = n;                  // predicted length excludes the EOS itself
                    if (n == NSLOT - 1 && L == 0) L = NSLOT - 1;
                    lens[b] = L;
                    if (n == NSLOT - 1 && p.lens_out) p.lens_out[t * B + b] = L;
                }
                if (tid < E) {                                  // note-embedding gather of the predicted token
                    const int pt = s_tok[0];
                    float v = p.emb_b[tid];
                    if (pt >= 0 && pt < P_RANGE) v += p.emb_wt[pt * E + tid];
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        if (s_tok[1 + k]) v += p.emb_wt[(P_RANGE + k) * E + tid];
                    pred[((long)b * NSLOT + n) * E + tid] = v;
                }
            }
            grid_sync(gb);
        }
        if (t == T_STEPS - 1) break;
        // ================= P4a: x-projections of the 16 predicted notes for both summary directions ==================
        if (!is_sum) {
            // activations: pred (B x 16 notes x 128) as MAXB*16 vectors, processed 16 vectors (one segment) at a time
            for (int b = 0; b < B; ++b) {
                __syncthreads();
                stage(act, E, pred + (long)b * NSLOT * E, E, NSLOT, E);
                __syncthreads();
                for (int j = gw; j < 2 * 3 * HE; j += n_gw) {
                    const int dir = j / (3 * HE), row = j % (3 * HE);
                    float a[MAXB];
#pragma unroll
                    for (int s = 0; s < MAXB; ++s) a[s] = 0.f;
                    dot_rows<false>(p.we_ih[dir] + (long)row * E, E, act, E, lane, a);
                    const float s = reduce_scatter<MAXB>(a, lane);
                    if (lane < NSLOT) gi_e[(((long)dir * MAXB + b) * NSLOT + lane) * 3 * HE + row] = s + p.be_ih[dir][row];
                }
            }
        }
        grid_sync(gb);
        // ================= P4b: variable-length bi-GRU(128) summary -> next time-step token ==========================
        if (is_sum) {
            const int s_id = cta - (G - N_SUM_CTAS), dir = s_id / 8, rg = s_id % 8;
            const int b0 = rg * 2;                              // rows b0, b0 + 1
            for (int i = tid; i < 2 * HE; i += NT) he_s[i] = 0.f;
            __syncthreads();
            const int len0 = b0 < B ? __ldcg(lens + b0) : 0, len1 = b0 + 1 < B ? __ldcg(lens + b0 + 1) : 0;
            const int lmax = max(len0, len1);
            for (int s = 0; s < lmax; ++s) {
                const int k = dir ? (lmax - 1 - s) : s;         // forward: 0.., reverse: from the longest sequence's end
                // gh = b_hh + W_hh h for both rows: thread j < 192 owns gate rows j and j + 192
                if (tid < 192) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int j = tid + rr * 192;
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
                        for (int q = 0; q < HE; q += 4) {
                            const float4 w = *reinterpret_cast<const float4*>(we_s + j * WE_LD + q);
                            const float4 x0 = *reinterpret_cast<const float4*>(he_s + q);
                            const float4 x1 = *reinterpret_cast<const float4*>(he_s + HE + q);
                            a0 = fmaf(w.x, x0.x, fmaf(w.y, x0.y, fmaf(w.z, x0.z, fmaf(w.w, x0.w, a0))));
                            a1 = fmaf(w.x, x1.x, fmaf(w.y, x1.y, fmaf(w.z, x1.z, fmaf(w.w, x1.w, a1))));
                        }
                        const float bb = p.be_hh[dir][j];
                        ghe_s[j] = a0 + bb;
                        ghe_s[3 * HE + j] = a1 + bb;
                    }
                }
                __syncthreads();
                {
                    const int r = tid >> 7, u = tid & 127;      // 256 threads = 2 rows x 128 units
                    const int b = b0 + r, len = r ? len1 : len0;
                    if (b < B && k < len) {                     // steps at or past a sequence's length carry its state
                        const float* gi = gi_e + (((long)dir * MAXB + b) * NSLOT + k) * 3 * HE;
                        const float* gh = ghe_s + r * 3 * HE;
                        const float rg_ = pd_sigmoid(__ldcg(gi + u) + gh[u]);
                        const float zg = pd_sigmoid(__ldcg(gi + HE + u) + gh[HE + u]);
                        const float ng = tanhf(__ldcg(gi + 2 * HE + u) + rg_ * gh[2 * HE + u]);
                        he_s[r * HE + u] = (1.0f - zg) * ng + zg * he_s[r * HE + u];
                    }
                }
                __syncthreads();
            }
            {
                const int r = tid >> 7, u = tid & 127;
                if (b0 + r < B) tok_time[(long)(b0 + r) * ZIN + dir * HE + u] = he_s[r * HE + u];
            }
        }
        grid_sync(gb);
    }
}

}  // namespace

static_assert(WS_FLOATS + 16 <= 310368, "PD_GREEDY_SMALL_WS_FLOATS in include/polydis_b200.h is too small");

// Whole greedy PianoTree decode (ptvae.py:430-491, inference=True) of B <= 16 segments in one cooperative launch.
// Weight pointers are the state-dict tensors (row-major, contiguous unless a stride is given); w_heads / b_heads are
// [pitch_out_linear | folded dur_hid_linear] (194 x 512); emb_wt = note_embedding.weight^T (135 x 128).
// tokens (32,15,B,6) int32; lens_out (32,B) int32 or NULL; ws: PD_GREEDY_SMALL_WS_FLOATS (310368) floats;
// bar: 2 x uint32 scratch.  Returns PD_BAD_ARG for B > 16 or a device that cannot co-schedule 144+ CTAs.
PD_API int pd_greedy_decode_small(int B, const float* h_time0, const float* gi_z, const float* wt_tok, long ld_wt,
                                  const float* wt_hh, const float* bt_hh, const float* init_tok, const float* w_t2n,
                                  const float* b_t2n, const float* wn_sum, long ld_wn, const float* bn_ih, const float* wn_tok,
                                  const float* wn_hh, const float* bn_hh, const float* w_heads, const float* b_heads,
                                  const float* d_wih, const float* d_bih, const float* d_whh, const float* d_bhh,
                                  const float* d_sos, const float* d_wout, const float* d_bout, const float* emb_wt,
                                  const float* emb_b, const float* we_ih_f, const float* we_hh_f, const float* be_ih_f,
                                  const float* be_hh_f, const float* we_ih_b, const float* we_hh_b, const float* be_ih_b,
                                  const float* be_hh_b, int* tokens, int* lens_out, float* ws, unsigned* bar, void* stream) {
    if (B <= 0) return 0;
    if (B > MAXB || (ld_wt & 3) || (ld_wn & 3)) return PD_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    constexpr int smem_work = (UNITS_PER_CTA * 3 * (HN + E) + 2 * HN + MAXB * (HT + ZIN)) * 4;
    constexpr int smem_sum = (3 * HE * WE_LD + 2 * HE + 2 * 3 * HE) * 4;
    constexpr int smem = smem_work > smem_sum ? smem_work : smem_sum;
    cudaError_t e = cudaFuncSetAttribute(greedy_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_small_kernel, NT, smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = sms;                                       // one CTA per SM
    if (!coop || per_sm < 1 || grid < N_NOTE_CTAS + N_SUM_CTAS) return PD_BAD_ARG;
    GreedyParams p;
    p.B = B; p.h_time0 = h_time0; p.gi_z = gi_z; p.wt_tok = wt_tok; p.ld_wt = ld_wt; p.wt_hh = wt_hh; p.bt_hh = bt_hh;
    p.init_tok = init_tok; p.w_t2n = w_t2n; p.b_t2n = b_t2n; p.wn_sum = wn_sum; p.ld_wn = ld_wn; p.bn_ih = bn_ih;
    p.wn_tok = wn_tok; p.wn_hh = wn_hh; p.bn_hh = bn_hh; p.w_heads = w_heads; p.b_heads = b_heads;
    p.d_wih = d_wih; p.d_bih = d_bih; p.d_whh = d_whh; p.d_bhh = d_bhh; p.d_sos = d_sos; p.d_wout = d_wout; p.d_bout = d_bout;
    p.emb_wt = emb_wt; p.emb_b = emb_b;
    p.we_ih[0] = we_ih_f; p.we_hh[0] = we_hh_f; p.be_ih[0] = be_ih_f; p.be_hh[0] = be_hh_f;
    p.we_ih[1] = we_ih_b; p.we_hh[1] = we_hh_b; p.be_ih[1] = be_ih_b; p.be_hh[1] = be_hh_b;
    p.tokens = tokens; p.lens_out = lens_out; p.ws = ws; p.bar = bar;
    cudaMemsetAsync(bar, 0, sizeof(unsigned), st);     // bar[1] (abort flag) is sticky: the caller zero-initialises it once
    void* args[] = {&p};
    e = cudaLaunchCooperativeKernel((const void*)greedy_small_kernel, dim3(grid), dim3(NT), args, smem, st);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return pd_launch_status();
}
