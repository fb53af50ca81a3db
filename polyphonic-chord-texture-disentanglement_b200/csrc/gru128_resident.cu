// Weight-resident variable-length GRU (hidden 128) -- the note-summary bi-GRU dec_notes_emb_gru of the PianoTree
// decoder (ptvae.py:446-453 ground-truth notes, :480-486 predicted notes; pack_padded_sequence semantics).
//
// The generic path runs it as 16 masked steps x 2 directions of (GEMM + gate kernel) over all 32*B sequences,
// although a sequence holds 4.5 notes on average: 2.4 ms of the 20.2 ms training step (tools/ablate_step.py).
// Here one CTA owns a tile of sequences for the WHOLE recurrence: W_hh (384x128 fp32, 198 KB) stays resident in
// shared memory, each warp owns 16 hidden units for all three gates and multiplies the tile's state by its
// W_hh slice with mma.sync m16n8k8 TF32 (accumulators in registers, so r, z, n of a unit meet in one thread and
// the gate math needs no staging), the new state goes back to shared memory for the next step, and the loop
// stops at the tile's longest sequence.  PASSES = 3 runs every product as hi*hi + hi*lo + lo*hi (error
// compensated TF32, fp32-class accuracy) for the token-parity inference mode.
// Outputs keep the masked-step semantics of pd_gru_gates_fwd: rows past their length carry their state.
// Gate functions are the MUFU forms of common.cuh (abs error ~1e-6): the kernels are instruction bound.
#include "common.cuh"

namespace {

constexpr int H = 128, G3 = 384, WS = 132;     // WS: padded row stride of W_hh in shared memory
constexpr int NTHR = 256;                      // 8 warps x 16 hidden units

__device__ __forceinline__ uint32_t tf32_of(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A * B with PASSES-fold error compensation; av are the fp32 A-fragment values, bv the B values (for
// PASSES == 1 the weights in shared memory are already rounded to TF32, so their bits go straight to the MMA)
template <int PASSES>
__device__ __forceinline__ void mma_comp(float (&c)[4], const float (&av)[4], float bv0, float bv1) {
    uint32_t ah[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ah[i] = tf32_of(av[i]);
    if (PASSES == 3) {
        const uint32_t b0 = tf32_of(bv0), b1 = tf32_of(bv1);
        uint32_t al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) al[i] = tf32_of(av[i] - __uint_as_float(ah[i]));
        mma8(c, ah, tf32_of(bv0 - __uint_as_float(b0)), tf32_of(bv1 - __uint_as_float(b1)));   // hi * lo
        mma8(c, al, b0, b1);                                                                    // lo * hi
        mma8(c, ah, b0, b1);
    } else {
        mma8(c, ah, __float_as_uint(bv0), __float_as_uint(bv1));
    }
}

struct FwdArgs {
    const float* gi; long ldr, ldt;      // (R, T, 384): gi[r*ldr + t*ldt + j]  (x-projection incl. b_ih)
    const int* lengths;                   // (R)
    const float* w_hh; const float* b_hh; // (384,128), (384)
    float* h_all; long hr, ht;            // (R, T, 128) states (carried past the end)
    float* rzn; long zr, zt;              // optional saves for the backward pass
    float* hn; long nr, nt;
    long R; int T; int reverse;
    const int* perm;                      // optional (R): tile i processes rows perm[16 i .. 16 i + 15] -- rows visited in
                                          // sorted-length order (a tile runs to ITS longest sequence) without moving data
};

template <int RT, int PASSES>
__global__ void __launch_bounds__(NTHR, 1) gru128_fwd_kernel(FwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float (*w_s)[WS] = reinterpret_cast<float (*)[WS]>(smem);                       // [384][132]
    float (*h_s)[WS] = reinterpret_cast<float (*)[WS]>(smem + G3 * WS);             // [2][16][132] double buffer
    __shared__ int len_s[16];
    __shared__ long row_s[16];                  // global row of every tile row
    __shared__ int tmax_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    for (int i = tid; i < G3 * (H / 4); i += NTHR) {
        const int r = i / (H / 4), c4 = (i % (H / 4)) * 4;
        float4 v = *reinterpret_cast<const float4*>(a.w_hh + r * H + c4);
        if (PASSES == 1) {
            v.x = __uint_as_float(tf32_of(v.x)); v.y = __uint_as_float(tf32_of(v.y));
            v.z = __uint_as_float(tf32_of(v.z)); v.w = __uint_as_float(tf32_of(v.w));
        }
        *reinterpret_cast<float4*>(&w_s[r][c4]) = v;
    }
    const int u0 = warp * 16;                     // this warp's hidden units [u0, u0+16)
    float bias[3][2][2];                          // [gate][n-tile][col]
#pragma unroll
    for (int gt = 0; gt < 3; ++gt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            bias[gt][nt][0] = a.b_hh[gt * H + u0 + nt * 8 + 2 * tig];
            bias[gt][nt][1] = a.b_hh[gt * H + u0 + nt * 8 + 2 * tig + 1];
        }
    const long n_tiles = (a.R + RT - 1) / RT;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long r0 = tile * RT;
        __syncthreads();
        if (tid < 16) {
            const bool ok = tid < RT && r0 + tid < a.R;
            const long rr = ok ? (a.perm ? (long)a.perm[r0 + tid] : r0 + tid) : 0;
            row_s[tid] = rr;
            len_s[tid] = ok ? min(a.lengths[rr], a.T) : 0;
        }
        for (int i = tid; i < 2 * 16 * WS; i += NTHR) (&h_s[0][0])[i] = 0.0f;
        __syncthreads();
        if (tid == 0) {
            int m = 0;
            for (int i = 0; i < 16; ++i) m = max(m, len_s[i]);
            tmax_s = m;
        }
        __syncthreads();
        const int tmax = tmax_s;
        int cur = 0;
        // rows of the m16 tile: g and g+8 (rows >= RT are padding)
        const int ra = g, rb = g + 8;
        const int la = len_s[ra], lb = len_s[rb];
        const long rga = row_s[ra], rgb = row_s[rb];
        // x-projections of the NEXT step are fetched while this step's matvec runs (they do not depend on h)
        float2 xi[2][2][3], xn[2][2][3];
        auto fetch = [&](float2 (&xi)[2][2][3], int t) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = half ? rb : ra, len = half ? lb : la;
                    if (t < len) {                              // len == 0 for padding rows
                        const float* gi = a.gi + (half ? rgb : rga) * a.ldr + (long)t * a.ldt + u0 + nt * 8 + 2 * tig;
#pragma unroll
                        for (int gt = 0; gt < 3; ++gt) xi[nt][half][gt] = __ldg(reinterpret_cast<const float2*>(gi + gt * H));
                    }
                }
        };
        fetch(xi, a.reverse ? a.T - 1 : 0);
        for (int s = 0; s < a.T; ++s) {
            const int t = a.reverse ? a.T - 1 - s : s;
            // nothing in this tile is active at step t: forward -> t >= tmax for the rest of the loop
            const bool tile_active = t < tmax;
            float (*hc)[WS] = h_s + cur * 16;
            float (*hx)[WS] = h_s + (cur ^ 1) * 16;
            if (s + 1 < a.T) fetch(xn, a.reverse ? t - 1 : t + 1);
            if (tile_active) {
                float acc[3][2][4];
#pragma unroll
                for (int gt = 0; gt < 3; ++gt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        acc[gt][nt][0] = acc[gt][nt][2] = bias[gt][nt][0];
                        acc[gt][nt][1] = acc[gt][nt][3] = bias[gt][nt][1];
                    }
#pragma unroll 4
                for (int kt = 0; kt < H / 8; ++kt) {
                    float av[4] = {hc[ra][8 * kt + tig], hc[rb][8 * kt + tig], hc[ra][8 * kt + tig + 4], hc[rb][8 * kt + tig + 4]};
#pragma unroll
                    for (int gt = 0; gt < 3; ++gt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int n = gt * H + u0 + nt * 8 + g;
                            mma_comp<PASSES>(acc[gt][nt], av, w_s[n][8 * kt + tig], w_s[n][8 * kt + tig + 4]);
                        }
                }
                // gate math in registers: thread holds (row ra | rb) x (unit pair) of r, z, n
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int row = half ? rb : ra, len = half ? lb : la;
                        const int u = u0 + nt * 8 + 2 * tig;
                        float2 hp = *reinterpret_cast<const float2*>(&hc[row][u]);
                        float2 ho = hp;
                        const long rr = half ? rgb : rga;
                        if (row < RT && r0 + row < a.R) {
                            if (t < len) {
                                const float2 ir = xi[nt][half][0], iz = xi[nt][half][1], in = xi[nt][half][2];
                                const float g0 = acc[2][nt][half * 2], g1 = acc[2][nt][half * 2 + 1];
                                const float r_0 = pd_sigmoid_fast(ir.x + acc[0][nt][half * 2]), r_1 = pd_sigmoid_fast(ir.y + acc[0][nt][half * 2 + 1]);
                                const float z_0 = pd_sigmoid_fast(iz.x + acc[1][nt][half * 2]), z_1 = pd_sigmoid_fast(iz.y + acc[1][nt][half * 2 + 1]);
                                const float n_0 = pd_tanh_fast(in.x + r_0 * g0), n_1 = pd_tanh_fast(in.y + r_1 * g1);
                                ho.x = (1.0f - z_0) * n_0 + z_0 * hp.x;
                                ho.y = (1.0f - z_1) * n_1 + z_1 * hp.y;
                                if (a.rzn) {
                                    float* sv = a.rzn + rr * a.zr + (long)t * a.zt + u;
                                    *reinterpret_cast<float2*>(sv) = make_float2(r_0, r_1);
                                    *reinterpret_cast<float2*>(sv + H) = make_float2(z_0, z_1);
                                    *reinterpret_cast<float2*>(sv + 2 * H) = make_float2(n_0, n_1);
                                }
                                if (a.hn) *reinterpret_cast<float2*>(a.hn + rr * a.nr + (long)t * a.nt + u) = make_float2(g0, g1);
                            }
                            *reinterpret_cast<float2*>(a.h_all + rr * a.hr + (long)t * a.ht + u) = ho;   // carried if masked
                        }
                        *reinterpret_cast<float2*>(&hx[row][u]) = ho;
                    }
                __syncthreads();
                cur ^= 1;
            } else {
                // whole tile past its end: just carry the states to h_all[:, t]
                for (int i = tid; i < RT * (H / 4); i += NTHR) {
                    const int row = i / (H / 4), c4 = (i % (H / 4)) * 4;
                    if (r0 + row < a.R)
                        *reinterpret_cast<float4*>(a.h_all + row_s[row] * a.hr + (long)t * a.ht + c4) =
                            *reinterpret_cast<const float4*>(&hc[row][c4]);
                }
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) (&xi[0][0][0])[i] = (&xn[0][0][0])[i];
        }
    }
}

struct BwdArgs {
    const float* dout; long dr, dt;       // (R, T, 128) gradient wrt h_all
    const float* h_all; long hr, ht;
    const float* rzn; long zr, zt;
    const float* hn; long nr, nt;
    const int* lengths;
    const float* w_hh;
    float* dgi; long gr, gt;              // (R, T, 384) [dr | dz | dn]      (zeros where masked)
    float* dgh; long qr, qt;              // (R, T, 384) [dr | dz | dn*r]
    long R; int T; int reverse;
    const int* cp;                        // packed note level (rows sorted by length): masked dgi entries of step t are
                                          // only zero-filled for rows < cp[t]; the consumers skip the rest.  nullptr: all
    int dout_step;                        // >= 0: dout is (R, 128) and is the gradient of step dout_step's output only (the
                                          // final state of a summariser); -1: dout is (R, T, 128)
};

// Backward of the recurrence for a tile of 16 sequences, walking the steps in reverse processing order.  The
// state gradient dh lives in REGISTERS in the mma accumulator layout (thread = rows g, g+8 x 4 units of its warp's
// 16), so the elementwise gate backward and dh_prev = dh * z + dgh . W_hh meet in one thread; only the step's
// gate gradients (the A operand, K = 384) are staged through shared memory.  W_hh sits in shared memory
// unpadded, TF32-rounded, with an XOR swizzle that makes the k-major B-fragment reads conflict-free.  The saves
// of the next step are prefetched while the current step's matvec runs.
constexpr int GS = 388;                                                             // row stride of the staged dgh
__device__ __forceinline__ int wsw(int j, int u) { return j * H + (u ^ ((j & 3) << 3)); }

__global__ void __launch_bounds__(NTHR, 1) gru128_bwd_kernel(BwdArgs a) {
    constexpr int RT = 16;
    extern __shared__ __align__(16) float smem[];
    float* w_s = smem;                                                              // [384][128] swizzled
    float (*g_s)[GS] = reinterpret_cast<float (*)[GS]>(smem + G3 * H);              // [16][388] dgh of the step
    __shared__ int len_s[RT];
    __shared__ int tmax_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    for (int i = tid; i < G3 * (H / 4); i += NTHR) {
        const int j = i / (H / 4), c4 = (i % (H / 4)) * 4;
        float4 v = *reinterpret_cast<const float4*>(a.w_hh + j * H + c4);
        v.x = __uint_as_float(tf32_of(v.x)); v.y = __uint_as_float(tf32_of(v.y));
        v.z = __uint_as_float(tf32_of(v.z)); v.w = __uint_as_float(tf32_of(v.w));
        *reinterpret_cast<float4*>(&w_s[wsw(j, c4)]) = v;
    }
    struct Saves { float2 d, r, z, n, hn, hp; };
    const int u0 = warp * 16;
    const long n_tiles = (a.R + RT - 1) / RT;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long r0 = tile * RT;
        __syncthreads();
        if (tid < RT) len_s[tid] = (r0 + tid < a.R) ? min(a.lengths[r0 + tid], a.T) : 0;
        __syncthreads();
        if (tid == 0) {
            int m = 0;
            for (int i = 0; i < RT; ++i) m = max(m, len_s[i]);
            tmax_s = m;
        }
        __syncthreads();
        const int tmax = tmax_s;
        const int rows[2] = {g, g + 8};
        const int lens[2] = {len_s[g], len_s[g + 8]};
        const bool live[2] = {r0 + g < a.R, r0 + g + 8 < a.R};
        float2 dh[2][2];                                                           // [n-tile][row half]
#pragma unroll
        for (int i = 0; i < 4; ++i) (&dh[0][0])[i] = make_float2(0.f, 0.f);
        Saves sv[2][2], sn[2][2];
        auto fetch = [&](Saves (&q)[2][2], int t) {
            const int tp = a.reverse ? t + 1 : t - 1;                               // step whose output was h_prev
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (!live[half]) continue;
                    const long rr = r0 + rows[half];
                    const int u = u0 + nt * 8 + 2 * tig;
                    Saves& x = q[nt][half];
                    if (a.dout_step < 0) x.d = __ldg(reinterpret_cast<const float2*>(a.dout + rr * a.dr + (long)t * a.dt + u));
                    else x.d = (t == a.dout_step) ? __ldg(reinterpret_cast<const float2*>(a.dout + rr * a.dr + u)) : make_float2(0.f, 0.f);
                    if (t < lens[half]) {
                        const float* p = a.rzn + rr * a.zr + (long)t * a.zt + u;
                        x.r = __ldg(reinterpret_cast<const float2*>(p));
                        x.z = __ldg(reinterpret_cast<const float2*>(p + H));
                        x.n = __ldg(reinterpret_cast<const float2*>(p + 2 * H));
                        x.hn = __ldg(reinterpret_cast<const float2*>(a.hn + rr * a.nr + (long)t * a.nt + u));
                        x.hp = (tp >= 0 && tp < a.T)
                                   ? __ldg(reinterpret_cast<const float2*>(a.h_all + rr * a.hr + (long)tp * a.ht + u))
                                   : make_float2(0.f, 0.f);
                    }
                }
        };
        fetch(sv, a.reverse ? 0 : a.T - 1);
        for (int s = a.T - 1; s >= 0; --s) {
            const int t = a.reverse ? a.T - 1 - s : s;            // same step order as forward, walked backwards
            const bool tile_active = t < tmax;                    // else every row is masked: dgh == 0
            if (s > 0) fetch(sn, a.reverse ? t + 1 : t - 1);
            // 1) gate gradients of this thread's (row, unit pair)s
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = rows[half], u = u0 + nt * 8 + 2 * tig;
                    float2 dr = make_float2(0.f, 0.f), dz = dr, dn = dr, dnr = dr;
                    if (live[half]) {
                        const Saves& x = sv[nt][half];
                        float2 d = make_float2(dh[nt][half].x + x.d.x, dh[nt][half].y + x.d.y);
                        if (t < lens[half]) {
                            dn = make_float2(d.x * (1.0f - x.z.x) * (1.0f - x.n.x * x.n.x), d.y * (1.0f - x.z.y) * (1.0f - x.n.y * x.n.y));
                            dz = make_float2(d.x * (x.hp.x - x.n.x) * x.z.x * (1.0f - x.z.x), d.y * (x.hp.y - x.n.y) * x.z.y * (1.0f - x.z.y));
                            dnr = make_float2(dn.x * x.r.x, dn.y * x.r.y);
                            dr = make_float2(dn.x * x.hn.x * x.r.x * (1.0f - x.r.x), dn.y * x.hn.y * x.r.y * (1.0f - x.r.y));
                            d = make_float2(d.x * x.z.x, d.y * x.z.y);
                        }
                        dh[nt][half] = d;                          // masked: pass-through; active: direct z path
                        const long rr = r0 + row;
                        if (t < lens[half] || a.cp == nullptr || rr < a.cp[t]) {
                            float* gi = a.dgi + rr * a.gr + (long)t * a.gt + u;
                            *reinterpret_cast<float2*>(gi) = dr;
                            *reinterpret_cast<float2*>(gi + H) = dz;
                            *reinterpret_cast<float2*>(gi + 2 * H) = dn;
                        }
                        float* gq = a.dgh + rr * a.qr + (long)t * a.qt + u;
                        *reinterpret_cast<float2*>(gq) = dr;
                        *reinterpret_cast<float2*>(gq + H) = dz;
                        *reinterpret_cast<float2*>(gq + 2 * H) = dnr;
                    }
                    if (tile_active) {
                        *reinterpret_cast<float2*>(&g_s[row][u]) = make_float2(__uint_as_float(tf32_of(dr.x)), __uint_as_float(tf32_of(dr.y)));
                        *reinterpret_cast<float2*>(&g_s[row][H + u]) = make_float2(__uint_as_float(tf32_of(dz.x)), __uint_as_float(tf32_of(dz.y)));
                        *reinterpret_cast<float2*>(&g_s[row][2 * H + u]) = make_float2(__uint_as_float(tf32_of(dnr.x)), __uint_as_float(tf32_of(dnr.y)));
                    }
                }
            // 2) dh += dgh . W_hh
            if (tile_active) {
                __syncthreads();
                float acc[2][2][4];                               // [k half][n-tile]: two independent chains per tile
#pragma unroll
                for (int i = 0; i < 16; ++i) (&acc[0][0][0])[i] = 0.f;
#pragma unroll 4
                for (int kt = 0; kt < G3 / 16; ++kt) {
#pragma unroll
                    for (int kh = 0; kh < 2; ++kh) {
                        const int k0 = 8 * (kt + kh * (G3 / 16));
                        const uint32_t av[4] = {__float_as_uint(g_s[g][k0 + tig]), __float_as_uint(g_s[g + 8][k0 + tig]),
                                                __float_as_uint(g_s[g][k0 + tig + 4]), __float_as_uint(g_s[g + 8][k0 + tig + 4])};
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int n = u0 + nt * 8 + g;
                            mma8(acc[kh][nt], av, __float_as_uint(w_s[wsw(k0 + tig, n)]), __float_as_uint(w_s[wsw(k0 + tig + 4, n)]));
                        }
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    dh[nt][0].x += acc[0][nt][0] + acc[1][nt][0];
                    dh[nt][0].y += acc[0][nt][1] + acc[1][nt][1];
                    dh[nt][1].x += acc[0][nt][2] + acc[1][nt][2];
                    dh[nt][1].y += acc[0][nt][3] + acc[1][nt][3];
                }
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) (&sv[0][0])[i] = (&sn[0][0])[i];
        }
    }
}

}  // namespace

// h_all (R,T,128) <- variable-length GRU over gi (R,T,384) with per-row lengths, h0 = 0; rows past their length
// carry their state (final state of row r is h_all[r, T-1], or h_all[r, 0] when reverse != 0).  rzn / hn optional
// saves for pd_gru128_bwd.  passes: 1 = TF32, 3 = error-compensated TF32 (fp32-class).  Strides in floats.
static int gru128_fwd_impl(const float* gi, long ldr, long ldt, const int* lengths, const float* w_hh, const float* b_hh,
                           float* h_all, long hr, long ht, float* rzn, long zr, long zt, float* hn, long nr, long nt, long R,
                           int T, int reverse, int passes, const int* perm, void* stream);

PD_API int pd_gru128_fwd(const float* gi, long ldr, long ldt, const int* lengths, const float* w_hh, const float* b_hh,
                         float* h_all, long hr, long ht, float* rzn, long zr, long zt, float* hn, long nr, long nt, long R,
                         int T, int reverse, int passes, void* stream) {
    return gru128_fwd_impl(gi, ldr, ldt, lengths, w_hh, b_hh, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, R, T, reverse, passes,
                           nullptr, stream);
}

// The same recurrence with the rows VISITED in the order perm (R) gives (e.g. pd_pack_order's: longest first): a tile of 16
// rows runs to its longest sequence, so with unsorted lengths most of a tile's steps serve one or two rows.  Data stays where
// it is: every array is still indexed by the original row.
PD_API int pd_gru128_fwd_perm(const float* gi, long ldr, long ldt, const int* lengths, const float* w_hh, const float* b_hh,
                              float* h_all, long hr, long ht, float* rzn, long zr, long zt, float* hn, long nr, long nt, long R,
                              int T, int reverse, int passes, const int* perm, void* stream) {
    if (perm == nullptr) return PD_BAD_ARG;
    return gru128_fwd_impl(gi, ldr, ldt, lengths, w_hh, b_hh, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, R, T, reverse, passes, perm,
                           stream);
}

static int gru128_fwd_impl(const float* gi, long ldr, long ldt, const int* lengths, const float* w_hh, const float* b_hh,
                           float* h_all, long hr, long ht, float* rzn, long zr, long zt, float* hn, long nr, long nt, long R,
                           int T, int reverse, int passes, const int* perm, void* stream) {
    if (R <= 0 || T <= 0) return 0;
    if (((uintptr_t)w_hh & 15) || ((uintptr_t)gi & 7) || (ldr & 1) || (ldt & 1) || ((uintptr_t)h_all & 15) || (hr & 3) ||
        (ht & 3) || (rzn && (((uintptr_t)rzn & 7) || (zr & 1) || (zt & 1))) || (hn && (((uintptr_t)hn & 7) || (nr & 1) || (nt & 1))))
        return PD_BAD_ARG;
    FwdArgs a{gi, ldr, ldt, lengths, w_hh, b_hh, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, R, T, reverse, perm};
    constexpr int smem = (G3 * WS + 2 * 16 * WS) * (int)sizeof(float);
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru128_fwd_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gru128_fwd_kernel<16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    long tiles = (R + 15) / 16;
    unsigned grid = (unsigned)(tiles < PD_NUM_SMS ? tiles : PD_NUM_SMS);
    if (passes == 3) gru128_fwd_kernel<16, 3><<<grid, NTHR, smem, (cudaStream_t)stream>>>(a);
    else gru128_fwd_kernel<16, 1><<<grid, NTHR, smem, (cudaStream_t)stream>>>(a);
    return pd_launch_status();
}

// dgi, dgh (R,T,384) <- dout (R,T,128) and the forward saves; BPTT with dh resident per tile (TF32 matvecs).
static int gru128_bwd_impl(const float* dout, long dr, long dt, const float* h_all, long hr, long ht, const float* rzn,
                           long zr, long zt, const float* hn, long nr, long nt, const int* lengths, const float* w_hh,
                           float* dgi, long gr, long gt, float* dgh, long qr, long qt, long R, int T, int reverse, const int* cp,
                           int dout_step, void* stream) {
    if (R <= 0 || T <= 0) return 0;
    if (dout_step >= T) return PD_BAD_ARG;
    const uintptr_t ptrs = (uintptr_t)dout | (uintptr_t)h_all | (uintptr_t)rzn | (uintptr_t)hn | (uintptr_t)dgi | (uintptr_t)dgh;
    if (((uintptr_t)w_hh & 15) || (ptrs & 7) || ((dr | dt | hr | ht | zr | zt | nr | nt | gr | gt | qr | qt) & 1)) return PD_BAD_ARG;
    BwdArgs a{dout, dr, dt, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, lengths, w_hh, dgi, gr, gt, dgh, qr, qt, R, T, reverse, cp, dout_step};
    constexpr int smem = (G3 * H + 16 * GS) * (int)sizeof(float);
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru128_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    long tiles = (R + 15) / 16;
    unsigned grid = (unsigned)(tiles < PD_NUM_SMS ? tiles : PD_NUM_SMS);
    gru128_bwd_kernel<<<grid, NTHR, smem, (cudaStream_t)stream>>>(a);
    return pd_launch_status();
}

PD_API int pd_gru128_bwd(const float* dout, long dr, long dt, const float* h_all, long hr, long ht, const float* rzn,
                         long zr, long zt, const float* hn, long nr, long nt, const int* lengths, const float* w_hh,
                         float* dgi, long gr, long gt, float* dgh, long qr, long qt, long R, int T, int reverse, void* stream) {
    return gru128_bwd_impl(dout, dr, dt, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, lengths, w_hh, dgi, gr, gt, dgh, qr, qt, R, T,
                           reverse, nullptr, -1, stream);
}

// Variant for the note summarisers.  cp (nullable; packed note level): rows sorted by length, dgi a slot-major gradient slab
// whose consumers skip dead rows -- the masked (row, step) entries of dgi are zero-filled only for rows < cp[t] (T ints,
// device), the others are left unwritten; dgh is written in full.  dout_step >= 0: only the state of that step was used
// (the summary), dout is its (R,128) gradient (row stride dr; dt ignored) -- no (R,T,128) gradient tensor of zeros.
PD_API int pd_gru128_bwd_rows(const float* dout, long dr, long dt, const float* h_all, long hr, long ht, const float* rzn,
                              long zr, long zt, const float* hn, long nr, long nt, const int* lengths, const float* w_hh,
                              float* dgi, long gr, long gt, float* dgh, long qr, long qt, long R, int T, int reverse,
                              const int* cp, int dout_step, void* stream) {
    return gru128_bwd_impl(dout, dr, dt, h_all, hr, ht, rzn, zr, zt, hn, nr, nt, lengths, w_hh, dgi, gr, gt, dgh, qr, qt, R, T,
                           reverse, cp, dout_step, stream);
}
