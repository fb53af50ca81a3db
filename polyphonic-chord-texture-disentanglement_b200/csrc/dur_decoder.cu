// Fused duration decoder (ptvae.py:345-367): per note, a 5-step GRU (input 5, hidden 64) started from
// dur_hid, duration head Linear(64->2) after every step, argmax bit fed back as a one-hot token.
// The reference issues 5 x (aten::gru + Linear + argmax + host-built one-hot) per note, 2,400 tiny calls
// per forward.  Here the whole 5-step recurrence of a tile of notes runs inside one CTA with W_hh
// RESIDENT IN REGISTERS (thread j keeps row j of the 192x64 matrix, 64 registers) and the hidden states
// in shared memory; the per-step matrix-vector products are register-FMA against shared-memory
// broadcasts, gate math and the greedy bit never leave the SM.  fp32 throughout (it is also the
// fp32-faithful path greedy decoding needs for token parity).
//
// State buffer S (Q, 6, 72) written by the forward for the backward pass and its weight-gradient GEMM:
//   slot s, cols 0..63 : hidden state ENTERING step s  (slot 0 = dur_hid output, slot s = h_{s-1})
//           cols 64..68: input token of step s (slot 0: dur_sos_token; slots 1..4: one-hot of the fed-back
//                        bit; slot 5: zeros)            col 69: 1.0      col 70: [s == 0]      col 71: 0
// Gradient buffer GX (Q, 6, 264) written by the backward:
//   slot s, cols 0..191: dL/d(W_hh h + b_hh) of step s = [dr | dz | dn*r]   cols 192..255: dn
//           cols 256..257: dL/dlogits of step s-1 (pairs with the state in S slot s)   cols 258..263: 0
// so that ONE tensor-core GEMM  GX^T (264 x 6Q) . S (6Q x 72)  yields every parameter gradient:
//   rows 0..191 x cols 0..63 -> dW_hh;  col 69 -> db_hh;  rows {0..127,192..255} x cols 64..68 -> dW_ih,
//   col 69 -> db_ih, col 70 -> sum of step-0 input-gate grads (for d dur_sos_token);
//   rows 256..257 x cols 0..63 -> dW_out, col 69 -> db_out.
//
// Two arithmetic modes (template TC): FFMA fp32 as described (the fp32-faithful mode greedy decoding needs
// for token parity) and, for training, TF32 tensor-core matvecs: each warp keeps its slice of W_hh as
// mma.sync m16n8k8 B-fragments in registers and multiplies the 32-note state tile straight out of shared
// memory.  (tcgen05 is not used here on purpose: the 64-wide, 5-step recurrence is bound by gate
// transcendentals and state traffic, the matvec is <10 % of the kernel once it is off the FFMA pipe, and
// warp-level MMA keeps accumulators in registers next to the gate math instead of a TMEM round trip.)
#include "common.cuh"

namespace {

constexpr int H = 64, G3 = 192, SW = 72, GXW = 264, NSLOT = 6, NSTEP = 5;
constexpr int RT = 32;            // notes per CTA tile
constexpr int HS = 68;            // padded row stride of 64-wide shared arrays (conflict-free LDS.128 per row)
constexpr int GS = 196;           // padded row stride of 192-wide shared arrays
constexpr int NTHR = 192;

struct DurParams {
    const float* w_ih;  // (192,5)
    const float* b_ih;  // (192)
    const float* w_hh;  // (192,64)
    const float* b_hh;  // (192)
    const float* sos;   // (5)
    const float* w_out; // (2,64)
    const float* b_out; // (2)
};

__device__ __forceinline__ float dot64(const float (&w)[H], const float* __restrict__ hrow) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(hrow + k);
        a0 = fmaf(w[k], v.x, a0); a1 = fmaf(w[k + 1], v.y, a1);
        a2 = fmaf(w[k + 2], v.z, a2); a3 = fmaf(w[k + 3], v.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// out[r][n] = bias[n] + sum_k in[r][k] * W[n][k] for a 32-row tile, W slice held as B fragments:
// warp w owns output columns [32w, 32w+32) (4 n-tiles), K = 64 (8 k-steps).
__device__ __forceinline__ void tile_matvec_rows_tc(const uint32_t (&bw)[4][8][2], const float (&bias)[4][2],
                                                    const float (*in_s)[68], float (*out_s)[196], int warp, int lane) {
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][2] = bias[nt][0]; acc[nt][1] = acc[nt][3] = bias[nt][1]; }
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            uint32_t a[4];
            a[0] = to_tf32(in_s[16 * m + g][8 * kt + tig]);
            a[1] = to_tf32(in_s[16 * m + g + 8][8 * kt + tig]);
            a[2] = to_tf32(in_s[16 * m + g][8 * kt + tig + 4]);
            a[3] = to_tf32(in_s[16 * m + g + 8][8 * kt + tig + 4]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], a, bw[nt][kt][0], bw[nt][kt][1]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int col = 32 * warp + 8 * nt + 2 * tig;
            *reinterpret_cast<float2*>(&out_s[16 * m + g][col]) = make_float2(acc[nt][0], acc[nt][1]);
            *reinterpret_cast<float2*>(&out_s[16 * m + g + 8][col]) = make_float2(acc[nt][2], acc[nt][3]);
        }
    }
}

__device__ __forceinline__ void load_row_frags(const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                                               uint32_t (&bw)[4][8][2], float (&bias)[4][2], int warp, int lane) {
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int n = 32 * warp + 8 * nt + g;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            bw[nt][kt][0] = to_tf32(w_hh[n * H + 8 * kt + tig]);
            bw[nt][kt][1] = to_tf32(w_hh[n * H + 8 * kt + tig + 4]);
        }
        bias[nt][0] = b_hh[32 * warp + 8 * nt + 2 * tig];
        bias[nt][1] = b_hh[32 * warp + 8 * nt + 2 * tig + 1];
    }
}

// gi tables: [0] = W_ih sos + b_ih (step 0), [1] = W_ih[:,0] + b_ih (fed-back bit 0), [2] = W_ih[:,1] + b_ih
__device__ __forceinline__ void build_gi_tables(const DurParams& p, float (*gi_t)[G3]) {
    for (int j = threadIdx.x; j < G3; j += blockDim.x) {
        float b = p.b_ih[j], s = b;
#pragma unroll
        for (int c = 0; c < 5; ++c) s = fmaf(p.w_ih[j * 5 + c], p.sos[c], s);
        gi_t[0][j] = s;
        gi_t[1][j] = p.w_ih[j * 5 + 0] + b;
        gi_t[2][j] = p.w_ih[j * 5 + 1] + b;
    }
}

template <bool TC>
__global__ void __launch_bounds__(NTHR) dur_fwd_kernel(const float* __restrict__ h0, long ldh0, long Q, DurParams p,
                                                       float* __restrict__ logits, float* __restrict__ S) {
    __shared__ __align__(16) float h_s[RT][HS];
    __shared__ __align__(16) float gh_s[RT][GS];
    __shared__ float gi_t[3][G3];
    __shared__ __align__(16) float wo_s[2][H];
    __shared__ float lg_s[RT][2];
    __shared__ int tok_s[RT];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    float w[H];                 // FFMA mode: row tid of W_hh        (dead in TC mode)
    uint32_t bw[4][8][2];       // TC mode: this warp's B fragments  (dead in FFMA mode)
    float bfrag[4][2];
    float bh = 0.0f;
    if (TC) {
        load_row_frags(p.w_hh, p.b_hh, bw, bfrag, warp, lane);
    } else {
#pragma unroll
        for (int k = 0; k < H; k += 4) {
            float4 v = *reinterpret_cast<const float4*>(p.w_hh + tid * H + k);
            w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
        }
        bh = p.b_hh[tid];
    }
    build_gi_tables(p, gi_t);
    if (tid < 2 * H) wo_s[tid / H][tid % H] = p.w_out[tid];
    const float bo0 = p.b_out[0], bo1 = p.b_out[1];

    const long n_tiles = (Q + RT - 1) / RT;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long q0 = tile * RT;
        const int rows = (int)min((long)RT, Q - q0);
        __syncthreads();
        for (int i = tid; i < RT * H; i += NTHR) {
            int r = i / H, u = i % H;
            h_s[r][u] = (r < rows) ? h0[(q0 + r) * ldh0 + u] : 0.0f;
        }
        if (tid < RT) tok_s[tid] = 0;     // table 0: dur_sos_token
        __syncthreads();
        for (int k = 0; k < NSTEP; ++k) {
            // state entering step k -> S slot k (with the step's input token)
            if (S) {
                for (int i = tid; i < rows * SW; i += NTHR) {
                    int r = i / SW, c = i % SW;
                    float v;
                    if (c < H) v = h_s[r][c];
                    else if (c < 69) v = (k == 0) ? p.sos[c - 64] : ((c - 64) == (tok_s[r] - 1) ? 1.0f : 0.0f);
                    else if (c == 69) v = 1.0f;
                    else if (c == 70) v = (k == 0) ? 1.0f : 0.0f;
                    else v = 0.0f;
                    S[((q0 + r) * NSLOT + k) * SW + c] = v;
                }
            }
            // phase A: gh[r][j] = b_hh[j] + W_hh[j] . h[r]
            if (TC) {
                tile_matvec_rows_tc(bw, bfrag, h_s, gh_s, warp, lane);
            } else {
#pragma unroll 2
                for (int r = 0; r < RT; ++r) gh_s[r][tid] = bh + dot64(w, h_s[r]);
            }
            __syncthreads();
            // phase B: gates
            for (int i = tid; i < RT * H; i += NTHR) {
                int r = i / H, u = i % H;
                const float* gi = gi_t[tok_s[r]];
                float rr = pd_sigmoid(gi[u] + gh_s[r][u]);
                float zz = pd_sigmoid(gi[H + u] + gh_s[r][H + u]);
                float nn = tanhf(gi[2 * H + u] + rr * gh_s[r][2 * H + u]);
                h_s[r][u] = (1.0f - zz) * nn + zz * h_s[r][u];
            }
            __syncthreads();
            // phase C: duration head
            if (tid < 2 * RT) {
                int r = tid >> 1, o = tid & 1;
                float a = o ? bo1 : bo0;
#pragma unroll
                for (int u = 0; u < H; u += 4) {
                    float4 hv = *reinterpret_cast<const float4*>(&h_s[r][u]);
                    float4 wv = *reinterpret_cast<const float4*>(&wo_s[o][u]);
                    a = fmaf(hv.x, wv.x, a); a = fmaf(hv.y, wv.y, a); a = fmaf(hv.z, wv.z, a); a = fmaf(hv.w, wv.w, a);
                }
                lg_s[r][o] = a;
                if (r < rows) logits[((q0 + r) * NSTEP + k) * 2 + o] = a;
            }
            __syncthreads();
            if (tid < RT) tok_s[tid] = (lg_s[tid][1] > lg_s[tid][0]) ? 2 : 1;   // table index of the fed-back bit
            __syncthreads();
        }
        if (S) {   // slot 5: final state, no token
            for (int i = tid; i < rows * SW; i += NTHR) {
                int r = i / SW, c = i % SW;
                S[((q0 + r) * NSLOT + NSTEP) * SW + c] = (c < H) ? h_s[r][c] : (c == 69 ? 1.0f : 0.0f);
            }
        }
    }
}

constexpr int WS = 72;            // padded row stride of the shared W_hh copy (conflict-free B-fragment loads)

template <bool TC>
__global__ void __launch_bounds__(NTHR) dur_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dlog,
                                                       long Q, DurParams p, float* __restrict__ GX,
                                                       float* __restrict__ dh0, long lddh0) {
    extern __shared__ __align__(16) float dyn_smem[];
    float (*hp_s)[HS] = reinterpret_cast<float (*)[HS]>(dyn_smem);                 // state entering the step
    float (*dh_s)[HS] = reinterpret_cast<float (*)[HS]>(dyn_smem + RT * HS);       // grad wrt the step's output state
    float (*dn_s)[HS] = reinterpret_cast<float (*)[HS]>(dyn_smem + 2 * RT * HS);
    float (*g_s)[GS] = reinterpret_cast<float (*)[GS]>(dyn_smem + 3 * RT * HS);    // gh, then [dr | dz | dn*r]
    // the tail of the dynamic buffer is the tf32 W_hh copy (TC mode) or the dh_prev partials (FFMA mode)
    float (*w_s)[WS] = reinterpret_cast<float (*)[WS]>(dyn_smem + 3 * RT * HS + RT * GS);
    float (*part_s)[RT][HS] = reinterpret_cast<float (*)[RT][HS]>(dyn_smem + 3 * RT * HS + RT * GS);
    __shared__ float gi_t[3][G3];
    __shared__ __align__(16) float wo_s[2][H];
    __shared__ float dl_s[RT][2];
    __shared__ int tok_s[RT];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    float w[H];      // FFMA: row tid of W_hh               (recompute gh)
    float wc[H];     // FFMA: column (tid%64), rows third*64.. (dh_prev = W_hh^T dgh)
    uint32_t bw[4][8][2];
    float bfrag[4][2];
    const int third = tid / H, kk = tid % H;
    float bh = 0.0f;
    if (TC) {
        load_row_frags(p.w_hh, p.b_hh, bw, bfrag, warp, lane);
        for (int i = tid; i < G3 * H; i += NTHR) w_s[i / H][i % H] = __uint_as_float(to_tf32(p.w_hh[i]));
    } else {
#pragma unroll
        for (int k = 0; k < H; k += 4) {
            float4 v = *reinterpret_cast<const float4*>(p.w_hh + tid * H + k);
            w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < H; ++j) wc[j] = p.w_hh[(third * H + j) * H + kk];
        bh = p.b_hh[tid];
    }
    build_gi_tables(p, gi_t);
    if (tid < 2 * H) wo_s[tid / H][tid % H] = p.w_out[tid];

    const long n_tiles = (Q + RT - 1) / RT;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long q0 = tile * RT;
        const int rows = (int)min((long)RT, Q - q0);
        __syncthreads();
        for (int i = tid; i < RT * H; i += NTHR) dh_s[i / H][i % H] = 0.0f;
        // GX slot 0 has no logit gradient; slot 5 has no gate gradient
        for (int i = tid; i < rows * 8; i += NTHR) GX[((q0 + i / 8) * NSLOT) * GXW + 256 + (i % 8)] = 0.0f;
        for (int i = tid; i < rows * 256; i += NTHR) GX[((q0 + i / 256) * NSLOT + NSTEP) * GXW + (i % 256)] = 0.0f;
        for (int k = NSTEP - 1; k >= 0; --k) {
            __syncthreads();
            for (int i = tid; i < RT * H; i += NTHR) {
                int r = i / H, u = i % H;
                hp_s[r][u] = (r < rows) ? S[((q0 + r) * NSLOT + k) * SW + u] : 0.0f;
            }
            if (tid < RT) {
                int t = 0;
                if (k > 0 && tid < rows) t = (S[((q0 + tid) * NSLOT + k) * SW + 65] > 0.5f) ? 2 : 1;
                tok_s[tid] = t;
            }
            if (tid < 2 * RT) {
                int r = tid >> 1, o = tid & 1;
                float v = (r < rows) ? dlog[((q0 + r) * NSTEP + k) * 2 + o] : 0.0f;
                dl_s[r][o] = v;
                if (r < rows) GX[((q0 + r) * NSLOT + k + 1) * GXW + 256 + o] = v;
            }
            if (tid >= 64 && tid < 64 + RT) {   // zero the pad columns of the slot that receives the logit gradient
                int r = tid - 64;
                if (r < rows)
                    for (int c = 258; c < GXW; ++c) GX[((q0 + r) * NSLOT + k + 1) * GXW + c] = 0.0f;
            }
            __syncthreads();
            // head backward: dh += W_out^T dlogits
            for (int i = tid; i < RT * H; i += NTHR) {
                int r = i / H, u = i % H;
                dh_s[r][u] += wo_s[0][u] * dl_s[r][0] + wo_s[1][u] * dl_s[r][1];
            }
            // recompute gh
            if (TC) {
                tile_matvec_rows_tc(bw, bfrag, hp_s, g_s, warp, lane);
            } else {
#pragma unroll 2
                for (int r = 0; r < RT; ++r) g_s[r][tid] = bh + dot64(w, hp_s[r]);
            }
            __syncthreads();
            // gate backward (in place: gh -> dgh)
            for (int i = tid; i < RT * H; i += NTHR) {
                int r = i / H, u = i % H;
                const float* gi = gi_t[tok_s[r]];
                float ghn = g_s[r][2 * H + u];
                float rr = pd_sigmoid(gi[u] + g_s[r][u]);
                float zz = pd_sigmoid(gi[H + u] + g_s[r][H + u]);
                float nn = tanhf(gi[2 * H + u] + rr * ghn);
                float d = dh_s[r][u];
                float dn = d * (1.0f - zz) * (1.0f - nn * nn);
                float dz = d * (hp_s[r][u] - nn) * zz * (1.0f - zz);
                g_s[r][u] = dn * ghn * rr * (1.0f - rr);
                g_s[r][H + u] = dz;
                g_s[r][2 * H + u] = dn * rr;
                dn_s[r][u] = dn;
                dh_s[r][u] = d * zz;                 // direct path to the previous state
            }
            __syncthreads();
            if (TC) {
                // dh_prev[r][kk] += sum_j dgh[r][j] W_hh[j][kk]: warps 0..3 own 16 output columns each
                // (2 n-tiles), K = 192 (24 k-steps), A from g_s, B fragments from the shared tf32 W copy;
                // every dh_s element is owned by exactly one thread, so the product is added in place.
                if (warp < 4) {
                    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
                        for (int kt = 0; kt < 24; ++kt) {
                            uint32_t a[4];
                            a[0] = to_tf32(g_s[16 * m + g][8 * kt + tig]);
                            a[1] = to_tf32(g_s[16 * m + g + 8][8 * kt + tig]);
                            a[2] = to_tf32(g_s[16 * m + g][8 * kt + tig + 4]);
                            a[3] = to_tf32(g_s[16 * m + g + 8][8 * kt + tig + 4]);
#pragma unroll
                            for (int nt = 0; nt < 2; ++nt) {
                                const int n = 16 * warp + 8 * nt + g;
                                mma_tf32(acc[nt], a, __float_as_uint(w_s[8 * kt + tig][n]),
                                         __float_as_uint(w_s[8 * kt + tig + 4][n]));
                            }
                        }
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const int col = 16 * warp + 8 * nt + 2 * tig;
                            dh_s[16 * m + g][col] += acc[nt][0];
                            dh_s[16 * m + g][col + 1] += acc[nt][1];
                            dh_s[16 * m + g + 8][col] += acc[nt][2];
                            dh_s[16 * m + g + 8][col + 1] += acc[nt][3];
                        }
                    }
                }
            } else {
                // dh_prev partials: part[third][r][kk] = sum_j W_hh[third*64+j][kk] * dgh[r][third*64+j]
#pragma unroll 2
                for (int r = 0; r < RT; ++r) part_s[third][r][kk] = dot64(wc, &g_s[r][third * H]);
            }
            // gate gradients -> GX slot k
            for (int i = tid; i < rows * 256; i += NTHR) {
                int r = i >> 8, c = i & 255;
                GX[((q0 + r) * NSLOT + k) * GXW + c] = (c < G3) ? g_s[r][c] : dn_s[r][c - G3];
            }
            if (!TC) {
                __syncthreads();
                for (int i = tid; i < RT * H; i += NTHR) {
                    int r = i / H, u = i % H;
                    dh_s[r][u] += part_s[0][r][u] + part_s[1][r][u] + part_s[2][r][u];
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < rows * H; i += NTHR) dh0[(q0 + i / H) * lddh0 + (i % H)] = dh_s[i / H][i % H];
    }
}

unsigned dur_grid(long Q) {
    long tiles = (Q + RT - 1) / RT;
    long cap = 4L * PD_NUM_SMS;
    return (unsigned)(tiles < cap ? tiles : cap);
}

}  // namespace

// logits (Q,5,2) <- 5-step greedy-feedback duration GRU from h0 (Q,64; row stride ldh0).  S (Q,6,72) may be
// NULL (inference).
PD_API int pd_dur_decode_fwd(const float* h0, long ldh0, long Q, const float* w_ih, const float* b_ih,
                             const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                             const float* b_out, float* logits, float* S, int tf32, void* stream) {
    if (Q <= 0) return 0;
    if (((uintptr_t)w_hh & 15)) return PD_BAD_ARG;
    DurParams p{w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out};
    if (tf32) dur_fwd_kernel<true><<<dur_grid(Q), NTHR, 0, (cudaStream_t)stream>>>(h0, ldh0, Q, p, logits, S);
    else dur_fwd_kernel<false><<<dur_grid(Q), NTHR, 0, (cudaStream_t)stream>>>(h0, ldh0, Q, p, logits, S);
    return pd_launch_status();
}

// GX (Q,6,264) and dh0 (Q,64) <- S, dlogits (Q,5,2).  Parameter gradients = GX^T . S (one GEMM by the caller).
PD_API int pd_dur_decode_bwd(const float* S, const float* dlogits, long Q, const float* w_ih, const float* b_ih,
                             const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                             const float* b_out, float* GX, float* dh0, long lddh0, int tf32, void* stream) {
    if (Q <= 0) return 0;
    if (((uintptr_t)w_hh & 15)) return PD_BAD_ARG;
    DurParams p{w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out};
    constexpr int smem_tc = (3 * RT * HS + RT * GS + G3 * WS) * (int)sizeof(float);
    constexpr int smem_ff = (3 * RT * HS + RT * GS + 3 * RT * HS) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(dur_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(dur_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ff);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    if (tf32) dur_bwd_kernel<true><<<dur_grid(Q), NTHR, smem_tc, (cudaStream_t)stream>>>(S, dlogits, Q, p, GX, dh0, lddh0);
    else dur_bwd_kernel<false><<<dur_grid(Q), NTHR, smem_ff, (cudaStream_t)stream>>>(S, dlogits, Q, p, GX, dh0, lddh0);
    return pd_launch_status();
}
