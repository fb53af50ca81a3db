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
// Two arithmetic modes.  fp32 (tf32 = 0): the FFMA kernels described above -- the fp32-faithful path greedy
// decoding needs for token parity.  TF32 (tf32 = 1, training): warp-autonomous kernels (dur_*_warp_kernel below):
// every warp owns 16 notes for the whole recurrence, multiplies its state tile by W_hh (TF32-rounded, resident in
// shared memory in a k-permuted layout so that one LDS.64 yields an mma B fragment) with mma.sync m16n8k8 into
// register accumulators whose layout puts r, z and n of a unit in the same thread, does the gate math, the head
// and the greedy bit in registers with quad shuffles, and synchronises with __syncwarp only -- no block barrier
// in the step loop.  (ncu, 245,760 notes: the block-synchronous TF32 variant this replaces ran 0.97 / 1.86 ms
// fwd / bwd at 43 % issue utilisation, stalled on barriers and fixed-latency waits with 11 % tensor-pipe use.)
#include "common.cuh"

namespace {

constexpr int H = 64, G3 = 192, SW = 72, GXW = 264, NSLOT = 6, NSTEP = 5;
constexpr int RT = 32;            // notes per CTA tile (FFMA kernels)
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

// ---- fp32 (FFMA) kernels ----------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR) dur_fwd_kernel(const float* __restrict__ h0, long ldh0, long Q, DurParams p,
                                                       float* __restrict__ logits, float* __restrict__ S) {
    __shared__ __align__(16) float h_s[RT][HS];
    __shared__ __align__(16) float gh_s[RT][GS];
    __shared__ float gi_t[3][G3];
    __shared__ __align__(16) float wo_s[2][H];
    __shared__ float lg_s[RT][2];
    __shared__ int tok_s[RT];
    const int tid = threadIdx.x;
    float w[H];                 // row tid of W_hh
#pragma unroll
    for (int k = 0; k < H; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(p.w_hh + tid * H + k);
        w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
    }
    const float bh = p.b_hh[tid];
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
#pragma unroll 2
            for (int r = 0; r < RT; ++r) gh_s[r][tid] = bh + dot64(w, h_s[r]);
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

__global__ void __launch_bounds__(NTHR) dur_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dlog,
                                                       long Q, DurParams p, float* __restrict__ GX,
                                                       float* __restrict__ dh0, long lddh0) {
    extern __shared__ __align__(16) float dyn_smem[];
    float (*hp_s)[HS] = reinterpret_cast<float (*)[HS]>(dyn_smem);                 // state entering the step
    float (*dh_s)[HS] = reinterpret_cast<float (*)[HS]>(dyn_smem + RT * HS);       // grad wrt the step's output state
    float (*dn_s)[HS] = reinterpret_cast<float (*)[HS]>(dyn_smem + 2 * RT * HS);
    float (*g_s)[GS] = reinterpret_cast<float (*)[GS]>(dyn_smem + 3 * RT * HS);    // gh, then [dr | dz | dn*r]
    float (*part_s)[RT][HS] = reinterpret_cast<float (*)[RT][HS]>(dyn_smem + 3 * RT * HS + RT * GS);   // dh_prev partials
    __shared__ float gi_t[3][G3];
    __shared__ __align__(16) float wo_s[2][H];
    __shared__ float dl_s[RT][2];
    __shared__ int tok_s[RT];
    const int tid = threadIdx.x;
    float w[H];      // row tid of W_hh               (recompute gh)
    float wc[H];     // column (tid%64), rows third*64.. (dh_prev = W_hh^T dgh)
    const int third = tid / H, kk = tid % H;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(p.w_hh + tid * H + k);
        w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < H; ++j) wc[j] = p.w_hh[(third * H + j) * H + kk];
    const float bh = p.b_hh[tid];
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
#pragma unroll 2
            for (int r = 0; r < RT; ++r) g_s[r][tid] = bh + dot64(w, hp_s[r]);
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
            // dh_prev partials: part[third][r][kk] = sum_j W_hh[third*64+j][kk] * dgh[r][third*64+j]
#pragma unroll 2
            for (int r = 0; r < RT; ++r) part_s[third][r][kk] = dot64(wc, &g_s[r][third * H]);
            // gate gradients -> GX slot k
            for (int i = tid; i < rows * 256; i += NTHR) {
                int r = i >> 8, c = i & 255;
                GX[((q0 + r) * NSLOT + k) * GXW + c] = (c < G3) ? g_s[r][c] : dn_s[r][c - G3];
            }
            __syncthreads();
            for (int i = tid; i < RT * H; i += NTHR) {
                int r = i / H, u = i % H;
                dh_s[r][u] += part_s[0][r][u] + part_s[1][r][u] + part_s[2][r][u];
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

// ---- TF32 warp-autonomous kernels ----------------------------------------------------------------------
constexpr int WM = 16;            // notes per warp tile (one m16 MMA tile)
constexpr int WSP = 72;           // row stride of the shared W_hh copy (== 8 mod 32: conflict-free fragment loads)
constexpr int TS = 200;           // row stride of the gi tables (rows of different tokens on different banks)
constexpr int FW_WARPS = 12, BW_WARPS = 8;

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
// position of column k inside a W_hh row of the shared copy: (k, k+4) of every 8-group become neighbours, so the
// B fragment {W[n][8kt+tig], W[n][8kt+tig+4]} of mma.m16n8k8 is ONE 8-byte shared load
__device__ __forceinline__ int kperm(int k) { return (k & ~7) | ((k & 3) << 1) | ((k >> 2) & 1); }

struct WarpShared {              // block-wide part of the dynamic shared memory
    float w[G3 * WSP];           // TF32-rounded W_hh (hi part), k-permuted rows
    float gi_t[3 * TS];          // x-projection per input token; b_hh of the r and z gates folded in
    float bhn[H];                // b_hh of the n gate (multiplied by r, cannot be folded)
    float wo[2 * H];
    float sos[8];
};

// w_lo (3-pass mode only): TF32-rounded residual W_hh - hi, same layout, placed behind the per-warp tiles
__device__ __forceinline__ void warp_setup(const DurParams& p, WarpShared* sh, float* w_lo = nullptr) {
    for (int i = threadIdx.x; i < G3 * H; i += blockDim.x) {
        const int n = i / H, k = i % H;
        const float v = p.w_hh[i], hi = __uint_as_float(to_tf32(v));
        sh->w[n * WSP + kperm(k)] = hi;
        if (w_lo) w_lo[n * WSP + kperm(k)] = __uint_as_float(to_tf32(v - hi));
    }
    for (int j = threadIdx.x; j < G3; j += blockDim.x) {
        const float bh = j < 2 * H ? p.b_hh[j] : 0.0f;
        float b = p.b_ih[j], s = b;
#pragma unroll
        for (int c = 0; c < 5; ++c) s = fmaf(p.w_ih[j * 5 + c], p.sos[c], s);
        sh->gi_t[j] = s + bh;
        sh->gi_t[TS + j] = p.w_ih[j * 5 + 0] + b + bh;
        sh->gi_t[2 * TS + j] = p.w_ih[j * 5 + 1] + b + bh;
        if (j >= 2 * H) sh->bhn[j - 2 * H] = p.b_hh[j];
    }
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sh->wo[i] = p.w_out[i];
    if (threadIdx.x < 8) sh->sos[threadIdx.x] = threadIdx.x < 5 ? p.sos[threadIdx.x] : 0.0f;
}

// gh (16 x 192, bias-free) = hw (16 x 64 state tile of this warp) . W_hh^T, as 24 n-tiles of accumulators:
// acc[j] / acc[8+j] / acc[16+j] hold r / z / n of units 8j..8j+7; element [2*half+c] is row g+8*half, unit 8j+2*tig+c
// PASSES = 3: error-compensated products hi*lo + lo*hi + hi*hi with a second shared copy holding the weights' low
// parts (fp32-class accuracy for the token-parity decode); PASSES = 1: hi parts only.
template <int PASSES = 1>
__device__ __forceinline__ void warp_matvec(const float* __restrict__ w_s, const float (*hw)[HS], float (&acc)[24][4],
                                            int g, int tig, const float* __restrict__ w_lo = nullptr) {
#pragma unroll
    for (int nt = 0; nt < 24; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
        const float av[4] = {hw[g][8 * kt + tig], hw[g + 8][8 * kt + tig], hw[g][8 * kt + tig + 4], hw[g + 8][8 * kt + tig + 4]};
        uint32_t a[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = to_tf32(av[i]);
            if (PASSES == 3) al[i] = to_tf32(av[i] - __uint_as_float(a[i]));
        }
#pragma unroll
        for (int nt = 0; nt < 24; ++nt) {
            const float2 b = *reinterpret_cast<const float2*>(w_s + (8 * nt + g) * WSP + 8 * kt + 2 * tig);
            if (PASSES == 3) {
                const float2 bl = *reinterpret_cast<const float2*>(w_lo + (8 * nt + g) * WSP + 8 * kt + 2 * tig);
                mma_tf32(acc[nt], a, __float_as_uint(bl.x), __float_as_uint(bl.y));
                mma_tf32(acc[nt], al, __float_as_uint(b.x), __float_as_uint(b.y));
                mma_tf32(acc[nt], a, __float_as_uint(b.x), __float_as_uint(b.y));
            } else {
                mma_tf32(acc[nt], a, __float_as_uint(b.x), __float_as_uint(b.y));
            }
        }
    }
}

// 16 rows x 64 floats (row stride ld) -> hw; rows >= rows are zero-filled.  al16: rows are 16-byte aligned
// (else 8-byte: the decoder reads dur_hid out of a wider head buffer at an even column offset)
__device__ __forceinline__ void warp_load_rows(const float* __restrict__ src, long ld, int rows, float (*hw)[HS], int lane,
                                               bool al16 = true) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i, r = idx >> 4, c4 = (idx & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
            const float* q = src + (long)r * ld + c4;
            if (al16) {
                v = __ldg(reinterpret_cast<const float4*>(q));
            } else {
                const float2 a = __ldg(reinterpret_cast<const float2*>(q)), b = __ldg(reinterpret_cast<const float2*>(q + 2));
                v = make_float4(a.x, a.y, b.x, b.y);
            }
        }
        *reinterpret_cast<float4*>(&hw[r][c4]) = v;
    }
}

// asynchronous variant (cp.async, 16-byte aligned rows): the copy lands in hw without passing through registers,
// so it can be issued a phase early; complete it with warp_async_wait() + __syncwarp()
__device__ __forceinline__ void warp_load_rows_async(const float* __restrict__ src, long ld, int rows, float (*hw)[HS], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i, r = idx >> 4, c4 = (idx & 15) * 4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&hw[r][c4]);
        const float* q = src + (long)(r < rows ? r : 0) * ld + c4;
        const int bytes = r < rows ? 16 : 0;               // src-size 0: zero-fill
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(q), "r"(bytes) : "memory");
    }
}
__device__ __forceinline__ void warp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// state tile + input tokens of step k -> S slot k (slot 5: final state, no token); 18 float4 per row
__device__ __forceinline__ void warp_store_slot(float* __restrict__ S, long q0, int rows, int k, const float (*hw)[HS],
                                                const int* tokw, const float* sos, int lane) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int idx = lane + 32 * i, r = idx / 18, c = idx - 18 * r;
        if (r >= rows) continue;
        float4 v;
        if (c < 16) {
            v = *reinterpret_cast<const float4*>(&hw[r][4 * c]);
        } else if (k == NSTEP) {
            v = (c == 16) ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(0.f, 1.f, 0.f, 0.f);
        } else if (k == 0) {
            v = (c == 16) ? make_float4(sos[0], sos[1], sos[2], sos[3]) : make_float4(sos[4], 1.f, 1.f, 0.f);
        } else {
            const int t = tokw[r];                                           // 1: bit 0, 2: bit 1
            v = (c == 16) ? make_float4(t == 1 ? 1.f : 0.f, t == 2 ? 1.f : 0.f, 0.f, 0.f) : make_float4(0.f, 1.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(S + ((q0 + r) * NSLOT + k) * SW + 4 * c) = v;
    }
}

// gate functions: MUFU forms (abs error ~1e-6, the size of fp32 reassociation differences against the reference)
constexpr bool FAST_GATES_3X = true;
template <int PASSES> __device__ __forceinline__ float gate_sigmoid(float x) { return PASSES == 1 || FAST_GATES_3X ? pd_sigmoid_fast(x) : pd_sigmoid(x); }
template <int PASSES> __device__ __forceinline__ float gate_tanh(float x) { return PASSES == 1 || FAST_GATES_3X ? pd_tanh_fast(x) : tanhf(x); }

template <int PASSES>
__global__ void __launch_bounds__(FW_WARPS * 32, 1) dur_fwd_warp_kernel(const float* __restrict__ h0, long ldh0, long Q,
                                                                         DurParams p, float* __restrict__ logits,
                                                                         float* __restrict__ S, PdRows live) {
    extern __shared__ __align__(16) float dyn_smem[];
    WarpShared* sh = reinterpret_cast<WarpShared*>(dyn_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    float* wbase = dyn_smem + sizeof(WarpShared) / 4 + warp * (WM * HS + WM);
    float (*hw)[HS] = reinterpret_cast<float (*)[HS]>(wbase);                 // this warp's state tile
    int* tokw = reinterpret_cast<int*>(wbase + WM * HS);                      // token (gi table index) per row
    float* w_lo = PASSES == 3 ? dyn_smem + sizeof(WarpShared) / 4 + FW_WARPS * (WM * HS + WM) : nullptr;
    warp_setup(p, sh, w_lo);
    __syncthreads();
    const bool al16 = (((uintptr_t)h0 & 15) == 0) && ((ldh0 & 3) == 0);
    const float bo0 = p.b_out[0], bo1 = p.b_out[1];
    const long n16 = (Q + WM - 1) / WM;
    for (long t16 = blockIdx.x + (long)gridDim.x * warp; t16 < n16; t16 += (long)gridDim.x * FW_WARPS) {   // tiles spread over CTAs first
        const long q0 = t16 * WM;
        const int rows = (int)min((long)WM, Q - q0);
        if (!pd_rows_live(live, q0, WM)) continue;         // packed note level: dead notes are neither read nor written
        __syncwarp();
        warp_load_rows(h0 + q0 * ldh0, ldh0, rows, hw, lane, al16);
        if (lane < WM) tokw[lane] = 0;
        int tok[2] = {0, 0};
        for (int k = 0; k < NSTEP; ++k) {
            __syncwarp();
            if (S) warp_store_slot(S, q0, rows, k, hw, tokw, sh->sos, lane);
            float acc[24][4];
            warp_matvec<PASSES>(sh->w, hw, acc, g, tig, w_lo);
            __syncwarp();                                  // every lane is done reading the old state
            float l0[2] = {0.f, 0.f}, l1[2] = {0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int u = 8 * j + 2 * tig;
                const float2 bn = *reinterpret_cast<const float2*>(&sh->bhn[u]);
                const float2 w0 = *reinterpret_cast<const float2*>(&sh->wo[u]);
                const float2 w1 = *reinterpret_cast<const float2*>(&sh->wo[H + u]);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = g + 8 * half;
                    const float* tb = sh->gi_t + tok[half] * TS + u;
                    const float2 gr = *reinterpret_cast<const float2*>(tb);
                    const float2 gz = *reinterpret_cast<const float2*>(tb + H);
                    const float2 gn = *reinterpret_cast<const float2*>(tb + 2 * H);
                    const float2 hp = *reinterpret_cast<const float2*>(&hw[row][u]);
                    const float r0 = gate_sigmoid<PASSES>(gr.x + acc[j][2 * half]), r1 = gate_sigmoid<PASSES>(gr.y + acc[j][2 * half + 1]);
                    const float z0 = gate_sigmoid<PASSES>(gz.x + acc[8 + j][2 * half]), z1 = gate_sigmoid<PASSES>(gz.y + acc[8 + j][2 * half + 1]);
                    const float n0 = gate_tanh<PASSES>(gn.x + r0 * (acc[16 + j][2 * half] + bn.x));
                    const float n1 = gate_tanh<PASSES>(gn.y + r1 * (acc[16 + j][2 * half + 1] + bn.y));
                    const float2 hn = make_float2((1.0f - z0) * n0 + z0 * hp.x, (1.0f - z1) * n1 + z1 * hp.y);
                    *reinterpret_cast<float2*>(&hw[row][u]) = hn;
                    l0[half] = fmaf(hn.x, w0.x, fmaf(hn.y, w0.y, l0[half]));
                    l1[half] = fmaf(hn.x, w1.x, fmaf(hn.y, w1.y, l1[half]));
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {         // duration head: finish the two dot products in the quad
                l0[half] += __shfl_xor_sync(0xffffffffu, l0[half], 1);
                l1[half] += __shfl_xor_sync(0xffffffffu, l1[half], 1);
                l0[half] += __shfl_xor_sync(0xffffffffu, l0[half], 2);
                l1[half] += __shfl_xor_sync(0xffffffffu, l1[half], 2);
                const float a0 = l0[half] + bo0, a1 = l1[half] + bo1;
                const int row = g + 8 * half;
                tok[half] = a1 > a0 ? 2 : 1;               // table index of the fed-back bit
                if (tig == 0) {
                    tokw[row] = tok[half];
                    if (row < rows) *reinterpret_cast<float2*>(logits + ((q0 + row) * NSTEP + k) * 2) = make_float2(a0, a1);
                }
            }
        }
        __syncwarp();
        if (S) warp_store_slot(S, q0, rows, NSTEP, hw, tokw, sh->sos, lane);
    }
}

// ---- small-Q inference form (TF32, no saves): ONE 16-note tile per CTA, spread over four warps -------------------------
// A greedy note slot of a few hundred rows (the no-grad greedy pass of free-running / scheduled-sampling training at batch
// 512, step-wise decodes) is a latency chain: with one warp per tile the five GRU steps take ~12 us behind a ~5 us
// shared-memory staging of W_hh (21 us per 512-note call, 27 % of a free-running training step).  Here warp w owns
// hidden units [16w, 16w+16) of all three gates -- 6 of the 24 n-tiles -- so a step's matvec and gate math are a quarter as
// long; its 96 B-fragment values of W_hh (L2-resident) live in REGISTERS for the five steps, so nothing is staged; the new
// state goes to the other half of a double-buffered shared tile and the duration head's partial dot products meet
// through shared memory, with one block barrier per step.
constexpr int QW = 4;
struct QuadShared {
    float gi_t[3 * TS];
    float bhn[H];
    float wo[2 * H];
    float hw[2][WM][HS];
    float lpart[2][QW][WM][2];
};

__global__ void __launch_bounds__(QW * 32) dur_fwd_quad_kernel(const float* __restrict__ h0, long ldh0, long Q, DurParams p,
                                                               float* __restrict__ logits) {
    __shared__ __align__(16) QuadShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const long q0 = (long)blockIdx.x * WM;
    const int rows = (int)min((long)WM, Q - q0);
    uint32_t bf[6][8][2];                     // [gate * 2 + unit block][k-tile][k, k+4]
#pragma unroll
    for (int gt = 0; gt < 3; ++gt)
#pragma unroll
        for (int ub = 0; ub < 2; ++ub) {
            const float* wr = p.w_hh + (long)(gt * H + (2 * warp + ub) * 8 + g) * H;
#pragma unroll
            for (int kt = 0; kt < 8; ++kt) {
                bf[gt * 2 + ub][kt][0] = to_tf32(__ldg(wr + 8 * kt + tig));
                bf[gt * 2 + ub][kt][1] = to_tf32(__ldg(wr + 8 * kt + tig + 4));
            }
        }
    for (int j = threadIdx.x; j < G3; j += QW * 32) {              // x-projection per input token (as warp_setup)
        const float bh = j < 2 * H ? p.b_hh[j] : 0.0f;
        float b = p.b_ih[j], sos = b;
#pragma unroll
        for (int c = 0; c < 5; ++c) sos = fmaf(p.w_ih[j * 5 + c], p.sos[c], sos);
        sh.gi_t[j] = sos + bh;
        sh.gi_t[TS + j] = p.w_ih[j * 5 + 0] + b + bh;
        sh.gi_t[2 * TS + j] = p.w_ih[j * 5 + 1] + b + bh;
        if (j >= 2 * H) sh.bhn[j - 2 * H] = p.b_hh[j];
    }
    for (int i = threadIdx.x; i < 2 * H; i += QW * 32) sh.wo[i] = p.w_out[i];
    const bool al16 = (((uintptr_t)h0 & 15) == 0) && ((ldh0 & 3) == 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {                                  // state tile: 16 rows x 16 float4
        const int idx = threadIdx.x + QW * 32 * i, r = idx >> 4, c4 = (idx & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
            const float* q = h0 + (q0 + r) * ldh0 + c4;
            if (al16) {
                v = __ldg(reinterpret_cast<const float4*>(q));
            } else {
                const float2 a = __ldg(reinterpret_cast<const float2*>(q)), b = __ldg(reinterpret_cast<const float2*>(q + 2));
                v = make_float4(a.x, a.y, b.x, b.y);
            }
        }
        *reinterpret_cast<float4*>(&sh.hw[0][r][c4]) = v;
    }
    __syncthreads();
    const float bo0 = p.b_out[0], bo1 = p.b_out[1];
    int tok[2] = {0, 0};
    int cur = 0;
    for (int k = 0; k < NSTEP; ++k) {
        const float (*hc)[HS] = sh.hw[cur];
        float (*hx)[HS] = sh.hw[cur ^ 1];
        float acc[6][4];
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
            const uint32_t a[4] = {to_tf32(hc[g][8 * kt + tig]), to_tf32(hc[g + 8][8 * kt + tig]),
                                   to_tf32(hc[g][8 * kt + tig + 4]), to_tf32(hc[g + 8][8 * kt + tig + 4])};
#pragma unroll
            for (int nt = 0; nt < 6; ++nt) mma_tf32(acc[nt], a, bf[nt][kt][0], bf[nt][kt][1]);
        }
        float l0[2] = {0.f, 0.f}, l1[2] = {0.f, 0.f};
#pragma unroll
        for (int ub = 0; ub < 2; ++ub) {
            const int u = (2 * warp + ub) * 8 + 2 * tig;
            const float2 bn = *reinterpret_cast<const float2*>(&sh.bhn[u]);
            const float2 w0 = *reinterpret_cast<const float2*>(&sh.wo[u]);
            const float2 w1 = *reinterpret_cast<const float2*>(&sh.wo[H + u]);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = g + 8 * half;
                const float* tb = sh.gi_t + tok[half] * TS + u;
                const float2 gr = *reinterpret_cast<const float2*>(tb);
                const float2 gz = *reinterpret_cast<const float2*>(tb + H);
                const float2 gn = *reinterpret_cast<const float2*>(tb + 2 * H);
                const float2 hp = *reinterpret_cast<const float2*>(&hc[row][u]);
                const float r0 = pd_sigmoid_fast(gr.x + acc[ub][2 * half]), r1 = pd_sigmoid_fast(gr.y + acc[ub][2 * half + 1]);
                const float z0 = pd_sigmoid_fast(gz.x + acc[2 + ub][2 * half]), z1 = pd_sigmoid_fast(gz.y + acc[2 + ub][2 * half + 1]);
                const float n0 = pd_tanh_fast(gn.x + r0 * (acc[4 + ub][2 * half] + bn.x));
                const float n1 = pd_tanh_fast(gn.y + r1 * (acc[4 + ub][2 * half + 1] + bn.y));
                const float2 hn = make_float2((1.0f - z0) * n0 + z0 * hp.x, (1.0f - z1) * n1 + z1 * hp.y);
                *reinterpret_cast<float2*>(&hx[row][u]) = hn;
                l0[half] = fmaf(hn.x, w0.x, fmaf(hn.y, w0.y, l0[half]));
                l1[half] = fmaf(hn.x, w1.x, fmaf(hn.y, w1.y, l1[half]));
            }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            l0[half] += __shfl_xor_sync(0xffffffffu, l0[half], 1);
            l1[half] += __shfl_xor_sync(0xffffffffu, l1[half], 1);
            l0[half] += __shfl_xor_sync(0xffffffffu, l0[half], 2);
            l1[half] += __shfl_xor_sync(0xffffffffu, l1[half], 2);
            if (tig == 0) *reinterpret_cast<float2*>(&sh.lpart[k & 1][warp][g + 8 * half][0]) = make_float2(l0[half], l1[half]);
        }
        __syncthreads();                       // new state + the four partial heads are in shared memory
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = g + 8 * half;
            float a0 = bo0, a1 = bo1;
#pragma unroll
            for (int w = 0; w < QW; ++w) {
                const float2 v = *reinterpret_cast<const float2*>(&sh.lpart[k & 1][w][row][0]);
                a0 += v.x; a1 += v.y;
            }
            tok[half] = a1 > a0 ? 2 : 1;           // every warp takes the same decision from the same numbers
            if (warp == 0 && tig == 0 && row < rows)
                *reinterpret_cast<float2*>(logits + ((q0 + row) * NSTEP + k) * 2) = make_float2(a0, a1);
        }
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(BW_WARPS * 32, 1) dur_bwd_warp_kernel(const float* __restrict__ S,
                                                                         const float* __restrict__ dlog, long Q, DurParams p,
                                                                         float* __restrict__ GX, float* __restrict__ dh0,
                                                                         long lddh0, PdRows live) {
    extern __shared__ __align__(16) float dyn_smem[];
    WarpShared* sh = reinterpret_cast<WarpShared*>(dyn_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    float* wbase = dyn_smem + sizeof(WarpShared) / 4 + warp * (WM * HS + WM * GS + 3 * WM);
    float (*hw)[HS] = reinterpret_cast<float (*)[HS]>(wbase);                 // state entering the step
    float (*gw)[GS] = reinterpret_cast<float (*)[GS]>(wbase + WM * HS);       // [dr | dz | dn*r] of the step
    int* tokw = reinterpret_cast<int*>(wbase + WM * HS + WM * GS);
    float2* dlw = reinterpret_cast<float2*>(wbase + WM * HS + WM * GS + WM);  // dL/dlogits of the step per row
    warp_setup(p, sh);
    __syncthreads();
    const long n16 = (Q + WM - 1) / WM;
    for (long t16 = blockIdx.x + (long)gridDim.x * warp; t16 < n16; t16 += (long)gridDim.x * BW_WARPS) {
        const long q0 = t16 * WM;
        const int rows = (int)min((long)WM, Q - q0);
        if (!pd_rows_live(live, q0, WM)) continue;
        float dh[8][4];                                    // grad wrt the step's output state, accumulator layout
#pragma unroll
        for (int j = 0; j < 8; ++j) dh[j][0] = dh[j][1] = dh[j][2] = dh[j][3] = 0.0f;
        warp_load_rows_async(S + (q0 * NSLOT + NSTEP - 1) * SW, (long)NSLOT * SW, rows, hw, lane);
        // GX slot 5: no gate gradient, logit gradient of step 4 in cols 256..257
#pragma unroll
        for (int i = 0; i < 33; ++i) {
            const int idx = lane + 32 * i, r = idx / 66, c = idx - 66 * r;
            if (r < rows) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c == 64) {
                    const float2 d = __ldg(reinterpret_cast<const float2*>(dlog + ((q0 + r) * NSTEP + NSTEP - 1) * 2));
                    v.x = d.x; v.y = d.y;
                }
                *reinterpret_cast<float4*>(GX + ((q0 + r) * NSLOT + NSTEP) * GXW + 4 * c) = v;
            }
        }
        for (int k = NSTEP - 1; k >= 0; --k) {
            if (lane < WM) {
                int t = 0;
                float2 d = make_float2(0.f, 0.f);
                if (lane < rows) {
                    if (k > 0) t = (__ldg(S + ((q0 + lane) * NSLOT + k) * SW + 65) > 0.5f) ? 2 : 1;
                    d = __ldg(reinterpret_cast<const float2*>(dlog + ((q0 + lane) * NSTEP + k) * 2));
                }
                tokw[lane] = t;
                dlw[lane] = d;
            }
            warp_async_wait();
            __syncwarp();
            float acc[24][4];
            warp_matvec(sh->w, hw, acc, g, tig);           // recompute gh of the step
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = g + 8 * half;
                const int tok = tokw[row];
                const float2 dl = dlw[row];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int u = 8 * j + 2 * tig;
                    const float2 bn = *reinterpret_cast<const float2*>(&sh->bhn[u]);
                    const float2 w0 = *reinterpret_cast<const float2*>(&sh->wo[u]);
                    const float2 w1 = *reinterpret_cast<const float2*>(&sh->wo[H + u]);
                    const float* tb = sh->gi_t + tok * TS + u;
                    const float2 gr = *reinterpret_cast<const float2*>(tb);
                    const float2 gz = *reinterpret_cast<const float2*>(tb + H);
                    const float2 gn = *reinterpret_cast<const float2*>(tb + 2 * H);
                    const float2 hp = *reinterpret_cast<const float2*>(&hw[row][u]);
                    float dr[2], dz[2], dnr[2], dn[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float ghn = acc[16 + j][2 * half + c] + (c ? bn.y : bn.x);
                        const float rr = pd_sigmoid_fast((c ? gr.y : gr.x) + acc[j][2 * half + c]);
                        const float zz = pd_sigmoid_fast((c ? gz.y : gz.x) + acc[8 + j][2 * half + c]);
                        const float nn = pd_tanh_fast((c ? gn.y : gn.x) + rr * ghn);
                        // head backward first: dh += W_out^T dlogits
                        const float d = dh[j][2 * half + c] + (c ? w0.y : w0.x) * dl.x + (c ? w1.y : w1.x) * dl.y;
                        dn[c] = d * (1.0f - zz) * (1.0f - nn * nn);
                        dz[c] = d * ((c ? hp.y : hp.x) - nn) * zz * (1.0f - zz);
                        dr[c] = dn[c] * ghn * rr * (1.0f - rr);
                        dnr[c] = dn[c] * rr;
                        dh[j][2 * half + c] = d * zz;      // direct path to the previous state
                    }
                    *reinterpret_cast<float2*>(&gw[row][u]) = make_float2(dr[0], dr[1]);
                    *reinterpret_cast<float2*>(&gw[row][H + u]) = make_float2(dz[0], dz[1]);
                    *reinterpret_cast<float2*>(&gw[row][2 * H + u]) = make_float2(dnr[0], dnr[1]);
                    if (row < rows)
                        *reinterpret_cast<float2*>(GX + ((q0 + row) * NSLOT + k) * GXW + G3 + u) = make_float2(dn[0], dn[1]);
                }
            }
            __syncwarp();
            // the state tile is dead until the next step: fetch the state entering step k-1 behind the rest of this one
            if (k > 0) warp_load_rows_async(S + (q0 * NSLOT + k - 1) * SW, (long)NSLOT * SW, rows, hw, lane);
            // GX slot k: cols 0..191 from the staged gate gradients; cols 256..263 = [dlogits of step k-1, 0...]
#pragma unroll
            for (int i = 0; i < 24; ++i) {
                const int idx = lane + 32 * i, r = idx / 48, c = idx - 48 * r;
                if (r < rows)
                    *reinterpret_cast<float4*>(GX + ((q0 + r) * NSLOT + k) * GXW + 4 * c) = *reinterpret_cast<const float4*>(&gw[r][4 * c]);
            }
            {
                const int r = lane >> 1, c = lane & 1;
                if (r < rows) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c == 0 && k > 0) {
                        const float2 d = __ldg(reinterpret_cast<const float2*>(dlog + ((q0 + r) * NSTEP + k - 1) * 2));
                        v.x = d.x; v.y = d.y;
                    }
                    *reinterpret_cast<float4*>(GX + ((q0 + r) * NSLOT + k) * GXW + 256 + 4 * c) = v;
                }
            }
            // dh += dgh . W_hh: 8 n-tiles (units), K = 192 gate rows; B fragment = W_hh[8kt+tig (+4)][8nt+g]
            float acc2[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc2[nt][0] = acc2[nt][1] = acc2[nt][2] = acc2[nt][3] = 0.0f;
            const int pg = kperm(g);                       // column 8nt+g sits at 8nt + kperm(g) in the shared copy
#pragma unroll 4
            for (int kt = 0; kt < 24; ++kt) {
                const uint32_t a[4] = {to_tf32(gw[g][8 * kt + tig]), to_tf32(gw[g + 8][8 * kt + tig]),
                                       to_tf32(gw[g][8 * kt + tig + 4]), to_tf32(gw[g + 8][8 * kt + tig + 4])};
                const float* wr = sh->w + (8 * kt + tig) * WSP + pg;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    mma_tf32(acc2[nt], a, __float_as_uint(wr[8 * nt]), __float_as_uint(wr[4 * WSP + 8 * nt]));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) dh[j][c] += acc2[j][c];
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = g + 8 * half;
            if (row < rows)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float2*>(dh0 + (q0 + row) * lddh0 + 8 * j + 2 * tig) = make_float2(dh[j][2 * half], dh[j][2 * half + 1]);
        }
    }
}

constexpr int FW_SMEM = (int)sizeof(WarpShared) + FW_WARPS * (WM * HS + WM) * 4;
constexpr int FW_SMEM3 = FW_SMEM + G3 * WSP * 4;          // + low parts of W_hh
constexpr int BW_SMEM = (int)sizeof(WarpShared) + BW_WARPS * (WM * HS + WM * GS + 3 * WM) * 4;

unsigned warp_grid(long Q) {
    long tiles = (Q + WM - 1) / WM;
    return (unsigned)(tiles < PD_NUM_SMS ? tiles : PD_NUM_SMS);
}

}  // namespace

// logits (Q,5,2) <- 5-step greedy-feedback duration GRU from h0 (Q,64; row stride ldh0).  S (Q,6,72) may be
// NULL (inference).  tf32: 0 = fp32 FFMA kernels, 1 = TF32 tensor-core matvecs, 3 = error-compensated 3xTF32 matvecs with
// expf / tanhf gates (fp32-class, forward only).  Nonzero: S must be 16-byte and h0 8-byte aligned, ldh0 even.
// calls of at most this many notes without saves take dur_fwd_quad_kernel (pd_dur_quad_max_notes: tuning / A-B switch)
static long g_dur_quad_max = 2048;
PD_API int pd_dur_quad_max_notes(int n) {
    g_dur_quad_max = n;
    return 0;
}

static int dur_fwd_impl(const float* h0, long ldh0, long Q, const float* w_ih, const float* b_ih,
                        const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                        const float* b_out, float* logits, float* S, int tf32, PdRows live, void* stream) {
    if (Q <= 0) return 0;
    if (live.cp != nullptr && tf32 != 1) return PD_BAD_ARG;      // the live-row table is a training-mode feature
    if (((uintptr_t)w_hh & 15)) return PD_BAD_ARG;
    DurParams p{w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out};
    // small fp32-class calls (step-wise decode of a few segments) are latency bound: one warp per 16 notes doing
    // three MMA passes loses to the FFMA kernel, which spreads a tile over six warps (16-segment decode: 82 vs 62 ms)
    if (tf32 == 3 && Q < 4096) tf32 = 0;
    if (tf32) {
        if (((uintptr_t)S & 15) || ((uintptr_t)h0 & 7) || (ldh0 & 1) || ((uintptr_t)logits & 7)) return PD_BAD_ARG;
        if (tf32 == 1 && S == nullptr && live.cp == nullptr && Q <= g_dur_quad_max) {
            // small inference call: one tile per CTA over four warps, W_hh fragments in registers (no staging)
            dur_fwd_quad_kernel<<<(unsigned)((Q + WM - 1) / WM), QW * 32, 0, (cudaStream_t)stream>>>(h0, ldh0, Q, p, logits);
            return pd_launch_status();
        }
        static unsigned long long attr = 0;
        if (pd_first_use_on_device(attr)) {
            cudaError_t e = cudaFuncSetAttribute(dur_fwd_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(dur_fwd_warp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM3);
            if (e != cudaSuccess) return (int)e;
        }
        if (tf32 == 3) dur_fwd_warp_kernel<3><<<warp_grid(Q), FW_WARPS * 32, FW_SMEM3, (cudaStream_t)stream>>>(h0, ldh0, Q, p, logits, S, live);
        else dur_fwd_warp_kernel<1><<<warp_grid(Q), FW_WARPS * 32, FW_SMEM, (cudaStream_t)stream>>>(h0, ldh0, Q, p, logits, S, live);
    } else {
        dur_fwd_kernel<<<dur_grid(Q), NTHR, 0, (cudaStream_t)stream>>>(h0, ldh0, Q, p, logits, S);
    }
    return pd_launch_status();
}

PD_API int pd_dur_decode_fwd(const float* h0, long ldh0, long Q, const float* w_ih, const float* b_ih,
                             const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                             const float* b_out, float* logits, float* S, int tf32, void* stream) {
    return dur_fwd_impl(h0, ldh0, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, logits, S, tf32, PdRows{nullptr, 0, 0}, stream);
}

// Packed note level: notes are slot-major rows (Q = n_slots * slot_rows) with a DEVICE live-row table cp (common.cuh
// PdRows); 16-note tiles without a live note are skipped (their logits / S rows are not written).  TF32 mode only.
PD_API int pd_dur_decode_fwd_rows(const float* h0, long ldh0, long Q, const float* w_ih, const float* b_ih,
                                  const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                                  const float* b_out, float* logits, float* S, const int* cp, int slot_rows, void* stream) {
    if (cp == nullptr || slot_rows <= 0) return PD_BAD_ARG;
    return dur_fwd_impl(h0, ldh0, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, logits, S, 1,
                        PdRows{cp, slot_rows, (int)((Q + slot_rows - 1) / slot_rows)}, stream);
}

// GX (Q,6,264) and dh0 (Q,64) <- S, dlogits (Q,5,2).  Parameter gradients = GX^T . S (one GEMM by the caller).
static int dur_bwd_impl(const float* S, const float* dlogits, long Q, const float* w_ih, const float* b_ih,
                        const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                        const float* b_out, float* GX, float* dh0, long lddh0, int tf32, PdRows live, void* stream) {
    if (Q <= 0) return 0;
    if (live.cp != nullptr && tf32 != 1) return PD_BAD_ARG;
    if (((uintptr_t)w_hh & 15)) return PD_BAD_ARG;
    DurParams p{w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out};
    constexpr int smem_ff = (3 * RT * HS + RT * GS + 3 * RT * HS) * (int)sizeof(float);
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(dur_bwd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dur_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ff);
        if (e != cudaSuccess) return (int)e;
    }
    if (tf32) {
        if ((((uintptr_t)S | (uintptr_t)GX) & 15) || (((uintptr_t)dlogits | (uintptr_t)dh0) & 7) || (lddh0 & 1)) return PD_BAD_ARG;
        dur_bwd_warp_kernel<<<warp_grid(Q), BW_WARPS * 32, BW_SMEM, (cudaStream_t)stream>>>(S, dlogits, Q, p, GX, dh0, lddh0, live);
    } else {
        dur_bwd_kernel<<<dur_grid(Q), NTHR, smem_ff, (cudaStream_t)stream>>>(S, dlogits, Q, p, GX, dh0, lddh0);
    }
    return pd_launch_status();
}

PD_API int pd_dur_decode_bwd(const float* S, const float* dlogits, long Q, const float* w_ih, const float* b_ih,
                             const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                             const float* b_out, float* GX, float* dh0, long lddh0, int tf32, void* stream) {
    return dur_bwd_impl(S, dlogits, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, GX, dh0, lddh0, tf32, PdRows{nullptr, 0, 0}, stream);
}

// Packed note level (see pd_dur_decode_fwd_rows): GX / dh0 rows of dead 16-note tiles are not written.
PD_API int pd_dur_decode_bwd_rows(const float* S, const float* dlogits, long Q, const float* w_ih, const float* b_ih,
                                  const float* w_hh, const float* b_hh, const float* sos, const float* w_out,
                                  const float* b_out, float* GX, float* dh0, long lddh0, const int* cp, int slot_rows,
                                  void* stream) {
    if (cp == nullptr || slot_rows <= 0) return PD_BAD_ARG;
    return dur_bwd_impl(S, dlogits, Q, w_ih, b_ih, w_hh, b_hh, sos, w_out, b_out, GX, dh0, lddh0, 1,
                        PdRows{cp, slot_rows, (int)((Q + slot_rows - 1) / slot_rows)}, stream);
}
