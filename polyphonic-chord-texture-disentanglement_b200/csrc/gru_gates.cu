// GRU cell gate math (PyTorch nn.GRU semantics, gate order r|z|n) around the per-step h-projection GEMM.
//
//   r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1-z) n + z h
//
// gi  = W_ih x + b_ih   (precomputed for all steps by one batched GEMM; row stride ldgi)
// gi2 = optional second x-projection term that is constant over the sequence (z_in / notes_summary /
//       z_chd_in halves of the reference's torch.cat inputs, ptvae.py:65,397,462), one row per sequence
// gh  = W_hh h + b_hh   (the per-step GEMM)
//
// Replaces the aten::gru step calls at ptvae.py:63-65, :359-360, :396-398, :461-462 and the packed
// bi-GRU at :446-453 / :480-486 (variable length handled by a per-row length mask: a row whose
// length <= t keeps its hidden state, which reproduces pack_padded_sequence's final hidden).
// HBM-bound elementwise kernels: float4 over the hidden dimension, one thread per 4 hidden units.
#include "common.cuh"

namespace {

struct GateFwd {
    const float* gi; long ldgi;
    const float* gi2; long ldgi2;
    const float* gh; long ldgh;
    const float* hprev; long ldhp;   // may be NULL => zeros
    float* hout; long ldho;
    float* rzn; long ldrzn;          // may be NULL
    float* hn; long ldhn;            // may be NULL
    const int* lengths; int t;       // may be NULL
    int B, H;
    float* h3; long ldh3;            // may be NULL: [hi | hi | lo] TF32 split of the new state (3xTF32 GEMM operand)
};

// the new state as the A operand of the single-launch 3xTF32 GEMM (pd_tf32_split3 order 0), written by the producer
__device__ __forceinline__ void store_split3(float* p, int H, float4 v) {
    float4 hi, lo;
#define PD_SPLIT(c) { uint32_t b; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.c)); hi.c = __uint_as_float(b); lo.c = v.c - hi.c; }
    PD_SPLIT(x) PD_SPLIT(y) PD_SPLIT(z) PD_SPLIT(w)
#undef PD_SPLIT
    *reinterpret_cast<float4*>(p) = hi;
    *reinterpret_cast<float4*>(p + H) = hi;
    *reinterpret_cast<float4*>(p + 2 * H) = lo;
}

__global__ void __launch_bounds__(256) gru_gates_fwd_kernel(GateFwd a) {
    const int hq = a.H >> 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.B * hq) return;
    const int b = (int)(idx / hq), j = (int)(idx % hq) * 4;
    float4 hp = a.hprev ? *reinterpret_cast<const float4*>(a.hprev + (long)b * a.ldhp + j)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.lengths && a.t >= a.lengths[b]) {   // past the end of this sequence: carry the state
        *reinterpret_cast<float4*>(a.hout + (long)b * a.ldho + j) = hp;
        if (a.h3) store_split3(a.h3 + (long)b * a.ldh3 + j, a.H, hp);
        return;
    }
    const float* gi = a.gi + (long)b * a.ldgi + j;
    const float* gh = a.gh + (long)b * a.ldgh + j;
    float4 ir = *reinterpret_cast<const float4*>(gi), iz = *reinterpret_cast<const float4*>(gi + a.H),
           in = *reinterpret_cast<const float4*>(gi + 2 * a.H);
    if (a.gi2) {
        const float* g2 = a.gi2 + (long)b * a.ldgi2 + j;
        float4 r2 = *reinterpret_cast<const float4*>(g2), z2 = *reinterpret_cast<const float4*>(g2 + a.H),
               n2 = *reinterpret_cast<const float4*>(g2 + 2 * a.H);
        ir.x += r2.x; ir.y += r2.y; ir.z += r2.z; ir.w += r2.w;
        iz.x += z2.x; iz.y += z2.y; iz.z += z2.z; iz.w += z2.w;
        in.x += n2.x; in.y += n2.y; in.z += n2.z; in.w += n2.w;
    }
    float4 hr = *reinterpret_cast<const float4*>(gh), hz = *reinterpret_cast<const float4*>(gh + a.H),
           hnn = *reinterpret_cast<const float4*>(gh + 2 * a.H);
    float4 r, z, n, ho;
#define PD_GATE(c)                                        \
    r.c = pd_sigmoid(ir.c + hr.c);                        \
    z.c = pd_sigmoid(iz.c + hz.c);                        \
    n.c = tanhf(in.c + r.c * hnn.c);                      \
    ho.c = (1.0f - z.c) * n.c + z.c * hp.c;
    PD_GATE(x) PD_GATE(y) PD_GATE(z) PD_GATE(w)
#undef PD_GATE
    *reinterpret_cast<float4*>(a.hout + (long)b * a.ldho + j) = ho;
    if (a.rzn) {
        float* s = a.rzn + (long)b * a.ldrzn + j;
        *reinterpret_cast<float4*>(s) = r;
        *reinterpret_cast<float4*>(s + a.H) = z;
        *reinterpret_cast<float4*>(s + 2 * a.H) = n;
    }
    if (a.hn) *reinterpret_cast<float4*>(a.hn + (long)b * a.ldhn + j) = hnn;
    if (a.h3) store_split3(a.h3 + (long)b * a.ldh3 + j, a.H, ho);
}

struct GateBwd {
    const float* dh; long lddh;        // grad wrt this step's output state (recurrent part); may be NULL
    const float* dh2; long lddh2;      // extra grad source for the same state (output use); may be NULL
    const float* dh3; long lddh3;      // third source (the dgh * W_hh product of the later step); may be NULL
    const float* rzn; long ldrzn;
    const float* hn; long ldhn;
    const float* hprev; long ldhp;     // may be NULL => zeros
    float* dgi; long lddgi;            // [dr, dz, dn]
    float* dgh; long lddgh;            // [dr, dz, dn*r]
    float* dhprev; long lddhp;         // dh * z  (caller's GEMM then adds dgh * W_hh)
    float* dgi2; long lddgi2;          // optional: accumulate dgi over the sequence (broadcast term)
    const int* lengths; int t;
    int B, H;
    const int* nrows;                  // packed note level: DEVICE count of live rows (a prefix); nullptr = all B
    float* zero_out; long ldzo;        // optional (B,H) buffer to clear: the output of the split-K dgh.W_hh GEMM that follows
    uint16_t* dgh_b; long lddghb;      // optional bf16 copy of dgh (B,3H): the A operand of a bf16 dgh.W_hh GEMM
};

__global__ void __launch_bounds__(256) gru_gates_bwd_kernel(GateBwd a) {
    const int hq = a.H >> 2;
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)a.B * hq) return;
    const int b = (int)(idx / hq), j = (int)(idx % hq) * 4;
    if (a.nrows != nullptr && b >= *a.nrows) return;
    if (a.zero_out) *reinterpret_cast<float4*>(a.zero_out + (long)b * a.ldzo + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 d = a.dh ? *reinterpret_cast<const float4*>(a.dh + (long)b * a.lddh + j) : make_float4(0, 0, 0, 0);
    if (a.dh2) {
        float4 e = *reinterpret_cast<const float4*>(a.dh2 + (long)b * a.lddh2 + j);
        d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
    }
    if (a.dh3) {
        float4 e = *reinterpret_cast<const float4*>(a.dh3 + (long)b * a.lddh3 + j);
        d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
    }
    float* dgi = a.dgi + (long)b * a.lddgi + j;
    float* dgh = a.dgh + (long)b * a.lddgh + j;
    if (a.lengths && a.t >= a.lengths[b]) {
        float4 zero = make_float4(0, 0, 0, 0);
        if (a.dgh_b) {
            uint16_t* q = a.dgh_b + (long)b * a.lddghb + j;
            *reinterpret_cast<uint2*>(q) = make_uint2(0u, 0u); *reinterpret_cast<uint2*>(q + a.H) = make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(q + 2 * a.H) = make_uint2(0u, 0u);
        }
        *reinterpret_cast<float4*>(dgi) = zero; *reinterpret_cast<float4*>(dgi + a.H) = zero;
        *reinterpret_cast<float4*>(dgi + 2 * a.H) = zero;
        *reinterpret_cast<float4*>(dgh) = zero; *reinterpret_cast<float4*>(dgh + a.H) = zero;
        *reinterpret_cast<float4*>(dgh + 2 * a.H) = zero;
        *reinterpret_cast<float4*>(a.dhprev + (long)b * a.lddhp + j) = d;
        return;
    }
    const float* s = a.rzn + (long)b * a.ldrzn + j;
    float4 r = *reinterpret_cast<const float4*>(s), z = *reinterpret_cast<const float4*>(s + a.H),
           n = *reinterpret_cast<const float4*>(s + 2 * a.H);
    float4 hn = *reinterpret_cast<const float4*>(a.hn + (long)b * a.ldhn + j);
    float4 hp = a.hprev ? *reinterpret_cast<const float4*>(a.hprev + (long)b * a.ldhp + j)
                        : make_float4(0, 0, 0, 0);
    float4 dr, dz, dn, dnr, dp;
#define PD_GB(c)                                              \
    dn.c = d.c * (1.0f - z.c) * (1.0f - n.c * n.c);           \
    dz.c = d.c * (hp.c - n.c) * z.c * (1.0f - z.c);           \
    dnr.c = dn.c * r.c;                                       \
    dr.c = dn.c * hn.c * r.c * (1.0f - r.c);                  \
    dp.c = d.c * z.c;
    PD_GB(x) PD_GB(y) PD_GB(z) PD_GB(w)
#undef PD_GB
    *reinterpret_cast<float4*>(dgi) = dr; *reinterpret_cast<float4*>(dgi + a.H) = dz;
    *reinterpret_cast<float4*>(dgi + 2 * a.H) = dn;
    *reinterpret_cast<float4*>(dgh) = dr; *reinterpret_cast<float4*>(dgh + a.H) = dz;
    *reinterpret_cast<float4*>(dgh + 2 * a.H) = dnr;
    if (a.dgh_b) {
        auto pack = [](float4 v) {
            uint2 o;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(v.y), "f"(v.x));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(v.w), "f"(v.z));
            return o;
        };
        uint16_t* q = a.dgh_b + (long)b * a.lddghb + j;
        *reinterpret_cast<uint2*>(q) = pack(dr); *reinterpret_cast<uint2*>(q + a.H) = pack(dz);
        *reinterpret_cast<uint2*>(q + 2 * a.H) = pack(dnr);
    }
    *reinterpret_cast<float4*>(a.dhprev + (long)b * a.lddhp + j) = dp;
    if (a.dgi2) {
        float* q = a.dgi2 + (long)b * a.lddgi2 + j;
        float4 q0 = *reinterpret_cast<float4*>(q), q1 = *reinterpret_cast<float4*>(q + a.H),
               q2 = *reinterpret_cast<float4*>(q + 2 * a.H);
        q0.x += dr.x; q0.y += dr.y; q0.z += dr.z; q0.w += dr.w;
        q1.x += dz.x; q1.y += dz.y; q1.z += dz.z; q1.w += dz.w;
        q2.x += dn.x; q2.y += dn.y; q2.z += dn.z; q2.w += dn.w;
        *reinterpret_cast<float4*>(q) = q0; *reinterpret_cast<float4*>(q + a.H) = q1;
        *reinterpret_cast<float4*>(q + 2 * a.H) = q2;
    }
}

inline bool al4(const void* p, long ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }

}  // namespace

static int gates_fwd_launch(const float* gi, long ldgi, const float* gi2, long ldgi2, const float* gh, long ldgh,
                            const float* hprev, long ldhp, float* hout, long ldho, float* rzn, long ldrzn, float* hn,
                            long ldhn, const int* lengths, int t, int B, int H, float* h3, long ldh3, void* stream) {
    if (B <= 0) return 0;
    if ((H & 3) || !al4(gi, ldgi) || !al4(gh, ldgh) || !al4(hout, ldho) || (gi2 && !al4(gi2, ldgi2)) ||
        (hprev && !al4(hprev, ldhp)) || (rzn && !al4(rzn, ldrzn)) || (hn && !al4(hn, ldhn)) ||
        (h3 && (!al4(h3, ldh3) || ldh3 < 3L * H)))
        return PD_BAD_ARG;
    GateFwd a{gi, ldgi, gi2, ldgi2, gh, ldgh, hprev, ldhp, hout, ldho, rzn, ldrzn, hn, ldhn, lengths, t, B, H, h3, ldh3};
    long n = (long)B * (H >> 2);
    gru_gates_fwd_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(a);
    return pd_launch_status();
}

PD_API int pd_gru_gates_fwd(const float* gi, long ldgi, const float* gi2, long ldgi2, const float* gh, long ldgh,
                            const float* hprev, long ldhp, float* hout, long ldho, float* rzn, long ldrzn,
                            float* hn, long ldhn, const int* lengths, int t, int B, int H, void* stream) {
    return gates_fwd_launch(gi, ldgi, gi2, ldgi2, gh, ldgh, hprev, ldhp, hout, ldho, rzn, ldrzn, hn, ldhn, lengths, t, B, H,
                            nullptr, 0, stream);
}

// Inference variant: additionally writes h3 (B, 3H; row stride ldh3) = [hi | hi | lo] of the new state, the A operand
// of the single-launch 3xTF32 GEMMs that consume it next (saves a pd_tf32_split3 pass per step of the greedy decode).
PD_API int pd_gru_gates_fwd_split3(const float* gi, long ldgi, const float* gi2, long ldgi2, const float* gh, long ldgh,
                                   const float* hprev, long ldhp, float* hout, long ldho, const int* lengths, int t, int B,
                                   int H, float* h3, long ldh3, void* stream) {
    return gates_fwd_launch(gi, ldgi, gi2, ldgi2, gh, ldgh, hprev, ldhp, hout, ldho, nullptr, 0, nullptr, 0, lengths, t, B, H,
                            h3, ldh3, stream);
}

static int gates_bwd_launch(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3,
                            long lddh3, const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp, float* dgi,
                            long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, float* dgi2,
                            long lddgi2, const int* lengths, int t, int B, int H, const int* nrows, void* stream,
                            float* zero_out = nullptr, long ldzo = 0, void* dgh_b = nullptr, long lddghb = 0) {
    if (B <= 0) return 0;
    if (zero_out && !al4(zero_out, ldzo)) return PD_BAD_ARG;
    if (dgh_b && (((uintptr_t)dgh_b & 7) || (lddghb & 3))) return PD_BAD_ARG;
    if ((H & 3) || (dh && !al4(dh, lddh)) || (dh2 && !al4(dh2, lddh2)) || (dh3 && !al4(dh3, lddh3)) ||
        !al4(rzn, ldrzn) || !al4(hn, ldhn) ||
        (hprev && !al4(hprev, ldhp)) || !al4(dgi, lddgi) || !al4(dgh, lddgh) || !al4(dhprev, lddhp) ||
        (dgi2 && !al4(dgi2, lddgi2)))
        return PD_BAD_ARG;
    GateBwd a{dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hprev, ldhp, dgi, lddgi, dgh, lddgh,
              dhprev, lddhp, dgi2, lddgi2, lengths, t, B, H, nrows, zero_out, ldzo, (uint16_t*)dgh_b, lddghb};
    long n = (long)B * (H >> 2);
    gru_gates_bwd_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(a);
    return pd_launch_status();
}

PD_API int pd_gru_gates_bwd(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3,
                            long lddh3, const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp, float* dgi,
                            long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, float* dgi2,
                            long lddgi2, const int* lengths, int t, int B, int H, void* stream) {
    return gates_bwd_launch(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hprev, ldhp, dgi, lddgi, dgh, lddgh, dhprev,
                            lddhp, dgi2, lddgi2, lengths, t, B, H, nullptr, stream);
}

// Packed note level: gate gradients of the first *nrows rows only (DEVICE count; the other rows are left untouched)
PD_API int pd_gru_gates_bwd_rows(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                                 const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp,
                                 float* dgi, long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, int B, int H,
                                 const int* nrows, void* stream) {
    if (nrows == nullptr) return PD_BAD_ARG;
    return gates_bwd_launch(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hprev, ldhp, dgi, lddgi, dgh, lddgh, dhprev,
                            lddhp, nullptr, 0, nullptr, 0, B, H, nrows, stream);
}

// pd_gru_gates_bwd that also clears zero_out (B,H; row stride ldzo): the accumulator of the split-K dgh . W_hh GEMM of the
// same step, which then runs in accumulate mode without its own zero-fill node (batch-sized recurrences: one graph node
// less on each of their serial backward steps)
PD_API int pd_gru_gates_bwd_z(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                              const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp,
                              float* dgi, long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, const int* lengths,
                              int t, int B, int H, float* zero_out, long ldzo, void* stream) {
    return gates_bwd_launch(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hprev, ldhp, dgi, lddgi, dgh, lddgh, dhprev,
                            lddhp, nullptr, 0, lengths, t, B, H, nullptr, stream, zero_out, ldzo);
}

// pd_gru_gates_bwd_z that additionally writes dgh_b, a bf16 copy of dgh (B,3H; row stride lddghb elements): the A operand
// of the bf16 dgh . W_hh GEMM of the batch-sized recurrences (pd_gemm_bf16)
PD_API int pd_gru_gates_bwd_zb(const float* dh, long lddh, const float* dh2, long lddh2, const float* dh3, long lddh3,
                               const float* rzn, long ldrzn, const float* hn, long ldhn, const float* hprev, long ldhp,
                               float* dgi, long lddgi, float* dgh, long lddgh, float* dhprev, long lddhp, const int* lengths,
                               int t, int B, int H, float* zero_out, long ldzo, void* dgh_b, long lddghb, void* stream) {
    if (dgh_b == nullptr) return PD_BAD_ARG;
    return gates_bwd_launch(dh, lddh, dh2, lddh2, dh3, lddh3, rzn, ldrzn, hn, ldhn, hprev, ldhp, dgi, lddgi, dgh, lddgh, dhprev,
                            lddhp, nullptr, 0, lengths, t, B, H, nullptr, stream, zero_out, ldzo, dgh_b, lddghb);
}
