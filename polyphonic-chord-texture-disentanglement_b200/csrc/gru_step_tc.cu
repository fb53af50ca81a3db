// Fused GRU step on the tensor cores: h' = GRUCell(gi, h) with the recurrent projection W_hh h computed by
// tcgen05.mma into TMEM and the gate math done in the epilogue, straight out of TMEM -- the (B,3H)
// h-projection never goes to HBM.  tools/gemm_dissect.py showed the per-step recurrent GEMMs are bound by
// their fp32 OUTPUT (100 MB per note-GRU step) which the gate kernel then reads straight back; this kernel
// removes both passes and one launch per step.
//
// Tile: 128 rows x 64 hidden units.  The B operand of a tile is three 64-row boxes of W_hh (rows u0.., H+u0..,
// 2H+u0..: the r, z and n gates of the same units) landing contiguously in one stage, so ONE M128 x N192
// accumulator holds gh_r | gh_z | gh_n of those units and the thread that owns a row (tcgen05.ld 32x32b) has
// all three pre-activations of every unit it needs.  Epilogue per 16-unit chunk: TMEM -> registers, add b_hh,
// read the row's gi (+ the sequence-constant gi2) and h_prev, apply the PyTorch gate equations, write h' and
// (training) the saved r|z|n and W_hn h + b_hn for the backward pass.
// Replaces aten::gru steps at ptvae.py:63-65, :396-398, :461-462 and inside the packed bi-GRUs (:446-453).
#include "tc_common.cuh"

namespace {

constexpr int UN = 64;            // hidden units per tile
constexpr int BN3 = 3 * UN;       // accumulator columns

struct StepArgs {
    const float* b_hh;
    const float* gi; long ldgi;
    const float* gi2; long ldgi2;
    const float* hprev; long ldhp;
    float* hout; long ldho;
    float* rzn; long ldrzn;
    float* hn; long ldhn;
    const int* lengths; int t;
    int B, H;
};

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
          "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ void ld16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        float4 q = *reinterpret_cast<const float4*>(p + i);
        v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
}
__device__ __forceinline__ void add16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        float4 q = *reinterpret_cast<const float4*>(p + i);
        v[i] += q.x; v[i + 1] += q.y; v[i + 2] += q.z; v[i + 3] += q.w;
    }
}
__device__ __forceinline__ void st16(float* p, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

template <int STAGES, int MINB>
__global__ void __launch_bounds__(NUM_THREADS, MINB)
gru_step_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, StepArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int A_BYTES = BM * 128, B_BYTES = BN3 * 128;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = (uint64_t*)(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, u0 = blockIdx.y * UN;
    const int nkb = (g.H + 31) / 32;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, k0 = i * 32;
                if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                tma_load_2d(&tmA, &full[s], sA + s * A_BYTES, k0, m0);
#pragma unroll
                for (int gate = 0; gate < 3; ++gate)      // r, z, n rows of the same 64 units
                    tma_load_2d(&tmB, &full[s], sB + s * B_BYTES + gate * (UN * 128), k0, gate * g.H + u0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN3 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint32_t a = smem_u32(sA + s * A_BYTES), b = smem_u32(sB + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_tf32(tmem_base, make_desc(a + k * 32, 16, 1024, 2), make_desc(b + k * 32, 16, 1024, 2), idesc,
                                (i > 0 || k > 0) ? 1u : 0u);
                tc_commit(&empty[s]);
            }
            tc_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const bool row_ok = m < g.B;
        const bool masked = row_ok && g.lengths && g.t >= g.lengths[m];
        const int H = g.H;
#pragma unroll 1
        for (int c = 0; c < UN / 16; ++c) {
            const int u = u0 + c * 16;
            if (u >= H) break;
            float ghr[16], ghz[16], ghn[16];
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + c * 16;
            tc_ld16(tbase, ghr);                 // all lanes take part in the TMEM loads (warp-collective)
            tc_ld16(tbase + UN, ghz);
            tc_ld16(tbase + 2 * UN, ghn);
            if (!row_ok) continue;
            float hp[16];
            ld16(g.hprev + (long)m * g.ldhp + u, hp);
            if (masked) {                        // past the end of this sequence: carry the state
                st16(g.hout + (long)m * g.ldho + u, hp);
                continue;
            }
            float ir[16], iz[16], in[16];
            const float* gi = g.gi + (long)m * g.ldgi + u;
            ld16(gi, ir); ld16(gi + H, iz); ld16(gi + 2 * H, in);
            if (g.gi2) {
                const float* g2 = g.gi2 + (long)m * g.ldgi2 + u;
                add16(g2, ir); add16(g2 + H, iz); add16(g2 + 2 * H, in);
            }
            add16(g.b_hh + u, ghr); add16(g.b_hh + H + u, ghz); add16(g.b_hh + 2 * H + u, ghn);
            float ho[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float r = pd_sigmoid(ir[i] + ghr[i]);
                const float z = pd_sigmoid(iz[i] + ghz[i]);
                const float n = tanhf(in[i] + r * ghn[i]);
                ho[i] = (1.0f - z) * n + z * hp[i];
                ir[i] = r; iz[i] = z; in[i] = n;
            }
            st16(g.hout + (long)m * g.ldho + u, ho);
            if (g.rzn) {
                float* s = g.rzn + (long)m * g.ldrzn + u;
                st16(s, ir); st16(s + H, iz); st16(s + 2 * H, in);
            }
            if (g.hn) st16(g.hn + (long)m * g.ldhn + u, ghn);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
    }
}

inline bool al16(const void* p, long ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }

}  // namespace

// hout (B,H) = GRU cell update of hprev (B,H) given gi (B,3H) [+ gi2] and W_hh (3H,H; row stride ldw), b_hh (3H):
// the W_hh h GEMM (TF32 tensor cores) and the gate math in one kernel.  rzn / hn as in pd_gru_gates_fwd.
// hout may alias hprev only if rzn/hn consumers do not need hprev (inference).  H must be a multiple of 64;
// all row strides multiples of 4 floats, bases 16-byte aligned; otherwise PD_BAD_ARG (use GEMM + gates).
PD_API int pd_gru_step_tf32(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* b_hh,
                            const float* gi, long ldgi, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn,
                            long ldrzn, float* hn, long ldhn, const int* lengths, int t, int B, int H, void* stream) {
    if (B <= 0) return 0;
    if (H % UN != 0 || hprev == nullptr || hout == hprev) return PD_BAD_ARG;
    if (!al16(hprev, ldhp) || !al16(w_hh, ldw) || !al16(gi, ldgi) || (gi2 && !al16(gi2, ldgi2)) || !al16(hout, ldho) ||
        (rzn && !al16(rzn, ldrzn)) || (hn && !al16(hn, ldhn)) || ((uintptr_t)b_hh & 15) || ldhp < 4 || ldw < 4)
        return PD_BAD_ARG;
    CUtensorMap ta, tb;
    int rc = make_map(&ta, hprev, 4, H, B, ldhp, BM, false);
    if (rc) return rc;
    rc = make_map(&tb, w_hh, 4, H, 3L * H, ldw, UN, false);
    if (rc) return rc;
    StepArgs g{b_hh, gi, ldgi, gi2, ldgi2, hprev, ldhp, hout, ldho, rzn, ldrzn, hn, ldhn, lengths, t, B, H};
    dim3 grid((B + BM - 1) / BM, H / UN);
    constexpr int STAGES = 2;
    constexpr int smem = STAGES * (BM * 128 + BN3 * 128) + 1024 + 256;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gru_step_tf32_kernel<STAGES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    gru_step_tf32_kernel<STAGES, 2><<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(ta, tb, g);
    return pd_launch_status();
}
