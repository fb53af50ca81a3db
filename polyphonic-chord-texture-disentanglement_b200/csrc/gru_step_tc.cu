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

// ---- persistent variant with TMA epilogue I/O (the fused step the training and decode paths run: ops.FUSED_GRU_STEP_TMA;
// validated on B200 in round 2, tests/test_gpu_kernels.py::test_gru_step_tma* and the whole-model parity tests) ---------
// Lessons applied: (1) the first fused kernel above lost to GEMM + gate kernel on its row-per-thread global I/O (seven
// arrays read / written as 64-byte pieces per lane) and its 2-stage, one-tile-per-CTA main loop; (2) the TMA bulk-store
// epilogue of the persistent GEMM took output-bound GEMMs to the HBM roofline.  Here: persistent CTAs, 4-stage
// TMA -> mbarrier ring, two TMEM accumulators (epilogue of item j overlaps the main loop of item j+1), and ALL
// epilogue traffic as 32-row x 16-column TMA boxes (64-byte swizzle): each epilogue warp TMA-loads gi (r|z|n), gi2
// (r|z|n) and h_prev of a 16-unit chunk into seven 2 KB buffers, computes the gates from TMEM + those buffers, writes
// r, z, n, h', W_hn h + b_hn back INTO the same buffers and hands them to TMA stores; the loads of the next chunk are
// issued once the stores have drained the buffers.  Unmasked steps only (lengths == NULL).
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct StepIo {
    const float* b_hh;
    int has_gi2, has_rzn, has_hn;
    int B, H;
    int KA;          // columns of the A operand: H (h_prev itself) or 3H ([hi | hi | lo] split of h_prev, 3xTF32 mode)
    int has_h3;      // also emit the [hi | hi | lo] split of the new state (operand of the next step's / the heads' GEMMs)
    int K2;          // SEG2 kernels: columns of the second A operand (the step's input x, e.g. the note embedding)
    const int* nrows;   // packed note level: DEVICE count of live rows of this step (rows are length-sorted, so they are a
                        // prefix; nullptr = all B).  Whole 128-row tiles up to the count are processed, the rest untouched.
    uint16_t* hb_out;   // bf16 variant: bf16 copy of the new state (row stride ldhb elements), the next step's A operand
    long ldhb;
};

constexpr int IOB = 2048;          // one epilogue buffer: 32 rows x 16 fp32 (64-byte rows, SWIZZLE_64B)
constexpr int N_IOB = 7;           // gi r,z,n | gi2 r,z,n | h_prev

// 16 floats of row `lane` of a 64-byte-swizzled 32 x 16 tile: 16-byte chunk j sits at j ^ ((row >> 1) & 3)
__device__ __forceinline__ void io_read(const uint8_t* buf, int lane, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 q = *reinterpret_cast<const float4*>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4));
        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
}
__device__ __forceinline__ void io_write(uint8_t* buf, int lane, const float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(buf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

constexpr int N_OUTB = 5;          // separate result buffers (OUTB variant): h' | r | z | n | W_hn h

// PRECISE: expf / tanhf gate math (the fp32-faithful greedy decode) instead of the MUFU forms (training).
// SEG2: the x-projection of the step is computed here too instead of being read from HBM: a second K segment
// A2 (B x K2: the step's input rows) . B2 (3H x K2: W_ih) follows the h segment through the same stage ring; its r / z
// products accumulate onto the h-projection's r / z columns, its n product goes to a fourth 64-column block (the GRU
// needs W_in x and W_hn h apart: n = tanh(i_n + r * h_n)).  Accumulator = 256 TMEM columns; no gi tensor exists.
// UNT: hidden units per tile (64; 32 gives the batch-sized recurrences -- 4 row tiles -- a full wave of CTAs)
// EB: bytes per element of the main-loop operands (4: fp32 multiplied as TF32; 2: bf16 copies of h_prev / W_hh,
// kind::f16 -- half the bytes and half the k-blocks of a step; accumulators, gate math and every epilogue array stay fp32)
template <int STAGES, int NSETS, bool OUTB, bool PRECISE, bool SEG2, int UNT = 64, int EB = 4>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gru_step_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmGi, const __grid_constant__ CUtensorMap tmGi2,
                    const __grid_constant__ CUtensorMap tmHp, const __grid_constant__ CUtensorMap tmHo,
                    const __grid_constant__ CUtensorMap tmRzn, const __grid_constant__ CUtensorMap tmHn,
                    const __grid_constant__ CUtensorMap tmH3, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB2, StepIo g, int tiles_m, int tiles_u) {
    constexpr int UN = UNT, BN3 = 3 * UNT;                             // (shadow the file-level tile constants)
    constexpr int ACC = SEG2 ? 4 * UN : BN3;                           // TMEM columns per accumulator
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int A_BYTES = BM * 128, B_BYTES = BN3 * 128;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint8_t* io = sB + STAGES * B_BYTES;                               // 4 warps x NSETS sets x 7 buffers x 2 KB
    uint8_t* io_out = io + 4 * NSETS * N_IOB * IOB;                    // OUTB: 4 warps x 5 result buffers x 2 KB
    uint64_t* full = (uint64_t*)(io_out + (OUTB ? 4 * N_OUTB * IOB : 0));
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;                              // [2]
    uint64_t* tmem_empty = tmem_full + 2;                              // [2]
    uint64_t* io_bar = tmem_empty + 2;                                 // [4][NSETS]: epilogue operands of a warp's set landed
    uint32_t* tmem_slot = (uint32_t*)(io_bar + 4 * NSETS);

    pd_grid_launch_dependents();        // (PDL launches: the next kernel of the chain may be scheduled while this one runs)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BKE = 128 / EB;                                      // elements per 128-byte k-block row
    static_assert(EB == 4 || !SEG2, "the folded x-projection is a TF32 path");
    const int nkb1 = (g.KA + BKE - 1) / BKE, nkb2 = SEG2 ? (g.K2 + 31) / 32 : 0;
    const int nkb = nkb1 + nkb2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        for (int q = 0; q < 4 * NSETS; ++q) mbar_init(&io_bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pd_grid_dependency_wait();          // prologue done; from here on the kernel touches data of its predecessors
    const uint32_t tmem_base = *tmem_slot;
    if (g.nrows != nullptr) {
        const int live = min(g.B, *g.nrows);
        tiles_m = (live + BM - 1) / BM;
    }
    const long n_items = (long)tiles_m * tiles_u;


    if (warp == 0) {
        if (lane == 0) {
            long cnt = 0;
            for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int m0 = (int)(item / tiles_u) * BM, u0 = (int)(item % tiles_u) * UN;
                for (int i = 0; i < nkb; ++i, ++cnt) {
                    const int s = (int)(cnt % STAGES), k0 = i * BKE;
                    if (cnt >= STAGES) mbar_wait(&empty[s], (uint32_t)((cnt / STAGES) - 1) & 1);
                    mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                    const bool seg2 = SEG2 && i >= nkb1;
                    const int kk = seg2 ? (i - nkb1) * 32 : k0;
                    tma_load_2d(seg2 ? &tmA2 : &tmA, &full[s], sA + s * A_BYTES, kk, m0);
#pragma unroll
                    for (int gate = 0; gate < 3; ++gate)      // r, z, n rows of the same 64 units
                        tma_load_2d(seg2 ? &tmB2 : &tmB, &full[s], sB + s * B_BYTES + gate * (UN * 128), kk, gate * g.H + u0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (Elem<EB>::FMT << 7) | (Elem<EB>::FMT << 10) | ((uint32_t)(BN3 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_rz = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * UN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            long cnt = 0, j = 0;
            for (long item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
                const int acc = (int)(j & 1);
                if (j >= 2) mbar_wait(&tmem_empty[acc], (uint32_t)((j >> 1) - 1) & 1);   // epilogue drained it
                tc_fence_after();
                for (int i = 0; i < nkb; ++i, ++cnt) {
                    const int s = (int)(cnt % STAGES);
                    mbar_wait(&full[s], (uint32_t)(cnt / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a = smem_u32(sA + s * A_BYTES), b = smem_u32(sB + s * B_BYTES);
                    if (SEG2 && i >= nkb1) {
                        // x segment: r | z onto the h-projection's columns, n into its own block (columns 192..255)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = make_desc(a + k * 32, 16, 1024, 2);
                            tc_mma_tf32(tmem_base + acc * ACC, ad, make_desc(b + k * 32, 16, 1024, 2), idesc_rz, 1u);
                            tc_mma_tf32(tmem_base + acc * ACC + 3 * UN, ad, make_desc(b + 2 * (UN * 128) + k * 32, 16, 1024, 2),
                                        idesc_n, (i > nkb1 || k > 0) ? 1u : 0u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma<EB>(tmem_base + acc * ACC, make_desc(a + k * 32, 16, 1024, 2), make_desc(b + k * 32, 16, 1024, 2),
                                       idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[s]);
                }
                tc_commit(&tmem_full[acc]);
            }
        }
    } else {
        const int q = warp & 3;                                        // TMEM lane quarter == rows q*32.. of the tile
        uint8_t* wbufs = io + q * (NSETS * N_IOB * IOB);
        const int H = g.H;
        constexpr int NCH = UN / 16;                                   // 16-unit chunks per item
        // This warp's chunk stream: chunk number k = j * NCH + c (j-th item of this CTA, chunk c).  Chunk k uses operand
        // set k % NSETS, whose mbarrier completes once per use: parity (k / NSETS) & 1.  Lane 0 is the producer of the
        // warp's own epilogue operands and keeps NSETS - 1 chunks of loads in flight behind the one being computed.
        const long n_mine = (n_items > (long)blockIdx.x) ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const long n_chunks = n_mine * NCH;
        auto issue_loads = [&](long k) {
            const long item = blockIdx.x + (k / NCH) * (long)gridDim.x;
            const int m0 = (int)(item / tiles_u) * BM, u0 = (int)(item % tiles_u) * UN;
            const int col = u0 + (int)(k % NCH) * 16, row = m0 + q * 32;
            const int set = (int)(k % NSETS);
            uint8_t* bufs = wbufs + set * (N_IOB * IOB);
            uint64_t* lbar = &io_bar[q * NSETS + set];
            mbar_expect_tx(lbar, (uint32_t)(((g.has_gi2 ? 3 : 0) + (SEG2 ? 0 : 3) + 1) * IOB));
            if (!SEG2) {
#pragma unroll
                for (int gate = 0; gate < 3; ++gate) tma_load_2d(&tmGi, lbar, bufs + gate * IOB, gate * H + col, row);
            }
            if (g.has_gi2) {
#pragma unroll
                for (int gate = 0; gate < 3; ++gate) tma_load_2d(&tmGi2, lbar, bufs + (3 + gate) * IOB, gate * H + col, row);
            }
            tma_load_2d(&tmHp, lbar, bufs + 6 * IOB, col, row);
        };
        if (lane == 0) {
            for (long k = 0; k < NSETS && k < n_chunks; ++k) issue_loads(k);
        }
        long k = 0;
        long j = 0;
        for (long item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
            const int m0 = (int)(item / tiles_u) * BM, u0 = (int)(item % tiles_u) * UN;
            const int acc = (int)(j & 1);
            mbar_wait(&tmem_full[acc], (uint32_t)(j >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++k) {
                const int col = u0 + c * 16, row = m0 + q * 32;
                const int set = (int)(k % NSETS);
                uint8_t* bufs = wbufs + set * (N_IOB * IOB);
                float ghr[16], ghz[16], ghn[16];
                const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC + c * 16;
                tc_ld16(tbase, ghr);
                tc_ld16(tbase + UN, ghz);
                tc_ld16(tbase + 2 * UN, ghn);
                float ir[16], iz[16], in[16], hp[16];
                if (SEG2) tc_ld16(tbase + 3 * UN, in);                 // W_in x of this chunk (r / z parts are already summed)
                if (c == NCH - 1) {                                    // last TMEM read of this item: hand the accumulator back
                    tc_fence_before();
                    if (lane == 0) mbar_arrive1(&tmem_empty[acc]);
                }
                mbar_wait(&io_bar[q * NSETS + set], (uint32_t)(k / NSETS) & 1);   // this chunk's gi / gi2 / h_prev boxes landed
                if (SEG2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) ir[i] = iz[i] = 0.0f;
                } else {
                    io_read(bufs, lane, ir);
                    io_read(bufs + IOB, lane, iz);
                    io_read(bufs + 2 * IOB, lane, in);
                }
                io_read(bufs + 6 * IOB, lane, hp);
                if (g.has_gi2) {
                    float t[16];
                    io_read(bufs + 3 * IOB, lane, t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) ir[i] += t[i];
                    io_read(bufs + 4 * IOB, lane, t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) iz[i] += t[i];
                    io_read(bufs + 5 * IOB, lane, t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) in[i] += t[i];
                }
                if (OUTB) {
                    // the operands are in registers: the set is free again -- start the next chunk's loads NOW, so their
                    // latency overlaps this chunk's gate math and stores (results leave through separate buffers)
                    __syncwarp();
                    if (lane == 0 && k + NSETS < n_chunks) issue_loads(k + NSETS);
                }
                add16(g.b_hh + col, ghr); add16(g.b_hh + H + col, ghz); add16(g.b_hh + 2 * H + col, ghn);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float r = PRECISE ? pd_sigmoid(ir[i] + ghr[i]) : pd_sigmoid_fast(ir[i] + ghr[i]);
                    const float z = PRECISE ? pd_sigmoid(iz[i] + ghz[i]) : pd_sigmoid_fast(iz[i] + ghz[i]);
                    const float n = PRECISE ? tanhf(in[i] + r * ghn[i]) : pd_tanh_fast(in[i] + r * ghn[i]);
                    hp[i] = (1.0f - z) * n + z * hp[i];
                    ir[i] = r; iz[i] = z; in[i] = n;
                }
                if (g.has_h3) {                                        // ir <- hi = rn_tf32(h'), iz <- lo = h' - hi
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t b;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(hp[i]));
                        ir[i] = __uint_as_float(b);
                        iz[i] = hp[i] - ir[i];
                    }
                }
                if (EB == 2 && g.hb_out != nullptr) {
                    // the next step's A operand: bf16 copy of h' (this thread: row `row + lane`, 16 columns = 32 bytes)
                    const long r_ = (long)row + lane;
                    if (r_ < g.B) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(hp[2 * i + 1]), "f"(hp[2 * i]));
                        uint4* dst = reinterpret_cast<uint4*>(g.hb_out + r_ * g.ldhb + col);
                        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                uint8_t* ob = OUTB ? io_out + q * (N_OUTB * IOB) : bufs;
                uint8_t* o_h = OUTB ? ob : bufs + 6 * IOB;
                uint8_t* o_r = OUTB ? ob + IOB : bufs;
                uint8_t* o_hn = OUTB ? ob + 4 * IOB : bufs + 3 * IOB;
                if (OUTB && k > 0) {                                   // the previous chunk's stores have read the result buffers
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
                // results leave through TMA stores (OUTB: from their own buffers; else from the buffers their operands came in)
                io_write(o_h, lane, hp);
                if (g.has_rzn) { io_write(o_r, lane, ir); io_write(o_r + IOB, lane, iz); io_write(o_r + 2 * IOB, lane, in); }
                if (g.has_hn) io_write(o_hn, lane, ghn);
                if (g.has_h3) { io_write(o_r, lane, ir); io_write(o_r + IOB, lane, iz); }      // (never together with rzn)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmHo), "r"(smem_u32(o_h)), "r"(col), "r"(row) : "memory");
                    if (g.has_rzn) {
#pragma unroll
                        for (int gate = 0; gate < 3; ++gate)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(&tmRzn), "r"(smem_u32(o_r + gate * IOB)), "r"(gate * H + col), "r"(row) : "memory");
                    }
                    if (g.has_hn)
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&tmHn), "r"(smem_u32(o_hn)), "r"(col), "r"(row) : "memory");
                    if (g.has_h3) {                                    // [hi | hi | lo] at column blocks 0, H, 2H
#pragma unroll
                        for (int part = 0; part < 3; ++part)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(&tmH3), "r"(smem_u32(o_r + (part == 2 ? IOB : 0))), "r"(part * H + col), "r"(row) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (!OUTB && k + NSETS < n_chunks) {
                        // this set is reloaded for chunk k + NSETS once its stores have read the buffers; meanwhile the
                        // loads of the NSETS - 1 chunks in between are already in flight
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        issue_loads(k + NSETS);
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// 2-D fp32 tensor [outer][inner] (row stride ld floats) as 16-column x 32-row boxes with the 64-byte swizzle:
// the epilogue operand / result tiles of gru_step_tma_kernel (plain FLOAT32: no TF32 rounding on the way)
int make_map_io(CUtensorMap* map, const void* base, long inner, long outer, long ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return 801;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 700 + (int)r;
}

int g_step_variant = 0;      // tuning switch of pd_gru_step_tma (pd_gru_step_tma_variant)

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
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru_step_tf32_kernel<STAGES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    gru_step_tf32_kernel<STAGES, 2><<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(ta, tb, g);
    return pd_launch_status();
}

// Persistent fused GRU step with TMA epilogue I/O (routed for recurrences of >= ops.FUSED_GRU_STEP_TMA_MIN_ROWS rows).
// Same contract as pd_gru_step_tf32 without the length mask; hout must not alias hprev.  gi: (B,3H) rows of the step
// (row stride ldgi), gi2 / rzn / hn optional.
PD_API int pd_gru_step_tma(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* b_hh, const float* gi,
                           long ldgi, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn, long ldrzn, float* hn,
                           long ldhn, int B, int H, void* stream) {
    if (B <= 0) return 0;
    if (H % UN != 0 || hprev == nullptr || hout == hprev) return PD_BAD_ARG;
    if (!al16(hprev, ldhp) || !al16(w_hh, ldw) || !al16(gi, ldgi) || (gi2 && !al16(gi2, ldgi2)) || !al16(hout, ldho) ||
        (rzn && !al16(rzn, ldrzn)) || (hn && !al16(hn, ldhn)) || ((uintptr_t)b_hh & 15) || ldhp < 4 || ldw < 4)
        return PD_BAD_ARG;
    CUtensorMap ta, tb, tgi, tgi2, thp, tho, trzn, thn;
    int rc = make_map(&ta, hprev, 4, H, B, ldhp, BM, false);
    if (!rc) rc = make_map(&tb, w_hh, 4, H, 3L * H, ldw, UN, false);
    if (!rc) rc = make_map_io(&tgi, gi, 3L * H, B, ldgi);
    if (!rc) rc = make_map_io(&thp, hprev, H, B, ldhp);
    if (!rc) rc = make_map_io(&tho, hout, H, B, ldho);
    tgi2 = tgi; trzn = tho; thn = tho;
    if (!rc && gi2) rc = make_map_io(&tgi2, gi2, 3L * H, B, ldgi2);
    if (!rc && rzn) rc = make_map_io(&trzn, rzn, 3L * H, B, ldrzn);
    if (!rc && hn) rc = make_map_io(&thn, hn, H, B, ldhn);
    if (rc) return rc;
    StepIo g{b_hh, gi2 != nullptr, rzn != nullptr, hn != nullptr, B, H, H, 0, 0, nullptr, nullptr, 0};
    const CUtensorMap th3 = tho;
    constexpr bool kPrecise = false;
    int tiles_m = (B + BM - 1) / BM, tiles_u = H / UN;
    long items = (long)tiles_m * tiles_u;
    if (g_step_variant == 0 && items < PD_NUM_SMS && H % 32 == 0) {
        // batch-sized recurrence (e.g. 512 x 1024: 64 tiles of 64 units): 32-unit tiles fill the machine in one wave and
        // leave room for a 4-stage ring
        rc = make_map(&tb, w_hh, 4, H, 3L * H, ldw, 32, false);
        if (rc) return rc;
        tiles_u = H / 32;
        items = (long)tiles_m * tiles_u;
        const int grid32 = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
        constexpr int ST = 4, NS = 1;
        constexpr int smem = ST * (BM * 128 + 96 * 128) + 4 * NS * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
        static unsigned long long attr32 = 0;
        if (pd_first_use_on_device(attr32)) {
            cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, true, kPrecise, false, 32>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
        }
        { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, true, kPrecise, false, 32>, dim3(grid32), dim3(NUM_THREADS), (size_t)smem, (cudaStream_t)stream,
            ta, tb, tgi, tgi2, thp, tho, trzn, thn, th3, ta, tb, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
        return pd_launch_status();
    }
    const int grid = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
    // variant 0 (default): 3-stage main loop, one operand set per epilogue warp reloaded as soon as its values are in
    // registers, results through separate buffers; variant 1: round-1 layout (4 stages, results written back into the
    // operand buffers); variant 2: 2 stages, double-buffered operand sets.  B200, 16384 x 512: v1 109 us, v2 125 us.
#define PD_STEP_LAUNCH(ST, NS, OB)                                                                                          \
    {                                                                                                                       \
        constexpr int smem = ST * (BM * 128 + BN3 * 128) + 4 * NS * N_IOB * IOB + (OB ? 4 * N_OUTB * IOB : 0) + 1024 + 256; \
        static unsigned long long attr = 0;                                                                                           \
        if (pd_first_use_on_device(attr)) {                                                                                                        \
            cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, OB, kPrecise, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            if (e != cudaSuccess) return (int)e;                                                                            \
        }                                                                                                                   \
        {                                                                                                                   \
            cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, OB, kPrecise, false>, dim3(grid), dim3(NUM_THREADS),     \
                                           (size_t)smem, (cudaStream_t)stream, ta, tb, tgi, tgi2, thp, tho, trzn, thn, th3, ta, \
                                           tb, g, tiles_m, tiles_u);                                                        \
            if (le != cudaSuccess) return (int)le;                                                                          \
        }                                                                                                                   \
        return pd_launch_status();                                                                                          \
    }
    if (g_step_variant == 1) PD_STEP_LAUNCH(4, 1, false)
    if (g_step_variant == 2) PD_STEP_LAUNCH(2, 2, false)
    PD_STEP_LAUNCH(3, 1, true)
#undef PD_STEP_LAUNCH
}

// bf16-operand form for the batch-sized recurrences (time GRU, encoder bi-GRUs, chord decoder; BASELINE configs[1]
// "bf16 / fp32-accumulate"): the A operand is a bf16 copy of h_prev (hb_prev: B x H, row stride ldhbp ELEMENTS), W a bf16
// copy of W_hh (3H x H) -- half the bytes and half the k-blocks of the latency-bound main loop -- while the accumulators,
// the gate math, h_prev in the blend and every saved array stay fp32.  Besides hout the kernel writes hb_out, the bf16
// copy of the new state that the next step multiplies.  32-unit tiles; H % 64 == 0.
// units: hidden units per tile, 32 (a 512-row recurrence fills the machine with one wave of 128 CTAs: the fastest single
// step) or 64 (64 CTAs per step: two independent recurrences -- the directions of a bi-GRU, the two encoders -- then run
// side by side on disjoint SMs instead of queueing behind each other's full wave).
PD_API int pd_gru_step_tma_bf16_units(const void* hb_prev, long ldhbp, const void* wb, long ldwb, const float* b_hh,
                                      const float* gi, long ldgi, const float* gi2, long ldgi2, const float* hprev, long ldhp,
                                      float* hout, long ldho, void* hb_out, long ldhbo, float* rzn, long ldrzn, float* hn,
                                      long ldhn, int B, int H, int units, void* stream) {
    if (B <= 0) return 0;
    if (units != 32 && units != 64) return PD_BAD_ARG;
    if (H % 64 != 0 || hb_prev == nullptr || hprev == nullptr || hout == hprev || hb_out == hb_prev || hb_out == nullptr)
        return PD_BAD_ARG;
    if (!al16(hprev, ldhp) || !al16(gi, ldgi) || (gi2 && !al16(gi2, ldgi2)) || !al16(hout, ldho) || (rzn && !al16(rzn, ldrzn)) ||
        (hn && !al16(hn, ldhn)) || ((uintptr_t)b_hh & 15) || ((uintptr_t)hb_prev & 15) || ((uintptr_t)wb & 15) ||
        ((uintptr_t)hb_out & 31) || (ldhbp & 7) || (ldwb & 7) || (ldhbo & 15))
        return PD_BAD_ARG;
    CUtensorMap ta, tb, tgi, tgi2, thp, tho, trzn, thn;
    int rc = make_map(&ta, hb_prev, 2, H, B, ldhbp, BM, false);
    if (!rc) rc = make_map(&tb, wb, 2, H, 3L * H, ldwb, units, false);
    if (!rc) rc = make_map_io(&tgi, gi, 3L * H, B, ldgi);
    if (!rc) rc = make_map_io(&thp, hprev, H, B, ldhp);
    if (!rc) rc = make_map_io(&tho, hout, H, B, ldho);
    tgi2 = tgi; trzn = tho; thn = tho;
    if (!rc && gi2) rc = make_map_io(&tgi2, gi2, 3L * H, B, ldgi2);
    if (!rc && rzn) rc = make_map_io(&trzn, rzn, 3L * H, B, ldrzn);
    if (!rc && hn) rc = make_map_io(&thn, hn, H, B, ldhn);
    if (rc) return rc;
    StepIo g{b_hh, gi2 != nullptr, rzn != nullptr, hn != nullptr, B, H, H, 0, 0, nullptr, (uint16_t*)hb_out, ldhbo};
    const int tiles_m = (B + BM - 1) / BM, tiles_u = H / units;
    const long items = (long)tiles_m * tiles_u;
    const int grid = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
    if (units == 64) {
        constexpr int ST = 3, NS = 1;
        constexpr int smem = ST * (BM * 128 + 192 * 128) + 4 * NS * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
        static unsigned long long attr64 = 0;
        if (pd_first_use_on_device(attr64)) {
            cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, true, false, false, 64, 2>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
        }
        { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, true, false, false, 64, 2>, dim3(grid), dim3(NUM_THREADS), (size_t)smem, (cudaStream_t)stream,
                ta, tb, tgi, tgi2, thp, tho, trzn, thn, tho, ta, tb, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
        return pd_launch_status();
    }
    constexpr int ST = 4, NS = 1;
    constexpr int smem = ST * (BM * 128 + 96 * 128) + 4 * NS * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, true, false, false, 32, 2>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, true, false, false, 32, 2>, dim3(grid), dim3(NUM_THREADS), (size_t)smem, (cudaStream_t)stream,
            ta, tb, tgi, tgi2, thp, tho, trzn, thn, tho, ta, tb, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
    return pd_launch_status();
}

PD_API int pd_gru_step_tma_bf16(const void* hb_prev, long ldhbp, const void* wb, long ldwb, const float* b_hh, const float* gi,
                                long ldgi, const float* gi2, long ldgi2, const float* hprev, long ldhp, float* hout, long ldho,
                                void* hb_out, long ldhbo, float* rzn, long ldrzn, float* hn, long ldhn, int B, int H,
                                void* stream) {
    return pd_gru_step_tma_bf16_units(hb_prev, ldhbp, wb, ldwb, b_hh, gi, ldgi, gi2, ldgi2, hprev, ldhp, hout, ldho, hb_out,
                                      ldhbo, rzn, ldrzn, hn, ldhn, B, H, 32, stream);
}

// Training form with the x-projection folded in (SEG2, TF32 single pass): gi = W_x x is computed inside the kernel as a
// second K segment (x: B x K2 rows of the step's input, w_x: 3H x K2), so the (B,T,3H) x-projection of the sequence is never
// materialised.  gi2 (B,3H) carries the rest of the input projection incl. b_ih.  Teacher-forced note GRU (ptvae.py:396-398):
// x = ground-truth note embedding of slot n (row stride 16*128), w_x = dec_notes_gru.weight_ih[:, 1024:].
static int gru_step_tmax_impl(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* x, long ldx, const float* w_x,
                              long ldwx, int K2, const float* b_hh, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn,
                              long ldrzn, float* hn, long ldhn, int B, int H, const int* nrows, void* stream) {
    if (B <= 0) return 0;
    if (H % UN != 0 || hprev == nullptr || hout == hprev || x == nullptr || gi2 == nullptr || K2 <= 0 || (K2 & 3)) return PD_BAD_ARG;
    if (!al16(hprev, ldhp) || !al16(w_hh, ldw) || !al16(x, ldx) || !al16(w_x, ldwx) || !al16(gi2, ldgi2) || !al16(hout, ldho) ||
        (rzn && !al16(rzn, ldrzn)) || (hn && !al16(hn, ldhn)) || ((uintptr_t)b_hh & 15) || ldhp < 4 || ldw < 4 || ldx < 4 || ldwx < 4)
        return PD_BAD_ARG;
    CUtensorMap ta, tb, ta2, tb2, tgi2, thp, tho, trzn, thn;
    int rc = make_map(&ta, hprev, 4, H, B, ldhp, BM, false);
    if (!rc) rc = make_map(&tb, w_hh, 4, H, 3L * H, ldw, UN, false);
    if (!rc) rc = make_map(&ta2, x, 4, K2, B, ldx, BM, false);
    if (!rc) rc = make_map(&tb2, w_x, 4, K2, 3L * H, ldwx, UN, false);
    if (!rc) rc = make_map_io(&tgi2, gi2, 3L * H, B, ldgi2);
    if (!rc) rc = make_map_io(&thp, hprev, H, B, ldhp);
    if (!rc) rc = make_map_io(&tho, hout, H, B, ldho);
    trzn = tho; thn = tho;
    if (!rc && rzn) rc = make_map_io(&trzn, rzn, 3L * H, B, ldrzn);
    if (!rc && hn) rc = make_map_io(&thn, hn, H, B, ldhn);
    if (rc) return rc;
    StepIo g{b_hh, 1, rzn != nullptr, hn != nullptr, B, H, H, 0, K2, nrows, nullptr, 0};
    int tiles_m = (B + BM - 1) / BM, tiles_u = H / UN;
    long items = (long)tiles_m * tiles_u;
    if (nrows == nullptr && 2 * items <= PD_NUM_SMS && H % 32 == 0) {
        // a few hundred rows (a note slot of the greedy pass at batch 512: 32 tiles of 64 units on 148 SMs): 32-unit tiles
        // double the CTAs and halve each CTA's weight stream and epilogue -- the step is a latency chain, not a throughput job
        rc = make_map(&tb, w_hh, 4, H, 3L * H, ldw, 32, false);
        if (!rc) rc = make_map(&tb2, w_x, 4, K2, 3L * H, ldwx, 32, false);
        if (rc) return rc;
        tiles_u = H / 32;
        items = (long)tiles_m * tiles_u;
        const int grid32 = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
        constexpr int ST32 = 4, NS32 = 1;
        constexpr int smem32 = ST32 * (BM * 128 + 96 * 128) + 4 * NS32 * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
        static unsigned long long attr32 = 0;
        if (pd_first_use_on_device(attr32)) {
            cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST32, NS32, true, false, true, 32>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
            if (e != cudaSuccess) return (int)e;
        }
        { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST32, NS32, true, false, true, 32>, dim3(grid32), dim3(NUM_THREADS), (size_t)smem32, (cudaStream_t)stream,
                ta, tb, tgi2, tgi2, thp, tho, trzn, thn, tho, ta2, tb2, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
        return pd_launch_status();
    }
    const int grid = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
    constexpr int ST = 3, NS = 1;
    constexpr int smem = ST * (BM * 128 + BN3 * 128) + 4 * NS * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, true, false, true>, dim3(grid), dim3(NUM_THREADS), (size_t)smem, (cudaStream_t)stream,
            ta, tb, tgi2, tgi2, thp, tho, trzn, thn, tho, ta2, tb2, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
    return pd_launch_status();
}

PD_API int pd_gru_step_tmax(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* x, long ldx, const float* w_x,
                            long ldwx, int K2, const float* b_hh, const float* gi2, long ldgi2, float* hout, long ldho, float* rzn,
                            long ldrzn, float* hn, long ldhn, int B, int H, void* stream) {
    return gru_step_tmax_impl(hprev, ldhp, w_hh, ldw, x, ldx, w_x, ldwx, K2, b_hh, gi2, ldgi2, hout, ldho, rzn, ldrzn, hn, ldhn, B, H,
                              nullptr, stream);
}

// Packed note level: the same step over the first *nrows rows only (a DEVICE count; rows are sorted by note count, so the
// live rows of note slot n are a prefix).  Rows beyond the last touched 128-row tile are neither read nor written.
PD_API int pd_gru_step_tmax_rows(const float* hprev, long ldhp, const float* w_hh, long ldw, const float* x, long ldx,
                                 const float* w_x, long ldwx, int K2, const float* b_hh, const float* gi2, long ldgi2, float* hout,
                                 long ldho, float* rzn, long ldrzn, float* hn, long ldhn, int B, int H, const int* nrows,
                                 void* stream) {
    if (nrows == nullptr) return PD_BAD_ARG;
    return gru_step_tmax_impl(hprev, ldhp, w_hh, ldw, x, ldx, w_x, ldwx, K2, b_hh, gi2, ldgi2, hout, ldho, rzn, ldrzn, hn, ldhn, B, H,
                              nrows, stream);
}

// Inference form of the fused step for the error-compensated 3xTF32 path (greedy decode at >= 512 rows): the A operand
// is the [hi | hi | lo] split of h_prev (a3: B x 3H), W3 the [hi | lo | hi] split of W_hh (3H x 3H), so the three TF32
// products accumulate in TMEM (fp32-class h-projection); gate math with expf / tanhf; the epilogue writes the new state
// (hout, may alias hprev) AND its [hi | hi | lo] split (h3out: B x 3H, must not alias a3) -- the operand of the next
// step's and the heads' GEMMs.  Replaces pd_gemm_tf32 (K = 3H) + pd_gru_gates_fwd_split3 per note slot.
PD_API int pd_gru_step_tma3(const float* a3, long lda3, const float* w3, long ldw3, const float* b_hh, const float* gi,
                            long ldgi, const float* gi2, long ldgi2, const float* hprev, long ldhp, float* hout, long ldho,
                            float* h3out, long ldh3, int B, int H, void* stream) {
    if (B <= 0) return 0;
    if (H % UN != 0 || a3 == nullptr || hprev == nullptr || h3out == nullptr || h3out == a3 || gi == nullptr) return PD_BAD_ARG;
    if (!al16(a3, lda3) || !al16(w3, ldw3) || !al16(gi, ldgi) || (gi2 && !al16(gi2, ldgi2)) || !al16(hprev, ldhp) ||
        !al16(hout, ldho) || !al16(h3out, ldh3) || ((uintptr_t)b_hh & 15) || lda3 < 4 || ldw3 < 4)
        return PD_BAD_ARG;
    CUtensorMap ta, tb, tgi, tgi2, thp, tho, trzn, thn, th3;
    int rc = make_map(&ta, a3, 4, 3L * H, B, lda3, BM, false);
    if (!rc) rc = make_map(&tb, w3, 4, 3L * H, 3L * H, ldw3, UN, false);
    if (!rc) rc = make_map_io(&tgi, gi, 3L * H, B, ldgi);
    if (!rc) rc = make_map_io(&thp, hprev, H, B, ldhp);
    if (!rc) rc = make_map_io(&tho, hout, H, B, ldho);
    if (!rc) rc = make_map_io(&th3, h3out, 3L * H, B, ldh3);
    tgi2 = tgi; trzn = tho; thn = tho;
    if (!rc && gi2) rc = make_map_io(&tgi2, gi2, 3L * H, B, ldgi2);
    if (rc) return rc;
    StepIo g{b_hh, gi2 != nullptr, 0, 0, B, H, 3 * H, 1, 0, nullptr, nullptr, 0};
    const int tiles_m = (B + BM - 1) / BM, tiles_u = H / UN;
    const long items = (long)tiles_m * tiles_u;
    const int grid = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
    constexpr int ST = 3, NS = 1;
    constexpr int smem = ST * (BM * 128 + BN3 * 128) + 4 * NS * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, true, true, false>, dim3(grid), dim3(NUM_THREADS), (size_t)smem, (cudaStream_t)stream,
            ta, tb, tgi, tgi2, thp, tho, trzn, thn, th3, ta, tb, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
    return pd_launch_status();
}

// Same step with the x-projection folded in (SEG2): instead of reading gi = W_ih x + b_ih from HBM, the kernel multiplies the
// step's input rows itself -- x3 (B, K2) = [hi | hi | lo] split of x (K2 = 3 * pad4(in_features)), wx3 (3H, K2) = [hi | lo | hi]
// split of W_ih's columns for x -- as a second K segment of the same tcgen05 main loop.  gi2 (B,3H) must carry everything
// else of the input projection (the sequence-constant part INCLUDING b_ih).  Note-GRU slot of the greedy decode:
// x = embedding of the previous token, gi2 = W_ih[:, :1024] summary + b_ih.  Removes one GEMM launch and the write + read
// of the (B,3H) x-projection per slot.
PD_API int pd_gru_step_tma3x(const float* a3, long lda3, const float* w3, long ldw3, const float* x3, long ldx3, const float* wx3,
                             long ldwx3, int K2, const float* b_hh, const float* gi2, long ldgi2, const float* hprev, long ldhp,
                             float* hout, long ldho, float* h3out, long ldh3, int B, int H, void* stream) {
    if (B <= 0) return 0;
    if (H % UN != 0 || a3 == nullptr || x3 == nullptr || hprev == nullptr || h3out == nullptr || h3out == a3 || gi2 == nullptr ||
        K2 <= 0 || (K2 & 3))
        return PD_BAD_ARG;
    if (!al16(a3, lda3) || !al16(w3, ldw3) || !al16(x3, ldx3) || !al16(wx3, ldwx3) || !al16(gi2, ldgi2) || !al16(hprev, ldhp) ||
        !al16(hout, ldho) || !al16(h3out, ldh3) || ((uintptr_t)b_hh & 15) || lda3 < 4 || ldw3 < 4 || ldx3 < 4 || ldwx3 < 4)
        return PD_BAD_ARG;
    CUtensorMap ta, tb, ta2, tb2, tgi2, thp, tho, th3;
    int rc = make_map(&ta, a3, 4, 3L * H, B, lda3, BM, false);
    if (!rc) rc = make_map(&tb, w3, 4, 3L * H, 3L * H, ldw3, UN, false);
    if (!rc) rc = make_map(&ta2, x3, 4, K2, B, ldx3, BM, false);
    if (!rc) rc = make_map(&tb2, wx3, 4, K2, 3L * H, ldwx3, UN, false);
    if (!rc) rc = make_map_io(&tgi2, gi2, 3L * H, B, ldgi2);
    if (!rc) rc = make_map_io(&thp, hprev, H, B, ldhp);
    if (!rc) rc = make_map_io(&tho, hout, H, B, ldho);
    if (!rc) rc = make_map_io(&th3, h3out, 3L * H, B, ldh3);
    if (rc) return rc;
    StepIo g{b_hh, 1, 0, 0, B, H, 3 * H, 1, K2, nullptr, nullptr, 0};
    const int tiles_m = (B + BM - 1) / BM, tiles_u = H / UN;
    const long items = (long)tiles_m * tiles_u;
    const int grid = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
    constexpr int ST = 3, NS = 1;
    constexpr int smem = ST * (BM * 128 + BN3 * 128) + 4 * NS * N_IOB * IOB + 4 * N_OUTB * IOB + 1024 + 256;
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gru_step_tma_kernel<ST, NS, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    { cudaError_t le = pd_launch_pdl(gru_step_tma_kernel<ST, NS, true, true, true>, dim3(grid), dim3(NUM_THREADS), (size_t)smem, (cudaStream_t)stream,
            ta, tb, tgi2, tgi2, thp, tho, tho, tho, th3, ta2, tb2, g, tiles_m, tiles_u); if (le != cudaSuccess) return (int)le; }
    return pd_launch_status();
}

// tuning / A-B switch for pd_gru_step_tma (see the variants above)
PD_API int pd_gru_step_tma_variant(int v) {
    if (v < 0 || v > 2) return PD_BAD_ARG;
    g_step_variant = v;
    return 0;
}
