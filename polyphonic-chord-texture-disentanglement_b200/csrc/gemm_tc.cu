// Tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32, fp32 operands straight from HBM, fp32
// accumulators in TMEM), operands staged by TMA through a 4-stage mbarrier ring, warp-specialised
// (1 TMA warp, 1 MMA-issuing warp, 4 epilogue warps reading TMEM with tcgen05.ld).
//
//   C[m, n] (+)= sum_k A(m,k) * B(k,n) (+ bias[n])         fp32 in HBM, TF32 multiply, fp32 accumulate
//
// Operand layouts (the three GEMMs of a Linear / GRU layer, no transposed copies anywhere):
//   A K-major : A[m*lda + k]   (activations as GEMM rows)       A MN-major: A[k*lda + m]  (dY^T for dW)
//   B K-major : B[n*ldb + k]   (nn.Linear weight [N,K])         B MN-major: B[k*ldb + n]  (W for dX, X for dW)
// K-major tiles are TMA boxes of 32 fp32 (128 B) x rows with the 128-byte swizzle; MN-major tiles are
// boxes of 32 fp32 along M/N x 32 k-rows with the 32-byte-atom 128-byte swizzle
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA layout SWIZZLE_128B_BASE32B), the only MN-major
// layout tcgen05 accepts for 4-byte operands.
//
// CTA tile 128 x BN (BN = 64/128/256, one tcgen05.mma M=128 N=BN K=8 per 32-byte k-slice), BK = 32.
// Split-K over blockIdx.z with a red.global.add epilogue for the weight-gradient GEMMs, whose M x N is
// tiny and whose K is the whole batch.
#include "tc_common.cuh"

namespace {

struct TcArgs {
    float* C; long ldc;
    const float* bias;
    int M, N, K;
    int kb_per_split;   // k-blocks (of BK) per blockIdx.z
    int atomic;         // accumulate into C (C += result, or split-K partial sums): red.global.add epilogue
    int dbg;            // tuning only: 1 = skip the epilogue stores, 2 = skip the MMAs, 3 = both
    int tma_store;      // persistent kernel: write C tiles with TMA bulk stores (plain stores, aligned C)
    PdRows rows;        // packed note level: live-row predicate (common.cuh) ...
    int pred;           // ... 0 = none, 1 = on the rows of A / C (dead M tiles are skipped), 2 = on K (dead k-blocks are skipped)
};

// MH = number of 128-row halves per CTA tile (1 or 2).  The per-step GEMMs of this path are L2->SM bandwidth
// bound with fp32 operands (a 128x256 tile needs 96 B/clk against ~43 B/clk/SM of L2 fabric); MH = 2 reuses
// every B (weight) stage for two accumulators in TMEM and cuts the traffic per MAC by a third.
template <int EB, int MH, int BN, int STAGES, int MINB, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, MINB)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int BKE = Elem<EB>::BKE, CW = 128 / EB;   // k-block elements; MN-major chunk width (elements per 128 B)
    constexpr int A_BYTES = MH * BM * 128, B_BYTES = BN * 128;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = (uint64_t*)(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * (MH * BM), n0 = blockIdx.y * BN;
    if (g.pred == 1 && !pd_rows_live(g.rows, m0, MH * BM)) return;      // dead row tile of the packed note level
    const int kb_total = (g.K + BKE - 1) / BKE;
    int kb_beg = blockIdx.z * g.kb_per_split;
    int kb_end = min(kb_total, kb_beg + g.kb_per_split);
    PdLiveBlocks lb{g.rows.cp, g.rows.slot_rows, g.rows.n_slots, BKE, 0, 0, 0};
    if (g.pred == 2) {
        // K runs over slot-major rows: this CTA takes an equal share of the LIVE k-blocks (kb_* count live blocks)
        const int total = lb.total();
        const int per = (total + (int)gridDim.z - 1) / (int)gridDim.z;
        kb_beg = min(total, (int)blockIdx.z * per);
        kb_end = min(total, kb_beg + per);
    }
    const int nkb = kb_end - kb_beg;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(MH * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            if (g.pred == 2 && nkb > 0) lb.seek(kb_beg);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                int k0 = (kb_beg + i) * BKE;
                if (g.pred == 2) { k0 = lb.row0(); lb.next(); }
                if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                uint8_t* a = sA + s * A_BYTES;
                uint8_t* b = sB + s * B_BYTES;
                if (A_MN) {
#pragma unroll
                    for (int c = 0; c < MH * BM / CW; ++c) tma_load_2d(&tmA, &full[s], a + c * (BKE * 128), m0 + CW * c, k0);
                } else {
                    tma_load_2d(&tmA, &full[s], a, k0, m0);
                }
                if (B_MN) {
#pragma unroll
                    for (int c = 0; c < BN / CW; ++c) tma_load_2d(&tmB, &full[s], b + c * (BKE * 128), n0 + CW * c, k0);
                } else {
                    tma_load_2d(&tmB, &full[s], b, k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (InstrDescriptor): D=f32 [4,6)=1, A/B format tf32=2 at [7,10)/[10,13),
            // a_major [15], b_major [16], N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (Elem<EB>::FMT << 7) | (Elem<EB>::FMT << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                mbar_wait(&full[s], (i / STAGES) & 1);
                tc_fence_after();
                const uint32_t a = smem_u32(sA + s * A_BYTES), b = smem_u32(sB + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // 4 MMA k-slices per 128-byte row
                    // K-major SW128: rows of 128 B, 8-row groups 1024 B apart, k-slice = +32 B inside the atom.
                    // MN-major SW128_BASE32B: [k][128 B of M/N]; 32-wide M/N chunks BK*128 B apart (LBO),
                    // 4-row k-atoms 512 B apart (SBO), k-slice of 8 rows = +1024 B.
                    const uint64_t bd = B_MN ? make_desc(b + k * Elem<EB>::MN_KSTEP, BKE * 128, Elem<EB>::MN_SBO, Elem<EB>::MN_LAYOUT)
                                             : make_desc(b + k * 32, 16, 1024, 2);
#pragma unroll
                    for (int h = 0; h < MH; ++h) {       // the two 128-row halves share this B stage
                        const uint32_t ah = a + h * (BM * 128);
                        const uint64_t ad = A_MN ? make_desc(ah + k * Elem<EB>::MN_KSTEP, BKE * 128, Elem<EB>::MN_SBO, Elem<EB>::MN_LAYOUT)
                                                 : make_desc(ah + k * 32, 16, 1024, 2);
                        if (!(g.dbg & 2)) tc_mma<EB>(tmem_base + h * BN, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit(&empty[s]);            // frees the smem slot once these MMAs have read it
            }
            tc_commit(tmem_full);                // accumulator complete
        }
    } else {
        // epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32).  tcgen05.ld hands each thread 32
        // consecutive columns of ONE row; writing those straight out touches 32 different rows per store
        // instruction.  Instead each warp transposes its 32x32 chunk through shared memory (the pipeline's
        // stage-0 buffer is free once the accumulator is complete) and stores full 128-byte row segments.
        const int q = warp & 3;
        if (nkb > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        constexpr int STG = 36;                                   // padded row stride (floats)
        float* stg = reinterpret_cast<float*>(sA) + q * (32 * STG);
        const bool add_bias = g.bias != nullptr && blockIdx.z == 0;
        const bool ldc_vec = ((g.ldc & 3) == 0) && ((((uintptr_t)g.C) & 15) == 0);
        const bool bias_vec = (((uintptr_t)g.bias) & 15) == 0;
        const int sub_r = lane >> 3, colq = (lane & 7) * 4;
#pragma unroll 1
        for (int hc = 0; hc < MH * (BN / 32); ++hc) {
            const int h = hc / (BN / 32), c = hc % (BN / 32);
            const int nb = n0 + c * 32;
            if (nb >= g.N || m0 + h * BM >= g.M || (g.dbg & 1)) continue;
            uint32_t r[32];
            if (nkb > 0) {
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * BN + c * 32, r);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(stg + lane * STG + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            __syncwarp();
            const int n = nb + colq;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add_bias) {
                if (n + 3 < g.N && bias_vec) bb = *reinterpret_cast<const float4*>(g.bias + n);
                else {
                    if (n < g.N) bb.x = g.bias[n];
                    if (n + 1 < g.N) bb.y = g.bias[n + 1];
                    if (n + 2 < g.N) bb.z = g.bias[n + 2];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = i * 4 + sub_r;
                const int m = m0 + h * BM + q * 32 + row;
                float4 v = *reinterpret_cast<const float4*>(stg + row * STG + colq);
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                if (m < g.M && n < g.N) {
                    float* dst = g.C + (long)m * g.ldc + n;
                    if (ldc_vec && n + 3 < g.N) {
                        if (g.atomic)
                            // accumulate / split-K: fire-and-forget vector reduction at L2 (a load+add+store
                            // epilogue here ran at ~30 GB/s)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                                         "f"(v.z), "f"(v.w) : "memory");
                        else
                            *reinterpret_cast<float4*>(dst) = v;
                    } else {
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n + j < g.N) {
                                if (g.atomic) atomicAdd(dst + j, e[j]);
                                else dst[j] = e[j];
                            }
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(MH * BN));
    }
}

// ---- persistent variant ---------------------------------------------------------------------------
// One CTA per SM loops over (m-tile, n-tile, k-split) work items.  Two accumulators live in TMEM
// (2 x BN columns): while the epilogue warps drain accumulator j%2 (tcgen05.ld -> smem transpose ->
// coalesced stores) the MMA warp already fills the other one from the next item's TMA stages, so the
// store phase -- 40 % of a K=512 tile in the one-shot kernel (tools/gemm_dissect.py) -- overlaps the
// main loop instead of following it.  Barrier phases run across items (global k-block counter).
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int EB, int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_persistent(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, TcArgs g, int tiles_m, int tiles_n, int n_split) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int BKE = Elem<EB>::BKE, CW = 128 / EB;
    constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, STG = 36;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    // epilogue staging: 4 warps x 2 buffers x (32 rows x 128 B) -- TMA-store tiles (128-byte swizzle), or one padded
    // 32 x STG transpose buffer per warp on the plain-store path
    float* stg_all = (float*)(sB + STAGES * B_BYTES);
    uint64_t* full = (uint64_t*)((uint8_t*)stg_all + 4 * 2 * 4096);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;                              // [2]
    uint64_t* tmem_empty = tmem_full + 2;                              // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_total = (g.K + BKE - 1) / BKE;
    const long n_items = (long)tiles_m * tiles_n * n_split;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item -> (m0, n0, k-block range); consecutive items share the A tile (same m, next n)
    auto decode = [&](long item, int& m0, int& n0, int& kb_beg, int& nkb, int& z) {
        z = (int)(item % n_split);
        long t = item / n_split;
        n0 = (int)(t % tiles_n) * BN;
        m0 = (int)(t / tiles_n) * BM;
        kb_beg = z * g.kb_per_split;
        nkb = min(kb_total, kb_beg + g.kb_per_split) - kb_beg;
    };

    if (warp == 0) {
        if (lane == 0) {
            long cnt = 0;
            for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
                int m0, n0, kb_beg, nkb, z;
                decode(item, m0, n0, kb_beg, nkb, z);
                if (g.pred == 1 && !pd_rows_live(g.rows, m0, BM)) continue;     // dead row tile (all three roles skip it)
                for (int i = 0; i < nkb; ++i, ++cnt) {
                    const int s = (int)(cnt % STAGES), k0 = (kb_beg + i) * BKE;
                    if (cnt >= STAGES) mbar_wait(&empty[s], (uint32_t)((cnt / STAGES) - 1) & 1);
                    mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                    uint8_t* a = sA + s * A_BYTES;
                    uint8_t* b = sB + s * B_BYTES;
                    if (A_MN) {
#pragma unroll
                        for (int c = 0; c < BM / CW; ++c) tma_load_2d(&tmA, &full[s], a + c * (BKE * 128), m0 + CW * c, k0);
                    } else {
                        tma_load_2d(&tmA, &full[s], a, k0, m0);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int c = 0; c < BN / CW; ++c) tma_load_2d(&tmB, &full[s], b + c * (BKE * 128), n0 + CW * c, k0);
                    } else {
                        tma_load_2d(&tmB, &full[s], b, k0, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (Elem<EB>::FMT << 7) | (Elem<EB>::FMT << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            long cnt = 0, j = 0;
            for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
                int m0, n0, kb_beg, nkb, z;
                decode(item, m0, n0, kb_beg, nkb, z);
                if (g.pred == 1 && !pd_rows_live(g.rows, m0, BM)) continue;
                const int acc = (int)(j & 1);
                if (j >= 2) mbar_wait(&tmem_empty[acc], (uint32_t)((j >> 1) - 1) & 1);   // epilogue drained it
                tc_fence_after();
                for (int i = 0; i < nkb; ++i, ++cnt) {
                    const int s = (int)(cnt % STAGES);
                    mbar_wait(&full[s], (uint32_t)(cnt / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a = smem_u32(sA + s * A_BYTES), b = smem_u32(sB + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {          // 4 MMA k-slices per 128-byte row
                        const uint64_t ad = A_MN ? make_desc(a + k * Elem<EB>::MN_KSTEP, BKE * 128, Elem<EB>::MN_SBO, Elem<EB>::MN_LAYOUT)
                                                : make_desc(a + k * 32, 16, 1024, 2);
                        const uint64_t bd = B_MN ? make_desc(b + k * Elem<EB>::MN_KSTEP, BKE * 128, Elem<EB>::MN_SBO, Elem<EB>::MN_LAYOUT)
                                             : make_desc(b + k * 32, 16, 1024, 2);
                        tc_mma<EB>(tmem_base + acc * BN, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[s]);
                }
                tc_commit(&tmem_full[acc]);
                ++j;
            }
        }
    } else {
        const int q = warp & 3;
        float* stg = stg_all + q * (32 * STG);
        const bool ldc_vec = ((g.ldc & 3) == 0) && ((((uintptr_t)g.C) & 15) == 0);
        const bool bias_vec = (((uintptr_t)g.bias) & 15) == 0;
        const int sub_r = lane >> 3, colq = (lane & 7) * 4;
        long j = 0;
        int n_st = 0;                                     // bulk stores issued by this warp
        for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
            int m0, n0, kb_beg, nkb, z;
            decode(item, m0, n0, kb_beg, nkb, z);
            if (g.pred == 1 && !pd_rows_live(g.rows, m0, BM)) continue;
            const int acc = (int)(j & 1);
            mbar_wait(&tmem_full[acc], (uint32_t)(j >> 1) & 1);
            tc_fence_after();
            const bool add_bias = g.bias != nullptr && z == 0;
            if (g.tma_store) {
                // TMEM -> registers (+ bias) -> swizzled shared tile -> one bulk tensor store per 32 x 32 chunk; two
                // tiles per warp so the next chunk's TMEM read overlaps the previous chunk's drain; M / N edges are
                // clipped by the tensor map
                uint8_t* tiles = (uint8_t*)stg_all + q * (2 * 4096);
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int nb = n0 + c * 32;
                    if (nb >= g.N) break;
                    uint32_t r[32];
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, r);
                    if (add_bias) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj)
                            if (nb + jj < g.N) r[jj] = __float_as_uint(__uint_as_float(r[jj]) + __ldg(g.bias + nb + jj));
                    }
                    uint8_t* tile = tiles + (n_st & 1) * 4096;
                    if (n_st >= 2) {                      // the store issued from this tile two chunks ago has read it
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && !(g.dbg & 1)) tma_store_2d(&tmC, tile, nb, m0 + q * 32);
                    ++n_st;
                }
                tc_fence_before();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                ++j;
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int nb = n0 + c * 32;
                if (nb >= g.N) break;
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, r);
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4)
                    *reinterpret_cast<uint4*>(stg + lane * STG + jj) = make_uint4(r[jj], r[jj + 1], r[jj + 2], r[jj + 3]);
                __syncwarp();
                const int n = nb + colq;
                float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                if (add_bias) {
                    if (n + 3 < g.N && bias_vec) bb = *reinterpret_cast<const float4*>(g.bias + n);
                    else {
                        if (n < g.N) bb.x = g.bias[n];
                        if (n + 1 < g.N) bb.y = g.bias[n + 1];
                        if (n + 2 < g.N) bb.z = g.bias[n + 2];
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = i * 4 + sub_r;
                    const int m = m0 + q * 32 + row;
                    float4 v = *reinterpret_cast<const float4*>(stg + row * STG + colq);
                    v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                    if (m < g.M && n < g.N) {
                        float* dst = g.C + (long)m * g.ldc + n;
                        if (ldc_vec && n + 3 < g.N) {
                            if (g.atomic)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                                             "f"(v.z), "f"(v.w) : "memory");
                            else
                                *reinterpret_cast<float4*>(dst) = v;
                        } else {
                            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (n + jj < g.N) {
                                    if (g.atomic) atomicAdd(dst + jj, e[jj]);
                                    else dst[jj] = e[jj];
                                }
                        }
                    }
                }
                __syncwarp();
            }
            // this warp's TMEM reads of accumulator `acc` are complete: hand it back to the MMA warp
            tc_fence_before();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            ++j;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all bulk stores of this warp landed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}

// zero-fill of a split-K output; with a live-row predicate only the row tiles the GEMM will write (dead rows keep their
// contents, as in the unsplit launch)
__global__ void zero_2d_tc_kernel(float* C, long ldc, int M, int N, PdRows rows, int tile_rows) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)M * N) return;
    const long m = i / N;
    if (rows.cp != nullptr && !pd_rows_live(rows, m / tile_rows * tile_rows, tile_rows)) return;
    C[m * ldc + (i % N)] = 0.0f;
}

template <int EB, int MH, int BN, int STAGES, int MINB, bool A_MN, bool B_MN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, dim3 grid, cudaStream_t st) {
    constexpr int smem = STAGES * (MH * BM * 128 + BN * 128) + 1024 + 256;
    // "background" launches (dbg bit 8: deferred weight-gradient GEMMs that run beside the step's latency-bound chain):
    // the shared-memory request is padded so that exactly ONE such CTA fits per SM and a chain CTA (<= 97.3 KB: the
    // 64-wide 4-stage tile of the per-step recurrent GEMMs) always finds room next to it -- 2 x 116 KB > 227 KB >= 116 + 98.3
    constexpr int kBgSmem = 116 * 1024;
    constexpr int smem_max = smem > kBgSmem ? smem : kBgSmem;
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<EB, MH, BN, STAGES, MINB, A_MN, B_MN>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) return (int)e;
    }
    gemm_tf32_kernel<EB, MH, BN, STAGES, MINB, A_MN, B_MN><<<grid, NUM_THREADS, (g.dbg & 8) ? smem_max : smem, st>>>(ta, tb, g);
    return pd_launch_status();
}

template <int EB, int MH, int BN, int STAGES, int MINB>
int launch_l(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, dim3 grid, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch<EB, MH, BN, STAGES, MINB, false, false>(ta, tb, g, grid, st);
    if (!a_mn && b_mn) return launch<EB, MH, BN, STAGES, MINB, false, true>(ta, tb, g, grid, st);
    if (a_mn && b_mn) return launch<EB, MH, BN, STAGES, MINB, true, true>(ta, tb, g, grid, st);
    return launch<EB, MH, BN, STAGES, MINB, true, false>(ta, tb, g, grid, st);
}

template <int EB, int BN, int STAGES, bool A_MN, bool B_MN>
int launch_p(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcArgs& g, int tiles_m, int tiles_n,
             int split, cudaStream_t st) {
    constexpr int smem = STAGES * (BM * 128 + BN * 128) + 4 * 2 * 4096 + 1024 + 256;
    static unsigned long long attr = 0;
    if (pd_first_use_on_device(attr)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_persistent<EB, BN, STAGES, A_MN, B_MN>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    long items = (long)tiles_m * tiles_n * split;
    int grid = (int)(items < PD_NUM_SMS ? items : PD_NUM_SMS);
    gemm_tf32_persistent<EB, BN, STAGES, A_MN, B_MN><<<grid, NUM_THREADS, smem, st>>>(ta, tb, tc, g, tiles_m, tiles_n, split);
    return pd_launch_status();
}

template <int EB, int BN, int STAGES>
int launch_pl(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const TcArgs& g,
              int tiles_m, int tiles_n, int split, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_p<EB, BN, STAGES, false, false>(ta, tb, tc, g, tiles_m, tiles_n, split, st);
    if (!a_mn && b_mn) return launch_p<EB, BN, STAGES, false, true>(ta, tb, tc, g, tiles_m, tiles_n, split, st);
    if (a_mn && b_mn) return launch_p<EB, BN, STAGES, true, true>(ta, tb, tc, g, tiles_m, tiles_n, split, st);
    return launch_p<EB, BN, STAGES, true, false>(ta, tb, tc, g, tiles_m, tiles_n, split, st);
}

// config id = (m_halves - 1) * 100000 + bn * 100 + stages * 10 + ctas_per_sm; 9xxxxx = persistent kernel
template <int EB>
int launch_cfg(int cfg, bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const TcArgs& g, dim3 grid,
               cudaStream_t st) {
    switch (cfg) {
        case 25622: return launch_l<EB, 1, 256, 2, 2>(a_mn, b_mn, ta, tb, g, grid, st);
        case 12823: return launch_l<EB, 1, 128, 2, 3>(a_mn, b_mn, ta, tb, g, grid, st);
        case 6441: return launch_l<EB, 1, 64, 4, 1>(a_mn, b_mn, ta, tb, g, grid, st);
        case 6433: return launch_l<EB, 1, 64, 3, 3>(a_mn, b_mn, ta, tb, g, grid, st);
        default: break;
    }
    if (EB == 4) {          // tuning-only configurations exist for the fp32/TF32 operand type
        switch (cfg) {
            case 25641: return launch_l<4, 1, 256, 4, 1>(a_mn, b_mn, ta, tb, g, grid, st);
            case 12841: return launch_l<4, 1, 128, 4, 1>(a_mn, b_mn, ta, tb, g, grid, st);
            case 12832: return launch_l<4, 1, 128, 3, 2>(a_mn, b_mn, ta, tb, g, grid, st);
            case 6442: return launch_l<4, 1, 64, 4, 2>(a_mn, b_mn, ta, tb, g, grid, st);
            case 6462: return launch_l<4, 1, 64, 6, 2>(a_mn, b_mn, ta, tb, g, grid, st);
            case 12842: return launch_l<4, 1, 128, 4, 2>(a_mn, b_mn, ta, tb, g, grid, st);
            case 125631: return launch_l<4, 2, 256, 3, 1>(a_mn, b_mn, ta, tb, g, grid, st);
            case 112841: return launch_l<4, 2, 128, 4, 1>(a_mn, b_mn, ta, tb, g, grid, st);
            case 112822: return launch_l<4, 2, 128, 2, 2>(a_mn, b_mn, ta, tb, g, grid, st);
            default: break;
        }
    }
    return PD_BAD_ARG;
}

// EB = 4: A, B fp32 (TF32 multiply).  EB = 2: A, B bf16.  Strides in elements; C / bias fp32.
template <int EB>
int gemm_tc_impl(const void* A, long sam, long sak, const void* B, long sbk, long sbn, float* C, long ldc,
                 const float* bias, int M, int N, int K, int accumulate, int cfg, cudaStream_t st, int pred = 0,
                 const int* cp = nullptr, int slot_rows = 0) {
    constexpr int BKE = 128 / EB, ALIGN = 16 / EB;       // k-block elements; stride alignment in elements
    const int dbg = cfg / 1000000;
    cfg %= 1000000;
    if (M <= 0 || N <= 0) return 0;
    if (K <= 0 || (sak != 1 && sam != 1) || (sbk != 1 && sbn != 1)) return PD_BAD_ARG;
    const bool a_mn = (sak != 1), b_mn = (sbk != 1);
    const long lda = a_mn ? sak : sam, ldb = b_mn ? sbk : sbn;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda % ALIGN) || (ldb % ALIGN) || lda < ALIGN || ldb < ALIGN)
        return PD_BAD_ARG;
    if (pred && (cp == nullptr || slot_rows <= 0 || (pred == 2 && slot_rows % BKE))) return PD_BAD_ARG;
    if (cfg == 0) {
        // measured on B200 (tools/gemm_tune.py, tools/gemm_dissect.py)
        const int tiles_m = (M + BM - 1) / BM;
        const int kb0 = (K + BKE - 1) / BKE;
        const long t256 = (long)tiles_m * ((N + 255) / 256), t128 = (long)tiles_m * ((N + 127) / 128);
        if (N <= 64) cfg = 6441;                                     // narrow heads
        else if (N <= 128) cfg = 12823;
        else if ((long)kb0 * BKE >= 32768 && t256 < PD_NUM_SMS) cfg = 12823;   // split-K weight gradients
        else if (t256 >= 2L * PD_NUM_SMS && kb0 >= 4) cfg = 925641;  // persistent, epilogue overlapped
        else if (t256 >= PD_NUM_SMS) cfg = 25622;                    // 2 CTAs/SM
        else if (t128 >= PD_NUM_SMS) cfg = 12823;
        else cfg = 6433;     // small per-step recurrent GEMMs (3 stages, 3 CTAs/SM: batch-512 step 9.59 -> 9.39 ms vs 6441)
        if (pred == 2 && cfg >= 900000) cfg = 12823;                 // live-k-block iteration lives in the one-shot kernel
    }
    const int mh = (cfg >= 100000 && cfg < 900000) ? 2 : 1;
    const int bn = (cfg % 100000) / 100;
    const int tiles_m = (M + mh * BM - 1) / (mh * BM);
    const int tiles_n = (N + bn - 1) / bn;
    const long tiles = (long)tiles_m * tiles_n;
    const int kb = (K + BKE - 1) / BKE;
    int split = 1;
    // split K until ~2 CTAs per SM exist: always for few-tile GEMMs, and for deep ones (K >= 4096: weight gradients such
    // as the time GRU's 3072x1024x16384, 192 tiles -> 318 us unsplit, 238 us split) also when tiles fill one wave only
    if ((tiles < PD_NUM_SMS && kb >= 32) || (tiles < 2 * PD_NUM_SMS && kb >= 4096 / BKE)) {
        split = (int)((2 * PD_NUM_SMS + tiles - 1) / tiles);
        if (split > kb / 8) split = kb / 8;
        if (split < 1) split = 1;
    }
    if ((dbg & 16) && kb > 128) {      // background launches with short-lived CTAs: at most 128 k-blocks each
        const int s2 = (kb + 127) / 128;
        if (s2 > split) split = s2;
    }
    int kb_per = (kb + split - 1) / split;
    split = (kb + kb_per - 1) / kb_per;
    TcArgs g{C, ldc, bias, M, N, K, kb_per, (split > 1 || accumulate || pred == 2) ? 1 : 0, dbg, 0,
             PdRows{cp, slot_rows, pred ? (int)(((pred == 1 ? (long)M : (long)K) + slot_rows - 1) / slot_rows) : 0}, pred};
    CUtensorMap ta, tb;
    int rc;
    // K-major: [rows][K] -> dims {K, rows}, box {128 B, BM|bn rows}.  MN-major: [K][rows] -> dims {rows, K},
    // box {128 B of M/N, BKE k-rows}.
    rc = a_mn ? make_map(&ta, A, EB, M, K, lda, BKE, true) : make_map(&ta, A, EB, K, M, lda, mh * BM, false);
    if (rc) return rc;
    rc = b_mn ? make_map(&tb, B, EB, N, K, ldb, BKE, true) : make_map(&tb, B, EB, K, N, ldb, bn, false);
    if (rc) return rc;
    if ((split > 1 || pred == 2) && !accumulate) {
        if (ldc == N && pred != 1) cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st);      // dense C: a memset node
        else zero_2d_tc_kernel<<<pd_blocks((long)M * N, 256), 256, 0, st>>>(C, ldc, M, N, pred == 1 ? g.rows : PdRows{nullptr, 0, 0},
                                                                            mh * BM);
    }
    if (cfg == 925641 || (cfg == 912861 && EB == 4)) {
        // C tiles leave through TMA bulk stores when they are plain stores into a 16-byte aligned matrix
        CUtensorMap tc = ta;
        if (!g.atomic && (((uintptr_t)C) & 15) == 0 && (ldc & 3) == 0 && ldc >= 4 && !(dbg & 4)) {
            rc = make_map(&tc, C, 4, N, M, ldc, 32, false, true);
            if (rc) return rc;
            g.tma_store = 1;
        }
        if (cfg == 925641) return launch_pl<EB, 256, 4>(a_mn, b_mn, ta, tb, tc, g, tiles_m, tiles_n, split, st);
        return launch_pl<4, 128, 6>(a_mn, b_mn, ta, tb, tc, g, tiles_m, tiles_n, split, st);
    }
    dim3 grid(tiles_m, tiles_n, split);
    return launch_cfg<EB>(cfg, a_mn, b_mn, ta, tb, g, grid, st);
}

// fp32 (rows, cols; row stride ldx) -> bf16 (row stride ldo), round to nearest even
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, long ldx, long rows, int cols, uint16_t* __restrict__ out,
                                   long ldo) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c2 = (cols + 1) >> 1;
    if (i >= rows * c2) return;
    const long r = i / c2;
    const int c = (int)(i % c2) * 2;
    const float a = x[r * ldx + c];
    const float b = (c + 1 < cols) ? x[r * ldx + c + 1] : 0.0f;
    uint32_t lo, hi;
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(*reinterpret_cast<uint16_t*>(&lo)) : "f"(a));
    asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(*reinterpret_cast<uint16_t*>(&hi)) : "f"(b));
    out[r * ldo + c] = (uint16_t)lo;
    if (c + 1 < cols) out[r * ldo + c + 1] = (uint16_t)hi;
}

}  // namespace

// Same contract as pd_gemm_f32 (strides in floats) with TF32 multiplies on the tensor cores.
// Requirements: operand base pointers 16-byte aligned, row strides multiples of 4 floats.  Returns
// PD_BAD_ARG (-22) when a requirement does not hold so the caller can route to pd_gemm_f32.
PD_API int pd_gemm_tf32(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                        const float* bias, int M, int N, int K, int accumulate, void* stream) {
    return gemm_tc_impl<4>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, 0, (cudaStream_t)stream);
}

// Tuning variant: cfg = BN*100 + stages*10 + CTAs/SM (one of the instantiated configurations), 0 = heuristic.
// Packed note level (ops.py): the same GEMM over slot-major row buffers with a DEVICE-side live-row table (common.cuh
// PdRows).  pred = 1: the rows of A and C are slot-major rows -- dead 128-row tiles are neither loaded, multiplied nor
// stored (their C rows keep whatever they held).  pred = 2: K runs over slot-major rows (weight gradients) -- only live
// 32-row k-blocks are accumulated, shared evenly between the split-K CTAs; slot_rows % 32 == 0.  cp: n_slots ints.
PD_API int pd_gemm_tf32_rows(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                             const float* bias, int M, int N, int K, int accumulate, int pred, const int* cp, int slot_rows,
                             void* stream) {
    return gemm_tc_impl<4>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, 0, (cudaStream_t)stream, pred, cp,
                           slot_rows);
}

// How many K splits pd_gemm_tf32 uses for this problem (1 = plain stores; > 1 = zero-fill + red.global.add epilogue):
// lets a caller that can clear C elsewhere pass accumulate = 1 and save the zero-fill node.
PD_API int pd_gemm_tf32_splits(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 1;
    const int tiles_m = (M + BM - 1) / BM, kb0 = (K + 31) / 32;
    const long t256 = (long)tiles_m * ((N + 255) / 256), t128 = (long)tiles_m * ((N + 127) / 128);
    int bn;
    if (N <= 64) bn = 64;
    else if (N <= 128) bn = 128;
    else if ((long)kb0 * 32 >= 32768 && t256 < PD_NUM_SMS) bn = 128;
    else if (t256 >= 2L * PD_NUM_SMS && kb0 >= 4) bn = 256;
    else if (t256 >= PD_NUM_SMS) bn = 256;
    else if (t128 >= PD_NUM_SMS) bn = 128;
    else bn = 64;
    const long tiles = (long)tiles_m * ((N + bn - 1) / bn);
    int split = 1;
    if ((tiles < PD_NUM_SMS && kb0 >= 32) || (tiles < 2 * PD_NUM_SMS && kb0 >= 128)) {
        split = (int)((2 * PD_NUM_SMS + tiles - 1) / tiles);
        if (split > kb0 / 8) split = kb0 / 8;
        if (split < 1) split = 1;
    }
    const int kb_per = (kb0 + split - 1) / split;
    return (kb0 + kb_per - 1) / kb_per;
}

PD_API int pd_gemm_tf32_cfg(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C, long ldc,
                            const float* bias, int M, int N, int K, int accumulate, int cfg, void* stream) {
    return gemm_tc_impl<4>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, cfg, (cudaStream_t)stream);
}

// bf16 operands (A, B: uint16 bf16 bit patterns; strides in elements, multiples of 8; bases 16-byte aligned),
// fp32 accumulate / output / bias: tcgen05.mma kind::f16.  Same layouts as pd_gemm_tf32.
PD_API int pd_gemm_bf16(const void* A, long sam, long sak, const void* B, long sbk, long sbn, float* C, long ldc,
                        const float* bias, int M, int N, int K, int accumulate, void* stream) {
    return gemm_tc_impl<2>(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate, 0, (cudaStream_t)stream);
}

PD_API int pd_f32_to_bf16(const float* x, long ldx, long rows, int cols, void* out, long ldo, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    long n = rows * ((cols + 1) / 2);
    f32_to_bf16_kernel<<<pd_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, (uint16_t*)out, ldo);
    return pd_launch_status();
}
