// Shared helpers for libpolydis_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PD_API extern "C" __attribute__((visibility("default")))

// Every entry point returns 0 on success or a cudaError_t / negative argument-error code.
#define PD_BAD_ARG (-22)

static inline int pd_launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

static inline unsigned pd_blocks(long n, int per) { return (unsigned)((n + per - 1) / per); }

// SM count of the CURRENT device (cached per device; grids are sized from it, nothing is compiled in)
static inline int pd_num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = cache[dev & 63];
    if (c == 0 && (cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || c <= 0)) c = 148;
    return c;
}
#define PD_NUM_SMS pd_num_sms()

// cudaFuncSetAttribute is per DEVICE: true the first time a kernel's launcher runs on the current device
static inline bool pd_first_use_on_device(unsigned long long& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

// Row predicate of the packed note level (length-sorted rows, slot-major buffers; ops.py "packed notes"): row q of a
// (n_slots * slot_rows)-row buffer belongs to slot q / slot_rows and is LIVE iff q % slot_rows < cp[slot].  cp is DEVICE
// data (it depends on the batch's note counts), so one captured CUDA graph serves every batch: kernels are launched
// for the full extent and skip dead tiles.  cp == nullptr: everything is live.
struct PdRows {
    const int* cp;
    int slot_rows, n_slots;
};
// any live row in [q0, q0 + n)?
__device__ __forceinline__ bool pd_rows_live(const PdRows& p, long q0, int n) {
    if (p.cp == nullptr) return true;
    const long q1 = q0 + n;
    int s = (int)(q0 / p.slot_rows);
    long base = (long)s * p.slot_rows;
    while (base < q1 && s < p.n_slots) {
        const long lo = (q0 > base ? q0 : base) - base;
        if (lo < p.cp[s]) return true;
        ++s;
        base += p.slot_rows;
    }
    return false;
}
// Live k-blocks (BK consecutive rows, slot_rows % BK == 0) of a slot-major K range, enumerated in order: the split-K
// weight-gradient GEMMs iterate these instead of [0, K / BK).
struct PdLiveBlocks {
    const int* cp;
    int slot_rows, n_slots, bk, s, j, ls;
    __device__ __forceinline__ int live_of(int slot) const {
        const int c = cp[slot] < slot_rows ? cp[slot] : slot_rows;
        return c <= 0 ? 0 : (c + bk - 1) / bk;
    }
    __device__ __forceinline__ int total() const {
        int t = 0;
        for (int i = 0; i < n_slots; ++i) t += live_of(i);
        return t;
    }
    __device__ __forceinline__ void seek(int a) {       // position at live block number a (a < total())
        s = 0;
        ls = 0;
        while (s < n_slots) {
            ls = live_of(s);
            if (a < ls) break;
            a -= ls;
            ++s;
        }
        j = a;
    }
    __device__ __forceinline__ int row0() const { return s * slot_rows + j * bk; }
    __device__ __forceinline__ void next() {
        if (++j >= ls) {
            j = 0;
            do { ++s; } while (s < n_slots && (ls = live_of(s)) == 0);
        }
    }
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------
// The step's critical path is a chain of ~150 dependent 10-20 us kernels; with PDL the CTAs of kernel N+1 are scheduled
// while kernel N drains and run their prologue (barrier init, TMEM allocation, tensor-map prefetch) before blocking in
// pd_grid_dependency_wait(), which returns once every prerequisite grid has completed and its writes are visible.
// Kernels that opt in call pd_grid_launch_dependents() first (lets the runtime schedule the dependent grid early) and
// pd_grid_dependency_wait() before their first access to global memory; both are no-ops for ordinary launches.
// Measured on B200 inside a captured graph (tools/pdl_bench.py): the 32-step chain of fused forward steps 13.1 -> 12.2 us
// per step; the backward chain (gate-gradient kernel + split-K GEMM) got SLOWER with it (15.9 -> 17.6 us: early-scheduled
// dependents hold SM slots the 320-CTA GEMM needs), and the whole step 8.3 -> 8.5 ms, so only the fused step kernel opts
// in and the switch (pd_set_pdl) stays off by default.
__device__ __forceinline__ void pd_grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pd_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
extern int g_pd_pdl;       // 0 = ordinary launches (default), 1 = launch the opted-in kernels with the PDL attribute

template <typename... KArgs, typename... Args>
static inline cudaError_t pd_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pd_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float pd_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU forms for the TF32-mode recurrent kernels, which are instruction-bound on the gate math: ex2.approx +
// rcp.approx, absolute error ~1e-6 (three orders below the TF32 operand rounding of the matvec next to them);
// saturate correctly (exp -> inf gives 1/inf = 0).  The fp32 / tf32x3 parity paths keep expf / tanhf.
__device__ __forceinline__ float pd_sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float pd_tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
