// Gradient exchange over NVLink peer memory: the data-parallel training step's one collective (the reference averages
// its replicas' gradients inside nn.DataParallel's backward, amc_dl/torch_plus/module.py:67-68, 152-157) as ONE kernel
// per gradient bucket instead of an NCCL call.
//
// Every rank owns a "symmetric" region (same layout on all ranks, allocated with cudaMalloc and opened by the peers
// through CUDA IPC): a small flag area followed by the flat gradient buckets.  The kernel on rank r
//   1. start barrier: block b tells block b of every peer "my bucket is written" and waits for theirs,
//   2. reduces ITS slice of the bucket (elements [r*n/W, (r+1)*n/W)): 16-byte loads from all W copies over NVLink, summed
//      in rank order, scaled by 1/W, and STORES the result into all W copies (two-shot all-reduce with the all-gather
//      done as peer stores, so one kernel and two barriers),
//   3. optionally accumulates the sum of squares of what it reduced (the clip's global norm: sum of the W ranks'
//      partials -- written to every peer, so all ranks add the same W numbers in the same order),
//   4. end barrier: signals "my stores are out" to every peer and waits for theirs; when the kernel retires, the local
//      copy of the bucket is the averaged gradient.
// Blocks pair up by index, so there is no grid-wide synchronisation: the data a block reads was written by the peer's
// PREVIOUS kernels in stream order, and any block's start flag proves those have completed; the stores a block waits
// for at the end are those of the peers' blocks with its own index, and the kernel retires when all blocks have.
// Flags carry a per-block epoch kept in device memory (the launch arguments of a captured graph are frozen), compared
// with >=, so replays need no reset.
// Traffic per rank and bucket of n floats: reads 4n(W-1)/W bytes from peers, writes the same -- bound by the NVLink
// port (900 GB/s per direction), latency-bound for the 8 MB buckets used here; grid and unroll are sized to keep
// ~2 MB of peer loads in flight.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

#define PD_AR_MAX_RANKS 8
#define PD_AR_MAX_BLOCKS 64
#define PD_AR_MAX_SRC 40
// 512-thread blocks: 32 of them keep ~1 MB of peer loads in flight per exchange (256-thread blocks halve the bandwidth of
// an exchange and did not make the step faster, profiles/r02_ddp_ab_2gpu.txt)
#define PD_AR_THREADS 512
// flag slot of ONE bucket (uint32 words): start[b][src], done[b][src], norm partials [src][b] (floats).  A slot per bucket,
// so exchanges of different buckets may be in flight at the same time (on different streams).
#define PD_AR_NORM_OFF (2 * PD_AR_MAX_BLOCKS * PD_AR_MAX_RANKS)
#define PD_AR_SLOT_WORDS (PD_AR_NORM_OFF + PD_AR_MAX_RANKS * PD_AR_MAX_BLOCKS)

namespace {

struct ArPeers {
    float* data[PD_AR_MAX_RANKS];
    unsigned* flags[PD_AR_MAX_RANKS];
};

// The bucket's gradients as their producers left them (one buffer per parameter): gathered into this rank's copy of the
// bucket by the exchange kernel itself.  off4 / n: slot start (in float4 of the bucket) and element count.
struct ArGather {
    const float* src[PD_AR_MAX_SRC];
    int off4[PD_AR_MAX_SRC];
    int n[PD_AR_MAX_SRC];
    int count;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Weak L1-bypassing accesses, ordered by the flag barriers around them (the acquire of the start barrier + __syncthreads
// before, __syncthreads + __threadfence_system + flag store of the end barrier after).  Sys-scope strong accesses
// measured the same (tools/ar_bench.py, profiles/r02_ar_bench_2gpu.txt).
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag >= e (wrap-safe); gives up after ~20 s and raises *err (a dead peer must not hang the GPU forever;
// ranks may legitimately be seconds apart in their first, eager steps)
__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned e, int* err) {
    if ((int)(ld_acquire_sys(flag) - e) >= 0) return;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(flag) - e) < 0) {
        if (globaltimer_ns() - t0 > 20000000000ull) {
            if (err) atomicExch(err, 1);
            return;
        }
    }
}

// Bucket-relative float4 index i belongs to thread (i mod stride) of the grid, in the gather and in the reduction, on
// every rank: what a thread reads from a peer's copy was gathered by the peer's thread with the same (block, lane), whose
// block signalled after it.
__device__ __forceinline__ long first_owned(long a, long lane, long stride) {
    long d = (lane - a) % stride;
    if (d < 0) d += stride;
    return a + d;
}

__device__ __forceinline__ void gather_local(const ArGather& G, float* dst, long lane, long stride) {
    for (int k = 0; k < G.count; ++k) {
        const float* src = G.src[k];
        const long o4 = G.off4[k], n = G.n[k], n4 = (n + 3) >> 2;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        for (long i = first_owned(o4, lane, stride); i < o4 + n4; i += stride) {
            const long e = (i - o4) * 4;
            float4 v;
            if (vec && e + 3 < n) {
                v = __ldcs(reinterpret_cast<const float4*>(src + e));
            } else {
                v.x = e < n ? src[e] : 0.0f;
                v.y = e + 1 < n ? src[e + 1] : 0.0f;
                v.z = e + 2 < n ? src[e + 2] : 0.0f;
                v.w = e + 3 < n ? src[e + 3] : 0.0f;
            }
            reinterpret_cast<float4*>(dst)[i] = v;
        }
    }
}

template <int W, int U>
__device__ __forceinline__ void reduce_slice(const ArPeers& P, long off4, long lo, long hi, float scale, float& ss) {
    const long stride = (long)gridDim.x * blockDim.x;
    long i = first_owned(lo, (long)blockIdx.x * blockDim.x + threadIdx.x, stride);
    for (; i + (U - 1) * stride < hi; i += U * stride) {
        float4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int p = 0; p < W; ++p) v[u][p] = ld_peer(reinterpret_cast<const float4*>(P.data[p]) + off4 + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float4 a = v[u][0];
#pragma unroll
            for (int p = 1; p < W; ++p) { a.x += v[u][p].x; a.y += v[u][p].y; a.z += v[u][p].z; a.w += v[u][p].w; }
            a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
            ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
#pragma unroll
            for (int p = 0; p < W; ++p) st_peer(reinterpret_cast<float4*>(P.data[p]) + off4 + i + u * stride, a);
        }
    }
    for (; i < hi; i += stride) {
        float4 a = ld_peer(reinterpret_cast<const float4*>(P.data[0]) + off4 + i);
#pragma unroll
        for (int p = 1; p < W; ++p) {
            float4 b = ld_peer(reinterpret_cast<const float4*>(P.data[p]) + off4 + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
#pragma unroll
        for (int p = 0; p < W; ++p) st_peer(reinterpret_cast<float4*>(P.data[p]) + off4 + i, a);
    }
}

// P.flags[p] = rank p's flag slot OF THIS BUCKET (start[b][src] | done[b][src] | norm partials [src][b]); epoch likewise.
// Only the W signalling threads of a block fence (after the block barrier that makes the block's writes theirs to
// publish -- the cooperative-groups grid-barrier pattern); a fence per thread costs ~10 us per exchange.
template <int W>
__global__ void __launch_bounds__(PD_AR_THREADS) allreduce_p2p_kernel(ArPeers P, int rank, long off4, long n4, float scale,
                                                            unsigned* epoch, int* err, int with_norm,
                                                            const __grid_constant__ ArGather G) {
    const int b = blockIdx.x;
    __shared__ unsigned e_s;
    __shared__ float red[16];
    if (threadIdx.x == 0) e_s = epoch[b] + 1;
    __syncthreads();
    const unsigned e = e_s;
    if (G.count > 0) {
        gather_local(G, P.data[rank] + off4 * 4, (long)b * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x);
        __syncthreads();
    }
    if (threadIdx.x < W) {
        const int p = threadIdx.x;
        if (G.count > 0) __threadfence_system();
        st_relaxed_sys(P.flags[p] + (b * PD_AR_MAX_RANKS + rank), e);
        wait_flag(P.flags[rank] + (b * PD_AR_MAX_RANKS + p), e, err);
    }
    __syncthreads();
    const long per = (n4 + W - 1) / W;
    const long lo = (long)rank * per, hi = (lo + per < n4) ? lo + per : n4;
    float ss = 0.0f;
    if (lo < hi) reduce_slice<W, (W <= 2 ? 4 : (W <= 4 ? 2 : 1))>(P, off4, lo, hi, scale, ss);
    if (with_norm) {
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.0f;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += red[k];
            for (int p = 0; p < W; ++p) {
                float* dst = reinterpret_cast<float*>(P.flags[p] + PD_AR_NORM_OFF) + rank * PD_AR_MAX_BLOCKS + b;
                asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(a) : "memory");
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < W) {
        const int p = threadIdx.x;
        __threadfence_system();
        st_relaxed_sys(P.flags[p] + ((PD_AR_MAX_BLOCKS + b) * PD_AR_MAX_RANKS + rank), e);
        wait_flag(P.flags[rank] + ((PD_AR_MAX_BLOCKS + b) * PD_AR_MAX_RANKS + p), e, err);
    }
    __syncthreads();
    if (threadIdx.x == 0) epoch[b] = e;
}

// sum of the norm partials of n_slots buckets (W ranks x nblocks each), fixed order -> identical on every rank
__global__ void __launch_bounds__(256) ar_norm_total_kernel(const unsigned* region, int n_slots, int W, int nblocks, float* out) {
    __shared__ float sh[256];
    float s = 0.0f;
    const int per_slot = W * PD_AR_MAX_BLOCKS;
    for (int i = threadIdx.x; i < n_slots * per_slot; i += 256) {
        const int slot = i / per_slot, j = i % per_slot;
        if (j % PD_AR_MAX_BLOCKS < nblocks)
            s += reinterpret_cast<const float*>(region + (long)slot * PD_AR_SLOT_WORDS + PD_AR_NORM_OFF)[j];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace

// bytes of the flag area in front of the gradient data of a symmetric region that serves n_buckets buckets
// (returns the byte count, not a status)
PD_API int pd_ar_flag_bytes(int n_buckets) {
    long words = (long)n_buckets * PD_AR_SLOT_WORDS;
    return (int)(((words * 4 + 255) / 256) * 256);
}

// cudaMalloc'ed, zero-filled region + its 64-byte CUDA IPC handle (cudaIpcMemHandle_t)
// compile-time limits (returns the value, not a status): 0 ranks, 1 blocks per launch, 2 gather sources per launch
PD_API int pd_ar_limit(int which) {
    return which == 0 ? PD_AR_MAX_RANKS : which == 1 ? PD_AR_MAX_BLOCKS : which == 2 ? PD_AR_MAX_SRC : PD_BAD_ARG;
}

PD_API int pd_ipc_alloc(long bytes, void** ptr, void* handle64) {
    if (bytes <= 0 || ptr == nullptr || handle64 == nullptr) return PD_BAD_ARG;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), *ptr);
}

PD_API int pd_ipc_open(const void* handle64, void** ptr) {
    if (ptr == nullptr || handle64 == nullptr) return PD_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

PD_API int pd_ipc_close(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
PD_API int pd_ipc_free(void* ptr) { return (int)cudaFree(ptr); }

// In-place average of n floats at element offset `off` of the data part of every rank's region.
// peers[p] = base address of rank p's region AS MAPPED IN THIS PROCESS (own region for p == rank); data starts
// flag_bytes after it.  off and n must be multiples of 4.  bucket: which flag slot / epoch row this exchange uses (every
// rank must issue the exchanges of one bucket in the same order; different buckets are independent).  epoch: n_buckets x
// pd_ar_limit(1) zero-initialised uint32 of THIS rank.
// n_src > 0: the bucket is first gathered from n_src (<= pd_ar_limit(2)) contiguous fp32 buffers: src[k] holds src_n[k]
// elements that go to bucket element src_off[k] (a multiple of 4; the padding up to the next slot is zero-filled).
PD_API int pd_allreduce_p2p(const void* const* peers, int rank, int world, long flag_bytes, long off, long n, float scale,
                            void* epoch, int* err, int bucket, int with_norm, int nblocks, const void* const* src,
                            const long* src_off, const long* src_n, int n_src, void* stream) {
    if (world < 2 || world > PD_AR_MAX_RANKS || rank < 0 || rank >= world || (off & 3) || (n & 3) || n <= 0 ||
        nblocks < 1 || nblocks > PD_AR_MAX_BLOCKS || n_src < 0 || n_src > PD_AR_MAX_SRC || bucket < 0 ||
        (long)(bucket + 1) * PD_AR_SLOT_WORDS * 4 > flag_bytes)
        return PD_BAD_ARG;
    ArGather G;
    G.count = n_src;
    for (int k = 0; k < n_src; ++k) {
        if ((src_off[k] & 3) || src_off[k] < 0 || src_n[k] < 0 || src_off[k] + src_n[k] > n ||
            (reinterpret_cast<uintptr_t>(src[k]) & 3))
            return PD_BAD_ARG;
        G.src[k] = reinterpret_cast<const float*>(src[k]);
        G.off4[k] = (int)(src_off[k] / 4);
        G.n[k] = (int)src_n[k];
    }
    ArPeers P;
    for (int p = 0; p < PD_AR_MAX_RANKS; ++p) {
        char* base = (char*)const_cast<void*>(peers[p < world ? p : 0]);
        P.flags[p] = reinterpret_cast<unsigned*>(base) + (long)bucket * PD_AR_SLOT_WORDS;
        P.data[p] = reinterpret_cast<float*>(base + flag_bytes);
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long off4 = off / 4, n4 = n / 4;
    // threads per block: 512 by default; PD_AR_THREADS=256 (with twice the blocks) lets an exchange block share an SM with a
    // CTA of the recurrent chain's tcgen05 kernels (192 threads x 255 registers leave 16 K registers) -- measured equal
    // (2 GPUs, GPU call 56: 7.98 ms/step with 32 x 512 threads on 4 streams, 8.00 with 64 x 256 on 2, 8.01 with 48 x 256 on 3)
    static int threads = 0;
    if (threads == 0) {
        const char* v = getenv("PD_AR_THREADS");
        threads = v ? atoi(v) : PD_AR_THREADS;
        if (threads != 128 && threads != 256 && threads != 512) threads = PD_AR_THREADS;
    }
    unsigned* ep = reinterpret_cast<unsigned*>(epoch) + (long)bucket * PD_AR_MAX_BLOCKS;
#define PD_AR_LAUNCH(Wv) allreduce_p2p_kernel<Wv><<<nblocks, threads, 0, s>>>(P, rank, off4, n4, scale, ep, err, with_norm, G)
    switch (world) {
        case 2: PD_AR_LAUNCH(2); break;
        case 3: PD_AR_LAUNCH(3); break;
        case 4: PD_AR_LAUNCH(4); break;
        case 5: PD_AR_LAUNCH(5); break;
        case 6: PD_AR_LAUNCH(6); break;
        case 7: PD_AR_LAUNCH(7); break;
        default: PD_AR_LAUNCH(8); break;
    }
#undef PD_AR_LAUNCH
    return pd_launch_status();
}

// out[0] = sum over the first n_slots buckets of the squared-norm partials pd_allreduce_p2p(with_norm) left in this
// rank's region (the squared global norm of the averaged gradient; same bits on every rank)
PD_API int pd_ar_norm_total(const void* region, int n_slots, int world, int nblocks, float* out, void* stream) {
    if (n_slots < 1 || world < 2 || world > PD_AR_MAX_RANKS) return PD_BAD_ARG;
    ar_norm_total_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned*>(region), n_slots, world, nblocks, out);
    return pd_launch_status();
}
