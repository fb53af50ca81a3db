// Microbenchmark: issue rate of legacy warp-level mma.sync on sm_100a (TF32 m16n8k8, BF16 m16n8k16) next to FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b0 = 0x3f800000u, b1 = 0x3f000000u;
    float f[32];
    for (int i = 0; i < 32; ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaf(f[i], 1.0001f, 0.5f);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    for (int i = 0; i < 32; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flop_per_inner, int inner) {
    float* out;
    cudaMalloc(&out, 148 * 4 * 256 * 4);
    const int iters = 20000;
    k<MODE><<<148 * 4, 256>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = 148.0 * 4 * 8, ops = warps * iters * inner;
    printf("%-28s %8.3f ms  %8.2f TFLOP/s   %.2f warp-instr/ns/SM\n", name, ms, ops * flop_per_inner / ms * 1e-9,
           ops / 148 / (ms * 1e6));
    cudaFree(out);
}

int main() {
    run<0>("mma.sync m16n8k8 tf32", 2.0 * 16 * 8 * 8, 8);
    run<1>("mma.sync m16n8k16 bf16", 2.0 * 16 * 8 * 16, 8);
    run<2>("ffma", 2.0 * 32, 32);
    return 0;
}
