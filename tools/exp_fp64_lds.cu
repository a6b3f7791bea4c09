// exp_fp64_lds.cu -- do the FP64 pipe and the shared-memory pipe of one SM overlap?
// mode 0: FP64 only, 1: LDS/STS only, 2: both in every warp (interleaved), 3: even warps FP64 / odd warps LDS
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_fp64_lds tools/exp_fp64_lds.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters) {
    extern __shared__ double2 sm[];
    const int tid = threadIdx.x, w = tid >> 5;
    double a[16];
    double2 v[4];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0 + tid * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = make_double2(tid, i);
    const bool do_f = MODE == 0 || MODE == 2 || (MODE == 3 && !(w & 1));
    const bool do_l = MODE == 1 || MODE == 2 || (MODE == 3 && (w & 1));
    const int work_f = MODE == 3 ? 2 : 1, work_l = MODE == 3 ? 2 : 1;  // mode 3: half the warps do twice the work each
    for (int it = 0; it < iters; it++) {
        if (do_f) {
#pragma unroll
            for (int r = 0; r < 4 * work_f; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] = fma(a[i], 1.0000001, 0.5);
        }
        if (do_l) {
#pragma unroll
            for (int r = 0; r < 2 * work_l; r++) {
#pragma unroll
                for (int i = 0; i < 4; i++) sm[(tid + 512 * i)] = v[i];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = sm[((tid + 32 * r) & 511) + 512 * i];
                __syncwarp();
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * 512 + tid] = s;
}

template <int MODE>
float run(double* d, int iters) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 512, 2048 * 16>>>(d, 10);
    cudaEventRecord(e0);
    k<MODE><<<148, 512, 2048 * 16>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* d; cudaMalloc(&d, 148 * 512 * 8);
    const int iters = 20000;
    // per iteration per warp: 64 DFMA (mode 0/2) ; 8 STS.128 + 8 LDS.128 (mode 1/2)
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    printf("FP64 only %.3f ms | LDS/STS only %.3f ms | both per warp %.3f ms | split warps %.3f ms\n", t0, t1, t2, t3);
    const double clk = 1.9e6;  // cycles per ms (approx)
    printf("cycles/iter/SM: fp64 %.0f (64 DFMA x 16 warps = 1024 warp-instr -> %.2f clk per instr per SMSP), lds %.0f (256 x 128-bit warp-instr -> %.2f clk each)\n",
           t0 * clk / iters, t0 * clk / iters / 256.0, t1 * clk / iters, t1 * clk / iters / 256.0);
    printf("overlap: both/max = %.2f, both/sum = %.2f\n", t2 / (t0 > t1 ? t0 : t1), t2 / (t0 + t1));
    return 0;
}
