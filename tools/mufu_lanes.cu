// mufu_lanes.cu -- does a MUFU.EX2 warp instruction cost less XU-pipe time when only part of the warp is active?
// (one 640x480 frame leaves 32 surplus warps; if a half-active warp cost half the XU time, 1-row tiles
// would be half-size work units.)  Every warp runs a chain of independent MUFU.EX2 with lanes >= nact switched off by
// divergence; 8 warps per SM sub-partition, so the XU pipe is the bound.  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mufu_lanes.cu -o tools/mufu_lanes
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k(float* out, float a, int nact, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a * (threadIdx.x + i);
    if ((threadIdx.x & 31) < nact) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 4096;
    printf("{");
    const int acts[5] = {32, 24, 16, 8, 1};
    for (int ai = 0; ai < 5; ++ai) {
        const int nact = acts[ai];
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<148, 1024>>>(out, 0.001f, nact, 64);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<<<148, 1024>>>(out, 0.001f, nact, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        // warp instructions per SM sub-partition: 8 warps x iters x 32
        const double winst = 8.0 * iters * 32;
        printf("%s\"active_%d\": {\"ms\": %.4f, \"cycles_per_warp_mufu_at_1965MHz\": %.3f}", ai ? ", " : "", nact, ms,
               ms * 1e-3 * 1.965e9 / winst);
    }
    printf("}\n");
    return 0;
}
