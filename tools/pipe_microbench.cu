// pipe_microbench.cu -- measures the per-pipe instruction throughput of this B200 that bounds the
// joint-bilateral kernels: FFMA (the FP32 roofline denominator, "of measured"), FADD, MUFU.EX2,
// VABSDIFF4, IDP.4A, the pass-1 / pass-2 tap bodies, and shared-memory LDS.128 bandwidth.
// MEASURED_PEAKS.json carries only HBM and bf16 tensor peaks; SURVEY.md 8(d) asks for this one.
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NCH 8

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}

template <int OP>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, float a, float b, uint32_t ua, int iters) {
    __shared__ float lut[256];   // per-channel colour-range LUT (OP 7): exp(-k^2/(2 sigma_c^2))
    if (OP == 7) {
        lut[threadIdx.x] = __expf(-(float)(threadIdx.x * threadIdx.x) / 5000.0f);
        __syncthreads();
    }
    float x[NCH];
    uint32_t u[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { x[i] = a * (threadIdx.x + i); u[i] = ua * (threadIdx.x + i + 1); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                if (OP == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
                if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                if (OP == 3) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(0));
                if (OP == 4) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(u[(i + 1) % NCH]));
                if (OP == 5) {  // pass-1 tap body: VABSDIFF4, IDP.4A, FADD, FFMA, MUFU.EX2, FFMA, FADD
                    uint32_t ad = __vabsdiffu4(u[i], ua + it * 4 + rep);
                    float cdf = __uint_as_float(__dp4a(ad, ad, 0x4B000000u)) - 8388608.0f;
                    float f;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(fmaf(cdf, a, b)));
                    x[i] = fmaf(f, b, x[i]);
                    x[(i + 1) % NCH] += f;
                }
                if (OP == 10) {  // int -> float conversion (I2FP.F32.S32): which pipe, what rate?
                    float t;
                    asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(u[i]));
                    u[i] = __float_as_uint(t) & 0x00ffffffu;
                }
                if (OP == 11) {  // I2FP interleaved 1:1 with FFMA: if they share a pipe the FFMA rate halves
                    float t;
                    asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(u[i]));
                    u[i] = __float_as_uint(t) & 0x00ffffffu;
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                }
                if (OP == 12) {  // IDP.4A interleaved 1:1 with FFMA
                    asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(u[(i + 1) % NCH]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                }
                if (OP == 13) {  // VABSDIFF4 interleaved 1:1 with FFMA
                    asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(0));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                }
                if (OP == 14) {  // FSETP (+ predicated FADD) interleaved with FFMA: does the compare share the FMA pipe?
                    asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, %2; @p fma.rn.f32 %0, %0, %3, %2; }"
                                 : "+f"(x[i]) : "f"(x[(i + 1) % NCH]), "f"(b), "f"(a));
                }
                if (OP == 15) {  // plain FSETP chain only (predicate consumed by a select on the ALU pipe)
                    asm volatile("{ .reg .pred p; setp.gt.f32 p, %1, %2; selp.u32 %0, %0, %3, p; }"
                                 : "+r"(u[i]) : "f"(x[i]), "f"(b), "r"(ua));
                }
                if (OP == 8) {  // packed fp32x2 FMA (Blackwell FFMA2): counts as ONE instruction, two FMAs
                    unsigned long long v = pack2(x[i], x[(i + 4) % NCH]);
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(pack2(a, a)), "l"(pack2(b, b)));
                    unpack2(v, x[i], x[(i + 4) % NCH]);
                }
                if (OP == 9 && (i & 1) == 0) {  // pass-1 body for TWO taps with FADD2/FFMA2 (one count = two taps)
                    uint32_t ad0 = __vabsdiffu4(u[i], ua + it * 4 + rep), ad1 = __vabsdiffu4(u[i + 1], ua + it * 4 + rep);
                    unsigned long long xx = pack2(__uint_as_float(__dp4a(ad0, ad0, 0x4B000000u)),
                                                  __uint_as_float(__dp4a(ad1, ad1, 0x4B000000u)));
                    unsigned long long cd, ar;
                    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(cd) : "l"(xx), "l"(pack2(-8388608.0f, -8388608.0f)));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ar) : "l"(cd), "l"(pack2(a, a)), "l"(pack2(b, b)));
                    float a0, a1, f0, f1;
                    unpack2(ar, a0, a1);
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f0) : "f"(a0));
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f1) : "f"(a1));
                    unsigned long long ff = pack2(f0, f1), acc = pack2(x[i], x[i + 1]), ws = pack2(x[(i + 2) % NCH], x[(i + 3) % NCH]);
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ff), "l"(pack2(b, b)));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(ws) : "l"(ff));
                    unpack2(acc, x[i], x[i + 1]);
                    unpack2(ws, x[(i + 2) % NCH], x[(i + 3) % NCH]);
                }
                if (OP == 7) {  // pass-1 tap body with per-channel colour LUTs instead of IDP.4A + MUFU
                    uint32_t ad = __vabsdiffu4(u[i], ua + it * 4 + rep);
                    float f = lut[ad & 0xffu] * lut[(ad >> 8) & 0xffu] * lut[(ad >> 16) & 0xffu] * a;
                    x[i] = fmaf(f, b, x[i]);
                    x[(i + 1) % NCH] += f;
                }
                if (OP == 6) {  // pass-2 tap body: + FADD, FSETP, predicated FFMA
                    uint32_t ad = __vabsdiffu4(u[i], ua + it * 4 + rep);
                    float cdf = __uint_as_float(__dp4a(ad, ad, 0x4B000000u)) - 8388608.0f;
                    float arg = fmaf(cdf, a, b);
                    float e = x[(i + 2) % NCH] - b;
                    if (!(fabsf(e) > 12.247f)) arg = fmaf(-e, e, arg);
                    float f;
                    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(arg));
                    x[i] = fmaf(f, e, x[i]);
                    x[(i + 1) % NCH] += f;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += x[i] + __uint_as_float(u[i] & 0x3fffffffu);
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) lds_kernel(float* out, int iters) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
            float4 v = buf[(idx + rep * 32) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 7;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

template <int OP>
static double run(int sms, int ctas_per_sm, double ops_per_inner, int dyn_smem = 0) {
    float* d;
    cudaMalloc(&d, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * ctas_per_sm;
    if (dyn_smem > 0) cudaFuncSetAttribute(pipe_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
    pipe_kernel<OP><<<grid, 256, dyn_smem>>>(d, 1.0001f, 0.5f, 0x01020304u, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        pipe_kernel<OP><<<grid, 256, dyn_smem>>>(d, 1.0001f, 0.5f, 0x01020304u, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    double inner = (double)grid * 256 * ITERS * 4 * NCH;
    return inner * ops_per_inner / (best * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ffma = run<0>(sms, 8, 1), fadd = run<1>(sms, 8, 1), mufu = run<2>(sms, 8, 1), vabs = run<3>(sms, 8, 1),
           idp = run<4>(sms, 8, 1), tap1 = run<5>(sms, 8, 1), tap2 = run<6>(sms, 8, 1), tap1lut = run<7>(sms, 8, 1),
           ffma2 = run<8>(sms, 8, 1), tap1x2 = run<9>(sms, 8, 1), i2fp = run<10>(sms, 8, 1), i2fp_ffma = run<11>(sms, 8, 1),
           idp_ffma = run<12>(sms, 8, 1), vabs_ffma = run<13>(sms, 8, 1), fsetp_ffma = run<14>(sms, 8, 1),
           fsetp_sel = run<15>(sms, 8, 1);
    // occupancy sensitivity of the MUFU-heavy tap bodies: 3 resident CTAs (24 warps/SM, as the filter kernel
    // runs) vs 8 (64 warps/SM); 72 KB of dynamic shared memory per CTA caps residency at 3
    double tap1x2_24w = run<9>(sms, 24, 1, 72 * 1024), tap2_24w = run<6>(sms, 24, 1, 72 * 1024),
           tap1x2_40w = run<9>(sms, 20, 1, 44 * 1024);
    // LDS.128
    float* d; cudaMalloc(&d, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    lds_kernel<<<sms * 8, 256>>>(d, 64); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); lds_kernel<<<sms * 8, 256>>>(d, 8192); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double lds_bytes = (double)sms * 8 * 256 * 8192.0 * 8 * 16 / (best * 1e-3);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, "
           "\"ffma_tflops\": %.2f, \"ffma_ginst_s\": %.1f, \"fadd_ginst_s\": %.1f, \"mufu_ex2_ginst_s\": %.1f, "
           "\"vabsdiff4_ginst_s\": %.1f, \"idp4a_ginst_s\": %.1f, \"tap_pass1_gtaps_s\": %.1f, \"tap_pass2_gtaps_s\": %.1f, "
           "\"tap_pass1_channel_lut_gtaps_s\": %.1f, \"ffma2_ginst_s\": %.1f, \"tap_pass1_f32x2_gtaps_s\": %.1f, \"lds128_tb_s\": %.2f, "
           "\"i2fp_ginst_s\": %.1f, \"i2fp_plus_ffma_gpairs_s\": %.1f, \"idp4a_plus_ffma_gpairs_s\": %.1f, \"vabsdiff4_plus_ffma_gpairs_s\": %.1f, \"fsetp_plus_pred_ffma_gpairs_s\": %.1f, \"fsetp_plus_sel_gpairs_s\": %.1f, \"tap_pass1_f32x2_gtaps_s_24warps\": %.1f, \"tap_pass2_gtaps_s_24warps\": %.1f, \"tap_pass1_f32x2_gtaps_s_40warps\": %.1f, "
           "\"ffma_per_clk_per_sm\": %.1f, \"mufu_per_clk_per_sm\": %.1f, \"vabsdiff4_per_clk_per_sm\": %.1f, "
           "\"idp4a_per_clk_per_sm\": %.1f, \"how\": \"8 CTAs x 256 thr per SM, 8 independent chains, best of 5, per-clk at max clock\"}\n",
           p.name, sms, clk_khz / 1e3, 2 * ffma / 1e12, ffma / 1e9, fadd / 1e9, mufu / 1e9, vabs / 1e9, idp / 1e9,
           tap1 / 1e9, tap2 / 1e9, tap1lut / 1e9, ffma2 / 1e9, tap1x2 / 1e9, lds_bytes / 1e12, i2fp / 1e9, i2fp_ffma / 1e9, idp_ffma / 1e9, vabs_ffma / 1e9, fsetp_ffma / 1e9, fsetp_sel / 1e9, tap1x2_24w / 1e9, tap2_24w / 1e9, tap1x2_40w / 1e9, ffma / (sms * clk_khz * 1e3), mufu / (sms * clk_khz * 1e3),
           vabs / (sms * clk_khz * 1e3), idp / (sms * clk_khz * 1e3));
    return 0;
}
