// pipe_microbench2.cu -- round-2 pipe experiments on B200 (register-only tap bodies, like pipe_microbench.cu):
//   * FP64 pipe: DFMA / DADD rate, F2F.F64.F32 rate and whether it shares the MUFU (XU) pipe,
//     IMAD.WIDE as a float->double bit conversion;
//   * pass-1 tap body with the depth-weighted sums accumulated in fp64 (otherwise idle pipe);
//   * pass-2 tap body (packed) with and without a second FADD2 for a (hi, lo) pass-1 mean;
//   * pass-1 tap body with a fraction of the 2^x evaluated by a Cody-Waite + polynomial FFMA sequence
//     instead of MUFU.EX2 (move work from the XU pipe to the FMA pipe).
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 1024
#define NCH 8

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// positive normal float -> double with the same value, by one IMAD.WIDE: bits * 2^29 + (896 << 52)
__device__ __forceinline__ double f2d_imad(float f) {
    unsigned long long r;
    asm volatile("mad.wide.u32 %0, %1, 536870912, %2;" : "=l"(r) : "r"(__float_as_uint(f)), "l"(0x3800000000000000ull));
    return __longlong_as_double((long long)r);
}
// 2^x for x in about [-60, 4]: Cody-Waite split with the 1.5*2^23 magic, degree-5 polynomial, exponent by integer add
__device__ __forceinline__ float ex2_poly(float x) {
    const float t = x + 12582912.0f;                 // FADD: round to nearest integer in the mantissa
    const float n = t - 12582912.0f;                 // FADD
    const float f = x - n;                           // FADD, f in [-0.5, 0.5]
    float p = 1.3333558e-3f;                         // 5 FFMA
    p = fmaf(p, f, 9.6181291e-3f);
    p = fmaf(p, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));   // SHL + IADD (ALU)
}

template <int OP>
__global__ void __launch_bounds__(256) k2(float* out, float a, float b, uint32_t ua, int iters) {
    float x[NCH];
    uint32_t u[NCH];
    double dd[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { x[i] = a * (threadIdx.x + i); u[i] = ua * (threadIdx.x + i + 1); dd[i] = (double)x[i]; }
    const double da = (double)a, db = (double)b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(da), "d"(db));
                if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(dd[i]) : "d"(db));
                if (OP == 2) {   // F2F.F64.F32 chain
                    double t;
                    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(x[i]));
                    x[i] = __uint_as_float(__double2hiint(t));
                }
                if (OP == 3) {   // F2F.F64.F32 + MUFU.EX2 1:1 -- same pipe => rate halves
                    double t;
                    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(x[i]));
                    x[i] = __uint_as_float(__double2hiint(t));
                    x[(i + 1) % NCH] = ex2a(x[(i + 1) % NCH]);
                }
                if (OP == 4) {   // IMAD.WIDE.U32 chain
                    unsigned long long r;
                    asm volatile("mad.wide.u32 %0, %1, 536870912, %2;" : "=l"(r) : "r"(u[i]), "l"(0x3800000000000000ull));
                    u[i] = (uint32_t)(r >> 32);
                }
                if (OP == 5) {   // DFMA + FFMA 1:1 (do they co-issue?)
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(da), "d"(db));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                }
                if (OP == 6) {   // DFMA + MUFU 1:1
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(da), "d"(db));
                    x[i] = ex2a(x[i]);
                }
                if (OP == 7 && (i & 1) == 0) {   // packed pass-1 body, fp32 accumulation (baseline; one count = two taps)
                    uint32_t ad0 = __vabsdiffu4(u[i], ua + it * 4 + rep), ad1 = __vabsdiffu4(u[i + 1], ua + it * 4 + rep);
                    f32x2 xx = pack2(__uint_as_float(__dp4a(ad0, ad0, 0x4B000000u)), __uint_as_float(__dp4a(ad1, ad1, 0x4B000000u)));
                    f32x2 ar = fma2(add2(xx, pack2(-8388608.0f, -8388608.0f)), pack2(a, a), pack2(b, b));
                    float a0, a1;
                    unpack2(ar, a0, a1);
                    f32x2 ff = pack2(ex2a(a0), ex2a(a1));
                    f32x2 acc = pack2(x[i], x[i + 1]), ws = pack2(x[(i + 2) % NCH], x[(i + 3) % NCH]);
                    acc = fma2(ff, pack2(b, b), acc);
                    ws = add2(ws, ff);
                    unpack2(acc, x[i], x[i + 1]);
                    unpack2(ws, x[(i + 2) % NCH], x[(i + 3) % NCH]);
                }
                if (OP == 8 && (i & 1) == 0) {   // packed pass-1 front end, fp64 sums via IMAD.WIDE conversion: DFMA + DADD per tap
                    uint32_t ad0 = __vabsdiffu4(u[i], ua + it * 4 + rep), ad1 = __vabsdiffu4(u[i + 1], ua + it * 4 + rep);
                    f32x2 xx = pack2(__uint_as_float(__dp4a(ad0, ad0, 0x4B000000u)), __uint_as_float(__dp4a(ad1, ad1, 0x4B000000u)));
                    f32x2 ar = fma2(add2(xx, pack2(-8388608.0f, -8388608.0f)), pack2(a, a), pack2(b, b));
                    float a0, a1;
                    unpack2(ar, a0, a1);
                    const double f0 = f2d_imad(ex2a(a0)), f1 = f2d_imad(ex2a(a1));
                    dd[i] = fma(f0, db, dd[i]);
                    dd[i + 1] = fma(f1, db, dd[i + 1]);
                    dd[(i + 2) % NCH] += f0;
                    dd[(i + 3) % NCH] += f1;
                }
                if (OP == 9 && (i & 1) == 0) {   // same with cvt.f64.f32 conversions
                    uint32_t ad0 = __vabsdiffu4(u[i], ua + it * 4 + rep), ad1 = __vabsdiffu4(u[i + 1], ua + it * 4 + rep);
                    f32x2 xx = pack2(__uint_as_float(__dp4a(ad0, ad0, 0x4B000000u)), __uint_as_float(__dp4a(ad1, ad1, 0x4B000000u)));
                    f32x2 ar = fma2(add2(xx, pack2(-8388608.0f, -8388608.0f)), pack2(a, a), pack2(b, b));
                    float a0, a1;
                    unpack2(ar, a0, a1);
                    const double f0 = (double)ex2a(a0), f1 = (double)ex2a(a1);
                    dd[i] = fma(f0, db, dd[i]);
                    dd[i + 1] = fma(f1, db, dd[i + 1]);
                    dd[(i + 2) % NCH] += f0;
                    dd[(i + 3) % NCH] += f1;
                }
                if ((OP == 10 || OP == 11) && (i & 1) == 0) {   // packed pass-2 body; OP 11: + one FADD2 ((hi, lo) mean)
                    uint32_t ad0 = __vabsdiffu4(u[i], ua + it * 4 + rep), ad1 = __vabsdiffu4(u[i + 1], ua + it * 4 + rep);
                    f32x2 xx = pack2(__uint_as_float(__dp4a(ad0, ad0, 0x4B000000u)), __uint_as_float(__dp4a(ad1, ad1, 0x4B000000u)));
                    f32x2 ar = fma2(add2(xx, pack2(-8388608.0f, -8388608.0f)), pack2(a, a), pack2(b, b));
                    f32x2 ee = add2(pack2(x[(i + 4) % NCH], x[(i + 4) % NCH]), pack2(b, a));
                    if (OP == 11) ee = add2(ee, pack2(a, b));
                    float a0, a1, e0, e1;
                    unpack2(ar, a0, a1);
                    unpack2(ee, e0, e1);
                    if (!(fabsf(e0) > 12.247f)) a0 = fmaf(-e0, e0, a0);
                    if (!(fabsf(e1) > 12.247f)) a1 = fmaf(-e1, e1, a1);
                    f32x2 ff = pack2(ex2a(a0), ex2a(a1));
                    f32x2 acc = pack2(x[i], x[i + 1]), ws = pack2(x[(i + 2) % NCH], x[(i + 3) % NCH]);
                    acc = fma2(ff, ee, acc);
                    ws = add2(ws, ff);
                    unpack2(acc, x[i], x[i + 1]);
                    unpack2(ws, x[(i + 2) % NCH], x[(i + 3) % NCH]);
                }
                if ((OP == 12 || OP == 13) && (i & 1) == 0) {   // packed pass-1 body with 1 of 8 (OP 12) / 1 of 4 (OP 13) tap pairs on the polynomial
                    uint32_t ad0 = __vabsdiffu4(u[i], ua + it * 4 + rep), ad1 = __vabsdiffu4(u[i + 1], ua + it * 4 + rep);
                    f32x2 xx = pack2(__uint_as_float(__dp4a(ad0, ad0, 0x4B000000u)), __uint_as_float(__dp4a(ad1, ad1, 0x4B000000u)));
                    f32x2 ar = fma2(add2(xx, pack2(-8388608.0f, -8388608.0f)), pack2(a, a), pack2(b, b));
                    float a0, a1;
                    unpack2(ar, a0, a1);
                    const bool poly = (OP == 12) ? (i == 0 && (rep & 1) == 0) : (i == 0);
                    f32x2 ff = poly ? pack2(ex2_poly(a0), ex2_poly(a1)) : pack2(ex2a(a0), ex2a(a1));
                    f32x2 acc = pack2(x[i], x[i + 1]), ws = pack2(x[(i + 2) % NCH], x[(i + 3) % NCH]);
                    acc = fma2(ff, pack2(b, b), acc);
                    ws = add2(ws, ff);
                    unpack2(acc, x[i], x[i + 1]);
                    unpack2(ws, x[(i + 2) % NCH], x[(i + 3) % NCH]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; i++) s += x[i] + __uint_as_float(u[i] & 0x3fffffffu) + (float)dd[i];
    if (s == 123.456f) out[0] = s;
}

template <int OP>
static double run(int sms, int ctas_per_sm, int dyn_smem = 0) {
    float* d;
    cudaMalloc(&d, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * ctas_per_sm;
    if (dyn_smem > 0) cudaFuncSetAttribute(k2<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
    k2<OP><<<grid, 256, dyn_smem>>>(d, 1.0001f, 0.5f, 0x01020304u, 32);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k2<OP><<<grid, 256, dyn_smem>>>(d, 1.0001f, 0.5f, 0x01020304u, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    double inner = (double)grid * 256 * ITERS * 4 * NCH;
    return inner / (best * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double perclk = 1.0 / ((double)sms * clk_khz * 1e3);
    double dfma = run<0>(sms, 8), dadd = run<1>(sms, 8), f2f = run<2>(sms, 8), f2f_mufu = run<3>(sms, 8),
           imadw = run<4>(sms, 8), dfma_ffma = run<5>(sms, 8), dfma_mufu = run<6>(sms, 8);
    // tap bodies at the filter kernel's occupancy (24 warps/SM: 3 CTAs of 256 threads, 72 KB dynamic smem each)
    double p1 = run<7>(sms, 24, 72 * 1024), p1_f64_imad = run<8>(sms, 24, 72 * 1024), p1_f64_cvt = run<9>(sms, 24, 72 * 1024),
           p2 = run<10>(sms, 24, 72 * 1024), p2_pair = run<11>(sms, 24, 72 * 1024), p1_poly8 = run<12>(sms, 24, 72 * 1024),
           p1_poly4 = run<13>(sms, 24, 72 * 1024);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, "
           "\"dfma_per_clk_per_sm\": %.2f, \"dadd_per_clk_per_sm\": %.2f, \"f2f_f64_f32_per_clk_per_sm\": %.2f, "
           "\"f2f_plus_mufu_pairs_per_clk_per_sm\": %.2f, \"imad_wide_per_clk_per_sm\": %.2f, "
           "\"dfma_plus_ffma_pairs_per_clk_per_sm\": %.2f, \"dfma_plus_mufu_pairs_per_clk_per_sm\": %.2f, "
           "\"tap_pass1_f32x2_gtaps_s\": %.1f, \"tap_pass1_f64sums_imadwide_gtaps_s\": %.1f, \"tap_pass1_f64sums_cvt_gtaps_s\": %.1f, "
           "\"tap_pass2_f32x2_gtaps_s\": %.1f, \"tap_pass2_f32x2_pair_mean_gtaps_s\": %.1f, "
           "\"tap_pass1_poly_1of8_gtaps_s\": %.1f, \"tap_pass1_poly_1of4_gtaps_s\": %.1f, "
           "\"how\": \"register-only bodies, 8 independent chains; pipe rates at 8 CTAs x 256 thr per SM, tap bodies at 24 warps/SM; best of 5\"}\n",
           p.name, sms, clk_khz / 1e3, dfma * perclk, dadd * perclk, f2f * perclk, f2f_mufu * perclk, imadw * perclk,
           dfma_ffma * perclk, dfma_mufu * perclk, p1 / 1e9, p1_f64_imad / 1e9, p1_f64_cvt / 1e9, p2 / 1e9, p2_pair / 1e9,
           p1_poly8 / 1e9, p1_poly4 / 1e9);
    return 0;
}
