// pipe_rates.cu -- instruction-throughput microbenchmark for sm_100a (development aid, not product code):
// warp-instructions per cycle per SM for FFMA, FFMA2 (packed fp32x2), FADD, FMUL, FSEL/ISETP (ALU pipe), FMNMX, MUFU.RCP, mixes.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, long long *cyc) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS / 16; ++i)
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        if (MODE == 0) {        // FFMA x8
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        } else if (MODE == 1) { // FFMA2 x4 (8 flops-pairs)
            float2 p0 = make_float2(x0, x1), p1 = make_float2(x2, x3), p2 = make_float2(x4, x5), p3 = make_float2(x6, x7);
            const float2 aa = make_float2(a, a), bb = make_float2(b, b);
            p0 = __ffma2_rn(p0, aa, bb); p1 = __ffma2_rn(p1, aa, bb); p2 = __ffma2_rn(p2, aa, bb); p3 = __ffma2_rn(p3, aa, bb);
            x0 = p0.x; x1 = p0.y; x2 = p1.x; x3 = p1.y; x4 = p2.x; x5 = p2.y; x6 = p3.x; x7 = p3.y;
        } else if (MODE == 2) { // FADD x8
            x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a;
        } else if (MODE == 3) { // FSEL-ish: compare + select x4 (8 ALU instrs)
            x0 = x0 > a ? x1 : b; x1 = x1 > a ? x2 : b; x2 = x2 > a ? x3 : b; x3 = x3 > a ? x0 : b;
        } else if (MODE == 4) { // FMNMX x8
            x0 = fminf(x0, a) ; x1 = fmaxf(x1, b); x2 = fminf(x2, a); x3 = fmaxf(x3, b);
            x4 = fminf(x4, a) ; x5 = fmaxf(x5, b); x6 = fminf(x6, a); x7 = fmaxf(x7, b);
            x0 += 1.f; x1 -= 1.f; x2 += 1.f; x3 -= 1.f; x4 += 1.f; x5 -= 1.f; x6 += 1.f; x7 -= 1.f;
        } else if (MODE == 5) { // 4 FFMA + 4 ALU(FMNMX) interleaved
            x0 = fmaf(x0, a, b); x1 = fminf(x1, x0); x2 = fmaf(x2, a, b); x3 = fmaxf(x3, x2);
            x4 = fmaf(x4, a, b); x5 = fminf(x5, x4); x6 = fmaf(x6, a, b); x7 = fmaxf(x7, x6);
        } else if (MODE == 6) { // MUFU.RCP x4 + 4 FFMA
            asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(x1));
            asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(x3));
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        } else if (MODE == 7) { // FADD.SAT x8
            x0 = __saturatef(x0 + a); x1 = __saturatef(x1 + a); x2 = __saturatef(x2 + a); x3 = __saturatef(x3 + a);
            x4 = __saturatef(x4 + a); x5 = __saturatef(x5 + a); x6 = __saturatef(x6 + a); x7 = __saturatef(x7 + a);
        } else if (MODE == 8) { // 4 FFMA2 + 4 FMNMX
            float2 p0 = make_float2(x0, x1), p1 = make_float2(x2, x3);
            const float2 aa = make_float2(a, a), bb = make_float2(b, b);
            p0 = __ffma2_rn(p0, aa, bb); p1 = __ffma2_rn(p1, aa, bb);
            x0 = p0.x; x1 = p0.y; x2 = p1.x; x3 = p1.y;
            x4 = fminf(x4, x0); x5 = fmaxf(x5, x1); x6 = fminf(x6, x2); x7 = fmaxf(x7, x3);
            p0 = make_float2(x4, x5); p1 = make_float2(x6, x7);
            p0 = __ffma2_rn(p0, aa, bb); p1 = __ffma2_rn(p1, aa, bb);
            x4 = p0.x; x5 = p0.y; x6 = p1.x; x7 = p1.y;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char *name, int instr_per_iter, float *out, long long *cyc) {
    // 4 CTAs x 256 threads per SM: 32 warps / SM (8 per scheduler)
    k<MODE><<<148 * 4, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double winstr_per_sm = 32.0 * ITERS * instr_per_iter;       // 32 warps per SM
    printf("%-28s %8.3f ms  %10lld cyc  -> %.2f warp-instr/cyc/SM (%.2f per scheduler)\n", name, ms, c, winstr_per_sm / c, winstr_per_sm / c / 4);
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&cyc, 8);
    run<0>("FFMA x8", 8, out, cyc);
    run<1>("FFMA2 x4 (+pack movs?)", 4, out, cyc);
    run<2>("FADD x8", 8, out, cyc);
    run<3>("FSETP+FSEL x4 (8 instr)", 8, out, cyc);
    run<4>("FMNMX x8 + FADD x8", 16, out, cyc);
    run<5>("FFMA x4 + FMNMX x4", 8, out, cyc);
    run<6>("MUFU.RCP x4 + FFMA x4", 8, out, cyc);
    run<7>("FADD.SAT x8", 8, out, cyc);
    run<8>("FFMA2 x4 + FMNMX x4", 8, out, cyc);
    return 0;
}
