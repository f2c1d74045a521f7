// Pipe-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2/FADD2/FMUL2, FMNMX, SHFL, and mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
#define U 8
template <int MODE>
__global__ void __launch_bounds__(1024) kern(float* out, float s) {
    float a[U], b[U];
    float2 p[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { a[k] = threadIdx.x * 0.001f + k; b[k] = s + k; p[k] = make_float2(a[k], b[k]); }
    const float2 m2 = make_float2(s, s * 0.5f), c2 = make_float2(0.25f * s, s);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (MODE == 0) a[k] = fmaf(a[k], s, b[k]);                       // FFMA
            if (MODE == 1) p[k] = __ffma2_rn(p[k], m2, c2);                   // FFMA2
            if (MODE == 2) a[k] = fmaxf(a[k], b[k] - (float)it);              // FMNMX + FADD
            if (MODE == 3) a[k] = __shfl_down_sync(0xffffffffu, a[k], 1);     // SHFL
            if (MODE == 4) { a[k] = fmaf(a[k], s, b[k]); b[k] = fminf(b[k], a[k]); }  // FFMA + FMNMX mix
            if (MODE == 5) { p[k] = __ffma2_rn(p[k], m2, c2); a[k] = fminf(a[k], p[k].x); }  // FFMA2 + FMNMX
            if (MODE == 6) p[k] = __fadd2_rn(p[k], c2);                       // FADD2
            if (MODE == 7) a[k] = a[k] + s;                                   // FADD
            if (MODE == 8) { p[k] = __ffma2_rn(p[k], m2, c2); a[k] = __shfl_down_sync(0xffffffffu, a[k], 1); }  // FFMA2 + SHFL
            if (MODE == 9) { a[k] = fminf(a[k], b[k]); b[k] = __shfl_down_sync(0xffffffffu, b[k], 1); }  // FMNMX + SHFL
        }
    }
    float r = 0;
#pragma unroll
    for (int k = 0; k < U; ++k) r += a[k] + b[k] + p[k].x + p[k].y;
    if (r == 123.456f) out[0] = r;
}
template <int MODE>
void run(const char* name, int inst_per_iter) {
    float* d; cudaMalloc(&d, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<sms * 2, 1024>>>(d, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<MODE><<<sms * 2, 1024>>>(d, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = (double)sms * 2 * 32 * ITER * U * inst_per_iter;
    double clk = ms * 1e-3 * khz * 1e3;
    printf("%-16s %8.3f ms  %6.2f warp-inst/clk/SM (nominal clock %d MHz)\n", name, ms, warp_inst / clk / sms, khz / 1000);
    cudaFree(d);
}
int main() {
    run<0>("FFMA", 1); run<1>("FFMA2", 1); run<2>("FMNMX+FADD", 2); run<3>("SHFL", 1); run<4>("FFMA+FMNMX", 2);
    run<5>("FFMA2+FMNMX", 2); run<6>("FADD2", 1); run<7>("FADD", 1); run<8>("FFMA2+SHFL", 2); run<9>("FMNMX+SHFL", 2);
    return 0;
}
