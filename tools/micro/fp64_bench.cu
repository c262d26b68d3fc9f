// Micro-benchmark: throughput of DFMA, DADD, F2F (f32<->f64), FFMA on this GPU (per SM per clock) and their dependent latency.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k_tp(double* out, float* outf, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float f0 = threadIdx.x * 1e-3f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    const double c = 1.0000001, d = 1e-9;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { a0 = fma(a0, c, d); a1 = fma(a1, c, d); a2 = fma(a2, c, d); a3 = fma(a3, c, d); a4 = fma(a4, c, d); a5 = fma(a5, c, d); a6 = fma(a6, c, d); a7 = fma(a7, c, d); }
        if (OP == 1) { a0 += d; a1 += d; a2 += d; a3 += d; a4 += d; a5 += d; a6 += d; a7 += d; }
        if (OP == 2) { a0 = (double)f0 + a0; a1 = (double)f1 + a1; a2 = (double)f2 + a2; a3 = (double)f3 + a3; f0 = (float)a0; f1 = (float)a1; f2 = (float)a2; f3 = (float)a3; }  // 4 F2F.F64.F32 + 4 F2F.F32.F64 + 4 DADD
        if (OP == 3) { f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f); f4 = fmaf(f4, 1.0000001f, 1e-9f); f5 = fmaf(f5, 1.0000001f, 1e-9f); f6 = fmaf(f6, 1.0000001f, 1e-9f); f7 = fmaf(f7, 1.0000001f, 1e-9f); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    outf[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
template <int OP>
__global__ void k_lat(double* out, float* outf, int iters, long long* cycles) {
    double a = threadIdx.x * 1e-3;
    float f = threadIdx.x * 1e-3f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) a = fma(a, 1.0000001, 1e-9);
        if (OP == 1) f = fmaf(f, 1.0000001f, 1e-9f);
        if (OP == 2) { a = (double)f * 1.0000001; f = (float)a; }     // F2F + DMUL + F2F
        if (OP == 3) { a = a * 1.0000001; }                          // DMUL
        if (OP == 4) { f = (float)((double)f); }                     // F2F + F2F (may be optimised away)
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; outf[threadIdx.x] = f;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    double* d; float* f; long long* cy;
    cudaMalloc(&d, 1 << 24); cudaMalloc(&f, 1 << 24); cudaMalloc(&cy, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("%s SMs %d clock %.3f GHz\n", p.name, sms, ghz);
    const int iters = 20000, blocks = sms * 8, threads = 256;
    const char* names[4] = {"DFMA x8", "DADD x8", "F2F x8 + DADD x4", "FFMA x8"};
    const double ops[4] = {8, 8, 12, 8};
    for (int op = 0; op < 4; ++op) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) k_tp<0><<<blocks, threads>>>(d, f, iters);
            if (op == 1) k_tp<1><<<blocks, threads>>>(d, f, iters);
            if (op == 2) k_tp<2><<<blocks, threads>>>(d, f, iters);
            if (op == 3) k_tp<3><<<blocks, threads>>>(d, f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double total = (double)blocks * threads * iters * ops[op];
        printf("%-20s %.2f ms  -> %.1f thread-ops/clk/SM\n", names[op], ms, total / (ms * 1e-3) / (ghz * 1e9) / sms);
    }
    const char* ln[5] = {"DFMA dependent", "FFMA dependent", "F2F+DMUL+F2F dependent", "DMUL dependent", "F2F+F2F dependent"};
    for (int op = 0; op < 5; ++op) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) k_lat<0><<<1, 32>>>(d, f, 10000, cy);
            if (op == 1) k_lat<1><<<1, 32>>>(d, f, 10000, cy);
            if (op == 2) k_lat<2><<<1, 32>>>(d, f, 10000, cy);
            if (op == 3) k_lat<3><<<1, 32>>>(d, f, 10000, cy);
            if (op == 4) k_lat<4><<<1, 32>>>(d, f, 10000, cy);
            cudaMemcpy(&h, cy, 8, cudaMemcpyDeviceToHost);
        }
        printf("%-26s %.1f cycles/iter\n", ln[op], h / 10000.0);
    }
    return 0;
}
