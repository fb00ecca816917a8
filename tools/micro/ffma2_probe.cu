// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) throughput, and a mixed issue test.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float b, float c) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    float2 bb = make_float2(b, b), cc = make_float2(c, c);
    int acc = threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c; a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c; }
        if (MODE == 1) { p0 = __ffma2_rn(p0, bb, cc); p1 = __ffma2_rn(p1, bb, cc); p2 = __ffma2_rn(p2, bb, cc); p3 = __ffma2_rn(p3, bb, cc); }
        if (MODE == 2) { // packed FMAs + as many integer ALU ops: does packing free issue slots?
            p0 = __ffma2_rn(p0, bb, cc); acc = acc * 3 + i; p1 = __ffma2_rn(p1, bb, cc); acc ^= (acc >> 3);
            p2 = __ffma2_rn(p2, bb, cc); acc += i & 7; p3 = __ffma2_rn(p3, bb, cc); acc = (acc << 1) | (acc >> 31);
        }
        if (MODE == 3) { // scalar FMAs + the same integer work
            a0 = a0 * b + c; a1 = a1 * b + c; acc = acc * 3 + i; a2 = a2 * b + c; a3 = a3 * b + c; acc ^= (acc >> 3);
            a4 = a4 * b + c; a5 = a5 * b + c; acc += i & 7; a6 = a6 * b + c; a7 = a7 * b + c; acc = (acc << 1) | (acc >> 31);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + acc;
}
template <int MODE> void run(const char* name, float* d, int blocks) {
    int iters = 1 << 15;
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 1e-3f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double fl = 2.0 * 8 * (double)iters * blocks * 256;
    printf("%-28s %.3f ms  %.1f TFLOP/s (FMA flops only)\n", name, best, fl / (best * 1e-3) / 1e12);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8; float* d; cudaMalloc(&d, sizeof(float) * blocks * 256);
    run<0>("scalar FFMA x8", d, blocks); run<1>("packed FFMA2 x4", d, blocks);
    run<2>("FFMA2 x4 + int ALU", d, blocks); run<3>("FFMA x8 + int ALU", d, blocks);
    return 0;
}
