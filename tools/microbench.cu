// microbench.cu — latency / throughput probes behind the DESIGN.md numbers (not part of the library).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I threshold_crypto_b200/csrc -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace tcb;

__global__ void k_chain_carry(u32 *out, int n, u32 a, u32 b) {      // one carry chain of IMAD.WIDE.X
    u32 lo = threadIdx.x, hi = 1;
    for (int i = 0; i < n; i++) {
        mad_pair_cc(lo, hi, a, b);
#pragma unroll
        for (int k = 0; k < 15; k++) madc_pair_cc(lo, hi, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = lo ^ hi;
}
__global__ void k_chain_acc(u64 *out, int n, u32 a, u32 b) {         // dependent through the accumulator only
    u64 acc = threadIdx.x;
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) acc = (u64)a * b + acc, a += (u32)acc;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int CH, bool DOT>
__global__ void k_mul(Fp *out, int n, const Fp *in) {
    Fp x[CH], y = in[1];
    for (int c = 0; c < CH; c++) { x[c] = in[0]; x[c].l[0] += c + threadIdx.x; x[c].l[11] &= 0x0fffffff; }
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = DOT ? dot2(x[c], y, y, x[c]) : x[c] * y;
    Fp s = x[0];
    for (int c = 1; c < CH; c++) s = s + x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double clk = p.clockRate * 1e3;
    printf("SMs %d clock %.0f MHz\n", sms, clk / 1e6);
    void *buf; cudaMalloc(&buf, 64 << 20);
    Fp h[2]; for (int i = 0; i < 12; i++) { h[0].l[i] = 0x01234567u * (i + 1); h[1].l[i] = 0x089abcdfu * (i + 3); } h[0].l[11] &= 0x0fffffff; h[1].l[11] &= 0x0fffffff;
    Fp *din; cudaMalloc(&din, sizeof h); cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
    {
        int n = 4096; float ms = timeit([&] { k_chain_carry<<<sms, 32>>>((u32 *)buf, n, 12345, 67890); });
        printf("IMAD.WIDE.X carry chain, 1 warp/SM: %.2f cycles per dependent step\n", ms * 1e-3 * clk / (n * 16.0));
        ms = timeit([&] { k_chain_acc<<<sms, 32>>>((u64 *)buf, n, 12345, 67890); });
        printf("IMAD.WIDE accumulator chain (+IADD), 1 warp/SM: %.2f cycles per step\n", ms * 1e-3 * clk / (n * 16.0));
    }
    int warps[] = {4, 8, 16, 32};
    for (int w : warps) {
        int n = 512;
        float m1 = timeit([&] { k_mul<1, false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float m2 = timeit([&] { k_mul<2, false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float d1 = timeit([&] { k_mul<1, true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float d2 = timeit([&] { k_mul<2, true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        auto rate = [&](float ms, int ch, int macs) { return (double)sms * w * 32 * n * ch * macs / (ms * 1e-3) / 1e12; };
        printf("warps/SM %2d: mul x1 %.2f  mul x2 %.2f  dot2 x1 %.2f  dot2 x2 %.2f  TMAC/s   (cycles/mul x1: %.0f)\n", w,
               rate(m1, 1, 300), rate(m2, 2, 300), rate(d1, 1, 444), rate(d2, 2, 444), m1 * 1e-3 * clk / n);
    }
    return 0;
}
