// microbench.cu — latency / throughput probes behind the DESIGN.md numbers (not part of the library).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I threshold_crypto_b200/csrc -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fp.cuh"
#include "mul28.cuh"
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
template <int CH, bool DOT>
__global__ void k_mul28(Fp *out, int n, const Fp *in) {
    Fp x[CH], y = in[1];
    y.l[1] ^= threadIdx.x;   // keep y out of the uniform registers
    for (int c = 0; c < CH; c++) { x[c] = in[0]; x[c].l[0] += c + threadIdx.x; x[c].l[11] &= 0x0fffffff; }
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = mul28<DOT>(x[c], y, y, x[c]);
    Fp s = x[0];
    for (int c = 1; c < CH; c++) s = s + x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// equality of the two multipliers on device
__global__ void k_cmp28(unsigned long long *bad, const Fp *in, int n) {
    Fp x = in[0], y = in[1];
    x.l[0] += threadIdx.x + blockIdx.x * 977; x.l[11] &= 0x0fffffff;
    int e = 0;
    for (int i = 0; i < n; i++) {
        Fp r1 = x * y, r2 = mul28<false>(x, y, x, y);
        if (r1 != r2) e++;
        Fp d1 = dot2(x, y, r1, x), d2 = mul28<true>(x, y, r1, x);
        if (d1 != d2) e++;
        x = r1; y = d1;
    }
    if (e) atomicAdd(bad, (unsigned long long)e);
}
// 8 independent carry chains of length 2 (IMAD.WIDE.U32.X with predicate carries), nothing else
__global__ void k_carry_tput(u32 *out, int n, u32 a, u32 b) {
    u32 lo[8], hi[8], lo2[8], hi2[8];
    for (int k = 0; k < 8; k++) { lo[k] = threadIdx.x + k; hi[k] = k; lo2[k] = 3 * k; hi2[k] = 7; }
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { mad_pair_cc(lo[k], hi[k], a, b + k); madc_pair_cc(lo2[k], hi2[k], a + k, b); }
    }
    u32 s = 0;
    for (int k = 0; k < 8; k++) s ^= lo[k] ^ hi[k] ^ lo2[k] ^ hi2[k];
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
    {
        int n = 8192; float ms = timeit([&] { k_carry_tput<<<sms * 4, 256>>>((u32 *)buf, n, 12345, 67890); });
        printf("carry-chained IMAD.WIDE pairs, full occupancy: %.2f TMAC/s (%.1f MAC/clk/SM)\n", (double)sms * 4 * 256 * n * 16 / (ms * 1e-3) / 1e12,
               (double)sms * 4 * 256 * n * 16 / (ms * 1e-3) / clk / sms);
        unsigned long long *bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
        k_cmp28<<<sms * 2, 128>>>(bad, din, 64);
        unsigned long long hb = 1; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
        printf("mul28 vs carry-chain mul mismatches on device: %llu\n", hb);
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
        float a1 = timeit([&] { k_mul28<1, false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float a2 = timeit([&] { k_mul28<2, false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float b1 = timeit([&] { k_mul28<1, true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        auto per = [&](float ms, int ch) { return (double)sms * w * 32 * n * ch / (ms * 1e-3) / 1e10; };
        printf("            muls/s (1e10): carry x1 %.2f x2 %.2f dot2 %.2f | mul28 x1 %.2f x2 %.2f dot2 %.2f\n", per(m1, 1), per(m2, 2), per(d1, 1), per(a1, 1), per(a2, 2), per(b1, 1));
    }
    return 0;
}
