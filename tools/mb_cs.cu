// mb_cs.cu — microbenchmark: carry-save (independent-MAC) Montgomery multiply (mulcs.cuh) against the
// carry-chained multiply of fp.cuh, at the occupancies the kernels run at.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I threshold_crypto_b200/csrc -I tools -o tools/mb_cs tools/mb_cs.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "mulcs.cuh"
using namespace tcb;

template <int CH, bool DOT, bool CS>
__global__ void k_mul(Fp *out, int n, const Fp *in) {
    Fp x[CH], y = in[1];
    for (int i = 0; i < 12; i++) y.l[i] ^= (threadIdx.x << (i & 3));   // keep y out of the uniform registers
    y.l[11] &= 0x0fffffff;
    for (int c = 0; c < CH; c++) { x[c] = in[0]; x[c].l[0] += c + threadIdx.x; x[c].l[11] &= 0x0fffffff; }
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (CS) x[c] = mont_mul_cs<FpParams, DOT>(x[c], y, y, x[c]);
            else x[c] = DOT ? dot2(x[c], y, y, x[c]) : x[c] * y;
        }
    Fp s = x[0];
    for (int c = 1; c < CH; c++) s = s + x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// a mix closer to the tower code: dot2, then an add and a sub on the result (non-multiply work between products)
template <bool CS>
__global__ void k_mix(Fp *out, int n, const Fp *in) {
    Fp x = in[0], y = in[1], z = in[0];
    for (int i = 0; i < 12; i++) y.l[i] ^= (threadIdx.x << (i & 3));
    y.l[11] &= 0x0fffffff; x.l[0] += threadIdx.x; x.l[11] &= 0x0fffffff; z.l[11] &= 0x0fffffff;
    for (int i = 0; i < n; i++) {
        Fp t = CS ? mont_mul_cs<FpParams, true>(x, y, z, x) : dot2(x, y, z, x);
        Fp u = t + x;
        z = u - y;
        x = t + z;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + z;
}
__global__ void k_cmp(unsigned long long *bad, const Fp *in, int n) {
    Fp x = in[0], y = in[1];
    x.l[0] += threadIdx.x + blockIdx.x * 977; x.l[11] &= 0x0fffffff;
    int e = 0;
    for (int i = 0; i < n; i++) {
        Fp r1 = x * y, r2 = mont_mul_cs<FpParams, false>(x, y, x, y);
        if (r1 != r2) e++;
        Fp d1 = dot2(x, y, r1, x), d2 = mont_mul_cs<FpParams, true>(x, y, r1, x);
        if (d1 != d2) e++;
        x = r1; y = d1;
    }
    if (e) atomicAdd(bad, (unsigned long long)e);
}
// independent carry-out-only MACs + counters, nothing else: the issue rate of that instruction form
__global__ void k_pout_tput(u32 *out, int n, u32 a, u32 b) {
    u32 lo[8], hi[8], cnt[8];
    for (int k = 0; k < 8; k++) { lo[k] = threadIdx.x + k; hi[k] = k; cnt[k] = 0; }
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) mac_cnt(lo[k], hi[k], cnt[k], a + k, b + r);
    }
    u32 s = 0;
    for (int k = 0; k < 8; k++) s ^= lo[k] ^ hi[k] ^ cnt[k];
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
        unsigned long long *bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
        k_cmp<<<sms * 2, 128>>>(bad, din, 64);
        unsigned long long hb = 1; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
        printf("carry-save vs carry-chain multiply mismatches on device: %llu (%s)\n", hb, cudaGetErrorString(cudaGetLastError()));
        for (int w : {4, 8, 32}) {
            int n = 4096; float ms = timeit([&] { k_pout_tput<<<sms, w * 32>>>((u32 *)buf, n, 12345, 67890); });
            printf("carry-out-only IMAD.WIDE + counter, %2d warps/SM: %.2f TMAC/s (%.1f MAC/clk/SM)\n", w, (double)sms * w * 32 * n * 16 / (ms * 1e-3) / 1e12,
                   (double)w * 32 * n * 16 / (ms * 1e-3) / clk);
        }
    }
    for (int w : {4, 8, 12, 16, 32}) {
        int n = 512;
        auto per = [&](float ms, int ch) { return (double)sms * w * 32 * n * ch / (ms * 1e-3) / 1e10; };
        float m1 = timeit([&] { k_mul<1, false, false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float d1 = timeit([&] { k_mul<1, true, false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float c1 = timeit([&] { k_mul<1, false, true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float c2 = timeit([&] { k_mul<2, false, true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float e1 = timeit([&] { k_mul<1, true, true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float x0 = timeit([&] { k_mix<false><<<sms, w * 32>>>((Fp *)buf, n, din); });
        float x1 = timeit([&] { k_mix<true><<<sms, w * 32>>>((Fp *)buf, n, din); });
        printf("warps/SM %2d  ops/s (1e10): chain mul %.2f dot2 %.2f | carry-save mul %.2f (x2 %.2f) dot2 %.2f | mix(dot2+3 add/sub) chain %.2f cs %.2f\n", w,
               per(m1, 1), per(d1, 1), per(c1, 1), per(c2, 2), per(e1, 1), per(x0, 1), per(x1, 1));
    }
    return 0;
}
