#include <stdint.h>
typedef uint32_t u32; typedef uint64_t u64;
__constant__ u32 P28c[14] = {0xfffaaab,0xfefffff,0x3ffffb9,0xfffeb15,0x6241eab,0xa0f6b0f,0xf6730d2,0xf38512b,0x4774b84,0x4bacd76,0xba7b643,0xe69a4b1,0x1ea397f,0x001a011};
#define INV28 0xffcfffdu
__device__ u32 g_zero_mem;
#define g_zero zero_
// native-28 representation: inputs/outputs are 14 x 28-bit limbs (no conversion), Montgomery R = 2^392
template<bool DOT2>
__device__ __forceinline__ void mul28n(u32* r, const u32* a, const u32* b, const u32* c, const u32* d, u32 zero_){
  u64 t[28];
  #pragma unroll
  for(int k=0;k<28;k++) t[k]=0;
  #pragma unroll
  for(int i=0;i<14;i++)
    #pragma unroll
    for(int j=0;j<14;j++){ t[i+j] += (u64)a[i]*b[j]; if(DOT2) t[i+j] += (u64)c[i]*d[j]; }
  u32 p[14];
  #pragma unroll
  for(int j=0;j<14;j++) p[j]=P28c[j] ^ (threadIdx.x & g_zero);
  #pragma unroll
  for(int i=0;i<14;i++){
    if(i) t[i] += t[i-1]>>28;
    u32 m = ((u32)t[i]*INV28) & 0x0fffffffu;
    asm volatile("" : "+r"(m));   // keep m a 32-bit value (otherwise the multiply is widened to 64x64)
    #pragma unroll
    for(int j=0;j<14;j++) t[i+j] += (u64)m*p[j];
  }
  u64 carry = t[13]>>28;
  #pragma unroll
  for(int k=0;k<14;k++){ u64 v=t[14+k]+carry; r[k]=(u32)v & 0x0fffffffu; carry=v>>28; }
}
template<bool DOT2>
__global__ void k_mul(u32* out, const u32* in, int n, u32 zero_){
  u32 x[14], y[14], z[14], w[14];
  for(int i=0;i<14;i++){ x[i]=in[threadIdx.x*28+i]&0xfffffff; y[i]=in[threadIdx.x*28+14+i]&0xfffffff; w[i]=(x[i]*7u)&0xfffffff; }
  for(int it=0; it<n; it++){ mul28n<DOT2>(z,x,y,w,x,zero_); for(int i=0;i<14;i++){ x[i]=z[i]; w[i]^=z[13-i];} }
  for(int i=0;i<14;i++) out[(blockIdx.x*blockDim.x+threadIdx.x)*14+i]=x[i];
}
template __global__ void k_mul<false>(u32*,const u32*,int,u32);
template __global__ void k_mul<true>(u32*,const u32*,int,u32);
#include <cstdio>
#include <cuda_runtime.h>
template <class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(){
  cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int sms=p.multiProcessorCount; double clk=p.clockRate*1e3;
  u32 *in,*out; cudaMalloc(&in, 1024*28*4); cudaMalloc(&out, (size_t)sms*1024*14*4*2);
  static u32 h[1024*28]; for(int i=0;i<1024*28;i++) h[i]=(u32)(i*2654435761u)>>4; cudaMemcpy(in,h,sizeof h,cudaMemcpyHostToDevice);
  int warps[]={4,8,16,32};
  for(int w: warps){ int n=512;
    float m=timeit([&]{k_mul<false><<<sms,w*32>>>(out,in,n,0);}); float d=timeit([&]{k_mul<true><<<sms,w*32>>>(out,in,n,0);});
    printf("native28 warps/SM %2d: mul %.2f e10/s  dot2 %.2f e10/s   (IMAD.WIDE/clk/SM: mul %.1f dot2 %.1f)\n", w, (double)sms*w*32*n/(m*1e-3)/1e10, (double)sms*w*32*n/(d*1e-3)/1e10,
      (double)w*32*n*400/(m*1e-3*clk), (double)w*32*n*596/(d*1e-3*clk)); }
  return 0; }
