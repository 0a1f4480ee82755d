// Microbenchmark: per-SM throughput of FFMA, DFMA, mma.sync TF32 (m16n8k8), mma.sync F64 (m8n8k4),
// shared-memory LDS.128, on this GPU. Development aid for DESIGN.md (which pipe bounds the LU).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e));return 1;}}while(0)

__global__ void k_ffma(float* out, int iters) {
  float a[8]; for (int i=0;i<8;++i) a[i]=threadIdx.x*0.001f+i;
  float b=1.0001f, c=0.5f;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<8;++i) a[i]=fmaf(a[i],b,c);
  }
  float s=0; for (int i=0;i<8;++i) s+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_dfma(double* out, int iters) {
  double a[8]; for (int i=0;i<8;++i) a[i]=threadIdx.x*0.001+i;
  double b=1.0001, c=0.5;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<8;++i) a[i]=fma(a[i],b,c);
  }
  double s=0; for (int i=0;i<8;++i) s+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_mma_tf32(float* out, int iters) {
  float c[4][4]; for (int i=0;i<4;++i) for (int j=0;j<4;++j) c[i][j]=0;
  unsigned a0=__float_as_uint(1.0f+threadIdx.x*1e-3f), a1=a0, a2=a0, a3=a0, b0=__float_as_uint(0.5f), b1=b0;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<4;++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[i][0]),"+f"(c[i][1]),"+f"(c[i][2]),"+f"(c[i][3]) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b0),"r"(b1));
  }
  float s=0; for (int i=0;i<4;++i) for (int j=0;j<4;++j) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_mma_bf16(float* out, int iters) {
  float c[4][4]; for (int i=0;i<4;++i) for (int j=0;j<4;++j) c[i][j]=0;
  unsigned a0=0x3f803f80u+threadIdx.x, a1=a0, a2=a0, a3=a0, b0=0x3f003f00u, b1=b0;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<4;++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[i][0]),"+f"(c[i][1]),"+f"(c[i][2]),"+f"(c[i][3]) : "r"(a0),"r"(a1),"r"(a2),"r"(a3),"r"(b0),"r"(b1));
  }
  float s=0; for (int i=0;i<4;++i) for (int j=0;j<4;++j) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_mma_f64(double* out, int iters) {
  double c[4][2]; for (int i=0;i<4;++i) for (int j=0;j<2;++j) c[i][j]=0;
  double a=1.0+threadIdx.x*1e-3, b=0.5;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<4;++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0; for (int i=0;i<4;++i) for (int j=0;j<2;++j) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_lds128(float* out, int iters) {
  extern __shared__ float4 sm4[];
  for (int i=threadIdx.x;i<4096;i+=blockDim.x) sm4[i]=make_float4(i,1,2,3);
  __syncthreads();
  float4 acc=make_float4(0,0,0,0);
  int idx=threadIdx.x;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<8;++i) { float4 v=sm4[(idx+i*512)&4095]; acc.x+=v.x; acc.y+=v.y; acc.z+=v.z; acc.w+=v.w; }
    idx=(idx+33)&4095;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc.x+acc.y+acc.z+acc.w;
}
template <typename F> float timeit(F f) { cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main() {
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  void* out; CK(cudaMalloc(&out, 148*4*1024*8));
  const int iters=20000; const int blocks=sms*2, threads=512;
  float ms;
  ms=timeit([&]{k_ffma<<<blocks,threads>>>((float*)out,iters);});
  printf("FFMA   : %.1f TFLOP/s  (%.1f FMA/clk/SM @%d MHz nominal)\n", 2.0*blocks*threads*8.0*iters/ms/1e9, blocks*threads*8.0*iters/(ms*1e-3)/sms/(clk*1e3), clk/1000);
  ms=timeit([&]{k_dfma<<<blocks,threads>>>((double*)out,iters);});
  printf("DFMA   : %.1f TFLOP/s  (%.1f FMA/clk/SM)\n", 2.0*blocks*threads*8.0*iters/ms/1e9, blocks*threads*8.0*iters/(ms*1e-3)/sms/(clk*1e3));
  ms=timeit([&]{k_mma_tf32<<<blocks,threads>>>((float*)out,iters);});
  { double macs=(double)blocks*(threads/32)*4.0*iters*16*8*8; printf("MMA.TF32 m16n8k8 : %.1f TFLOP/s (%.1f MAC/clk/SM)\n", 2*macs/ms/1e9, macs/(ms*1e-3)/sms/(clk*1e3)); }
  ms=timeit([&]{k_mma_bf16<<<blocks,threads>>>((float*)out,iters);});
  { double macs=(double)blocks*(threads/32)*4.0*iters*16*8*16; printf("MMA.BF16 m16n8k16: %.1f TFLOP/s (%.1f MAC/clk/SM)\n", 2*macs/ms/1e9, macs/(ms*1e-3)/sms/(clk*1e3)); }
  ms=timeit([&]{k_mma_f64<<<blocks,threads>>>((double*)out,iters);});
  { double macs=(double)blocks*(threads/32)*4.0*iters*8*8*4; printf("MMA.F64 m8n8k4   : %.1f TFLOP/s (%.1f MAC/clk/SM)\n", 2*macs/ms/1e9, macs/(ms*1e-3)/sms/(clk*1e3)); }
  ms=timeit([&]{k_lds128<<<sms,512,65536>>>((float*)out,iters);});
  { double bytes=(double)sms*512*8.0*iters*16; printf("LDS.128: %.1f TB/s (%.1f B/clk/SM)\n", bytes/ms/1e9, bytes/(ms*1e-3)/sms/(clk*1e3)); }
  return 0;
}
