// Micro-benchmarks that size the kernel designs: FP32 scalar vs packed (FFMA2) issue
// rate, MUFU.LG2 rate, shared-memory bandwidth and a float4 streaming copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

template<int ILP> __global__ void k_ffma(float* out, int iters, float a, float b){
  float v[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) v[i]=threadIdx.x*0.001f+i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) v[i]=fmaf(v[i],a,b);
  }
  float s=0; 
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_ffma2(float2* out, int iters, float2 a, float2 b){
  float2 v[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) v[i]=make_float2(threadIdx.x*0.001f+i, i);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) v[i]=__ffma2_rn(v[i],a,b);
  }
  float2 s=make_float2(0,0);
  #pragma unroll
  for(int i=0;i<ILP;i++) s=__fadd2_rn(s,v[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// 3 distinct register sources (no immediates / constants)
template<int ILP> __global__ void k_ffma3(float* out, int iters, const float* in){
  float v[ILP], w[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++){ v[i]=in[threadIdx.x+i]; w[i]=in[threadIdx.x+i+32]; }
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) v[i]=fmaf(v[i],w[i],w[(i+1)%ILP]);
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_lg2(float* out, int iters){
  float v[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) v[i]=threadIdx.x+2.0f+i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) v[i]=__log2f(v[i])+3.0f;
  }
  float s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_smem(float* out, int iters){
  extern __shared__ float4 sm[];
  int t=threadIdx.x;
  for(int i=t;i<4096;i+=blockDim.x) sm[i]=make_float4(i,i,i,i);
  __syncthreads();
  float4 acc=make_float4(0,0,0,0);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int j=0;j<8;j++){
      float4 v=sm[(t+j*blockDim.x+it)&4095];
      acc.x+=v.x;acc.y+=v.y;acc.z+=v.z;acc.w+=v.w;
    }
  }
  out[blockIdx.x*blockDim.x+t]=acc.x+acc.y+acc.z+acc.w;
}
__global__ void k_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n){
  size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x, stride=(size_t)gridDim.x*blockDim.x;
  for(;i+3*stride<n;i+=4*stride){
    float4 a=__ldcs(in+i),b=__ldcs(in+i+stride),c=__ldcs(in+i+2*stride),d=__ldcs(in+i+3*stride);
    __stcs(out+i,a);__stcs(out+i+stride,b);__stcs(out+i+2*stride,c);__stcs(out+i+3*stride,d);
  }
  for(;i<n;i+=stride) out[i]=in[i];
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int clk; cudaDeviceGetAttribute(&clk,cudaDevAttrClockRate,0);
  printf("dev %s sms %d clock %d kHz\n",p.name,p.multiProcessorCount,clk);
  float* out; CK(cudaMalloc(&out,148*8*1024*8*4));
  float* in; CK(cudaMalloc(&in,4096)); CK(cudaMemset(in,0,4096));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid=148*4, block=512, iters=20000; float ms;
#define TIME(name, launch, opsPerThreadIter) \
  launch; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); launch; cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
  cudaEventElapsedTime(&ms,e0,e1); printf("%-28s %8.3f ms  %8.2f Tops/s (lane-ops)\n", name, ms, (double)grid*block*iters*(opsPerThreadIter)/ms/1e9);
  TIME("ffma imm/const ILP8", (k_ffma<8><<<grid,block>>>(out,iters,1.0001f,0.5f)), 8)
  TIME("ffma 3reg ILP8", (k_ffma3<8><<<grid,block>>>(out,iters,in)), 8)
  TIME("ffma2 ILP8 (x2 counted)", (k_ffma2<8><<<grid,block>>>((float2*)out,iters,make_float2(1.0001f,1.0002f),make_float2(0.5f,0.25f))), 16)
  TIME("ffma2 ILP4 (x2 counted)", (k_ffma2<4><<<grid,block>>>((float2*)out,iters,make_float2(1.0001f,1.0002f),make_float2(0.5f,0.25f))), 8)
  iters=4000;
  TIME("lg2+fadd ILP8 (lg2 count)", (k_lg2<8><<<grid,block>>>(out,iters)), 8)
  iters=2000; grid=148*2; block=512;
  CK(cudaFuncSetAttribute(k_smem,cudaFuncAttributeMaxDynamicSharedMemorySize,65536));

  k_smem<<<grid,block,65536>>>(out,iters); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); k_smem<<<grid,block,65536>>>(out,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms,e0,e1);
  printf("smem LDS.128: %.3f ms %.1f TB/s  (%.1f B/clk/SM at %d kHz)\n",ms,(double)grid*block*iters*8*16/ms/1e9, (double)grid*block*iters*8*16/(ms*1e-3)/148/(clk*1e3),clk);
  size_t n=(size_t)1<<28; // 256M float4 = 4 GiB
  n = (size_t)1<<27; // 2 GiB each
  float4 *a,*b; CK(cudaMalloc(&a,n*16)); CK(cudaMalloc(&b,n*16)); CK(cudaMemset(a,1,n*16));
  for(int g=1; g<=16; g*=2){
    k_copy<<<148*g,512>>>(a,b,n); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for(int r=0;r<5;r++) k_copy<<<148*g,512>>>(a,b,n); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms,e0,e1); ms/=5;
    printf("copy grid=148*%d: %.3f ms %.1f GB/s (r+w)\n",g,ms,2.0*n*16/ms/1e6);
  }
  CK(cudaMemcpy(b,a,n*16,cudaMemcpyDeviceToDevice)); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); for(int r=0;r<5;r++) cudaMemcpyAsync(b,a,n*16,cudaMemcpyDeviceToDevice); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms,e0,e1); ms/=5; printf("cudaMemcpy D2D: %.1f GB/s (r+w)\n",2.0*n*16/ms/1e6);
  // pinned host <-> device
  float* h; size_t hb=(size_t)1<<30; CK(cudaMallocHost(&h,hb));
  cudaEventRecord(e0); cudaMemcpyAsync(a,h,hb,cudaMemcpyHostToDevice); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms,e0,e1); printf("H2D pinned: %.1f GB/s\n",hb/ms/1e6);
  cudaEventRecord(e0); cudaMemcpyAsync(h,a,hb,cudaMemcpyDeviceToHost); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms,e0,e1); printf("D2H pinned: %.1f GB/s\n",hb/ms/1e6);
  cudaStream_t s1,s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEventRecord(e0); cudaMemcpyAsync(a,h,hb/2,cudaMemcpyHostToDevice,s1); cudaMemcpyAsync(h+hb/8,b,hb/2,cudaMemcpyDeviceToHost,s2);
  CK(cudaDeviceSynchronize()); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  cudaEventElapsedTime(&ms,e0,e1); printf("H2D+D2H concurrent: %.1f GB/s each dir\n",hb/2/ms/1e6);
  return 0;
}
