// FP64 pipe microbenchmarks for B200 (sm_100a): DMMA m8n8k4 / m16n8k16 peak, DFMA peak, DMMA+DADD mix,
// HBM streaming read.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double &c0,double &c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4],const double (&a)[8],const double (&b)[4]){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double (&c)[4],const double (&a)[4],const double (&b)[2]){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}

template<int NACC>
__global__ void k_dmma884(double *out,int iters,double a,double b){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma884(c[i][0],c[i][1],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void k_dmma16816(double *out,int iters,double a0,double b0){
  double c[NACC][4]; double a[8],b[4];
  #pragma unroll
  for(int i=0;i<8;i++)a[i]=a0+i;
  #pragma unroll
  for(int i=0;i<4;i++)b[i]=b0+i;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=threadIdx.x;c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma16816(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void k_dmma1688(double *out,int iters,double a0,double b0){
  double c[NACC][4]; double a[4],b[2];
  #pragma unroll
  for(int i=0;i<4;i++)a[i]=a0+i;
  #pragma unroll
  for(int i=0;i<2;i++)b[i]=b0+i;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=threadIdx.x;c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma1688(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void k_dfma(double *out,int iters,double a,double b){
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++)c[i]=threadIdx.x+i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// mix: NACC dmma884 + NADD dadd per iteration
template<int NACC,int NADD>
__global__ void k_mix(double *out,int iters,double a,double b){
  double c[NACC][2]; double e[NADD+1];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  #pragma unroll
  for(int i=0;i<NADD;i++) e[i]=i+threadIdx.x;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma884(c[i][0],c[i][1],a,b);
    #pragma unroll
    for(int i=0;i<NADD;i++) e[i]=__dadd_rn(e[i],a);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  #pragma unroll
  for(int i=0;i<NADD;i++) s+=e[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_read(const double2* __restrict__ x,size_t n,double *out){
  size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  double s=0;
  for(;i+3*stride<n;i+=4*stride){
    double2 v0=x[i],v1=x[i+stride],v2=x[i+2*stride],v3=x[i+3*stride];
    s+=v0.x+v0.y+v1.x+v1.y+v2.x+v2.y+v3.x+v3.y;
  }
  for(;i<n;i+=stride){double2 v=x[i];s+=v.x+v.y;}
  if(s==1.2345) out[0]=s;
}
__global__ void k_copy(const double2* __restrict__ x,double2* __restrict__ y,size_t n){
  size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  for(;i<n;i+=stride) y[i]=x[i];
}
template<class F> float timeit(F f,int rep=5){
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<rep;r++){cudaEventRecord(e0);f();cudaEventRecord(e1);CK(cudaEventSynchronize(e1));float ms;cudaEventElapsedTime(&ms,e0,e1);if(ms<best)best=ms;}
  return best;
}
int main(){
  cudaDeviceProp p;CK(cudaGetDeviceProperties(&p,0));
  printf("device %s sms=%d clock=%d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  int sms=p.multiProcessorCount;
  double *out;CK(cudaMalloc(&out,sizeof(double)*sms*8*1024));
  int iters=20000;
  for(int warps=4;warps<=16;warps*=2){
    int thr=warps*32; int blocks=sms*2;
    { float ms=timeit([&]{k_dmma884<8><<<blocks,thr>>>(out,iters,1.0,1e-9);});
      double fl=(double)blocks*warps*iters*8*(8*8*4*2); printf("dmma884   acc=8  warps/blk=%2d blocks=%d: %.2f TFLOP/s (%.3f ms)\n",warps,blocks,fl/ms*1e-9,ms);}
    { float ms=timeit([&]{k_dmma884<16><<<blocks,thr>>>(out,iters,1.0,1e-9);});
      double fl=(double)blocks*warps*iters*16*(8*8*4*2); printf("dmma884   acc=16 warps/blk=%2d blocks=%d: %.2f TFLOP/s (%.3f ms)\n",warps,blocks,fl/ms*1e-9,ms);}
    { float ms=timeit([&]{k_dmma16816<4><<<blocks,thr>>>(out,iters/4,1.0,1e-9);});
      double fl=(double)blocks*warps*(iters/4)*4*(16.0*8*16*2); printf("dmma16816 acc=4  warps/blk=%2d blocks=%d: %.2f TFLOP/s (%.3f ms)\n",warps,blocks,fl/ms*1e-9,ms);}
    { float ms=timeit([&]{k_dmma16816<8><<<blocks,thr>>>(out,iters/4,1.0,1e-9);});
      double fl=(double)blocks*warps*(iters/4)*8*(16.0*8*16*2); printf("dmma16816 acc=8  warps/blk=%2d blocks=%d: %.2f TFLOP/s (%.3f ms)\n",warps,blocks,fl/ms*1e-9,ms);}
    { float ms=timeit([&]{k_dmma1688<8><<<blocks,thr>>>(out,iters/2,1.0,1e-9);});
      double fl=(double)blocks*warps*(iters/2)*8*(16.0*8*8*2); printf("dmma1688  acc=8  warps/blk=%2d blocks=%d: %.2f TFLOP/s (%.3f ms)\n",warps,blocks,fl/ms*1e-9,ms);}
    { float ms=timeit([&]{k_dfma<16><<<blocks,thr>>>(out,iters,1.0000001,1e-9);});
      double fl=(double)blocks*thr*(double)iters*16*2; printf("dfma      acc=16 warps/blk=%2d blocks=%d: %.2f TFLOP/s (%.3f ms)\n",warps,blocks,fl/ms*1e-9,ms);}
    { float ms0=timeit([&]{k_mix<8,1><<<blocks,thr>>>(out,iters,1.0,1e-9);});
      float ms1=timeit([&]{k_mix<8,8><<<blocks,thr>>>(out,iters,1.0,1e-9);});
      float ms2=timeit([&]{k_mix<8,32><<<blocks,thr>>>(out,iters,1.0,1e-9);});
      printf("mix 8 dmma884 + {1,8,32} dadd warps/blk=%2d: %.3f / %.3f / %.3f ms\n",warps,ms0,ms1,ms2);}
  }
  size_t n=(size_t)1<<28; // 268M double2 = 4 GiB
  double2 *x,*y;CK(cudaMalloc(&x,n*16));CK(cudaMalloc(&y,n*16));CK(cudaMemset(x,0,n*16));CK(cudaMemset(y,0,n*16));
  for(int bps=2;bps<=16;bps*=2){
    float ms=timeit([&]{k_read<<<sms*bps,256>>>(x,n,out);});
    printf("read  4GiB blocks/sm=%2d: %.1f GB/s\n",bps,n*16.0/ms*1e-6);
    ms=timeit([&]{k_copy<<<sms*bps,256>>>(x,y,n);});
    printf("copy  4GiB blocks/sm=%2d: %.1f GB/s (r+w)\n",bps,2*n*16.0/ms*1e-6);
  }
  { float ms=timeit([&]{cudaMemcpyAsync(y,x,n*16,cudaMemcpyDeviceToDevice);});
    printf("cudaMemcpy D2D 4GiB: %.1f GB/s (r+w)\n",2*n*16.0/ms*1e-6);}
  return 0;
}
