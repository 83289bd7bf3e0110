// micro-benchmarks: fr_mul throughput, raw IMAD / IMAD.WIDE / DFMA issue rates
#include "/root/repo/aztec-2.0_b200/csrc/field.cuh"
#include <cstdio>
using namespace bbg;
template<int ILP> __global__ void __launch_bounds__(256) k_mulchain(fr_t* o, int iters){
  fr_t x[ILP]; fr_t y;
  for(int k=0;k<ILP;k++) for(int i=0;i<8;i++) x[k].l[i]=threadIdx.x*7+i+k*13+1;
  for(int i=0;i<8;i++) y.l[i]=blockIdx.x+i*3+5;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int k=0;k<ILP;k++) x[k]=fe_mul(x[k],y);
  }
  fr_t s=x[0]; for(int k=1;k<ILP;k++) s=fe_add(s,x[k]);
  if(s.l[0]==0x12345678) fe_store(o+threadIdx.x,s);
}
__global__ void __launch_bounds__(256) k_imad(uint32_t* o,int iters){
  uint32_t a[8]; for(int i=0;i<8;i++) a[i]=threadIdx.x+i; uint32_t b=blockIdx.x*3+1;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int r=0;r<16;r++){
#pragma unroll
    for(int i=0;i<8;i++) a[i]=a[i]*b+a[(i+1)&7];
    }
  }
  uint32_t s=0; for(int i=0;i<8;i++) s+=a[i]; if(s==0x1234567) o[threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k_imadwide(uint64_t* o,int iters){
  uint64_t a[8]; for(int i=0;i<8;i++) a[i]=threadIdx.x+i; uint32_t b=blockIdx.x*3+1;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int r=0;r<16;r++){
#pragma unroll
    for(int i=0;i<8;i++) a[i]=(uint64_t)((uint32_t)a[i])*b+a[i];
    }
  }
  uint64_t s=0; for(int i=0;i<8;i++) s+=a[i]; if(s==0x1234567) o[threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k_dfma(double* o,int iters){
  double a[8]; for(int i=0;i<8;i++) a[i]=threadIdx.x+i; double b=blockIdx.x*3+1.000001;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int r=0;r<16;r++){
#pragma unroll
    for(int i=0;i<8;i++) a[i]=fma(a[i],b,a[i]);
    }
  }
  double s=0; for(int i=0;i<8;i++) s+=a[i]; if(s==0.1234567) o[threadIdx.x]=s;
}
template<class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(){
  void* o; cudaMalloc(&o, 1<<20);
  int blocks=148*8, iters=2000;
  for(int occ=0;occ<1;occ++){
  float ms;
  ms=timeit([&]{k_mulchain<1><<<blocks,256>>>((fr_t*)o,iters);}); printf("fr_mul ILP1: %.3f ms %.1f Gmul/s\n",ms,(double)blocks*256*iters*1/ms/1e6);
  ms=timeit([&]{k_mulchain<2><<<blocks,256>>>((fr_t*)o,iters);}); printf("fr_mul ILP2: %.3f ms %.1f Gmul/s\n",ms,(double)blocks*256*iters*2/ms/1e6);
  ms=timeit([&]{k_mulchain<4><<<blocks,256>>>((fr_t*)o,iters);}); printf("fr_mul ILP4: %.3f ms %.1f Gmul/s\n",ms,(double)blocks*256*iters*4/ms/1e6);
  ms=timeit([&]{k_imad<<<blocks,256>>>((uint32_t*)o,iters);}); printf("IMAD: %.3f ms %.2f Tops/s\n",ms,(double)blocks*256*iters*128/ms/1e9);
  ms=timeit([&]{k_imadwide<<<blocks,256>>>((uint64_t*)o,iters);}); printf("IMAD.WIDE: %.3f ms %.2f Tops/s\n",ms,(double)blocks*256*iters*128/ms/1e9);
  ms=timeit([&]{k_dfma<<<blocks,256>>>((double*)o,iters);}); printf("DFMA: %.3f ms %.2f Tops/s\n",ms,(double)blocks*256*iters*128/ms/1e9);
  }
  return 0;
}
