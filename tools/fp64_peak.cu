// FP64 peak micro-benchmark for B200 (sm_100a): DFMA pipe, DMMA (mma.sync f64) pipe, smem-fed DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// Prints one JSON line. Roofline denominator for the FP64 kernels (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters) {
  double a0=threadIdx.x*1e-9, a1=a0+1, a2=a0+2, a3=a0+3, a4=a0+4, a5=a0+5, a6=a0+6, a7=a0+7;
  double b=1.0000001, c=1e-9;
  for (int i=0;i<iters;i++){
    a0=fma(a0,b,c); a1=fma(a1,b,c); a2=fma(a2,b,c); a3=fma(a3,b,c);
    a4=fma(a4,b,c); a5=fma(a5,b,c); a6=fma(a6,b,c); a7=fma(a7,b,c);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}

__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
   : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma884(double* out, int iters) {
  double c[16]; for(int i=0;i<16;i++) c[i]=0;
  double a=threadIdx.x*1e-9, b=1.0000001;
  for (int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<8;j++) dmma884(c[2*j],c[2*j+1],a,b);
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
#ifdef TRY_M16
__device__ __forceinline__ void dmma1688(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
   : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__global__ void k_dmma1688(double* out, int iters) {
  double c[16]; for(int i=0;i<16;i++) c[i]=0;
  double a[4]={threadIdx.x*1e-9,1,2,3}, b[2]={1.0000001,0.5};
  for (int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<4;j++) dmma1688(c+4*j,a,b);
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
#endif
// DFMA with one operand from shared memory (broadcast LDS.64 per FMA) and with LDS.128 per 2 FMA
__global__ void k_dfma_lds(double* out, int iters) {
  __shared__ double v[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) v[i]=1.0+1e-9*i;
  __syncthreads();
  double a0=threadIdx.x*1e-9, a1=a0+1, a2=a0+2, a3=a0+3, a4=a0+4, a5=a0+5, a6=a0+6, a7=a0+7;
  double c=1e-9;
  for (int i=0;i<iters;i++){
    const double2* vv = (const double2*)&v[(i*8)&1023];
    double2 p=vv[0], q=vv[1], r=vv[2], s=vv[3];
    a0=fma(a0,p.x,c); a1=fma(a1,p.y,c); a2=fma(a2,q.x,c); a3=fma(a3,q.y,c);
    a4=fma(a4,r.x,c); a5=fma(a5,r.y,c); a6=fma(a6,s.x,c); a7=fma(a7,s.y,c);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}

template<class F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  int iters=20000;
  double res[8]; const char* names[8]; int n=0;
  for (int tpb : {256, 512, 1024}) {
    int blocks=sms*(2048/tpb);
    float ms=timeit([&]{k_dfma<<<blocks,tpb>>>(out,iters);});
    double tf=2.0*8*iters*(double)blocks*tpb/ms/1e9;
    printf("# dfma tpb=%d blocks=%d ms=%.3f TFLOPs=%.2f\n",tpb,blocks,ms,tf);
    if(tpb==512){res[n]=tf;names[n++]="dfma_tflops";}
  }
  for (int tpb : {128, 256, 512}) {
    int blocks=sms*(2048/tpb);
    float ms=timeit([&]{k_dmma884<<<blocks,tpb>>>(out,iters);});
    double tf=2.0*8*8*4*8*iters*(double)blocks*(tpb/32)/ms/1e9;
    printf("# dmma m8n8k4 tpb=%d ms=%.3f TFLOPs=%.2f\n",tpb,ms,tf);
    if(tpb==256){res[n]=tf;names[n++]="dmma884_tflops";}
  }
#ifdef TRY_M16
  for (int tpb : {128, 256, 512}) {
    int blocks=sms*(2048/tpb);
    float ms=timeit([&]{k_dmma1688<<<blocks,tpb>>>(out,iters);});
    double tf=2.0*16*8*8*4*iters*(double)blocks*(tpb/32)/ms/1e9;
    printf("# dmma m16n8k8 tpb=%d ms=%.3f TFLOPs=%.2f\n",tpb,ms,tf);
    if(tpb==256){res[n]=tf;names[n++]="dmma1688_tflops";}
  }
#endif
  for (int tpb : {256, 512}) {
    int blocks=sms*(2048/tpb);
    float ms=timeit([&]{k_dfma_lds<<<blocks,tpb>>>(out,iters);});
    double tf=2.0*8*iters*(double)blocks*tpb/ms/1e9;
    printf("# dfma+lds128 tpb=%d ms=%.3f TFLOPs=%.2f\n",tpb,ms,tf);
    if(tpb==512){res[n]=tf;names[n++]="dfma_lds128_tflops";}
  }
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_mhz\":%d",p.name,sms,p.clockRate/1000);
  for(int i=0;i<n;i++) printf(",\"%s\":%.2f",names[i],res[i]);
  printf("}\n");
  return 0;
}
