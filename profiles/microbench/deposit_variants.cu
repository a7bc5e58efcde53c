// Microbenchmark (round 1): which scatter primitive should the TSC deposit be built on?
// Synthetic: L^3 dense grid, L^3 particles laid out tile by tile (T^3 cells per tile, T^3 particles
// per tile, uniform inside the tile) = the locality a Hilbert-sorted particle array gives.
//   A  REDG.F32       27 scalar global reductions per particle
//   B  REDG.F32x4     9 rows x (1 or 2) aligned float4 global reductions per particle
//   C  ATOMS.ADD u32  shared-memory fixed-point tile (T+2)^3, flushed with REDG.F32
//   D  ATOMS.CAS f32  shared-memory float tile (CAS loop), flushed with REDG.F32
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o deposit_variants deposit_variants.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

template<int T>
__global__ void gen(float4* p, int L){
  long i = blockIdx.x*(long)blockDim.x+threadIdx.x; long n=(long)L*L*L; if(i>=n) return;
  const int TT=T*T*T; long tile=i/TT; int nt=L/T;
  int tx=tile%nt, ty=(tile/nt)%nt, tz=tile/((long)nt*nt);
  uint32_t h=hash32((uint32_t)i*3u+1u), h2=hash32((uint32_t)i*3u+2u), h3=hash32((uint32_t)i*3u+3u);
  float fx=(tx*T + (h>>8)*(1.0f/16777216.0f)*T)/L, fy=(ty*T+(h2>>8)*(1.0f/16777216.0f)*T)/L, fz=(tz*T+(h3>>8)*(1.0f/16777216.0f)*T)/L;
  p[i]=make_float4(fx,fy,fz,1.0f);
}

__device__ __forceinline__ void tsc(float x, int L, int& c, float w[3]){
  float t=x*L; c=(int)t; if(c>L-1) c=L-1; float s=t-(c+0.5f);
  w[0]=0.5f*(0.5f-s)*(0.5f-s); w[1]=0.75f-s*s; w[2]=0.5f*(0.5f+s)*(0.5f+s);
}

__global__ void dep_A(const float4* __restrict__ p, float* __restrict__ g, int L, long n){
  long i = blockIdx.x*(long)blockDim.x+threadIdx.x; if(i>=n) return;
  float4 q=p[i]; int cx,cy,cz; float wx[3],wy[3],wz[3]; tsc(q.x,L,cx,wx); tsc(q.y,L,cy,wy); tsc(q.z,L,cz,wz);
  #pragma unroll
  for(int k=0;k<3;k++){ int z=(cz+k-1+L)&(L-1);
    #pragma unroll
    for(int j=0;j<3;j++){ int y=(cy+j-1+L)&(L-1); float wyz=wy[j]*wz[k]; long row=((long)z*L+y)*L;
      #pragma unroll
      for(int a=0;a<3;a++){ int x=(cx+a-1+L)&(L-1); atomicAdd(&g[row+x], wyz*wx[a]); } } }
}

__global__ void dep_B(const float4* __restrict__ p, float* __restrict__ g, int L, long n){
  long i = blockIdx.x*(long)blockDim.x+threadIdx.x; if(i>=n) return;
  float4 q=p[i]; int cx,cy,cz; float wx[3],wy[3],wz[3]; tsc(q.x,L,cx,wx); tsc(q.y,L,cy,wy); tsc(q.z,L,cz,wz);
  // x cells cx-1..cx+1 -> aligned groups of 4
  int x0=(cx-1+L)&(L-1), g0=x0&~3, o=x0&3;          // o in 0..3 ; cells o,o+1,o+2 relative to g0 (may spill into next group)
  float a0[4]={0,0,0,0}, a1[4]={0,0,0,0};
  #pragma unroll
  for(int a=0;a<3;a++){ int r=o+a; if(r<4) a0[r]=wx[a]; else a1[r-4]=wx[a]; }
  int g1=(g0+4)&(L-1); bool two=(o>=2);
  #pragma unroll
  for(int k=0;k<3;k++){ int z=(cz+k-1+L)&(L-1);
    #pragma unroll
    for(int j=0;j<3;j++){ int y=(cy+j-1+L)&(L-1); float wyz=wy[j]*wz[k]; long row=((long)z*L+y)*L;
      atomicAdd((float4*)&g[row+g0], make_float4(a0[0]*wyz,a0[1]*wyz,a0[2]*wyz,a0[3]*wyz));
      if(two) atomicAdd((float4*)&g[row+g1], make_float4(a1[0]*wyz,a1[1]*wyz,a1[2]*wyz,a1[3]*wyz));
    } }
}

// tile kernels: one CTA per tile of T^3 cells; particles [tile*T^3, (tile+1)*T^3)
template<int T, bool FIXED>
__global__ void dep_tile(const float4* __restrict__ p, float* __restrict__ g, int L){
  constexpr int H=T+2, HH=H*H*H, TT=T*T*T;
  __shared__ uint32_t s[HH];
  for(int i=threadIdx.x;i<HH;i+=blockDim.x) s[i]=0;
  __syncthreads();
  int nt=L/T; long tile=blockIdx.x; int tx=tile%nt, ty=(tile/nt)%nt, tz=tile/((long)nt*nt);
  const float4* pp=p+tile*TT;
  for(int i=threadIdx.x;i<TT;i+=blockDim.x){
    float4 q=pp[i]; int cx,cy,cz; float wx[3],wy[3],wz[3]; tsc(q.x,L,cx,wx); tsc(q.y,L,cy,wy); tsc(q.z,L,cz,wz);
    int lx=min(max(cx-tx*T,0),T-1), ly=min(max(cy-ty*T,0),T-1), lz=min(max(cz-tz*T,0),T-1);     // 0..T-1 ; halo index +1-1 => lx+a
    #pragma unroll
    for(int k=0;k<3;k++)
      #pragma unroll
      for(int j=0;j<3;j++){ float wyz=wy[j]*wz[k]; int base=((lz+k)*H+(ly+j))*H+lx;
        #pragma unroll
        for(int a=0;a<3;a++){
          float v=wyz*wx[a];
          if(FIXED) atomicAdd(&s[base+a], (uint32_t)__float2uint_rn(v*4194304.0f));
          else atomicAdd((float*)&s[base+a], v);
        } }
  }
  __syncthreads();
  for(int i=threadIdx.x;i<HH;i+=blockDim.x){
    int hx=i%H, hy=(i/H)%H, hz=i/(H*H);
    int x=(tx*T+hx-1+L)&(L-1), y=(ty*T+hy-1+L)&(L-1), z=(tz*T+hz-1+L)&(L-1);
    float v = FIXED ? (float)s[i]*(1.0f/4194304.0f) : __uint_as_float(s[i]);
    if(v!=0.f) atomicAdd(&g[((long)z*L+y)*L+x], v);
  }
}

template<typename F> float timeit(F f, int reps, float* g, size_t gb){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best=1e30f;
  for(int r=0;r<reps;r++){ CK(cudaMemset(g,0,gb)); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(r>0 && ms<best) best=ms; }
  return best;
}

double checksum(const float* g, long n){ float* h=(float*)malloc(n*4); CK(cudaMemcpy(h,g,n*4,cudaMemcpyDeviceToHost)); double s=0; for(long i=0;i<n;i++) s+=h[i]; free(h); return s; }

int main(int argc,char**argv){
  int L = argc>1? atoi(argv[1]) : 256; long n=(long)L*L*L;
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0)); printf("device %s SMs %d L2 %d MB smem/SM %zu\n",pr.name,pr.multiProcessorCount,pr.l2CacheSize>>20,pr.sharedMemPerMultiprocessor);
  float4* p; float* g; CK(cudaMalloc(&p,n*16)); CK(cudaMalloc(&g,n*4));
  const int B=256; int nb=(int)((n+B-1)/B);
  double bytes = 16.0*n + 4.0*n;
  for(int pass=0; pass<2; pass++){
    if(pass==0) gen<16><<<nb,B>>>(p,L); else gen<8><<<nb,B>>>(p,L);
    CK(cudaDeviceSynchronize());
    int T = pass==0?16:8;
    float tA=timeit([&]{dep_A<<<nb,B>>>(p,g,L,n);},4,g,n*4); double cA=checksum(g,n); printf("A done %f\n",tA); fflush(stdout);
    float tB=timeit([&]{dep_B<<<nb,B>>>(p,g,L,n);},4,g,n*4); double cB=checksum(g,n); printf("B done %f\n",tB); fflush(stdout);
    float tC,tD; double cC,cD;
    if(T==16){ long nt=n/4096; tC=timeit([&]{dep_tile<16,true><<<nt,512>>>(p,g,L);},4,g,n*4); cC=checksum(g,n); tD=timeit([&]{dep_tile<16,false><<<nt,512>>>(p,g,L);},4,g,n*4); cD=checksum(g,n);} 
    else { long nt=n/512; tC=timeit([&]{dep_tile<8,true><<<nt,256>>>(p,g,L);},4,g,n*4); cC=checksum(g,n); tD=timeit([&]{dep_tile<8,false><<<nt,256>>>(p,g,L);},4,g,n*4); cD=checksum(g,n);} 
    printf("L=%d tile-locality T=%d  n=%ld  roofline bytes %.1f MB\n",L,T,n,bytes/1e6);
    printf("  A REDG.F32 x27        %8.3f ms  %7.2f Gpart/s  %6.1f GB/s-alg  sum %.1f\n",tA,n/tA/1e6,bytes/tA/1e6,cA);
    printf("  B REDG.F32x4 x9..18   %8.3f ms  %7.2f Gpart/s  %6.1f GB/s-alg  sum %.1f\n",tB,n/tB/1e6,bytes/tB/1e6,cB);
    printf("  C smem u32 ATOMS.ADD  %8.3f ms  %7.2f Gpart/s  %6.1f GB/s-alg  sum %.1f\n",tC,n/tC/1e6,bytes/tC/1e6,cC);
    printf("  D smem f32 CAS        %8.3f ms  %7.2f Gpart/s  %6.1f GB/s-alg  sum %.1f\n",tD,n/tD/1e6,bytes/tD/1e6,cD);
  }
  return 0;
}
