// Micro-benchmarks that decide the paint/read kernel design on B200 (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_smem micro_smem.cu && ./micro_smem
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

constexpr int TILE = 17*17*17;   // 4913 floats
__device__ __forceinline__ unsigned hash(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// MODE 0: smem float atomicAdd (CAS loop); 1: plain LDS+FADD+STS (racy upper bound); 2: smem int atomicAdd;
// MODE 3: LDS gather only (read-side)
template<int MODE, bool CLUSTERED>
__global__ void __launch_bounds__(256) k_smem(float* out, int iters){
  __shared__ float tile[TILE];
  for(int i=threadIdx.x;i<TILE;i+=256) tile[i]=0.f;
  __syncthreads();
  unsigned s = hash(blockIdx.x*256+threadIdx.x+1);
  float acc=0.f;
  for(int it=0; it<iters; ++it){
    s = hash(s);
    // a "particle": base cell in 16^3, 8 corners
    int cx = CLUSTERED ? (s&3) : (s&15), cy = CLUSTERED ? ((s>>4)&3) : ((s>>4)&15), cz=(s>>8)&15;
    float w = (float)(s>>24)*(1.f/256.f);
    #pragma unroll
    for(int c=0;c<8;++c){
      int a = ((cx+(c&1))*17 + (cy+((c>>1)&1)))*17 + cz+((c>>2)&1);
      if(MODE==0) atomicAdd(&tile[a], w);
      else if(MODE==1) tile[a] += w;
      else if(MODE==2) atomicAdd((int*)&tile[a], 1);
      else acc += tile[a]*w;
    }
  }
  __syncthreads();
  if(MODE==3) out[blockIdx.x*256+threadIdx.x]=acc;
  else for(int i=threadIdx.x;i<TILE;i+=256) if(tile[i]!=0.f) atomicAdd(&out[i], tile[i]);
}

// 64-bit fixed-point accumulation in shared memory (ATOMS.ADD.64?)
template<bool CLUSTERED>
__global__ void __launch_bounds__(256) k_smem64(float* out, int iters){
  __shared__ unsigned long long tile[TILE];
  for(int i=threadIdx.x;i<TILE;i+=256) tile[i]=0ull;
  __syncthreads();
  unsigned s = hash(blockIdx.x*256+threadIdx.x+1);
  for(int it=0; it<iters; ++it){
    s = hash(s);
    int cx = CLUSTERED ? (s&3) : (s&15), cy = CLUSTERED ? ((s>>4)&3) : ((s>>4)&15), cz=(s>>8)&15;
    float w = (float)(s>>24)*(1.f/256.f);
    #pragma unroll
    for(int c=0;c<8;++c){
      int a = ((cx+(c&1))*17 + (cy+((c>>1)&1)))*17 + cz+((c>>2)&1);
      atomicAdd(&tile[a], (unsigned long long)__float2ll_rn(w*4294967296.f));
    }
  }
  __syncthreads();
  for(int i=threadIdx.x;i<TILE;i+=256) if(tile[i]!=0ull) atomicAdd(&out[i], (float)((double)(long long)tile[i]*2.3283064365386963e-10));
}

// global atomics: REDG f32 / int with return, scattered over `span` floats (L2-resident or DRAM-resident)
template<int MODE>
__global__ void __launch_bounds__(256) k_glob(float* mesh, size_t span, int iters, int* sink){
  unsigned s = hash(blockIdx.x*256+threadIdx.x+1);
  int r=0;
  for(int it=0; it<iters; ++it){
    s = hash(s);
    size_t a = ((size_t)s * 2654435761ULL) % span;
    if(MODE==0) atomicAdd(&mesh[a], 1.0f);
    else if(MODE==1) r += atomicAdd((int*)&mesh[a], 1);
    else if(MODE==2) { // 8-corner pattern: 4 rows x 2 adjacent
      size_t b = a % (span-600000);
      #pragma unroll
      for(int c=0;c<8;++c) atomicAdd(&mesh[b + (c&1) + ((c>>1)&1)*512 + ((c>>2)&1)*262144], 1.0f);
    }
  }
  if(MODE==1 && r==0x7fffffff) *sink=r;
}

template<class F> float timeit(F f, int rep=5){
  f(); CK(cudaDeviceSynchronize());
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); for(int i=0;i<rep;++i) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms,a,b); return ms/rep;
}

int main(){
  float* out; CK(cudaMalloc(&out, 1<<26));
  int* sink; CK(cudaMalloc(&sink,4));
  const int blocks=148*8, iters=2000;
  const double nupd = (double)blocks*256*iters*8;
  #define RUN(M,C,name) { float ms=timeit([&]{k_smem<M,C><<<blocks,256>>>(out,iters);}); \
     printf("%-44s %8.3f ms  %7.2f G upd/s  %6.2f cyc/particle/SM @1.9GHz\n", name, ms, nupd/ms*1e-6, ms*1e-3*1.9e9*148/(nupd/8)); }
  RUN(0,false,"smem atomicAdd f32 (CAS loop), random 16^3");
  RUN(0,true, "smem atomicAdd f32 (CAS loop), clustered 4x4x16");
  RUN(1,false,"smem plain RMW (racy bound), random");
  RUN(2,false,"smem atomicAdd int (native), random");
  RUN(2,true, "smem atomicAdd int (native), clustered");
  RUN(3,false,"smem LDS gather x8, random");
  { float ms=timeit([&]{k_smem64<false><<<blocks,256>>>(out,iters);}); printf("%-44s %8.3f ms  %7.2f G upd/s  %6.2f cyc/particle/SM\n","smem atomicAdd u64 fixed-point, random",ms,nupd/ms*1e-6, ms*1e-3*1.9e9*148/(nupd/8)); }
  { float ms=timeit([&]{k_smem64<true><<<blocks,256>>>(out,iters);}); printf("%-44s %8.3f ms  %7.2f G upd/s  %6.2f cyc/particle/SM\n","smem atomicAdd u64 fixed-point, clustered",ms,nupd/ms*1e-6, ms*1e-3*1.9e9*148/(nupd/8)); }
  size_t big=(size_t)512*512*512, small=(size_t)256*256*64;
  float* mesh; CK(cudaMalloc(&mesh,big*4)); CK(cudaMemset(mesh,0,big*4));
  const int gb=148*16, gi=200; double ng=(double)gb*256*gi;
  #define RUNG(M,span,mult,name) { float ms=timeit([&]{k_glob<M><<<gb,256>>>(mesh,span,gi,sink);}); \
     printf("%-44s %8.3f ms  %7.2f G atom/s\n", name, ms, ng*mult/ms*1e-6); }
  RUNG(0,small,1,"REDG f32 random, 16 MiB span (L2)");
  RUNG(0,big,1,  "REDG f32 random, 512 MiB span (DRAM)");
  RUNG(1,small,1,"ATOMG int+return random, 16 MiB span");
  RUNG(1,big,1,  "ATOMG int+return random, 512 MiB span");
  RUNG(2,small,8,"REDG f32 8-corner pattern, 16 MiB span");
  RUNG(2,big,8,  "REDG f32 8-corner pattern, 512 MiB span");
  return 0;
}
