// Micro-probe for TMA tensor load / reduce-add on sm_100a (debug tool, not part of the library).
// usage: tma_probe <mode> <bz> <by> <bx> <cz> <cy> <cx> <n>
//   mode 0 = load 3d, 1 = reduce-add 3d, 2 = 1-D bulk reduce-add (rows), 3 = load 4d
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int mode, int nbox, int c0, int c1, int c2, float* out) {
  extern __shared__ __align__(128) float box[];
  __shared__ __align__(8) unsigned long long bar;
  if (mode == 0 || mode == 3) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbox * 4) : "memory");
      if (mode == 0)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          ::"r"(smem_u32(box)), "l"((unsigned long long)&tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
          ::"r"(smem_u32(box)), "l"((unsigned long long)&tm), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    unsigned ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < nbox; i += blockDim.x) out[i] = box[i];
  } else if (mode == 1) {
    for (int i = threadIdx.x; i < nbox; i += blockDim.x) box[i] = 1.0f + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
        ::"l"((unsigned long long)&tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(box)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else if (mode == 2) {
    // 1-D bulk reduce: out (global, 16-B aligned) += box[0:nbox]
    for (int i = threadIdx.x; i < nbox; i += blockDim.x) box[i] = 1.0f + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
        ::"l"(out), "r"(smem_u32(box)), "r"(nbox * 4) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

int main(int argc, char** argv) {
  if (argc < 9) { printf("args\n"); return 2; }
  int mode = atoi(argv[1]), bz = atoi(argv[2]), by = atoi(argv[3]), bx = atoi(argv[4]);
  int cz = atoi(argv[5]), cy = atoi(argv[6]), cx = atoi(argv[7]), n = atoi(argv[8]);
  const int rank = mode == 3 ? 4 : 3;
  const long long nc = (long long)n * n * n;
  float* d; CK(cudaMalloc(&d, 3 * nc * 4));
  std::vector<float> h(3 * nc);
  for (long long i = 0; i < 3 * nc; ++i) h[i] = (float)(i % 1000);
  CK(cudaMemcpy(d, h.data(), 3 * nc * 4, cudaMemcpyHostToDevice));
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n, 3};
  cuuint64_t str[3] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4, (cuuint64_t)nc * 4};
  cuuint32_t box[4] = {(cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, 3}, es[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeFn)sym)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int nbox = bz * by * bx * (mode == 3 ? 3 : 1);
  float* out; CK(cudaMalloc(&out, nbox * 4)); CK(cudaMemset(out, 0, nbox * 4));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe<<<1, 256, nbox * 4>>>(tm, mode, nbox, cz, cy, cx, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d box (%d,%d,%d) at (%d,%d,%d) n=%d: %s", mode, bz, by, bx, cz, cy, cx, n, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    // verify
    std::vector<float> ho(nbox), hm(3 * nc);
    cudaMemcpy(ho.data(), out, nbox * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hm.data(), d, 3 * nc * 4, cudaMemcpyDeviceToHost);
    long long bad = 0;
    if (mode == 0 || mode == 3) {
      for (int f = 0; f < (mode == 3 ? 3 : 1); ++f)
      for (int x = 0; x < bx; ++x) for (int y = 0; y < by; ++y) for (int z = 0; z < bz; ++z) {
        int gx = cx + x, gy = cy + y, gz = cz + z;
        float want = (gx < 0 || gy < 0 || gz < 0 || gx >= n || gy >= n || gz >= n) ? 0.f : h[f * nc + ((long long)gx * n + gy) * n + gz];
        if (ho[((f * bx + x) * by + y) * bz + z] != want) ++bad;
      }
    } else if (mode == 1) {
      for (int x = 0; x < n; ++x) for (int y = 0; y < n; ++y) for (int z = 0; z < n; ++z) {
        int lx = x - cx, ly = y - cy, lz = z - cz;
        float add = (lx >= 0 && ly >= 0 && lz >= 0 && lx < bx && ly < by && lz < bz) ? 1.0f + ((lx * by + ly) * bz + lz) : 0.f;
        if (hm[((long long)x * n + y) * n + z] != h[((long long)x * n + y) * n + z] + add) ++bad;
      }
    } else {
      for (int i = 0; i < nbox; ++i) if (ho[i] != 1.0f + i) ++bad;
    }
    printf("  mismatches=%lld", bad);
  }
  printf("\n");
  return 0;
}
