// tma_probe.cu -- bisect harness for the TMA box load used by hg_resample_tma.cu (one CTA, one box).
//   ./tma_probe <variant> <BW> <BH> <col0> <row0>
// variant bit0: skip prefetch.tensormap, bit1: drop ".tile", bit2: issue from a warp-uniform branch (whole warp 0, elected lane)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int BW, int BH, int col0, int row0, int variant) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((BW * BH * 4 + 127) / 128) * 128);
  const bool issuer = (variant & 4) ? (threadIdx.x < 32) : (threadIdx.x == 0);
  if (threadIdx.x == 0) {
    if (!(variant & 1)) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (issuer) {
    bool lead = threadIdx.x == 0;
    if (variant & 4) {
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      lead = pred != 0;
    }
    if (lead) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BW * BH * 4) : "memory");
      if (variant & 2)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(bar)), "r"(col0), "r"(row0), "r"(0) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(bar)), "r"(col0), "r"(row0), "r"(0) : "memory");
    }
  }
  uint32_t ok = 0;
  long long spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    if (++spins > (1ll << 26)) { if (threadIdx.x == 0) printf("probe: mbarrier never completed\n"); return; }
  }
  const float* t = reinterpret_cast<const float*>(smem);
  for (int e = threadIdx.x; e < BW * BH; e += blockDim.x) out[e] = t[e];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0, BW = argc > 2 ? atoi(argv[2]) : 132, BH = argc > 3 ? atoi(argv[3]) : 35;
  const int col0 = argc > 4 ? atoi(argv[4]) : -1, row0 = argc > 5 ? atoi(argv[5]) : 0;
  const int W = 300, H = 200, P = 2;
  std::vector<float> h((size_t)W * H * P);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, (size_t)BW * BH * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  if (!fp) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((PFN_encodeTiled)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  if (r != CUDA_SUCCESS) return 3;
  const int smem = ((BW * BH * 4 + 127) / 128) * 128 + 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 256, smem>>>(tm, o, BW, BH, col0, row0, variant);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d box %dx%d at (%d,%d): %s\n", variant, BW, BH, col0, row0, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> res((size_t)BW * BH);
  cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < BH; ++y)
    for (int x = 0; x < BW; ++x) {
      const int gx = col0 + x, gy = row0 + y;
      const float exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)gy * W + gx] : 0.f;
      if (res[(size_t)y * BW + x] != exp) ++bad;
    }
  printf("mismatches: %d\n", bad);
  return bad ? 4 : 0;
}
