// sts_probe.cu -- how does a warp-wide STS.128 split into shared-memory wavefronts?  Each lane owns two neighbouring 16-byte
// items (2L, 2L+1) and stores them with two instructions; which lanes store their odd item first decides the conflicts.
//   nvcc -arch=sm_100a -O3 -o sts_probe tools/probes/sts_probe.cu && ./sts_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(long long* out, int iters) {
  __shared__ uint4 buf[2048];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* base = buf + warp * 64;
  uint4 a = make_uint4(lane, 1, 2, 3), b = make_uint4(lane, 5, 6, 7);
  int first, second;
  if (MODE == 0) { first = lane; second = 32 + lane; }                       // reference: consecutive items (ideal, 4 wavefronts)
  else {
    const bool sw = MODE == 1 ? false : MODE == 2 ? ((lane >> 2) & 1) : MODE == 3 ? ((lane >> 3) & 1) : MODE == 4 ? ((lane >> 4) & 1) : (lane & 1);
    first = 2 * lane + (sw ? 1 : 0);
    second = 2 * lane + (sw ? 0 : 1);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared(base + first)), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared(base + second)), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[MODE] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 4096, threads = 256;
  probe<0><<<1, threads>>>(d, iters); probe<1><<<1, threads>>>(d, iters); probe<2><<<1, threads>>>(d, iters);
  probe<3><<<1, threads>>>(d, iters); probe<4><<<1, threads>>>(d, iters); probe<5><<<1, threads>>>(d, iters);
  long long h[8];
  cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  const char* names[6] = {"consecutive items (reference)", "pairs, even item first in every lane", "pairs, odd first where lane bit 2",
                          "pairs, odd first where lane bit 3", "pairs, odd first where lane bit 4", "pairs, odd first where lane bit 0"};
  for (int m = 0; m < 6; ++m)
    printf("{\"pattern\": \"%s\", \"cycles_per_warp_store\": %.2f}\n", names[m], (double)h[m] / (2.0 * iters * (threads / 32)));
  return cudaGetLastError() != cudaSuccess;
}
