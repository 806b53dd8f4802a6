// hg_conv_wgrad_umma.cu -- weight / bias gradient of the 7-tap hex convolution on tcgen05 (sm_100a).
//
//   gw[co,ci,k] += sum_{n,R,q} gy[n,co,R,q] * P[n,ci, R + ro[k], q + co[R&1][k]]     (P = padded x, hg_conv.cuh)
//   gb[co]      += sum_{n,R,q} gy[n,co,R,q]
//
// GEMM view per tap: D_k[co, ci] += A[co, pixels] * B_k[ci, pixels]^T with the PIXELS as the reduction
// dimension.  Both operands are kept in shared memory exactly like the forward kernel's input rows --
// [channels/8][pixel][8 channels], one 16-byte unit per pixel -- but are handed to the tensor core as
// MN-major matrices (channels = M/N, contiguous 8-channel units; pixels = K, 16 bytes apart), so a tap's
// column shift is again a plain 16-byte offset of the descriptor start address and one staged x row
// serves all 7 taps of the three gy rows that touch it.  The 7 accumulators (7 x Cin fp32 columns) plus a
// 16-column bias accumulator (B = a block of ones) live in TMEM for the whole kernel; each persistent CTA
// reduces its share of (image, row band, column tile) items and adds its partial to gw / gb with one
// round of fp32 atomics at the end.  UMMA M is 128: rows 64..127 of the A view run past the Cout = 64
// channel groups into the neighbouring ring memory; those accumulator lanes are never read.
//
// CTA = 14 warps: warps 0-7 loaders (x rows and gy rows -> bf16 -> rings), warps 8-11 final epilogue,
// warp 12 MMA issuer + TMEM allocator, warp 13 idle (reserved for a TMA producer).
#include "hg_conv.cuh"
#include "hg_ptx.cuh"

namespace hg {

constexpr int kWuTile = 128;
constexpr int kWuPW = 144;
constexpr int kWuLoaders = 256;
constexpr int kWuThreads = 448;
constexpr int kWuBand = 32;
constexpr int kWuMaxQ = 5;
constexpr int kWuTaps = 7;
constexpr int kWuGSlots = 3;

struct WgParams {
  int N, Cin, Cout, H, W, Ho, Wo;
  int row0, col0;                // x row / col of (gy 0,0) for row slot 0 / shift 0
  int ra[kWuTaps];
  int sh[2][kWuTaps];
  int pad;
  float pad_value;
  int xslots, bands, ctiles, has_bias;
  long long items;
};

template <typename T> __device__ __forceinline__ float wu_ld(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ float wu_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

__device__ __forceinline__ uint32_t wu_pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// MN-major, no swizzle: element (mn, k) at start + (mn/8)*SBO + (mn%8)*2 + (k/8)*LBO + (k%8)*16   (bf16)
__device__ __forceinline__ uint64_t wu_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return ptx::umma_desc_kmajor_noswizzle(addr, lbo, sbo);   // same bit fields; the major-ness is in the instruction descriptor
}
__host__ __device__ constexpr uint32_t wu_idesc(int M, int N) {   // bf16 x bf16 -> fp32, A and B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <typename TX, typename TG>
__global__ void __launch_bounds__(kWuThreads, 1)
hexconv_wgrad_umma_kernel(const TX* __restrict__ x, const TG* __restrict__ gy, float* __restrict__ gw, float* __restrict__ gb,
                          WgParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gslot_bytes = P.Cout * kWuTile * 2;         // [Cout/8][128 px][16 B]
  const int xslot_bytes = P.Cin * kWuPW * 2;            // [Cin/8][PW px][16 B]
  unsigned char* gring = smem;                          // gy ring first: its M = 128 view may run into the x ring
  unsigned char* xring = gring + kWuGSlots * gslot_bytes;
  unsigned char* ones = xring + P.xslots * xslot_bytes; // 512 B of bf16 1.0 (bias accumulator operand)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 512);
  uint64_t* xfull = bars;                   // [xslots]
  uint64_t* xempty = xfull + P.xslots;      // [xslots]
  uint64_t* gfull = xempty + P.xslots;      // [kWuGSlots]
  uint64_t* gempty = gfull + kWuGSlots;     // [kWuGSlots]
  uint64_t* done = gempty + kWuGSlots;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  for (int e = tid; e < 256; e += kWuThreads) reinterpret_cast<__nv_bfloat16*>(ones)[e] = __float2bfloat16_rn(1.f);
  if (tid == 0) {
    for (int s = 0; s < P.xslots; ++s) { ptx::mbar_init(&xfull[s], kWuLoaders); ptx::mbar_init(&xempty[s], 1); }
    for (int s = 0; s < kWuGSlots; ++s) { ptx::mbar_init(&gfull[s], kWuLoaders); ptx::mbar_init(&gempty[s], 1); }
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 12) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the issue loop
  const int per_n = P.bands * P.ctiles;
  const int bias_col = kWuTaps * P.Cin;

  if (warp < 8) {
    // ===== loaders =========================================================================================
    const size_t xplane = (size_t)P.H * P.W, gplane = (size_t)P.Ho * P.Wo;
    const int xtasks = (P.Cin >> 3) * kWuPW, gtasks = (P.Cout >> 3) * kWuTile;
    long long xt = 0, gt = 0;
    auto load_x = [&](const TX* __restrict__ xn, int i, int c0) {
      const int slot = (int)(xt % P.xslots);
      ptx::mbar_wait(&xempty[slot], (uint32_t)(((xt / P.xslots) & 1) ^ 1));
      const bool row_in = i >= 0 && i < P.H, row_frame = i >= -P.pad && i < P.H + P.pad;
      unsigned char* sb = xring + (size_t)slot * xslot_bytes;
      float v[kWuMaxQ][8];
#pragma unroll
      for (int q = 0; q < kWuMaxQ; ++q) {
        const int task = tid + q * kWuLoaders;
        if (task < xtasks) {
          const int kc = task / kWuPW, p = task - kc * kWuPW;
          const int j = c0 + P.col0 + p;
          const bool col_in = j >= 0 && j < P.W;
          const float fill = (row_frame && j >= -P.pad && j < P.W + P.pad) ? P.pad_value : 0.f;
          const TX* __restrict__ src = xn + (size_t)(kc * 8) * xplane + (size_t)i * P.W + j;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[q][e] = (row_in && col_in) ? wu_ld(src + (size_t)e * xplane) : fill;
        }
      }
#pragma unroll
      for (int q = 0; q < kWuMaxQ; ++q) {
        const int task = tid + q * kWuLoaders;
        if (task < xtasks) {
          uint4 pk;
          pk.x = wu_pack(v[q][0], v[q][1]); pk.y = wu_pack(v[q][2], v[q][3]);
          pk.z = wu_pack(v[q][4], v[q][5]); pk.w = wu_pack(v[q][6], v[q][7]);
          *reinterpret_cast<uint4*>(sb + (size_t)task * 16) = pk;
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&xfull[slot]);
      ++xt;
    };
    auto load_g = [&](const TG* __restrict__ gn, int R, int c0) {
      const int slot = (int)(gt % kWuGSlots);
      ptx::mbar_wait(&gempty[slot], (uint32_t)(((gt / kWuGSlots) & 1) ^ 1));
      unsigned char* sb = gring + (size_t)slot * gslot_bytes;
      for (int base = 0; base < gtasks; base += kWuMaxQ * kWuLoaders) {   // Cout = 128 needs two passes
        float v[kWuMaxQ][8];
#pragma unroll
        for (int q = 0; q < kWuMaxQ; ++q) {
          const int task = base + tid + q * kWuLoaders;
          if (task < gtasks) {
            const int kc = task / kWuTile, p = task - kc * kWuTile;
            const int c = c0 + p;
            const TG* __restrict__ src = gn + (size_t)(kc * 8) * gplane + (size_t)R * P.Wo + c;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q][e] = c < P.Wo ? wu_ld(src + (size_t)e * gplane) : 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < kWuMaxQ; ++q) {
          const int task = base + tid + q * kWuLoaders;
          if (task < gtasks) {
            uint4 pk;
            pk.x = wu_pack(v[q][0], v[q][1]); pk.y = wu_pack(v[q][2], v[q][3]);
            pk.z = wu_pack(v[q][4], v[q][5]); pk.w = wu_pack(v[q][6], v[q][7]);
            *reinterpret_cast<uint4*>(sb + (size_t)task * 16) = pk;
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&gfull[slot]);
      ++gt;
    };
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int n = (int)(item / per_n);
      const int rem = (int)(item - (long long)n * per_n);
      const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
      const int r0 = band * kWuBand, rows = min(kWuBand, P.Ho - r0), c0 = ct * kWuTile;
      const TX* __restrict__ xn = x + (size_t)n * P.Cin * xplane;
      const TG* __restrict__ gn = gy + (size_t)n * P.Cout * gplane;
      load_x(xn, r0 + P.row0 + 0, c0);
      load_x(xn, r0 + P.row0 + 1, c0);
      for (int rr = 0; rr < rows; ++rr) {
        load_x(xn, r0 + P.row0 + rr + 2, c0);
        load_g(gn, r0 + rr, c0);
      }
    }
  } else if (warp == 12) {
    // ===== MMA issuer ======================================================================================
    const uint32_t idesc = wu_idesc(128, P.Cin), idesc_b = wu_idesc(128, 16);
    const uint32_t g_addr = ptx::smem_u32(gring), x_addr = ptx::smem_u32(xring), o_addr = ptx::smem_u32(ones);
    const uint32_t sbo_g = kWuTile * 16, sbo_x = kWuPW * 16;
    // descriptor = constant high word (SBO, version) | low word (start >> 4, LBO = 128 B between 8-pixel groups)
    const uint32_t g_hi = (sbo_g >> 4) | (1u << 14), x_hi = (sbo_x >> 4) | (1u << 14), o_hi = (256u >> 4) | (1u << 14);
    const uint32_t lo_const = (128u >> 4) << 16;
    const uint64_t od = ((uint64_t)o_hi << 32) | ((o_addr >> 4) + lo_const);
    uint32_t slot0 = 0, phase0 = 0, gs = 0, gphase = 0;
    auto next_slot = [&](uint32_t& sl, uint32_t& ph) { if (++sl == (uint32_t)P.xslots) { sl = 0; ph ^= 1; } };
    uint32_t started = 0;
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int rem = (int)(item % per_n);
      const int band = rem / P.ctiles;
      const int r0 = band * kWuBand, rows = min(kWuBand, P.Ho - r0);
      for (int rr = 0; rr < rows; ++rr) {
        uint32_t s1 = slot0, p1 = phase0; next_slot(s1, p1);
        uint32_t s2 = s1, p2 = p1; next_slot(s2, p2);
        ptx::mbar_wait(&xfull[slot0], phase0);
        ptx::mbar_wait(&xfull[s1], p1);
        ptx::mbar_wait(&xfull[s2], p2);
        ptx::mbar_wait(&gfull[gs], gphase);
        ptx::tc_fence_after_sync();
        {
          const int par = (r0 + rr) & 1;
          const uint32_t a_lo0 = ((g_addr + gs * (uint32_t)gslot_bytes) >> 4) + lo_const;
          const uint32_t xb0 = (x_addr + slot0 * (uint32_t)xslot_bytes) >> 4;
          const uint32_t xb1 = (x_addr + s1 * (uint32_t)xslot_bytes) >> 4;
          const uint32_t xb2 = (x_addr + s2 * (uint32_t)xslot_bytes) >> 4;
#pragma unroll
          for (int k = 0; k < kWuTaps; ++k) {
            const int ra = P.ra[k];
            uint32_t b_lo = (ra == 0 ? xb0 : (ra == 1 ? xb1 : xb2)) + (uint32_t)P.sh[par][k] + lo_const;
            uint32_t a_lo = a_lo0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(k * P.Cin);
#pragma unroll
            for (int j = 0; j < kWuTile / 16; ++j) {
              ptx::umma_bf16_elect(d_tmem, ((uint64_t)g_hi << 32) | a_lo, ((uint64_t)x_hi << 32) | b_lo, idesc, started | (uint32_t)j);
              a_lo += 16; b_lo += 16;                       // 16 pixels = 256 bytes
            }
          }
          if (P.has_bias) {
            uint32_t a_lo = a_lo0;
#pragma unroll
            for (int j = 0; j < kWuTile / 16; ++j) {
              ptx::umma_bf16_elect(tmem_base + (uint32_t)bias_col, ((uint64_t)g_hi << 32) | a_lo, od, idesc_b, started | (uint32_t)j);
              a_lo += 16;
            }
          }
          started = 1;
          ptx::umma_commit_elect(&gempty[gs]);
          ptx::umma_commit_elect(&xempty[slot0]);
          if (rr == rows - 1) {
            ptx::umma_commit_elect(&xempty[s1]);
            ptx::umma_commit_elect(&xempty[s2]);
          }
        }
        __syncwarp();
        next_slot(slot0, phase0);
        if (++gs == (uint32_t)kWuGSlots) { gs = 0; gphase ^= 1; }
      }
      next_slot(slot0, phase0);
      next_slot(slot0, phase0);
    }
    ptx::umma_commit_elect(done);
    __syncwarp();
  } else if (warp >= 8 && warp < 12) {
    // ===== final epilogue: TMEM partials -> fp32 atomics ====================================================
    const int q4 = warp & 3;
    const int co = q4 * 32 + lane;
    ptx::mbar_wait(done, 0);
    ptx::tc_fence_after_sync();
    for (int k = 0; k < kWuTaps; ++k) {
      for (int cb = 0; cb < P.Cin; cb += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(k * P.Cin + cb), v);
        ptx::tmem_ld_wait();
        if (co < P.Cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cb + j < P.Cin) atomicAdd(gw + ((size_t)co * P.Cin + cb + j) * kWuTaps + k, __uint_as_float(v[j]));
        }
      }
    }
    if (P.has_bias) {
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)bias_col, v);
      ptx::tmem_ld_wait();
      if (co < P.Cout) atomicAdd(gb + co, __uint_as_float(v[0]));
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side --------------------------------------------------------------------------------------------
static int g_wu_sms = 0, g_wu_smem_max = 0;

static size_t wu_smem_bytes(int Cin, int Cout, int xslots) {
  return (size_t)kWuGSlots * Cout * kWuTile * 2 + (size_t)xslots * Cin * kWuPW * 2 + 512 + (size_t)(2 * xslots + 2 * kWuGSlots + 1) * 8 + 16;
}

static int wu_pick_slots(int Cin, int Cout) {
  if (g_wu_smem_max == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_wu_smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&g_wu_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaGetLastError() != cudaSuccess || g_wu_smem_max <= 0) { g_wu_smem_max = 0; return 0; }
  }
  for (int s = 6; s >= 4; --s)
    if (wu_smem_bytes(Cin, Cout, s) <= (size_t)g_wu_smem_max) return s;
  return 0;
}

bool conv_wgrad_umma_eligible(const hg_conv_desc* d) {
  if (d->radius != 2 || d->stride != 1 || d->dilation != 1 || d->groups != 1) return false;
  if (d->Cin % 16 != 0 || d->Cin < 16 || d->Cin > 64) return false;        // 7*Cin + 16 TMEM columns <= 512
  if (d->Cout % 8 != 0 || d->Cout < 8 || d->Cout > 128) return false;      // UMMA M = 128 rows of 8-channel groups
  // the M = 128 view of the last gy slot runs (16 - Cout/8) channel groups past it: that must stay inside the x ring
  const int xslots = wu_pick_slots((int)d->Cin, (int)d->Cout);
  if (xslots == 0) return false;
  if ((size_t)(16 - d->Cout / 8) * kWuTile * 16 > (size_t)xslots * d->Cin * kWuPW * 2) return false;
  if (d->algo == 0 && (d->x_dtype != HG_BF16 || d->Cin * d->Cout < 32 * 32)) return false;
  return true;
}

template <typename TX, typename TG>
static int launch_wu(const void* x, const void* gy, float* gw, float* gb, const WgParams& P, cudaStream_t st) {
  const size_t smem = wu_smem_bytes(P.Cin, P.Cout, P.xslots);
  auto kern = hexconv_wgrad_umma_kernel<TX, TG>;
  static thread_local size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("hexconv_wgrad_umma: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    configured = smem;
  }
  long long grid = g_wu_sms > 0 ? g_wu_sms : 148;
  if (grid > P.items) grid = P.items;
  kern<<<(unsigned)grid, kWuThreads, smem, st>>>((const TX*)x, (const TG*)gy, gw, gb, P);
  return finish_launch("hexconv_wgrad_umma");
}

int conv_wgrad_umma(const hg_conv_desc* d, const ConvGeom& g, const ConvTaps& tp, const void* x, const void* gy, float* gw,
                    float* gbias, cudaStream_t st) {
  WgParams P{};
  P.N = g.N; P.Cin = g.Cin; P.Cout = g.Cout; P.H = g.H; P.W = g.W; P.Ho = g.Ho; P.Wo = g.Wo;
  int cmin = 1 << 30, cmax = -(1 << 30);
  for (int par = 0; par < 2; ++par)
    for (int k = 0; k < kWuTaps; ++k) { cmin = min(cmin, tp.co[par][k]); cmax = max(cmax, tp.co[par][k]); }
  HG_REQUIRE(tp.K == kWuTaps && cmax - cmin <= kWuPW - kWuTile, HG_E_UNSUPPORTED, "hexconv_wgrad_umma: unexpected tap geometry");
  for (int k = 0; k < kWuTaps; ++k) {
    P.ra[k] = tp.ro[k];
    for (int par = 0; par < 2; ++par) P.sh[par][k] = tp.co[par][k] - cmin;
  }
  P.row0 = -g.pad; P.col0 = cmin - g.pad;
  P.pad = g.pad; P.pad_value = g.pad_value;
  P.has_bias = gbias != nullptr;
  P.bands = (int)ceil_div(g.Ho, kWuBand);
  P.ctiles = (int)ceil_div(g.Wo, kWuTile);
  P.items = (long long)g.N * P.bands * P.ctiles;
  P.xslots = wu_pick_slots(g.Cin, g.Cout);
  HG_REQUIRE(P.xslots > 0, HG_E_UNSUPPORTED, "hexconv_wgrad_umma: shared memory does not fit");
  const int xdt = d->x_dtype, gdt = d->y_dtype;
  if (xdt == HG_F32 && gdt == HG_F32) return launch_wu<float, float>(x, gy, gw, gbias, P, st);
  if (xdt == HG_BF16 && gdt == HG_F32) return launch_wu<__nv_bfloat16, float>(x, gy, gw, gbias, P, st);
  if (xdt == HG_F32 && gdt == HG_BF16) return launch_wu<float, __nv_bfloat16>(x, gy, gw, gbias, P, st);
  if (xdt == HG_BF16 && gdt == HG_BF16) return launch_wu<__nv_bfloat16, __nv_bfloat16>(x, gy, gw, gbias, P, st);
  set_error("hexconv_wgrad_umma: unsupported dtypes x=%d gy=%d", xdt, gdt);
  return HG_E_DTYPE;
}

}  // namespace hg
