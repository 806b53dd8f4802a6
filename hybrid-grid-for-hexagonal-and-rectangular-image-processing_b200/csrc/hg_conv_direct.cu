// hg_conv_direct.cu -- hex convolution as a direct sparse stencil on CUDA cores (sm_100a): forward,
// data gradient, weight / bias gradient.  Any radius / stride / dilation / groups; fp32 or bf16
// activations, fp32 weights and accumulation.  This is the path for small channel counts (where the
// op is HBM-bound) and the general fallback; the dense 64x64-and-up contraction runs on tcgen05
// (hg_conv_umma.cu).  Geometry: hg_conv.cuh (ref: HexFrames.py:96-169).
//
// Forward / dgrad decomposition: CTA = 8 warps; warp w owns one output row, lanes run along the row
// (coalesced loads of every tap: neighbouring lanes read neighbouring cells), each thread keeps
// PIX x CT accumulators (PIX columns 32 apart, CT channels); the CTA's weight slab sits in shared
// memory as [c_red][tap][CT] and is read as broadcast float4.
#include "hg_conv.cuh"

namespace hg {

constexpr int kConvThreads = 256;
constexpr int kPix = 2;           // output columns per thread (32 apart)
constexpr int kCT = 16;           // output channels per thread
constexpr int kSmemFloats = 8192; // weight slab per reduction chunk (32 KB)

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------------
// forward
// grid: x = ceil(Wo / 64), y = ceil(Ho / 8), z = N * (Cout / CT tiles)
// ---------------------------------------------------------------------------------------------------
template <typename TX, typename TY>
__global__ void __launch_bounds__(kConvThreads)
hexconv_fwd_direct(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ bias, TY* __restrict__ y,
                   ConvGeom g, ConvTaps tp, int ci_chunk) {
  __shared__ __align__(16) float ws[kSmemFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co_tiles = (g.cout_g + kCT - 1) / kCT;        // tiles never straddle a group
  const int tiles_per_n = co_tiles * g.groups;
  const int n = blockIdx.z / tiles_per_n;
  const int tile = blockIdx.z % tiles_per_n;
  const int grp = tile / co_tiles;
  const int co0 = grp * g.cout_g + (tile % co_tiles) * kCT;
  const int co_n = min(kCT, (grp + 1) * g.cout_g - co0);
  const int R = blockIdx.y * 8 + warp;
  const int q0 = blockIdx.x * (32 * kPix) + lane;
  const bool row_ok = R < g.Ho;
  const int par = R & 1;
  const int K = tp.K;

  float acc[kPix][kCT];
#pragma unroll
  for (int p = 0; p < kPix; ++p)
#pragma unroll
    for (int c = 0; c < kCT; ++c) acc[p][c] = 0.f;

  const TX* __restrict__ xn = x + ((int64_t)n * g.Cin + (int64_t)grp * g.cin_g) * g.H * g.W;
  for (int cb = 0; cb < g.cin_g; cb += ci_chunk) {
    const int cn = min(ci_chunk, g.cin_g - cb);
    __syncthreads();
    // ws[(ci*K + k)*CT + c] = w[co0 + c, cb + ci, k]
    for (int e = threadIdx.x; e < cn * K * kCT; e += kConvThreads) {
      const int c = e % kCT, k = (e / kCT) % K, ci = e / (kCT * K);
      ws[e] = c < co_n ? __ldg(w + ((int64_t)(co0 + c) * g.cin_g + cb + ci) * K + k) * (scale ? __ldg(scale + co0 + c) : 1.f) : 0.f;
    }
    __syncthreads();
    if (!row_ok) continue;
    for (int ci = 0; ci < cn; ++ci) {
      const TX* __restrict__ xc = xn + (int64_t)(cb + ci) * g.H * g.W;
      for (int k = 0; k < K; ++k) {
        const int i = g.s * R + tp.ro[k] - g.pad;         // warp-uniform
        const bool in_r = i >= 0 && i < g.H;
        const int jb = tp.co[par][k] - g.pad;
        float xv[kPix];
#pragma unroll
        for (int p = 0; p < kPix; ++p) {
          const int j = g.s * (q0 + 32 * p) + jb;
          float v = 0.f;                                   // beyond the padded frame: literal zero
          if (j < g.W + g.pad) {
            v = g.pad_value;                               // inside the frame, outside the image
            if (in_r && j >= 0 && j < g.W) v = ldf(xc + (int64_t)i * g.W + j);
            else if (g.pad_mode)                           // reflect / replicate / circular frame: read the image in place
              v = ldf(xc + (int64_t)conv_pad_remap(i, g.H, g.pad_mode) * g.W + conv_pad_remap(j, g.W, g.pad_mode));
          }
          xv[p] = v;
        }
        const float4* wk = reinterpret_cast<const float4*>(ws + (ci * K + k) * kCT);
#pragma unroll
        for (int c4 = 0; c4 < kCT / 4; ++c4) {
          const float4 wv = wk[c4];
#pragma unroll
          for (int p = 0; p < kPix; ++p) {
            acc[p][4 * c4 + 0] = fmaf(xv[p], wv.x, acc[p][4 * c4 + 0]);
            acc[p][4 * c4 + 1] = fmaf(xv[p], wv.y, acc[p][4 * c4 + 1]);
            acc[p][4 * c4 + 2] = fmaf(xv[p], wv.z, acc[p][4 * c4 + 2]);
            acc[p][4 * c4 + 3] = fmaf(xv[p], wv.w, acc[p][4 * c4 + 3]);
          }
        }
      }
    }
  }
  if (!row_ok) return;
#pragma unroll
  for (int c = 0; c < kCT; ++c) {
    if (c >= co_n) break;
    const float b = bias ? __ldg(bias + co0 + c) : 0.f;
    TY* __restrict__ yr = y + (((int64_t)n * g.Cout + co0 + c) * g.Ho + R) * g.Wo;
#pragma unroll
    for (int p = 0; p < kPix; ++p) {
      const int q = q0 + 32 * p;
      if (q < g.Wo) {
        float v = acc[p][c] + b;
        if (g.relu) v = fmaxf(v, 0.f);
        stf(yr + q, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// data gradient:  gx[n,ci,i,j] = sum_{co,k} w[co,ci,k] * gy[n,co,R,q]  over the (R,q) with
//   s*R + ro[k] - pad == i  and  s*q + co[R&1][k] - pad == j.
// grid: x = ceil(W / 64), y = ceil(H / 8), z = N * (Cin / CT tiles)
// ---------------------------------------------------------------------------------------------------
template <typename TG, typename TX>
__global__ void __launch_bounds__(kConvThreads)
hexconv_dgrad_direct(const TG* __restrict__ gy, const float* __restrict__ w, TX* __restrict__ gx, ConvGeom g, ConvTaps tp,
                     int co_chunk) {
  __shared__ __align__(16) float ws[kSmemFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci_tiles = (g.cin_g + kCT - 1) / kCT;
  const int tiles_per_n = ci_tiles * g.groups;
  const int n = blockIdx.z / tiles_per_n;
  const int tile = blockIdx.z % tiles_per_n;
  const int grp = tile / ci_tiles;
  const int cl0 = (tile % ci_tiles) * kCT;                 // channel offset inside the group
  const int ci_n = min(kCT, g.cin_g - cl0);
  const int i = blockIdx.y * 8 + warp;
  const int j0 = blockIdx.x * (32 * kPix) + lane;
  const bool row_ok = i < g.H;
  const int K = tp.K;

  float acc[kPix][kCT];
#pragma unroll
  for (int p = 0; p < kPix; ++p)
#pragma unroll
    for (int c = 0; c < kCT; ++c) acc[p][c] = 0.f;

  const TG* __restrict__ gn = gy + ((int64_t)n * g.Cout + (int64_t)grp * g.cout_g) * g.Ho * g.Wo;
  for (int cb = 0; cb < g.cout_g; cb += co_chunk) {
    const int cn = min(co_chunk, g.cout_g - cb);
    __syncthreads();
    // ws[(co*K + k)*CT + c] = w[grp*cout_g + cb + co, cl0 + c, k]
    for (int e = threadIdx.x; e < cn * K * kCT; e += kConvThreads) {
      const int c = e % kCT, k = (e / kCT) % K, co = e / (kCT * K);
      ws[e] = c < ci_n ? __ldg(w + ((int64_t)(grp * g.cout_g + cb + co) * g.cin_g + cl0 + c) * K + k) : 0.f;
    }
    __syncthreads();
    if (!row_ok) continue;
    for (int co = 0; co < cn; ++co) {
      const TG* __restrict__ gc = gn + (int64_t)(cb + co) * g.Ho * g.Wo;
      for (int k = 0; k < K; ++k) {
        const int rn = i + g.pad - tp.ro[k];               // = s*R, warp-uniform
        if (rn < 0 || rn % g.s != 0) continue;
        const int R = rn / g.s;
        if (R >= g.Ho) continue;
        const int cofs = g.pad - tp.co[R & 1][k];
        float gv[kPix];
#pragma unroll
        for (int p = 0; p < kPix; ++p) {
          const int qn = j0 + 32 * p + cofs;               // = s*q
          float v = 0.f;
          if (qn >= 0 && qn % g.s == 0) {
            const int q = qn / g.s;
            if (q < g.Wo) v = ldf(gc + (int64_t)R * g.Wo + q);
          }
          gv[p] = v;
        }
        const float4* wk = reinterpret_cast<const float4*>(ws + (co * K + k) * kCT);
#pragma unroll
        for (int c4 = 0; c4 < kCT / 4; ++c4) {
          const float4 wv = wk[c4];
#pragma unroll
          for (int p = 0; p < kPix; ++p) {
            acc[p][4 * c4 + 0] = fmaf(gv[p], wv.x, acc[p][4 * c4 + 0]);
            acc[p][4 * c4 + 1] = fmaf(gv[p], wv.y, acc[p][4 * c4 + 1]);
            acc[p][4 * c4 + 2] = fmaf(gv[p], wv.z, acc[p][4 * c4 + 2]);
            acc[p][4 * c4 + 3] = fmaf(gv[p], wv.w, acc[p][4 * c4 + 3]);
          }
        }
      }
    }
  }
  if (!row_ok) return;
#pragma unroll
  for (int c = 0; c < kCT; ++c) {
    if (c >= ci_n) break;
    TX* __restrict__ xr = gx + (((int64_t)n * g.Cin + (int64_t)grp * g.cin_g + cl0 + c) * g.H + i) * g.W;
#pragma unroll
    for (int p = 0; p < kPix; ++p) {
      const int j = j0 + 32 * p;
      if (j < g.W) stf(xr + j, acc[p][c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient: gw[co,ci,k] += sum_{n,R,q} gy[n,co,R,q] * P[n,ci, s*R+ro[k], s*q+co[R&1][k]]
// CTA = (image n, band of kBand output rows, 4 co x 4 ci); lanes run along q, warps along rows; each
// lane keeps 4x4xKT partial sums, reduced by shuffles -> shared -> one atomicAdd per weight per CTA.
// grid: x = bands, y = (Cout/4) * (cin_g/4) tiles, z = N
// ---------------------------------------------------------------------------------------------------
constexpr int kWgC = 4;      // co and ci per CTA
constexpr int kWgKT = 4;     // taps per pass (4 x 4 x 4 = 64 accumulators per thread: no spills, 2 CTAs / SM)
constexpr int kWgBand = 64;  // output rows per CTA

template <typename TX, typename TG>
__global__ void __launch_bounds__(kConvThreads, 2)
hexconv_wgrad_direct(const TX* __restrict__ x, const TG* __restrict__ gy, float* __restrict__ gw, ConvGeom g, ConvTaps tp) {
  __shared__ float red[kWgC * kWgC * kWgKT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci_tiles = (g.cin_g + kWgC - 1) / kWgC, co_tiles = (g.cout_g + kWgC - 1) / kWgC;
  const int ci_t = blockIdx.y % ci_tiles;
  const int co_t = (blockIdx.y / ci_tiles) % co_tiles;
  const int grp = blockIdx.y / (ci_tiles * co_tiles);
  const int col0 = co_t * kWgC;                             // output channel inside the group
  const int co0 = grp * g.cout_g + col0;                    // global output channel
  const int cl0 = ci_t * kWgC;                              // input channel inside the group
  const int n = blockIdx.z;
  const int R0 = blockIdx.x * kWgBand;
  const int R1 = min(R0 + kWgBand, g.Ho);
  const int K = tp.K;

  for (int kb = 0; kb < K; kb += kWgKT) {
    const int kn = min(kWgKT, K - kb);
    float acc[kWgC][kWgC][kWgKT];
#pragma unroll
    for (int a = 0; a < kWgC; ++a)
#pragma unroll
      for (int b = 0; b < kWgC; ++b)
#pragma unroll
        for (int k = 0; k < kWgKT; ++k) acc[a][b][k] = 0.f;

    for (int R = R0 + warp; R < R1; R += 8) {
      const int par = R & 1;
      for (int q = lane; q < g.Wo; q += 32) {
        float gv[kWgC];
#pragma unroll
        for (int a = 0; a < kWgC; ++a) {
          const int co = co0 + a;
          gv[a] = (col0 + a < g.cout_g) ? ldf(gy + (((int64_t)n * g.Cout + co) * g.Ho + R) * g.Wo + q) : 0.f;
        }
#pragma unroll
        for (int b = 0; b < kWgC; ++b) {
          // no `break` in these loops: they must unroll completely or acc[][][] lands in local memory
          const int cl = min(cl0 + b, g.cin_g - 1);
          const bool cl_ok = cl0 + b < g.cin_g;
          const TX* __restrict__ xc = x + ((int64_t)n * g.Cin + (int64_t)grp * g.cin_g + cl) * g.H * g.W;
#pragma unroll
          for (int k = 0; k < kWgKT; ++k) {
            const int kk = min(kb + k, K - 1);
            const int i = g.s * R + tp.ro[kk] - g.pad;
            const int j = g.s * q + tp.co[par][kk] - g.pad;
            float v = 0.f;
            if (cl_ok && k < kn && j < g.W + g.pad) {
              v = g.pad_value;
              if (i >= 0 && i < g.H && j >= 0 && j < g.W) v = ldf(xc + (int64_t)i * g.W + j);
              else if (g.pad_mode) v = ldf(xc + (int64_t)conv_pad_remap(i, g.H, g.pad_mode) * g.W + conv_pad_remap(j, g.W, g.pad_mode));
            }
#pragma unroll
            for (int a = 0; a < kWgC; ++a) acc[a][b][k] = fmaf(gv[a], v, acc[a][b][k]);
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < kWgC * kWgC * kWgKT) red[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < kWgC; ++a)
#pragma unroll
      for (int b = 0; b < kWgC; ++b)
#pragma unroll
        for (int k = 0; k < kWgKT; ++k) {
          float v = acc[a][b][k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) atomicAdd(&red[(a * kWgC + b) * kWgKT + k], v);
        }
    __syncthreads();
    if (threadIdx.x < kWgC * kWgC * kWgKT) {
      const int k = threadIdx.x % kWgKT, b = (threadIdx.x / kWgKT) % kWgC, a = threadIdx.x / (kWgKT * kWgC);
      const int co = co0 + a, cl = cl0 + b;
      if (k < kn && col0 + a < g.cout_g && cl < g.cin_g)
        atomicAdd(gw + ((int64_t)co * g.cin_g + cl) * K + kb + k, red[threadIdx.x]);
    }
  }
}

// gbias[co] += sum_{n,R,q} gy[n,co,R,q];  grid: x = Cout, y = N
template <typename TG>
__global__ void __launch_bounds__(kConvThreads)
hexconv_bgrad(const TG* __restrict__ gy, float* __restrict__ gb, int Cout, int64_t plane) {
  __shared__ float part[8];
  const TG* __restrict__ p = gy + ((int64_t)blockIdx.y * Cout + blockIdx.x) * plane;
  float s = 0.f;
  for (int64_t e = threadIdx.x; e < plane; e += kConvThreads) s += ldf(p + e);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) s += part[k];
    atomicAdd(gb + blockIdx.x, s);
  }
}

// ---- launchers (called from hg_conv.cu) -------------------------------------------------------------------
static int red_chunk(int K) {
  int c = kSmemFloats / (K * kCT);
  return c < 1 ? 1 : c;
}

int conv_fwd_direct(const ConvGeom& g, const ConvTaps& tp, const void* x, int xdt, const float* w, const float* scale, const float* bias,
                    void* y, int ydt, cudaStream_t st) {
  const int co_tiles = (g.cout_g + kCT - 1) / kCT;
  dim3 grid((unsigned)ceil_div(g.Wo, 32 * kPix), (unsigned)ceil_div(g.Ho, 8), (unsigned)((int64_t)g.N * co_tiles * g.groups));
  HG_REQUIRE((int64_t)g.N * co_tiles * g.groups <= 65535 && grid.y <= 65535, HG_E_SHAPE, "hexconv_fwd: batch x channel tiles exceed the grid");
  const int ch = red_chunk(tp.K);
#define HG_CASE(XD, TX, YD, TY)                                                                                             \
  if (xdt == XD && ydt == YD) {                                                                                             \
    hexconv_fwd_direct<TX, TY><<<grid, kConvThreads, 0, st>>>((const TX*)x, w, scale, bias, (TY*)y, g, tp, ch);                     \
    return finish_launch("hexconv_fwd_direct");                                                                             \
  }
  HG_CASE(HG_F32, float, HG_F32, float)
  HG_CASE(HG_BF16, __nv_bfloat16, HG_F32, float)
  HG_CASE(HG_F32, float, HG_BF16, __nv_bfloat16)
  HG_CASE(HG_BF16, __nv_bfloat16, HG_BF16, __nv_bfloat16)
#undef HG_CASE
  set_error("hexconv_fwd: unsupported dtypes x=%d y=%d", xdt, ydt);
  return HG_E_DTYPE;
}

int conv_dgrad_direct(const ConvGeom& g, const ConvTaps& tp, const void* gy, int gdt, const float* w, void* gx, int xdt,
                      cudaStream_t st) {
  const int ci_tiles = (g.cin_g + kCT - 1) / kCT;
  dim3 grid((unsigned)ceil_div(g.W, 32 * kPix), (unsigned)ceil_div(g.H, 8), (unsigned)((int64_t)g.N * ci_tiles * g.groups));
  HG_REQUIRE((int64_t)g.N * ci_tiles * g.groups <= 65535 && grid.y <= 65535, HG_E_SHAPE, "hexconv_dgrad: batch x channel tiles exceed the grid");
  const int ch = red_chunk(tp.K);
#define HG_CASE(GD, TG, XD, TX)                                                                                             \
  if (gdt == GD && xdt == XD) {                                                                                             \
    hexconv_dgrad_direct<TG, TX><<<grid, kConvThreads, 0, st>>>((const TG*)gy, w, (TX*)gx, g, tp, ch);                        \
    return finish_launch("hexconv_dgrad_direct");                                                                           \
  }
  HG_CASE(HG_F32, float, HG_F32, float)
  HG_CASE(HG_BF16, __nv_bfloat16, HG_F32, float)
  HG_CASE(HG_F32, float, HG_BF16, __nv_bfloat16)
  HG_CASE(HG_BF16, __nv_bfloat16, HG_BF16, __nv_bfloat16)
#undef HG_CASE
  set_error("hexconv_dgrad: unsupported dtypes gy=%d gx=%d", gdt, xdt);
  return HG_E_DTYPE;
}

int conv_wgrad_direct(const ConvGeom& g, const ConvTaps& tp, const void* x, int xdt, const void* gy, int gdt, float* gw,
                      float* gbias, cudaStream_t st) {
  const int ci_tiles = (g.cin_g + kWgC - 1) / kWgC;
  const int co_tiles = (g.cout_g + kWgC - 1) / kWgC;
  dim3 grid((unsigned)ceil_div(g.Ho, kWgBand), (unsigned)((int64_t)g.groups * co_tiles * ci_tiles), (unsigned)g.N);
  HG_REQUIRE(grid.y <= 65535 && g.N <= 65535, HG_E_SHAPE, "hexconv_wgrad: channel tiles exceed the grid");
  int rc = HG_OK;
#define HG_CASE(XD, TX, GD, TG)                                                                                  \
  if (xdt == XD && gdt == GD) {                                                                                  \
    hexconv_wgrad_direct<TX, TG><<<grid, kConvThreads, 0, st>>>((const TX*)x, (const TG*)gy, gw, g, tp);         \
    rc = finish_launch("hexconv_wgrad_direct");                                                                  \
    if (rc == HG_OK && gbias) {                                                                                  \
      hexconv_bgrad<TG><<<dim3((unsigned)g.Cout, (unsigned)g.N), kConvThreads, 0, st>>>((const TG*)gy, gbias, g.Cout, (int64_t)g.Ho * g.Wo); \
      rc = finish_launch("hexconv_bgrad");                                                                       \
    }                                                                                                            \
    return rc;                                                                                                   \
  }
  HG_CASE(HG_F32, float, HG_F32, float)
  HG_CASE(HG_BF16, __nv_bfloat16, HG_F32, float)
  HG_CASE(HG_F32, float, HG_BF16, __nv_bfloat16)
  HG_CASE(HG_BF16, __nv_bfloat16, HG_BF16, __nv_bfloat16)
#undef HG_CASE
  set_error("hexconv_wgrad: unsupported dtypes x=%d gy=%d", xdt, gdt);
  return HG_E_DTYPE;
}

}  // namespace hg
