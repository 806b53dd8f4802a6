// hg_resample.cu -- rect<->hex resampling gathers and the lattice index kernels (sm_100a).
//
// Every kernel here is HBM-bound: one pass over the source planes, one pass over the
// destination planes.  Work decomposition shared by all of them:
//   CTA = 256 threads = 8 warps; tile = 32 output rows x 128 output columns;
//   warp w owns rows 4w..4w+3, lane l owns columns l, l+32, l+64, l+96 of the tile
//   (lane-consecutive columns: every warp store is one full 128-byte line for fp32, and the
//   gathered taps of neighbouring lanes fall in the same / adjacent 128-byte source lines).
// The sampling geometry depends only on the shapes, never on the image: each thread derives
// its column (and row) tables once and re-uses them for every plane it processes.
#include "hg_common.cuh"
#include "hg_hexgeom.cuh"
#include <type_traits>
#include <stdlib.h>

namespace hg {

constexpr int kTileW = 128;
constexpr int kTileH = 32;
constexpr int kThreads = 256;
constexpr int kColsPerThread = 4;
constexpr int kRowsPerWarp = 4;

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

// ==========================================================================================
// R1  rect -> hex   (ref: geometry_np.py:358-519)
// ==========================================================================================
struct RectCol {  // per output column (ref: geometry_np.py:441-449)
  int j_n;
  double j_f;
};

__device__ __forceinline__ void rect_axis(double coord, int n, int& idx, double& frac) {
  // i_ = x_ + (h-1)*0.5 ; i_n = trunc(i_) ; i_f = i_ - float32(i_n)
  double c = dadd(coord, (double)(n - 1) * 0.5);
  idx = trunc_i32(c);
  frac = dsub(c, (double)(float)idx);
}

__global__ void rect2hex_index_kernel(const double* __restrict__ xs, const double* __restrict__ ys, int h, int w,
                                      int h1, int w1, int32_t* i_n, double* i_f, int32_t* j_n, double* j_f) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < h1) {
    int n; double f;
    rect_axis(xs[t], h, n, f);
    i_n[t] = n; i_f[t] = f;
  }
  if (t < w1) {
    int n; double f;
    rect_axis(ys[t], w, n, f);
    j_n[t] = n; j_f[t] = f;
  }
}

// (One plane per CTA: the separable geometry is 8 float64 operations, and the chunked variant measured 2.3x slower
// on the 2:1 down-sampling of C2.)
// Bilinear blend, literal operation order of geometry_np.py:515-517.
template <typename TS, typename TD, bool EXACT>
__global__ void __launch_bounds__(kThreads)
rect2hex_bilinear_kernel(const TS* __restrict__ src, TD* __restrict__ dst, const double* __restrict__ xs,
                         const double* __restrict__ ys, int h, int w, int h1, int w1, int tiles_x, int tiles_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int64_t plane = tile / tiles_y;
  const TS* __restrict__ sp = src + plane * (int64_t)h * w;
  TD* __restrict__ dp = dst + plane * (int64_t)h1 * w1;

  int jn[kColsPerThread];
  double jf[kColsPerThread];
  bool cok[kColsPerThread];
#pragma unroll
  for (int k = 0; k < kColsPerThread; ++k) {
    const int b = tx * kTileW + lane + 32 * k;
    cok[k] = b < w1;
    rect_axis(cok[k] ? ys[b] : 0.0, w, jn[k], jf[k]);
  }
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int a = ty * kTileH + warp * kRowsPerWarp + rr;
    if (a >= h1) break;
    int in; double u;
    rect_axis(xs[a], h, in, u);
    const bool r0 = in >= 0 && in < h, r1 = in + 1 >= 0 && in + 1 < h;
    const TS* row0 = sp + (int64_t)in * w;
    const TS* row1 = row0 + w;
    TS p[kColsPerThread][4];
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
      const bool c0 = jn[k] >= 0 && jn[k] < w, c1 = jn[k] + 1 >= 0 && jn[k] + 1 < w;
      p[k][0] = (r0 && c0 && cok[k]) ? ldg(row0 + jn[k]) : TS(0);
      p[k][1] = (r0 && c1 && cok[k]) ? ldg(row0 + jn[k] + 1) : TS(0);
      p[k][2] = (r1 && c0 && cok[k]) ? ldg(row1 + jn[k]) : TS(0);
      p[k][3] = (r1 && c1 && cok[k]) ? ldg(row1 + jn[k] + 1) : TS(0);
    }
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
      if (!cok[k]) continue;
      const int b = tx * kTileW + lane + 32 * k;
      TD o;
      if (EXACT) {
        const double v = jf[k];
        const double u1 = dsub(1.0, u), v1 = dsub(1.0, v);
        const double t1 = dadd(dmul(u, to_f64(p[k][2])), dmul(u1, to_f64(p[k][0])));
        const double t2 = dadd(dmul(u, to_f64(p[k][3])), dmul(u1, to_f64(p[k][1])));
        o = (TD)dadd(dmul(v, t2), dmul(v1, t1));
      } else {
        const float uf = (float)u, vf = (float)jf[k];
        const float t1 = fmaf(uf, to_f32(p[k][2]) - to_f32(p[k][0]), to_f32(p[k][0]));
        const float t2 = fmaf(uf, to_f32(p[k][3]) - to_f32(p[k][1]), to_f32(p[k][1]));
        o = (TD)fmaf(vf, t2 - t1, t1);
      }
      st_stream(dp + (int64_t)a * w1 + b, o);
    }
  }
}

// Nearest: literal 4-way argmin of geometry_np.py:499-512 (distances between the centred sample
// coordinates and the un-centred corner indices; first minimum wins), then one gather.  The selected corner
// does not depend on the plane: a CTA resolves the source offsets of its tile once (float64) and then streams
// `chunk` planes through them.
template <typename T>
__global__ void __launch_bounds__(kThreads)
rect2hex_nearest_kernel(const T* __restrict__ src, T* __restrict__ dst, const double* __restrict__ xs,
                        const double* __restrict__ ys, int64_t planes, int chunk, int h, int w, int h1, int w1, int tiles_x,
                        int tiles_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int64_t p0 = (int64_t)(tile / tiles_y) * chunk;
  const int np = (int)min((int64_t)chunk, planes - p0);
  const int64_t sps = (int64_t)h * w, dps = (int64_t)h1 * w1;

  int jn[kColsPerThread];
  double dy0[kColsPerThread], dy1[kColsPerThread];
#pragma unroll
  for (int k = 0; k < kColsPerThread; ++k) {
    const int b = tx * kTileW + lane + 32 * k;
    const double y = b < w1 ? ys[b] : 0.0;
    double f;
    rect_axis(y, w, jn[k], f);
    const double e0 = dsub(y, (double)jn[k]), e1 = dsub(y, (double)(jn[k] + 1));
    dy0[k] = dmul(e0, e0);
    dy1[k] = dmul(e1, e1);
  }
  int off[kRowsPerWarp][kColsPerThread];   // source offset, -1 = outside the image (zero), -2 = no such output
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int a = ty * kTileH + warp * kRowsPerWarp + rr;
    const double x = xs[a < h1 ? a : h1 - 1];
    int in; double f;
    rect_axis(x, h, in, f);
    const double e0 = dsub(x, (double)in), e1 = dsub(x, (double)(in + 1));
    const double dx0 = dmul(e0, e0), dx1 = dmul(e1, e1);
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
      const double d1 = dadd(dx0, dy0[k]), d2 = dadd(dx0, dy1[k]), d3 = dadd(dx1, dy0[k]), d4 = dadd(dx1, dy1[k]);
      int sel = 0; double best = d1;
      if (d2 < best) { best = d2; sel = 1; }
      if (d3 < best) { best = d3; sel = 2; }
      if (d4 < best) { best = d4; sel = 3; }
      const int i = in + (sel >> 1), j = jn[k] + (sel & 1);
      const bool live = a < h1 && tx * kTileW + lane + 32 * k < w1;
      off[rr][k] = !live ? -2 : ((i >= 0 && i < h && j >= 0 && j < w) ? i * w + j : -1);
    }
  }
  const T* __restrict__ sp = src + p0 * sps;
  T* __restrict__ dp = dst + p0 * dps + (int64_t)(ty * kTileH + warp * kRowsPerWarp) * w1 + tx * kTileW + lane;
  for (int p = 0; p < np; ++p, sp += sps, dp += dps) {
    T v[kRowsPerWarp][kColsPerThread];
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr)
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k) v[rr][k] = off[rr][k] >= 0 ? ldg(sp + off[rr][k]) : T(0);
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr)
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k)
        if (off[rr][k] != -2) st_stream(dp + (int64_t)rr * w1 + 32 * k, v[rr][k]);
  }
}

// ==========================================================================================
// R2/R3/R4  sampling a hex-lattice image (ref: geometry_np.py:276-354, geometry_torch.py:278-356)
// ==========================================================================================
// Arith / HexSample / hex_locate / hex_locate_fast: hg_hexgeom.cuh (shared with hg_hexsrc_tma.cu)
// ---- coordinate sources ------------------------------------------------------------------
struct CoordTables {  // separable: hex->rect, hexresize (ref geometry_np.py:253-254 linspace)
  using CT = double;
  const double* xs; const double* ys;
  __device__ __forceinline__ void get(int a, int b, int w1, double& x, double& y) const { x = xs[a]; y = ys[b]; }
};
template <typename CT_>
struct CoordPlanes {  // warp with host-evaluated inverse map (bit-identical to the reference einsum)
  using CT = CT_;
  const CT_* cx; const CT_* cy;
  __device__ __forceinline__ void get(int a, int b, int w1, CT_& x, CT_& y) const {
    x = cx[(int64_t)a * w1 + b]; y = cy[(int64_t)a * w1 + b];
  }
};
template <typename CT_>
struct CoordAffine {  // warp with the inverse map evaluated in-kernel
  using CT = CT_;
  double m[6]; double row0, col0;
  __device__ __forceinline__ void get(int a, int b, int w1, CT_& x, CT_& y) const {
    const double X = dadd(row0, (double)a);
    const double Y = dadd(dadd(col0, (double)b), (a & 1) ? 0.5 : 0.0);
    x = (CT_)dadd(dadd(dmul(m[0], X), dmul(m[1], Y)), m[2]);
    y = (CT_)dadd(dadd(dmul(m[3], X), dmul(m[4], Y)), m[5]);
  }
};

struct HexConsts { double hx, wy, ci, cj; };
static HexConsts hex_consts(int64_t h, int64_t w) {
  // the reference evaluates these python-float expressions in double
  return HexConsts{(h - 1) / 2.0, (w - 0.5) / 2.0, (h - 1) * 0.5, (w - 0.5) * 0.5};
}

// Blend type of `alpha * p1 + beta * p2 + gamma * p3` under numpy/torch promotion:
// float32 weights stay float32 unless the image is float64.
template <typename CT, typename TS> struct BlendT { using type = double; };
template <> struct BlendT<float, float> { using type = float; };
template <> struct BlendT<float, uint8_t> { using type = float; };

// Planes per CTA.  The per-sample geometry (float64, ~200 instructions in exact mode) does not depend on the
// plane, so a CTA evaluates it once per output cell and streams `chunk` planes through it; the chunk is as
// large as possible while the launch still has ~12 waves of CTAs (tail effect < 10 %).
static int plane_chunk(int64_t planes, int64_t tiles) {
  const int64_t want_blocks = 148 * 8 * 12;
  int64_t groups = ceil_div(want_blocks, tiles);
  const int64_t max_groups = ceil_div(planes, 8);
  if (groups > max_groups) groups = max_groups;
  if (groups < 1) groups = 1;
  return (int)ceil_div(planes, groups);
}

template <typename TS, typename TD, typename Coord, bool FAST>
__global__ void __launch_bounds__(kThreads)
hexsrc_linear_kernel(const TS* __restrict__ src, TD* __restrict__ dst, Coord coord, HexConsts hc, int64_t planes, int chunk,
                     int h, int w, int h1, int w1, int tiles_x, int tiles_y) {
  using CT = typename Coord::CT;
  using WT = typename std::conditional<FAST, float, CT>::type;
  using BT = typename std::conditional<FAST, float, typename BlendT<CT, TS>::type>::type;
  using AB = Arith<BT>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int64_t p0 = (int64_t)(tile / tiles_y) * chunk;
  const int np = (int)min((int64_t)chunk, planes - p0);
  const int64_t sps = (int64_t)h * w, dps = (int64_t)h1 * w1;

  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int a = ty * kTileH + warp * kRowsPerWarp + rr;
    if (a >= h1) break;
    HexSample<WT> s[kColsPerThread];
    bool cok[kColsPerThread];
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
      const int b = tx * kTileW + lane + 32 * k;
      cok[k] = b < w1;
      CT x, y;
      coord.get(a, cok[k] ? b : 0, w1, x, y);
      if (FAST) hex_locate_fast((double)x, (double)y, h, w, hc.ci, hc.cj, (HexSample<float>&)s[k]);
      else hex_locate<CT, true, false>(x, y, h, w, (CT)hc.hx, (CT)hc.wy, (CT)hc.ci, (CT)hc.cj, (HexSample<CT>&)s[k]);
    }
    const TS* __restrict__ sp = src + p0 * sps;
    TD* __restrict__ dp = dst + p0 * dps + (int64_t)a * w1 + tx * kTileW + lane;
    for (int p = 0; p < np; ++p, sp += sps, dp += dps) {
      TS v[kColsPerThread][3];
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k)
#pragma unroll
        for (int t = 0; t < 3; ++t) v[k][t] = (cok[k] && s[k].off[t] >= 0) ? ldg(sp + s[k].off[t]) : TS(0);
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k) {
        if (!cok[k]) continue;
        BT o;
        if (FAST) {
          o = fmaf((float)s[k].wgt[2], to_f32(v[k][2]),
                   fmaf((float)s[k].wgt[1], to_f32(v[k][1]), (float)s[k].wgt[0] * to_f32(v[k][0])));
        } else {
          // ref geometry_np.py:354  alpha * p1 + beta * p2 + gamma * p3
          o = AB::add(AB::add(AB::mul((BT)s[k].wgt[0], (BT)v[k][0]), AB::mul((BT)s[k].wgt[1], (BT)v[k][1])),
                      AB::mul((BT)s[k].wgt[2], (BT)v[k][2]));
        }
        st_stream(dp + 32 * k, (TD)o);
      }
    }
  }
}

// Plane loop outermost: the 8 samples of a thread (3 offsets + 3 weights each) stay in registers and `chunk` planes
// stream through them, so the per-sample geometry -- ~250 float64 instructions in exact mode -- is paid once per
// `chunk` planes (the row-outer kernel above pays it once per 8 and keeps only 4 samples' loads in flight).
// Measured on C4: 2:1 hexresize fast 0.68 -> 1.06 of the HBM copy rate.
constexpr int kFastRows = 2;                 // output rows per warp of the plane-outer linear kernel (8 samples per thread in registers)
constexpr int kFastTileH = 8 * kFastRows;

// (float32 coordinates -- the torch twin's warps -- fit 80 registers without spilling: three blocks per SM instead of two)
template <typename TS, typename TD, typename Coord, bool FAST>
__global__ void __launch_bounds__(kThreads, (!FAST && sizeof(typename Coord::CT) == 4 && sizeof(TD) == 4) ? 3 : 2)
hexsrc_linear_fast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, Coord coord, HexConsts hc, int64_t planes, int chunk,
                          int h, int w, int h1, int w1, int tiles_x, int tiles_y) {
  using CT = typename Coord::CT;
  using WT = typename std::conditional<FAST, float, CT>::type;
  using BT = typename std::conditional<FAST, float, typename BlendT<CT, TS>::type>::type;
  using AB = Arith<BT>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int64_t p0 = (int64_t)(tile / tiles_y) * chunk;
  const int np = (int)min((int64_t)chunk, planes - p0);
  const int64_t sps = (int64_t)h * w, dps = (int64_t)h1 * w1;
  int off[kFastRows][kColsPerThread][3];
  WT wgt[kFastRows][kColsPerThread][3];
  bool live[kFastRows][kColsPerThread];
#pragma unroll
  for (int rr = 0; rr < kFastRows; ++rr) {
    const int a = ty * kFastTileH + warp * kFastRows + rr;
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
      const int b = tx * kTileW + lane + 32 * k;
      live[rr][k] = a < h1 && b < w1;
      CT x, y;
      coord.get(live[rr][k] ? a : 0, live[rr][k] ? b : 0, w1, x, y);
      HexSample<WT> s;
      if (FAST) hex_locate_fast((double)x, (double)y, h, w, hc.ci, hc.cj, (HexSample<float>&)s);
      else hex_locate<CT, true, false>(x, y, h, w, (CT)hc.hx, (CT)hc.wy, (CT)hc.ci, (CT)hc.cj, (HexSample<CT>&)s);
#pragma unroll
      for (int t = 0; t < 3; ++t) { off[rr][k][t] = live[rr][k] ? s.off[t] : -1; wgt[rr][k][t] = s.wgt[t]; }
    }
  }
  const TS* __restrict__ sp = src + p0 * sps;
  TD* __restrict__ dp = dst + p0 * dps + (int64_t)(ty * kFastTileH + warp * kFastRows) * w1 + tx * kTileW + lane;
  // (two planes per trip -- 48 gathers in flight per thread -- spills: 8 samples x (3 offsets + 3 weights) are 48 registers already)
  for (int p = 0; p < np; ++p, sp += sps, dp += dps) {
#pragma unroll
    for (int rr = 0; rr < kFastRows; ++rr) {
      TS v[kColsPerThread][3];
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k)
#pragma unroll
        for (int t = 0; t < 3; ++t) v[k][t] = off[rr][k][t] >= 0 ? ldg(sp + off[rr][k][t]) : TS(0);
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k) {
        if (!live[rr][k]) continue;
        BT o;
        if (FAST) {
          o = fmaf((float)wgt[rr][k][2], to_f32(v[k][2]), fmaf((float)wgt[rr][k][1], to_f32(v[k][1]), (float)wgt[rr][k][0] * to_f32(v[k][0])));
        } else {   // ref geometry_np.py:354  alpha * p1 + beta * p2 + gamma * p3, one rounding per operation
          o = AB::add(AB::add(AB::mul((BT)wgt[rr][k][0], (BT)v[k][0]), AB::mul((BT)wgt[rr][k][1], (BT)v[k][1])),
                      AB::mul((BT)wgt[rr][k][2], (BT)v[k][2]));
        }
        st_stream(dp + (int64_t)rr * w1 + 32 * k, (TD)o);
      }
    }
  }
}

// The closest lattice point of a sample does not depend on the plane: resolve the tile's 16 source offsets per
// thread once (float64 / float32 distances exactly as the reference evaluates them), then stream `chunk` planes.
template <typename T, typename Coord>
__global__ void __launch_bounds__(kThreads)
hexsrc_nearest_kernel(const T* __restrict__ src, T* __restrict__ dst, Coord coord, HexConsts hc, int64_t planes, int chunk,
                      int h, int w, int h1, int w1, int tiles_x, int tiles_y) {
  using CT = typename Coord::CT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int64_t p0 = (int64_t)(tile / tiles_y) * chunk;
  const int np = (int)min((int64_t)chunk, planes - p0);
  const int64_t sps = (int64_t)h * w, dps = (int64_t)h1 * w1;
  int off[kRowsPerWarp][kColsPerThread];   // -1 = outside the lattice (zero), -2 = no such output cell
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int a = ty * kTileH + warp * kRowsPerWarp + rr;
#pragma unroll
    for (int k = 0; k < kColsPerThread; ++k) {
      const int b = tx * kTileW + lane + 32 * k;
      off[rr][k] = -2;
      if (a < h1 && b < w1) {
        CT x, y;
        coord.get(a, b, w1, x, y);
        HexSample<CT> s;
        hex_locate<CT, false, true>(x, y, h, w, (CT)hc.hx, (CT)hc.wy, (CT)hc.ci, (CT)hc.cj, s);
        off[rr][k] = s.off[s.nearest];
      }
    }
  }
  const T* __restrict__ sp = src + p0 * sps;
  T* __restrict__ dp = dst + p0 * dps + (int64_t)(ty * kTileH + warp * kRowsPerWarp) * w1 + tx * kTileW + lane;
  for (int p = 0; p < np; ++p, sp += sps, dp += dps) {
    T v[kRowsPerWarp][kColsPerThread];
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr)
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k) v[rr][k] = off[rr][k] >= 0 ? ldg(sp + off[rr][k]) : T(0);
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr)
#pragma unroll
      for (int k = 0; k < kColsPerThread; ++k)
        if (off[rr][k] != -2) st_stream(dp + (int64_t)rr * w1 + 32 * k, v[rr][k]);
  }
}

template <typename Coord>
__global__ void hexsrc_index_kernel(Coord coord, HexConsts hc, int h, int w, int h1, int w1, int32_t* i_n,
                                    int32_t* j_n, uint8_t* tri, int32_t* off) {
  using CT = typename Coord::CT;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)h1 * w1) return;
  const int a = (int)(t / w1), b = (int)(t % w1);
  CT x, y;
  coord.get(a, b, w1, x, y);
  HexSample<CT> s;
  hex_locate<CT, false, false>(x, y, h, w, (CT)hc.hx, (CT)hc.wy, (CT)hc.ci, (CT)hc.cj, s);
  i_n[t] = s.i_n; j_n[t] = s.j_n;
  tri[t] = (uint8_t)((s.flag ? 1 : 0) | (s.off[0] >= 0 ? 2 : 0) | (s.off[1] >= 0 ? 4 : 0) | (s.off[2] >= 0 ? 8 : 0));
  const int64_t n = (int64_t)h1 * w1;
  off[t] = s.off[0]; off[n + t] = s.off[1]; off[2 * n + t] = s.off[2];
}

__global__ void axial_offset_kernel(const int32_t* i, const int32_t* jin, int32_t* jout, int64_t n, int sign) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) jout[t] = jin[t] + sign * trunc_half(i[t] + 1);
}

// ---- host-side launch helpers -------------------------------------------------------------
struct Tiling { int tx, ty; int64_t blocks; };
static int make_tiling(int64_t h1, int64_t w1, int64_t plane_groups, Tiling& t) {
  t.tx = (int)ceil_div(w1, kTileW);
  t.ty = (int)ceil_div(h1, kTileH);
  t.blocks = (int64_t)t.tx * t.ty * plane_groups;
  HG_REQUIRE(t.blocks > 0 && t.blocks < (1ll << 31), HG_E_SHAPE, "grid of %lld tiles is out of range", (long long)t.blocks);
  return HG_OK;
}
static int check_plane(int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1) {
  HG_REQUIRE(planes >= 0 && h > 0 && w > 0 && h1 >= 0 && w1 >= 0, HG_E_SHAPE, "bad shape planes=%lld h=%lld w=%lld h1=%lld w1=%lld",
             (long long)planes, (long long)h, (long long)w, (long long)h1, (long long)w1);
  HG_REQUIRE(h * w < (1ll << 31) && h1 * w1 < (1ll << 31), HG_E_SHAPE, "a single plane must have < 2^31 cells");
  return HG_OK;
}

// hg_resample_stream.cu: same contract; float32 weights between lattices of the same pitch
int try_hexsrc_linear_stream(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs, const double* host_ys,
                             int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt, int ddt, int math, cudaStream_t st);
// hg_hexsrc_tma.cu: same contract
int try_hexsrc_linear_tma(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs, const double* host_ys,
                          int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt, int ddt, int math, cudaStream_t st);
// hg_resample_stream.cu: same contract; lattices of (almost) the same pitch (hex_dsize = None)
int try_rect2hex_bilinear_stream(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs,
                                 const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt,
                                 int ddt, int math, cudaStream_t st);
// hg_resample_tma.cu: HG_OK launched, 1 not applicable (fall back to the direct gather), else error
int try_rect2hex_bilinear_tma(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs,
                              const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt,
                              int ddt, int math, cudaStream_t st);

template <typename TS, typename TD>
static int launch_rect2hex_bilinear(const void* src, void* dst, const double* xs, const double* ys, int64_t planes,
                                    int64_t h, int64_t w, int64_t h1, int64_t w1, int math, cudaStream_t st) {
  Tiling t;
  int rc = make_tiling(h1, w1, planes, t);
  if (rc) return rc;
  if (math == HG_MATH_EXACT)
    rect2hex_bilinear_kernel<TS, TD, true><<<(unsigned)t.blocks, kThreads, 0, st>>>(
        (const TS*)src, (TD*)dst, xs, ys, (int)h, (int)w, (int)h1, (int)w1, t.tx, t.ty);
  else
    rect2hex_bilinear_kernel<TS, TD, false><<<(unsigned)t.blocks, kThreads, 0, st>>>(
        (const TS*)src, (TD*)dst, xs, ys, (int)h, (int)w, (int)h1, (int)w1, t.tx, t.ty);
  return finish_launch("rect2hex_bilinear");
}

template <typename T>
static int launch_rect2hex_nearest(const void* src, void* dst, const double* xs, const double* ys, int64_t planes,
                                   int64_t h, int64_t w, int64_t h1, int64_t w1, cudaStream_t st) {
  Tiling t;
  const int chunk = plane_chunk(planes, ceil_div(h1, kTileH) * ceil_div(w1, kTileW));
  int rc = make_tiling(h1, w1, ceil_div(planes, chunk), t);
  if (rc) return rc;
  rect2hex_nearest_kernel<T><<<(unsigned)t.blocks, kThreads, 0, st>>>((const T*)src, (T*)dst, xs, ys, planes, chunk, (int)h,
                                                                      (int)w, (int)h1, (int)w1, t.tx, t.ty);
  return finish_launch("rect2hex_nearest");
}

template <typename TS, typename TD, typename Coord, bool FAST>
static int launch_hexsrc_linear(const void* src, void* dst, const Coord& c, int64_t planes, int64_t h, int64_t w,
                                int64_t h1, int64_t w1, cudaStream_t st) {
  Tiling t;
  // 32-bit weights (HG_MATH_FAST, or the torch twin's float32 coordinates): plane loop outermost.  With float64 weights the
  // 8 samples of a thread do not fit the register budget (C4: 14.0 ms vs 9.2 ms row-outer, 7.1 ms TMA tiles).
  if (FAST || std::is_same<typename Coord::CT, float>::value) {
    static const bool rows_outer = [] { const char* e = getenv("HG_HEXSRC_ROWS_OUTER"); return e && e[0] == '1'; }();
    if (!rows_outer) {
      const int pc = plane_chunk(planes, ceil_div(h1, kFastTileH) * ceil_div(w1, kTileW));
      t.tx = (int)ceil_div(w1, kTileW);
      t.ty = (int)ceil_div(h1, kFastTileH);
      t.blocks = (int64_t)t.tx * t.ty * ceil_div(planes, pc);
      HG_REQUIRE(t.blocks > 0 && t.blocks < (1ll << 31), HG_E_SHAPE, "grid of %lld tiles is out of range", (long long)t.blocks);
      hexsrc_linear_fast_kernel<TS, TD, Coord, FAST><<<(unsigned)t.blocks, kThreads, 0, st>>>(
          (const TS*)src, (TD*)dst, c, hex_consts(h, w), planes, pc, (int)h, (int)w, (int)h1, (int)w1, t.tx, t.ty);
      return finish_launch("hexsrc_linear_fast");
    }
  }
  const int chunk = 8;   // row-outer loop: larger chunks lose the L1 reuse of source rows shared by consecutive output rows (measured)
  int rc = make_tiling(h1, w1, ceil_div(planes, chunk), t);
  if (rc) return rc;
  hexsrc_linear_kernel<TS, TD, Coord, FAST><<<(unsigned)t.blocks, kThreads, 0, st>>>(
      (const TS*)src, (TD*)dst, c, hex_consts(h, w), planes, chunk, (int)h, (int)w, (int)h1, (int)w1, t.tx, t.ty);
  return finish_launch("hexsrc_linear");
}

template <typename Coord, bool FAST>
static int dispatch_hexsrc_linear(const void* src, void* dst, const Coord& c, int64_t planes, int64_t h, int64_t w,
                                  int64_t h1, int64_t w1, int sdt, int ddt, cudaStream_t st) {
#define HG_CASE(S, TS, D, TD) \
  if (sdt == S && ddt == D) return launch_hexsrc_linear<TS, TD, Coord, FAST>(src, dst, c, planes, h, w, h1, w1, st);
  HG_CASE(HG_U8, uint8_t, HG_F32, float)
  HG_CASE(HG_U8, uint8_t, HG_F64, double)
  HG_CASE(HG_F32, float, HG_F32, float)
  HG_CASE(HG_F32, float, HG_F64, double)
  HG_CASE(HG_F64, double, HG_F32, float)
  HG_CASE(HG_F64, double, HG_F64, double)
#undef HG_CASE
  set_error("hexsrc_linear: unsupported dtypes src=%d dst=%d", sdt, ddt);
  return HG_E_DTYPE;
}

template <typename Coord>
static int dispatch_hexsrc_nearest(const void* src, void* dst, const Coord& c, int64_t planes, int64_t h, int64_t w,
                                   int64_t h1, int64_t w1, int elem, cudaStream_t st) {
  Tiling t;
  const int chunk = plane_chunk(planes, ceil_div(h1, kTileH) * ceil_div(w1, kTileW));
  int rc = make_tiling(h1, w1, ceil_div(planes, chunk), t);
  if (rc) return rc;
  const HexConsts hc = hex_consts(h, w);
#define HG_CASE(E, T)                                                                                          \
  if (elem == E) {                                                                                             \
    hexsrc_nearest_kernel<T, Coord><<<(unsigned)t.blocks, kThreads, 0, st>>>((const T*)src, (T*)dst, c, hc, planes, chunk, \
                                                                            (int)h, (int)w, (int)h1, (int)w1, t.tx, t.ty); \
    return finish_launch("hexsrc_nearest");                                                                    \
  }
  HG_CASE(1, uint8_t)
  HG_CASE(2, uint16_t)
  HG_CASE(4, uint32_t)
  HG_CASE(8, uint64_t)
#undef HG_CASE
  set_error("hexsrc_nearest: unsupported element size %d", elem);
  return HG_E_DTYPE;
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_axial_to_offset_i32(const int32_t* i, const int32_t* j_ax, int32_t* j_off, int64_t n, hg_stream_t stream) {
  HG_REQUIRE(n >= 0, HG_E_SHAPE, "n < 0");
  if (n == 0) return HG_OK;
  axial_offset_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(i, j_ax, j_off, n, -1);
  return finish_launch("axial_to_offset");
}
int hg_offset_to_axial_i32(const int32_t* i, const int32_t* j_off, int32_t* j_ax, int64_t n, hg_stream_t stream) {
  HG_REQUIRE(n >= 0, HG_E_SHAPE, "n < 0");
  if (n == 0) return HG_OK;
  axial_offset_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(i, j_off, j_ax, n, +1);
  return finish_launch("offset_to_axial");
}

int hg_rect2hex_index(const double* xs, const double* ys, int64_t h, int64_t w, int64_t h1, int64_t w1, int32_t* i_n,
                      double* i_f, int32_t* j_n, double* j_f, hg_stream_t stream) {
  int rc = check_plane(0, h, w, h1, w1);
  if (rc) return rc;
  const int64_t n = h1 > w1 ? h1 : w1;
  if (n == 0) return HG_OK;
  rect2hex_index_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(xs, ys, (int)h, (int)w, (int)h1,
                                                                                  (int)w1, i_n, i_f, j_n, j_f);
  return finish_launch("rect2hex_index");
}

int hg_hexsrc_index(const void* xs, const void* ys, int coords_2d, int coord_f32, int64_t h, int64_t w, int64_t h1,
                    int64_t w1, int32_t* i_n, int32_t* j_n, uint8_t* tri, int32_t* off, hg_stream_t stream) {
  int rc = check_plane(0, h, w, h1, w1);
  if (rc) return rc;
  const int64_t n = h1 * w1;
  if (n == 0) return HG_OK;
  const unsigned g = (unsigned)ceil_div(n, 256);
  const HexConsts hc = hex_consts(h, w);
  cudaStream_t st = as_stream(stream);
  if (!coords_2d) {
    HG_REQUIRE(!coord_f32, HG_E_UNSUPPORTED, "separable coordinate tables are float64");
    hexsrc_index_kernel<<<g, 256, 0, st>>>(CoordTables{(const double*)xs, (const double*)ys}, hc, (int)h, (int)w,
                                           (int)h1, (int)w1, i_n, j_n, tri, off);
  } else if (coord_f32) {
    hexsrc_index_kernel<<<g, 256, 0, st>>>(CoordPlanes<float>{(const float*)xs, (const float*)ys}, hc, (int)h, (int)w,
                                           (int)h1, (int)w1, i_n, j_n, tri, off);
  } else {
    hexsrc_index_kernel<<<g, 256, 0, st>>>(CoordPlanes<double>{(const double*)xs, (const double*)ys}, hc, (int)h,
                                           (int)w, (int)h1, (int)w1, i_n, j_n, tri, off);
  }
  return finish_launch("hexsrc_index");
}

int hg_rect2hex_nearest(const void* src, void* dst, const double* xs, const double* ys, int64_t planes, int64_t h,
                        int64_t w, int64_t h1, int64_t w1, int elem_size, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  switch (elem_size) {
    case 1: return launch_rect2hex_nearest<uint8_t>(src, dst, xs, ys, planes, h, w, h1, w1, st);
    case 2: return launch_rect2hex_nearest<uint16_t>(src, dst, xs, ys, planes, h, w, h1, w1, st);
    case 4: return launch_rect2hex_nearest<uint32_t>(src, dst, xs, ys, planes, h, w, h1, w1, st);
    case 8: return launch_rect2hex_nearest<uint64_t>(src, dst, xs, ys, planes, h, w, h1, w1, st);
  }
  set_error("rect2hex_nearest: unsupported element size %d", elem_size);
  return HG_E_DTYPE;
}

int hg_rect2hex_bilinear(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs,
                         const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int src_dtype,
                         int dst_dtype, int math, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  HG_REQUIRE(math == HG_MATH_EXACT || math == HG_MATH_FAST, HG_E_ARG, "bad math mode %d", math);
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  // same-pitch lattices (hex_dsize = None): row-streaming kernel, no staging at all
  rc = try_rect2hex_bilinear_stream(src, dst, xs, ys, host_xs, host_ys, planes, h, w, h1, w1, src_dtype, dst_dtype, math, st);
  if (rc != 1) return rc;
  // TMA-staged tiles when the host copies of the tables are available and the lattices have similar pitch
  rc = try_rect2hex_bilinear_tma(src, dst, xs, ys, host_xs, host_ys, planes, h, w, h1, w1, src_dtype, dst_dtype, math, st);
  if (rc != 1) return rc;
#define HG_CASE(S, TS, D, TD) \
  if (src_dtype == S && dst_dtype == D) return launch_rect2hex_bilinear<TS, TD>(src, dst, xs, ys, planes, h, w, h1, w1, math, st);
  HG_CASE(HG_U8, uint8_t, HG_F32, float)
  HG_CASE(HG_U8, uint8_t, HG_F64, double)
  HG_CASE(HG_F32, float, HG_F32, float)
  HG_CASE(HG_F32, float, HG_F64, double)
  HG_CASE(HG_F64, double, HG_F32, float)
  HG_CASE(HG_F64, double, HG_F64, double)
#undef HG_CASE
  set_error("rect2hex_bilinear: unsupported dtypes src=%d dst=%d", src_dtype, dst_dtype);
  return HG_E_DTYPE;
}

int hg_hex2rect_nearest(const void* src, void* dst, const double* xs, const double* ys, int64_t planes, int64_t h,
                        int64_t w, int64_t h1, int64_t w1, int elem_size, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  return dispatch_hexsrc_nearest(src, dst, CoordTables{xs, ys}, planes, h, w, h1, w1, elem_size, as_stream(stream));
}

int hg_hex2rect_linear(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs,
                       const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int src_dtype,
                       int dst_dtype, int math, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  HG_REQUIRE(math == HG_MATH_EXACT || math == HG_MATH_FAST, HG_E_ARG, "bad math mode %d", math);
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  rc = try_hexsrc_linear_stream(src, dst, xs, ys, host_xs, host_ys, planes, h, w, h1, w1, src_dtype, dst_dtype, math, as_stream(stream));
  if (rc != 1) return rc;
  rc = try_hexsrc_linear_tma(src, dst, xs, ys, host_xs, host_ys, planes, h, w, h1, w1, src_dtype, dst_dtype, math, as_stream(stream));
  if (rc != 1) return rc;
  CoordTables c{xs, ys};
  if (math == HG_MATH_EXACT)
    return dispatch_hexsrc_linear<CoordTables, false>(src, dst, c, planes, h, w, h1, w1, src_dtype, dst_dtype, as_stream(stream));
  return dispatch_hexsrc_linear<CoordTables, true>(src, dst, c, planes, h, w, h1, w1, src_dtype, dst_dtype, as_stream(stream));
}

int hg_hexwarp_nearest(const void* src, void* dst, const void* cx, const void* cy, int coord_f32, int64_t planes,
                       int64_t h, int64_t w, int64_t h1, int64_t w1, int elem_size, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  if (coord_f32)
    return dispatch_hexsrc_nearest(src, dst, CoordPlanes<float>{(const float*)cx, (const float*)cy}, planes, h, w, h1,
                                   w1, elem_size, as_stream(stream));
  return dispatch_hexsrc_nearest(src, dst, CoordPlanes<double>{(const double*)cx, (const double*)cy}, planes, h, w, h1,
                                 w1, elem_size, as_stream(stream));
}

int hg_hexwarp_linear(const void* src, void* dst, const void* cx, const void* cy, int coord_f32, int64_t planes,
                      int64_t h, int64_t w, int64_t h1, int64_t w1, int src_dtype, int dst_dtype, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  if (coord_f32)
    return dispatch_hexsrc_linear<CoordPlanes<float>, false>(src, dst, CoordPlanes<float>{(const float*)cx, (const float*)cy},
                                                             planes, h, w, h1, w1, src_dtype, dst_dtype, as_stream(stream));
  return dispatch_hexsrc_linear<CoordPlanes<double>, false>(src, dst, CoordPlanes<double>{(const double*)cx, (const double*)cy},
                                                            planes, h, w, h1, w1, src_dtype, dst_dtype, as_stream(stream));
}

int hg_hexwarp_affine(const void* src, void* dst, const double* host_hinv, double row0, double col0, int coord_f32,
                      int interp, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int src_dtype,
                      int dst_dtype, hg_stream_t stream) {
  int rc = check_plane(planes, h, w, h1, w1);
  if (rc) return rc;
  HG_REQUIRE(host_hinv != nullptr, HG_E_ARG, "host_hinv is NULL");
  HG_REQUIRE(interp == 0 || interp == 1, HG_E_ARG, "interp must be 0 (nearest) or 1 (linear)");
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  if (coord_f32) {
    CoordAffine<float> c;
    for (int k = 0; k < 6; ++k) c.m[k] = host_hinv[k];
    c.row0 = row0; c.col0 = col0;
    if (interp == 1) return dispatch_hexsrc_linear<CoordAffine<float>, false>(src, dst, c, planes, h, w, h1, w1, src_dtype, dst_dtype, st);
    HG_REQUIRE(src_dtype == dst_dtype, HG_E_DTYPE, "nearest keeps the element type");
    return dispatch_hexsrc_nearest(src, dst, c, planes, h, w, h1, w1, dtype_size(src_dtype), st);
  }
  CoordAffine<double> c;
  for (int k = 0; k < 6; ++k) c.m[k] = host_hinv[k];
  c.row0 = row0; c.col0 = col0;
  if (interp == 1) return dispatch_hexsrc_linear<CoordAffine<double>, false>(src, dst, c, planes, h, w, h1, w1, src_dtype, dst_dtype, st);
  HG_REQUIRE(src_dtype == dst_dtype, HG_E_DTYPE, "nearest keeps the element type");
  return dispatch_hexsrc_nearest(src, dst, c, planes, h, w, h1, w1, dtype_size(src_dtype), st);
}

}  // extern "C"
