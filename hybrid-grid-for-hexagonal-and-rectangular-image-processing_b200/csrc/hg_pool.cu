// hg_pool.cu -- hex-lattice pooling, forward and backward (sm_100a).  HBM-bound.
//
//   ref: HexFrames.py:255-341 HexPool2d, :344-401 HexAdaptivePool2d, :402-414 HexGlobalPool2d,
//        reductions :461-479 (max: NaN -> -inf, min: NaN -> +inf, average: mean of the non-NaN cells,
//        all-NaN -> NaN).
// The reference builds int64 index tables on the CPU and gathers a (hn*wn*kh*kw, B, C) tensor; here the
// window of out(I, J) -- rows sh*I + a, cols ((I%2)*shift)/2 + J*sw + b -- is pure index arithmetic and
// padding / ceil-mode tails are virtual (never materialised).
// Forward: one thread per output cell, lanes along J (a warp reads one contiguous span per window row).
// Backward: one thread per *input* cell gathers from the (few) windows covering it -- deterministic, no
// atomics, gx written exactly once.
#include "hg_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace hg {

constexpr int kPoolThreads = 256;

struct PoolGeom {
  int H, W, hn, wn, kh, kw, sh, sw, shift, pad, tail_h, tail_w;
};

template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type ld_acc(const T* p) { return (typename Acc<T>::type)__ldg(p); }
template <> __device__ __forceinline__ float ld_acc<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T, typename A> __device__ __forceinline__ T st_acc(A v) { return (T)v; }
template <> __device__ __forceinline__ __nv_bfloat16 st_acc<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }

// rows = planes * hn output rows; grid.x = rows * jtiles (one CTA = one output row segment of kPoolThreads cells).
template <typename T, int METHOD, typename AUX>
__global__ void __launch_bounds__(kPoolThreads)
hexpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, AUX* __restrict__ aux, unsigned jtiles, unsigned nrows, int tw_log2,
                   PoolGeom g, typename Acc<T>::type pad_value, typename Acc<T>::type tail_value) {
  using A = typename Acc<T>::type;
  // CTA = (256 >> tw_log2) rows x (1 << tw_log2) columns, so narrow maps still fill the block
  const unsigned rgrp = blockIdx.x / jtiles;
  const unsigned row = rgrp * (kPoolThreads >> tw_log2) + (threadIdx.x >> tw_log2);
  const int J = (int)((blockIdx.x - rgrp * jtiles) << tw_log2) + (int)(threadIdx.x & ((1u << tw_log2) - 1));
  if (J >= g.wn || row >= nrows) return;
  const unsigned plane = row / (unsigned)g.hn;
  const int I = (int)(row - plane * (unsigned)g.hn);
  const T* __restrict__ xp = x + (size_t)plane * g.H * g.W;
  const size_t t = ((size_t)plane * g.hn + I) * g.wn + J;
  const int r0 = g.sh * I, c0 = ((I & 1) * g.shift) / 2 + J * g.sw;
  const int Hp = g.H + 2 * g.pad, Wp = g.W + 2 * g.pad;
  const A inf = (A)CUDART_INF;
  A best = METHOD == HG_POOL_MAX ? -inf : inf;
  A sum = 0;
  int slot = 0, cnt = 0;
  bool best_masked = false;
  // interior window (no padding / tail cell touched): no per-cell bounds logic
  const bool interior = g.pad == 0 && r0 + g.kh <= g.H && c0 + g.kw <= g.W;
  for (int a = 0; a < g.kh; ++a) {
    const int rr = r0 + a;
    for (int b = 0; b < g.kw; ++b) {
      const int cc = c0 + b;
      A v;
      if (interior) v = ld_acc(xp + (size_t)rr * g.W + cc);
      else if (rr >= Hp || cc >= Wp) v = tail_value;
      else {
        const int i = rr - g.pad, j = cc - g.pad;
        v = (i >= 0 && i < g.H && j >= 0 && j < g.W) ? ld_acc(xp + (size_t)i * g.W + j) : pad_value;
      }
      const bool nan = v != v;
      if (METHOD == HG_POOL_AVG) {
        if (!nan) { sum += v; ++cnt; }
      } else {
        const A m = nan ? (METHOD == HG_POOL_MAX ? -inf : inf) : v;
        const bool first = (a == 0 && b == 0);
        const bool better = METHOD == HG_POOL_MAX ? (m > best) : (m < best);
        if (first || better) { best = m; slot = a * g.kw + b; best_masked = nan; }
      }
    }
  }
  if (METHOD == HG_POOL_AVG) {
    y[t] = st_acc<T, A>(cnt ? sum / (A)cnt : (A)CUDART_NAN);
    if (aux) aux[t] = (AUX)cnt;
  } else {
    y[t] = st_acc<T, A>(best);
    if (aux) aux[t] = best_masked ? (AUX)-1 : (AUX)slot;  // a NaN "winner" passes no gradient (masked_fill)
  }
}

__device__ __forceinline__ int ceil_div_i(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }
__device__ __forceinline__ int floor_div_i(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// rows = planes * H input rows; grid.x = rows * jtiles.  DISJOINT: windows do not overlap (kh <= sh, kw <= sw), so an
// input cell belongs to at most one window and the gather is a single candidate.
template <typename T, int METHOD, typename AUX, bool DISJOINT>
__global__ void __launch_bounds__(kPoolThreads)
hexpool_bwd_kernel(const T* __restrict__ gy, const AUX* __restrict__ aux, const T* __restrict__ x, T* __restrict__ gx,
                   unsigned jtiles, unsigned nrows, int tw_log2, PoolGeom g) {
  using A = typename Acc<T>::type;
  const unsigned rgrp = blockIdx.x / jtiles;
  const unsigned row = rgrp * (kPoolThreads >> tw_log2) + (threadIdx.x >> tw_log2);
  const int j = (int)((blockIdx.x - rgrp * jtiles) << tw_log2) + (int)(threadIdx.x & ((1u << tw_log2) - 1));
  if (j >= g.W || row >= nrows) return;
  const unsigned plane = row / (unsigned)g.H;
  const int i = (int)(row - plane * (unsigned)g.H);
  const size_t t = ((size_t)plane * g.H + i) * g.W + j;
  const size_t ob = (size_t)plane * g.hn * g.wn;
  const int rr = i + g.pad, cc = j + g.pad;
  A acc = 0;
  bool live = true;
  if (METHOD == HG_POOL_AVG && x != nullptr) { const A v = ld_acc(x + t); live = (v == v); }
  if (live) {
    if (DISJOINT) {
      const int I = rr / g.sh;                       // warp-uniform
      const int a = rr - I * g.sh;
      if (I < g.hn && a < g.kh) {
        const int cj = cc - ((I & 1) * g.shift) / 2;
        if (cj >= 0) {
          const int J = cj / g.sw, b = cj - J * g.sw;
          if (J < g.wn && b < g.kw) {
            const size_t o = ob + (size_t)I * g.wn + J;
            if (METHOD == HG_POOL_AVG) {
              const int cnt = (int)aux[o];
              if (cnt > 0) acc = ld_acc(gy + o) / (A)cnt;
            } else if ((int)aux[o] == a * g.kw + b) acc = ld_acc(gy + o);
          }
        }
      }
    } else {
      const int I_lo = max(ceil_div_i(rr - g.kh + 1, g.sh), 0), I_hi = min(floor_div_i(rr, g.sh), g.hn - 1);
      for (int I = I_lo; I <= I_hi; ++I) {
        const int off = ((I & 1) * g.shift) / 2;
        const int J_lo = max(ceil_div_i(cc - off - g.kw + 1, g.sw), 0), J_hi = min(floor_div_i(cc - off, g.sw), g.wn - 1);
        for (int J = J_lo; J <= J_hi; ++J) {
          const size_t o = ob + (size_t)I * g.wn + J;
          if (METHOD == HG_POOL_AVG) {
            const int cnt = (int)aux[o];
            if (cnt > 0) acc += ld_acc(gy + o) / (A)cnt;
          } else {
            const int slot = (rr - g.sh * I) * g.kw + (cc - off - J * g.sw);
            if ((int)aux[o] == slot) acc += ld_acc(gy + o);
          }
        }
      }
    }
  }
  gx[t] = st_acc<T, A>(acc);
}

// ---- 2 x 2 / stride 2 fast path (HexPool2d(method, 2, 2): the C4 pyramid and the C5 network) --------------
// kh = kw = sh = sw = shift = 2, no padding, no ceil-mode tail, float32, W % 4 == 0, 16-byte aligned planes:
//   out(I, J) = reduce x[2I + a, (I & 1) + 2J + b],  a, b in {0, 1};  hn = H / 2, wn = (W - 1) / 2.
// One warp owns a 128-column segment of one *pair* of output rows (I = 2g even, 2g + 1 odd): every lane reads
// one aligned float4 of each of the four input rows (4 x LDG.128 in flight per lane, each byte of x read
// once); the odd row's windows straddle the quads by one cell, which comes from the next lane by shuffle.
// The reduction visits the cells in the reference's window order, so results (and max / min slots) are
// bit-identical to the generic kernel.
template <int METHOD>
__device__ __forceinline__ void pool4(float v0, float v1, float v2, float v3, float& out, int& aux) {
  const float v[4] = {v0, v1, v2, v3};
  if (METHOD == HG_POOL_AVG) {
    float sum = 0.f; int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (v[i] == v[i]) { sum += v[i]; ++cnt; }
    out = cnt ? sum / (float)cnt : CUDART_NAN_F;
    aux = cnt;
  } else {
    const float inf = CUDART_INF_F;
    float best = 0.f; int slot = 0; bool masked = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool nan = v[i] != v[i];
      const float m = nan ? (METHOD == HG_POOL_MAX ? -inf : inf) : v[i];
      const bool better = METHOD == HG_POOL_MAX ? (m > best) : (m < best);
      if (i == 0 || better) { best = m; slot = i; masked = nan; }
    }
    out = best;
    aux = masked ? -1 : slot;
  }
}

template <int METHOD>
__global__ void __launch_bounds__(kPoolThreads)
hexpool2x2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int8_t* __restrict__ aux, long long items, int pairs,
                      int segs, int H, int W, int hn, int wn) {
  const long long item = (long long)blockIdx.x * (kPoolThreads / 32) + (threadIdx.x >> 5);
  if (item >= items) return;                              // warp-uniform
  const int lane = threadIdx.x & 31;
  const long long pr = item / segs;
  const int seg = (int)(item - pr * segs);
  const long long plane = pr / pairs;
  const int g = (int)(pr - plane * pairs);
  const int c4 = seg * 128 + 4 * lane;                    // first input column of this lane's quad
  const bool live = c4 < W;
  const int I0 = 2 * g;
  const bool has_odd = I0 + 1 < hn;
  const float* __restrict__ xp = x + ((size_t)plane * H + (size_t)2 * I0) * W + c4;
  float4 r[4];
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) r[k] = (live && (k < 2 || has_odd)) ? __ldg(reinterpret_cast<const float4*>(xp + (size_t)k * W)) : z;
  // odd output row: the window J = 2c + 1 ends in the first cell of the next quad
  float n2 = __shfl_down_sync(0xffffffffu, r[2].x, 1), n3 = __shfl_down_sync(0xffffffffu, r[3].x, 1);
  if (lane == 31 && has_odd && c4 + 4 < W) { n2 = __ldg(xp + (size_t)2 * W + 4); n3 = __ldg(xp + (size_t)3 * W + 4); }
  if (!live) return;
  const int J = c4 >> 1;                                  // outputs J, J + 1
  float o; int a;
  const size_t t0 = ((size_t)plane * hn + I0) * wn + J;
  if (J < wn) {
    pool4<METHOD>(r[0].x, r[0].y, r[1].x, r[1].y, o, a);
    __stcs(y + t0, o);
    if (aux) aux[t0] = (int8_t)a;
  }
  if (J + 1 < wn) {
    pool4<METHOD>(r[0].z, r[0].w, r[1].z, r[1].w, o, a);
    __stcs(y + t0 + 1, o);
    if (aux) aux[t0 + 1] = (int8_t)a;
  }
  if (has_odd) {
    const size_t t1 = t0 + wn;
    if (J < wn) {
      pool4<METHOD>(r[2].y, r[2].z, r[3].y, r[3].z, o, a);
      __stcs(y + t1, o);
      if (aux) aux[t1] = (int8_t)a;
    }
    if (J + 1 < wn) {
      pool4<METHOD>(r[2].w, n2, r[3].w, n3, o, a);
      __stcs(y + t1 + 1, o);
      if (aux) aux[t1 + 1] = (int8_t)a;
    }
  }
}

// Backward of the same configuration: one warp writes a 128-column segment of FOUR input rows (two output
// rows) of gx as aligned float4 stores; gy / aux of the (at most three) windows touching a quad are read once.
template <int METHOD>
__device__ __forceinline__ float pool4_grad(float gy, int aux, int slot) {
  if (METHOD == HG_POOL_AVG) return aux > 0 ? gy / (float)aux : 0.f;
  return aux == slot ? gy : 0.f;
}

template <int METHOD>
__global__ void __launch_bounds__(kPoolThreads)
hexpool2x2_bwd_kernel(const float* __restrict__ gy, const int8_t* __restrict__ aux, const float* __restrict__ x,
                      float* __restrict__ gx, long long items, int pairs, int segs, int H, int W, int hn, int wn) {
  const long long item = (long long)blockIdx.x * (kPoolThreads / 32) + (threadIdx.x >> 5);
  if (item >= items) return;                              // warp-uniform
  const int lane = threadIdx.x & 31;
  const long long pr = item / segs;
  const int seg = (int)(item - pr * segs);
  const long long plane = pr / pairs;
  const int g = (int)(pr - plane * pairs);
  const int c4 = seg * 128 + 4 * lane;
  const bool live = c4 < W;
  const int I0 = 2 * g, J = c4 >> 1;
  const int row0 = 4 * g;                                 // first of the (up to) four input rows
  // windows J, J + 1 of the even row; J - 1, J, J + 1 of the odd row
  float ge[2] = {0.f, 0.f}, go[3] = {0.f, 0.f, 0.f};
  int ae[2] = {METHOD == HG_POOL_AVG ? 0 : -1, METHOD == HG_POOL_AVG ? 0 : -1};
  int ao[3] = {ae[0], ae[0], ae[0]};
  if (live && I0 < hn) {
    const size_t t0 = ((size_t)plane * hn + I0) * wn + J;
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (J + e < wn) { ge[e] = __ldg(gy + t0 + e); ae[e] = (int)aux[t0 + e]; }
    if (I0 + 1 < hn) {
      const size_t t1 = t0 + wn;
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (J + e < wn) { go[1 + e] = __ldg(gy + t1 + e); ao[1 + e] = (int)aux[t1 + e]; }
    }
  }
  go[0] = __shfl_up_sync(0xffffffffu, go[2], 1);
  ao[0] = __shfl_up_sync(0xffffffffu, ao[2], 1);
  if (lane == 0) {
    go[0] = 0.f; ao[0] = METHOD == HG_POOL_AVG ? 0 : -1;
    if (live && I0 + 1 < hn && J >= 1) {
      const size_t t = ((size_t)plane * hn + I0 + 1) * wn + J - 1;
      go[0] = __ldg(gy + t); ao[0] = (int)aux[t];
    }
  }
  if (!live) return;
  float4 o[4];
  // even output row -> input rows row0, row0 + 1; slots a * 2 + b
  o[0] = make_float4(pool4_grad<METHOD>(ge[0], ae[0], 0), pool4_grad<METHOD>(ge[0], ae[0], 1),
                     pool4_grad<METHOD>(ge[1], ae[1], 0), pool4_grad<METHOD>(ge[1], ae[1], 1));
  o[1] = make_float4(pool4_grad<METHOD>(ge[0], ae[0], 2), pool4_grad<METHOD>(ge[0], ae[0], 3),
                     pool4_grad<METHOD>(ge[1], ae[1], 2), pool4_grad<METHOD>(ge[1], ae[1], 3));
  // odd output row -> input rows row0 + 2, row0 + 3; columns shifted right by one cell
  o[2] = make_float4(pool4_grad<METHOD>(go[0], ao[0], 1), pool4_grad<METHOD>(go[1], ao[1], 0),
                     pool4_grad<METHOD>(go[1], ao[1], 1), pool4_grad<METHOD>(go[2], ao[2], 0));
  o[3] = make_float4(pool4_grad<METHOD>(go[0], ao[0], 3), pool4_grad<METHOD>(go[1], ao[1], 2),
                     pool4_grad<METHOD>(go[1], ao[1], 3), pool4_grad<METHOD>(go[2], ao[2], 2));
  const size_t base = ((size_t)plane * H + row0) * W + c4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (row0 + k >= H) break;
    float4 v = o[k];
    if (METHOD == HG_POOL_AVG && x != nullptr) {          // NaN cells took no part in the mean: no gradient
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base + (size_t)k * W));
      if (xv.x != xv.x) v.x = 0.f;
      if (xv.y != xv.y) v.y = 0.f;
      if (xv.z != xv.z) v.z = 0.f;
      if (xv.w != xv.w) v.w = 0.f;
    }
    __stcs(reinterpret_cast<float4*>(gx + base + (size_t)k * W), v);
  }
}

// Same configuration for ANY width / alignment (the deeper pyramid levels are 1919, 959, 479 ... cells wide):
// lanes run along the input columns, so all 16 loads of a warp (4 rows x 128 columns) are fully coalesced
// scalar loads; the horizontal neighbour of a window comes from lane + 1 by shuffle (from the first lane of the
// next 32-column chunk, or one extra scalar load at the segment end).  Windows of the even output row start on
// even lanes, those of the odd output row on odd lanes.
template <int METHOD>
__global__ void __launch_bounds__(kPoolThreads)
hexpool2x2c_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int8_t* __restrict__ aux, long long items, int pairs,
                       int segs, int H, int W, int hn, int wn) {
  const long long item = (long long)blockIdx.x * (kPoolThreads / 32) + (threadIdx.x >> 5);
  if (item >= items) return;                              // warp-uniform
  const int lane = threadIdx.x & 31;
  const long long pr = item / segs;
  const int seg = (int)(item - pr * segs);
  const long long plane = pr / pairs;
  const int g = (int)(pr - plane * pairs);
  const int I0 = 2 * g;
  const bool has_odd = I0 + 1 < hn;
  const int col0 = seg * 128 + lane;
  const float* __restrict__ xp = x + ((size_t)plane * H + (size_t)2 * I0) * W;
  float v[4][4];                                          // [chunk][row]
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int r = 0; r < 4; ++r)
      v[k][r] = (col0 + 32 * k < W && (r < 2 || has_odd)) ? __ldg(xp + (size_t)r * W + col0 + 32 * k) : 0.f;
  float edge2 = 0.f, edge3 = 0.f;                         // column seg*128 + 128 of rows 2, 3 (lane 31 only)
  if (lane == 31 && has_odd && seg * 128 + 128 < W) {
    edge2 = __ldg(xp + (size_t)2 * W + seg * 128 + 128);
    edge3 = __ldg(xp + (size_t)3 * W + seg * 128 + 128);
  }
  const size_t yb = ((size_t)plane * hn + I0) * wn;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int col = col0 + 32 * k;
    float n[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) n[r] = __shfl_down_sync(0xffffffffu, v[k][r], 1);
    if (k < 3) {
      const float f2 = __shfl_sync(0xffffffffu, v[k < 3 ? k + 1 : 3][2], 0), f3 = __shfl_sync(0xffffffffu, v[k < 3 ? k + 1 : 3][3], 0);
      if (lane == 31) { n[2] = f2; n[3] = f3; }
    } else if (lane == 31) { n[2] = edge2; n[3] = edge3; }
    // even lanes own a window of the even output row (rows 0, 1), odd lanes one of the odd output row (rows 2, 3):
    // operands are selected first so that all 32 lanes run ONE reduction and one store (no divergent halves)
    const bool odd = lane & 1;
    const float a0 = odd ? v[k][2] : v[k][0], a1 = odd ? n[2] : n[0];
    const float b0 = odd ? v[k][3] : v[k][1], b1 = odd ? n[3] : n[1];
    const int J = (col - (odd ? 1 : 0)) >> 1;
    float o; int a;
    pool4<METHOD>(a0, a1, b0, b1, o, a);
    if (J < wn && (!odd || has_odd)) {
      const size_t t = yb + (odd ? wn : 0) + J;
      __stcs(y + t, o);
      if (aux) aux[t] = (int8_t)a;
    }
  }
}

template <int METHOD>
__global__ void __launch_bounds__(kPoolThreads)
hexpool2x2c_bwd_kernel(const float* __restrict__ gy, const int8_t* __restrict__ aux, const float* __restrict__ x,
                       float* __restrict__ gx, long long items, int pairs, int segs, int H, int W, int hn, int wn) {
  const long long item = (long long)blockIdx.x * (kPoolThreads / 32) + (threadIdx.x >> 5);
  if (item >= items) return;
  const int lane = threadIdx.x & 31;
  const long long pr = item / segs;
  const int seg = (int)(item - pr * segs);
  const long long plane = pr / pairs;
  const int g = (int)(pr - plane * pairs);
  const int I0 = 2 * g, row0 = 4 * g;
  const size_t ob = (size_t)plane * hn * wn;
  const size_t base = ((size_t)plane * H + row0) * W;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int col = seg * 128 + 32 * k + lane;
    if (col >= W) continue;
    float ge = 0.f, go = 0.f;
    int ae = METHOD == HG_POOL_AVG ? 0 : -1, ao = ae;
    const int Je = col >> 1, be = col & 1;                // even output row: window / slot column of this cell
    const int Jo = (col - 1) >> 1, bo = (col - 1) & 1;    // odd output row (cells shifted right by one)
    if (I0 < hn && Je < wn) { const size_t o = ob + (size_t)I0 * wn + Je; ge = __ldg(gy + o); ae = (int)aux[o]; }
    if (I0 + 1 < hn && col >= 1 && Jo < wn) { const size_t o = ob + (size_t)(I0 + 1) * wn + Jo; go = __ldg(gy + o); ao = (int)aux[o]; }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (row0 + r >= H) break;
      float v = r < 2 ? pool4_grad<METHOD>(ge, ae, r * 2 + be) : pool4_grad<METHOD>(go, ao, (r - 2) * 2 + bo);
      if (METHOD == HG_POOL_AVG && x != nullptr) { const float xv = __ldg(x + base + (size_t)r * W + col); if (xv != xv) v = 0.f; }
      __stcs(gx + base + (size_t)r * W + col, v);
    }
  }
}

static bool pool2x2_ok(const PoolGeom& g, int tail_h, int tail_w) {
  static const bool off = [] { const char* e = getenv("HG_POOL_GENERIC"); return e && e[0] == '1'; }();
  return !off && g.kh == 2 && g.kw == 2 && g.sh == 2 && g.sw == 2 && g.shift == 2 && g.pad == 0 && tail_h == 0 && tail_w == 0 &&
         g.hn == g.H / 2 && g.wn == (g.W - 1) / 2 && g.hn > 0 && g.wn > 0;
}
static bool pool2x2_vec_ok(const PoolGeom& g, const void* a, const void* b) {   // float4 rows
  return g.W % 4 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0;
}

// ---- global pooling: x [planes, L] -> y [planes] ------------------------------------------------------
template <typename A> struct Red { A v; int idx; int cnt; };

template <typename A, int METHOD>
__device__ __forceinline__ void red_merge(Red<A>& a, const Red<A>& b) {
  if (METHOD == HG_POOL_AVG) { a.v += b.v; a.cnt += b.cnt; }
  else {
    const bool better = METHOD == HG_POOL_MAX ? (b.v > a.v) : (b.v < a.v);
    if (better || (b.v == a.v && b.idx < a.idx)) { a.v = b.v; a.idx = b.idx; }
  }
}

template <typename T, int METHOD, int GROUP>  // GROUP threads cooperate on one plane (32 or 256)
__global__ void __launch_bounds__(256)
globalpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int32_t* __restrict__ aux, int64_t planes, int64_t L) {
  using A = typename Acc<T>::type;
  const int lane = threadIdx.x % GROUP;
  const int64_t plane = (int64_t)blockIdx.x * (256 / GROUP) + threadIdx.x / GROUP;
  const bool active = plane < planes;
  const A inf = (A)CUDART_INF;
  Red<A> r;
  r.v = METHOD == HG_POOL_AVG ? (A)0 : (METHOD == HG_POOL_MAX ? -inf : inf);
  r.idx = 0x7fffffff; r.cnt = 0;
  if (active) {
    const T* __restrict__ xp = x + plane * L;
    for (int64_t l = lane; l < L; l += GROUP) {
      const A v = ld_acc(xp + l);
      const bool nan = v != v;
      if (METHOD == HG_POOL_AVG) { if (!nan) { r.v += v; ++r.cnt; } }
      else {
        const A m = nan ? (METHOD == HG_POOL_MAX ? -inf : inf) : v;
        const bool better = METHOD == HG_POOL_MAX ? (m > r.v) : (m < r.v);
        if (better || (m == r.v && (int)l < r.idx)) { r.v = m; r.idx = (int)l; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Red<A> b;
    b.v = __shfl_xor_sync(0xffffffffu, r.v, o);
    b.idx = __shfl_xor_sync(0xffffffffu, r.idx, o);
    b.cnt = __shfl_xor_sync(0xffffffffu, r.cnt, o);
    red_merge<A, METHOD>(r, b);
  }
  if (GROUP > 32) {
    __shared__ Red<A> part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) red_merge<A, METHOD>(r, part[w]);
    }
  }
  if (active && lane == 0) {
    if (METHOD == HG_POOL_AVG) {
      y[plane] = st_acc<T, A>(r.cnt ? r.v / (A)r.cnt : (A)CUDART_NAN);
      if (aux) aux[plane] = r.cnt;
    } else {
      y[plane] = st_acc<T, A>(r.v);
      if (aux) aux[plane] = r.idx;
    }
  }
}

template <typename T, int METHOD>
__global__ void __launch_bounds__(kPoolThreads)
globalpool_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, const int32_t* __restrict__ aux, T* __restrict__ gx,
                      int64_t total, int64_t L) {
  using A = typename Acc<T>::type;
  const int64_t t = (int64_t)blockIdx.x * kPoolThreads + threadIdx.x;
  if (t >= total) return;
  const int64_t plane = t / L;
  const int64_t l = t - plane * L;
  bool nan = false;
  if (x != nullptr) { const A v = ld_acc(x + t); nan = v != v; }
  A g = 0;
  if (!nan) {
    if (METHOD == HG_POOL_AVG) { const int cnt = aux[plane]; if (cnt > 0) g = ld_acc(gy + plane) / (A)cnt; }
    else if ((int64_t)aux[plane] == l) g = ld_acc(gy + plane);
  }
  gx[t] = st_acc<T, A>(g);
}

// ---- host dispatch ------------------------------------------------------------------------------------
static int check_geom(int64_t planes, int64_t H, int64_t W, int64_t hn, int64_t wn, int kh, int kw, int sh, int sw, int shift,
                      int pad, int tail_h, int tail_w, PoolGeom& g) {
  HG_REQUIRE(planes >= 0 && H > 0 && W > 0 && hn >= 0 && wn >= 0, HG_E_SHAPE, "bad pool shape");
  HG_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && shift >= 0 && pad >= 0 && tail_h >= 0 && tail_w >= 0, HG_E_ARG, "bad pool parameters");
  HG_REQUIRE(H < (1 << 30) && W < (1 << 30) && hn < (1 << 30) && wn < (1 << 30), HG_E_SHAPE, "pool plane too large");
  const int64_t Hv = H + 2 * pad + tail_h, Wv = W + 2 * pad + tail_w;
  if (hn > 0 && wn > 0) {
    const int64_t max_r = (int64_t)sh * (hn - 1) + kh - 1;
    const int64_t max_c = (hn > 1 ? shift / 2 : 0) + (int64_t)sw * (wn - 1) + kw - 1;
    HG_REQUIRE(max_r < Hv && max_c < Wv, HG_E_SHAPE,
               "pool window leaves the image (row %lld of %lld, col %lld of %lld): the reference raises IndexError here",
               (long long)max_r, (long long)Hv, (long long)max_c, (long long)Wv);
  }
  g = PoolGeom{(int)H, (int)W, (int)hn, (int)wn, kh, kw, sh, sw, shift, pad, tail_h, tail_w};
  return HG_OK;
}

template <typename T, typename AUX>
static int launch_pool_fwd(const void* x, void* y, void* aux, int64_t planes, const PoolGeom& g, double pv, double tv, int method, cudaStream_t st) {
  using A = typename Acc<T>::type;
  int tw_log2 = 5;
  while ((1 << tw_log2) < g.wn && tw_log2 < 8) ++tw_log2;
  const unsigned jtiles = (unsigned)ceil_div(g.wn, 1 << tw_log2);
  const int64_t nrows64 = planes * g.hn;
  const int64_t blocks = ceil_div(nrows64, kPoolThreads >> tw_log2) * jtiles;
  HG_REQUIRE(blocks < (1ll << 31) && nrows64 < (1ll << 32), HG_E_SHAPE, "hexpool_fwd: too many rows for one launch");
  const unsigned grid = (unsigned)blocks, nrows = (unsigned)nrows64;
  switch (method) {
    case HG_POOL_MAX: hexpool_fwd_kernel<T, HG_POOL_MAX, AUX><<<grid, kPoolThreads, 0, st>>>((const T*)x, (T*)y, (AUX*)aux, jtiles, nrows, tw_log2, g, (A)pv, (A)tv); break;
    case HG_POOL_MIN: hexpool_fwd_kernel<T, HG_POOL_MIN, AUX><<<grid, kPoolThreads, 0, st>>>((const T*)x, (T*)y, (AUX*)aux, jtiles, nrows, tw_log2, g, (A)pv, (A)tv); break;
    default: hexpool_fwd_kernel<T, HG_POOL_AVG, AUX><<<grid, kPoolThreads, 0, st>>>((const T*)x, (T*)y, (AUX*)aux, jtiles, nrows, tw_log2, g, (A)pv, (A)tv); break;
  }
  return finish_launch("hexpool_fwd");
}
template <typename T, typename AUX>
static int launch_pool_bwd(const void* gy, const void* aux, const void* x, void* gx, int64_t planes, const PoolGeom& g, int method, cudaStream_t st) {
  int tw_log2 = 5;
  while ((1 << tw_log2) < g.W && tw_log2 < 8) ++tw_log2;
  const unsigned jtiles = (unsigned)ceil_div(g.W, 1 << tw_log2);
  const int64_t nrows64 = planes * g.H;
  const int64_t blocks = ceil_div(nrows64, kPoolThreads >> tw_log2) * jtiles;
  HG_REQUIRE(blocks < (1ll << 31) && nrows64 < (1ll << 32), HG_E_SHAPE, "hexpool_bwd: too many rows for one launch");
  const unsigned grid = (unsigned)blocks, nrows = (unsigned)nrows64;
  const bool disjoint = g.kh <= g.sh && g.kw <= g.sw;
#define HG_BWD(M)                                                                                                                  \
  if (disjoint) hexpool_bwd_kernel<T, M, AUX, true><<<grid, kPoolThreads, 0, st>>>((const T*)gy, (const AUX*)aux, (const T*)x, (T*)gx, jtiles, nrows, tw_log2, g); \
  else hexpool_bwd_kernel<T, M, AUX, false><<<grid, kPoolThreads, 0, st>>>((const T*)gy, (const AUX*)aux, (const T*)x, (T*)gx, jtiles, nrows, tw_log2, g);
  switch (method) {
    case HG_POOL_MAX: HG_BWD(HG_POOL_MAX) break;
    case HG_POOL_MIN: HG_BWD(HG_POOL_MIN) break;
    default: HG_BWD(HG_POOL_AVG) break;
  }
#undef HG_BWD
  return finish_launch("hexpool_bwd");
}

template <typename T>
static int launch_gpool_fwd(const void* x, void* y, int32_t* aux, int64_t planes, int64_t L, int method, cudaStream_t st) {
#define HG_GP(M)                                                                                                         \
  if (L >= 1024) globalpool_fwd_kernel<T, M, 256><<<(unsigned)planes, 256, 0, st>>>((const T*)x, (T*)y, aux, planes, L); \
  else globalpool_fwd_kernel<T, M, 32><<<(unsigned)ceil_div(planes, 8), 256, 0, st>>>((const T*)x, (T*)y, aux, planes, L);
  switch (method) {
    case HG_POOL_MAX: HG_GP(HG_POOL_MAX) break;
    case HG_POOL_MIN: HG_GP(HG_POOL_MIN) break;
    default: HG_GP(HG_POOL_AVG) break;
  }
#undef HG_GP
  return finish_launch("hexglobalpool_fwd");
}
template <typename T>
static int launch_gpool_bwd(const void* gy, const void* x, const int32_t* aux, void* gx, int64_t planes, int64_t L, int method, cudaStream_t st) {
  const int64_t total = planes * L;
  const unsigned grid = (unsigned)ceil_div(total, kPoolThreads);
  switch (method) {
    case HG_POOL_MAX: globalpool_bwd_kernel<T, HG_POOL_MAX><<<grid, kPoolThreads, 0, st>>>((const T*)gy, (const T*)x, aux, (T*)gx, total, L); break;
    case HG_POOL_MIN: globalpool_bwd_kernel<T, HG_POOL_MIN><<<grid, kPoolThreads, 0, st>>>((const T*)gy, (const T*)x, aux, (T*)gx, total, L); break;
    default: globalpool_bwd_kernel<T, HG_POOL_AVG><<<grid, kPoolThreads, 0, st>>>((const T*)gy, (const T*)x, aux, (T*)gx, total, L); break;
  }
  return finish_launch("hexglobalpool_bwd");
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_hexpool_fwd(const void* x, void* y, void* aux, int aux_bytes, int64_t planes, int64_t H, int64_t W, int64_t hn,
                   int64_t wn, int kh, int kw, int sh, int sw, int shift, int pad, double pad_value, int tail_h, int tail_w,
                   double tail_value, int method, int dtype, hg_stream_t stream) {
  PoolGeom g;
  int rc = check_geom(planes, H, W, hn, wn, kh, kw, sh, sw, shift, pad, tail_h, tail_w, g);
  if (rc) return rc;
  HG_REQUIRE(method >= HG_POOL_MAX && method <= HG_POOL_AVG, HG_E_ARG, "bad pool method %d", method);
  HG_REQUIRE(aux == nullptr || aux_bytes == 1 || aux_bytes == 4, HG_E_ARG, "aux_bytes must be 1 or 4");
  HG_REQUIRE(aux == nullptr || aux_bytes == 4 || (int64_t)kh * kw <= 127, HG_E_ARG, "windows of more than 127 cells need a 4-byte aux");
  const int64_t total = planes * hn * wn;
  if (total == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == HG_F32 && (aux == nullptr || aux_bytes == 1) && pool2x2_ok(g, tail_h, tail_w)) {
    const int pairs = (g.hn + 1) / 2, segs = (int)ceil_div(g.W, 128);
    const long long items = (long long)planes * pairs * segs;
    const long long blocks = ceil_div(items, kPoolThreads / 32);
    HG_REQUIRE(blocks < (1ll << 31), HG_E_SHAPE, "hexpool_fwd: too many rows for one launch");
    const bool vec = pool2x2_vec_ok(g, x, x);
#define HG_P2(M)                                                                                                                          \
  if (vec) hexpool2x2_fwd_kernel<M><<<(unsigned)blocks, kPoolThreads, 0, st>>>((const float*)x, (float*)y, (int8_t*)aux, items, pairs, segs, g.H, g.W, g.hn, g.wn); \
  else hexpool2x2c_fwd_kernel<M><<<(unsigned)blocks, kPoolThreads, 0, st>>>((const float*)x, (float*)y, (int8_t*)aux, items, pairs, segs, g.H, g.W, g.hn, g.wn)
    switch (method) {
      case HG_POOL_MAX: HG_P2(HG_POOL_MAX); break;
      case HG_POOL_MIN: HG_P2(HG_POOL_MIN); break;
      default: HG_P2(HG_POOL_AVG); break;
    }
#undef HG_P2
    return finish_launch("hexpool2x2_fwd");
  }
  const bool wide = aux != nullptr && aux_bytes == 4;
#define HG_CASE(D, T)                                                                                          \
  if (dtype == D) return wide ? launch_pool_fwd<T, int32_t>(x, y, aux, planes, g, pad_value, tail_value, method, st) \
                              : launch_pool_fwd<T, int8_t>(x, y, aux, planes, g, pad_value, tail_value, method, st);
  HG_CASE(HG_F32, float)
  HG_CASE(HG_F64, double)
  HG_CASE(HG_BF16, __nv_bfloat16)
#undef HG_CASE
  set_error("hexpool_fwd: unsupported dtype %d", dtype);
  return HG_E_DTYPE;
}

int hg_hexpool_bwd(const void* gy, const void* aux, int aux_bytes, const void* x, void* gx, int64_t planes, int64_t H,
                   int64_t W, int64_t hn, int64_t wn, int kh, int kw, int sh, int sw, int shift, int pad, int method,
                   int dtype, hg_stream_t stream) {
  PoolGeom g;
  int rc = check_geom(planes, H, W, 0, 0, kh, kw, sh, sw, shift, pad, 0, 0, g);
  if (rc) return rc;
  g.hn = (int)hn; g.wn = (int)wn;
  HG_REQUIRE(hn >= 0 && wn >= 0 && hn < (1 << 30) && wn < (1 << 30), HG_E_SHAPE, "bad pool output shape");
  HG_REQUIRE(method >= HG_POOL_MAX && method <= HG_POOL_AVG, HG_E_ARG, "bad pool method %d", method);
  HG_REQUIRE(aux != nullptr && (aux_bytes == 1 || aux_bytes == 4), HG_E_ARG, "hexpool_bwd needs the aux buffer written by the forward");
  const int64_t total = planes * H * W;
  if (total == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == HG_F32 && aux_bytes == 1 && pool2x2_ok(g, 0, 0)) {
    const int pairs = (int)ceil_div(g.H, 4), segs = (int)ceil_div(g.W, 128);
    const long long items = (long long)planes * pairs * segs;
    const long long blocks = ceil_div(items, kPoolThreads / 32);
    HG_REQUIRE(blocks < (1ll << 31), HG_E_SHAPE, "hexpool_bwd: too many rows for one launch");
    const bool vec = pool2x2_vec_ok(g, gx, x ? x : gx);
#define HG_P2(M)                                                                                                                          \
  if (vec) hexpool2x2_bwd_kernel<M><<<(unsigned)blocks, kPoolThreads, 0, st>>>((const float*)gy, (const int8_t*)aux, (const float*)x, (float*)gx, items, pairs, segs, g.H, g.W, g.hn, g.wn); \
  else hexpool2x2c_bwd_kernel<M><<<(unsigned)blocks, kPoolThreads, 0, st>>>((const float*)gy, (const int8_t*)aux, (const float*)x, (float*)gx, items, pairs, segs, g.H, g.W, g.hn, g.wn)
    switch (method) {
      case HG_POOL_MAX: HG_P2(HG_POOL_MAX); break;
      case HG_POOL_MIN: HG_P2(HG_POOL_MIN); break;
      default: HG_P2(HG_POOL_AVG); break;
    }
#undef HG_P2
    return finish_launch("hexpool2x2_bwd");
  }
#define HG_CASE(D, T)                                                                              \
  if (dtype == D) return aux_bytes == 4 ? launch_pool_bwd<T, int32_t>(gy, aux, x, gx, planes, g, method, st) \
                                        : launch_pool_bwd<T, int8_t>(gy, aux, x, gx, planes, g, method, st);
  HG_CASE(HG_F32, float)
  HG_CASE(HG_F64, double)
  HG_CASE(HG_BF16, __nv_bfloat16)
#undef HG_CASE
  set_error("hexpool_bwd: unsupported dtype %d", dtype);
  return HG_E_DTYPE;
}

int hg_hexglobalpool_fwd(const void* x, void* y, int32_t* aux_idx, int64_t planes, int64_t L, int method, int dtype,
                         hg_stream_t stream) {
  HG_REQUIRE(planes >= 0 && L > 0 && L < (1ll << 31) && planes < (1ll << 31), HG_E_SHAPE, "bad global pool shape");
  HG_REQUIRE(method >= HG_POOL_MAX && method <= HG_POOL_AVG, HG_E_ARG, "bad pool method %d", method);
  if (planes == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == HG_F32) return launch_gpool_fwd<float>(x, y, aux_idx, planes, L, method, st);
  if (dtype == HG_F64) return launch_gpool_fwd<double>(x, y, aux_idx, planes, L, method, st);
  if (dtype == HG_BF16) return launch_gpool_fwd<__nv_bfloat16>(x, y, aux_idx, planes, L, method, st);
  set_error("hexglobalpool_fwd: unsupported dtype %d", dtype);
  return HG_E_DTYPE;
}

int hg_hexglobalpool_bwd(const void* gy, const void* x, const int32_t* aux_idx, void* gx, int64_t planes, int64_t L,
                         int method, int dtype, hg_stream_t stream) {
  HG_REQUIRE(planes >= 0 && L > 0 && L < (1ll << 31), HG_E_SHAPE, "bad global pool shape");
  HG_REQUIRE(method >= HG_POOL_MAX && method <= HG_POOL_AVG, HG_E_ARG, "bad pool method %d", method);
  HG_REQUIRE(aux_idx != nullptr, HG_E_ARG, "hexglobalpool_bwd needs the aux buffer written by the forward");
  if (planes == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  if (dtype == HG_F32) return launch_gpool_bwd<float>(gy, x, aux_idx, gx, planes, L, method, st);
  if (dtype == HG_F64) return launch_gpool_bwd<double>(gy, x, aux_idx, gx, planes, L, method, st);
  if (dtype == HG_BF16) return launch_gpool_bwd<__nv_bfloat16>(gy, x, aux_idx, gx, planes, L, method, st);
  set_error("hexglobalpool_bwd: unsupported dtype %d", dtype);
  return HG_E_DTYPE;
}

}  // extern "C"
