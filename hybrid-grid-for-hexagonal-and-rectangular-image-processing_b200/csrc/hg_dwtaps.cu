// hg_dwtaps.cu -- depthwise weighted tap gathers on doubled coordinates: the learned lattice resamplers the reference
// retired into "HyGrid/codes in old versions.txt" (SURVEY.md section 8f rank 3):
//   Hex_to_Square_Conv2d_by_Double_Stride   (:1-66)     hex lattice -> square grid, f x f rhombus window, stride (f, 2f-1)
//   Square_to_Hex_Conv2d_by_Double_Stride   (:421-493)  square grid -> hex lattice, 2 x 2 window, odd rows start one pixel right
//   Hex_to_Square_original_resolution       (:587-636)  odd rows re-interpolated from a 4-cell rhombus, even rows kept
// The reference materialises the doubled ("type1") image, unfolds the window with one strided slice + torch.cat per
// tap and multiplies channel by channel in a Python loop.  All three are the same operator:
//
//   y[n,c,R,J] = sum_t w_s[c,t] * A_s(sy*R + ry_s[t], sx_s*J + ex_s[t]),      s = (R odd and R < odd_limit) ? 1 : 0
//
// with two tap sets (even / odd output rows), each reading either the doubled view of the (virtually padded) hex
// lattice -- A(i,c) = P[i, (c - s_i) >> 1] for s_i <= c < 2*Wp + s_i, s_i = (i + parity) & 1, literal 0 elsewhere -- or
// the plain padded image.  One thread per output cell (lanes along J: coalesced stores, taps of neighbouring lanes in
// the same source lines); HBM-bound streaming.  The data gradient is the adjoint scatter (fp32 atomics into a zeroed
// gx), the weight gradient a per-channel reduction (warp shuffles -> shared -> one atomic per block and tap).
#include "hg_common.cuh"

namespace hg {

constexpr int kDwThreads = 256;

struct DwGeom {
  int N, C, H, W, Ho, Wo, Hp, Wp;
  int sy, odd_limit, parity, pad;
  float pad_value;
  hg_taps_set set[2];
};

// value of the (virtually padded) source at padded row r, coordinate col of tap set s; `idx` = offset inside the plane or -1
__device__ __forceinline__ float dw_fetch(const float* __restrict__ plane, const DwGeom& g, int doubled, int r, int col, int& idx) {
  idx = -1;
  if (r < 0 || r >= g.Hp) return 0.f;
  int pc = col;
  if (doubled) {
    const int si = (r + g.parity) & 1;
    if (col < si || col >= 2 * g.Wp + si) return 0.f;      // the doubled image's extra zero column
    pc = (col - si) >> 1;
  } else if (col < 0 || col >= g.Wp) {
    return 0.f;
  }
  const int i = r - g.pad, j = pc - g.pad;
  if (i < 0 || i >= g.H || j < 0 || j >= g.W) return g.pad_value;
  idx = i * g.W + j;
  return __ldg(plane + idx);
}

__global__ void __launch_bounds__(kDwThreads)
dwtaps_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w0, const float* __restrict__ w1, float* __restrict__ y, DwGeom g) {
  const long long total = (long long)g.N * g.C * g.Ho * g.Wo;
  for (long long e = (long long)blockIdx.x * kDwThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kDwThreads) {
    const int J = (int)(e % g.Wo);
    long long q = e / g.Wo;
    const int R = (int)(q % g.Ho);
    q /= g.Ho;
    const int c = (int)(q % g.C);
    const int s = ((R & 1) && R < g.odd_limit) ? 1 : 0;
    const hg_taps_set& ts = g.set[s];
    const float* __restrict__ wt = s ? w1 : w0;
    const float* __restrict__ plane = x + q * (long long)g.H * g.W;
    float acc = 0.f;
    for (int t = 0; t < ts.T; ++t) {
      int idx;
      const float v = dw_fetch(plane, g, ts.doubled, g.sy * R + ts.ry[t], ts.sx * J + ts.ex[t], idx);
      acc = fmaf(wt ? __ldg(wt + (long long)c * ts.T + t) : 1.f, v, acc);
    }
    y[e] = acc;
  }
}

__global__ void __launch_bounds__(kDwThreads)
dwtaps_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w0, const float* __restrict__ w1, float* __restrict__ gx, DwGeom g) {
  const long long total = (long long)g.N * g.C * g.Ho * g.Wo;
  for (long long e = (long long)blockIdx.x * kDwThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kDwThreads) {
    const int J = (int)(e % g.Wo);
    long long q = e / g.Wo;
    const int R = (int)(q % g.Ho);
    q /= g.Ho;
    const int c = (int)(q % g.C);
    const int s = ((R & 1) && R < g.odd_limit) ? 1 : 0;
    const hg_taps_set& ts = g.set[s];
    const float* __restrict__ wt = s ? w1 : w0;
    float* __restrict__ plane = gx + q * (long long)g.H * g.W;
    const float go = gy[e];
    for (int t = 0; t < ts.T; ++t) {
      int idx;
      dw_fetch(plane, g, ts.doubled, g.sy * R + ts.ry[t], ts.sx * J + ts.ex[t], idx);
      if (idx >= 0) atomicAdd(plane + idx, go * (wt ? __ldg(wt + (long long)c * ts.T + t) : 1.f));
    }
  }
}

// grid = (chunks, C): block (k, c) reduces its share of the (n, R, J) cells of channel c for tap set `s`
__global__ void __launch_bounds__(kDwThreads)
dwtaps_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw, DwGeom g, int s) {
  __shared__ float part[64];
  const hg_taps_set& ts = g.set[s];
  const int c = blockIdx.y, lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < 64; t += kDwThreads) part[t] = 0.f;
  __syncthreads();
  const long long per_c = (long long)g.N * g.Ho * g.Wo;
  for (int t = 0; t < ts.T; ++t) {
    float acc = 0.f;
    for (long long e = (long long)blockIdx.x * kDwThreads + threadIdx.x; e < per_c; e += (long long)gridDim.x * kDwThreads) {
      const int J = (int)(e % g.Wo);
      long long q = e / g.Wo;
      const int R = (int)(q % g.Ho);
      const long long n = q / g.Ho;
      if ((((R & 1) && R < g.odd_limit) ? 1 : 0) != s) continue;
      const long long pl = n * g.C + c;
      int idx;
      const float v = dw_fetch(x + pl * (long long)g.H * g.W, g, ts.doubled, g.sy * R + ts.ry[t], ts.sx * J + ts.ex[t], idx);
      acc = fmaf(gy[(pl * g.Ho + R) * (long long)g.Wo + J], v, acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(&part[t], acc);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ts.T; t += kDwThreads) atomicAdd(gw + (long long)c * ts.T + t, part[t]);
}

static int dw_geom(const hg_dwtaps_desc* d, DwGeom& g) {
  HG_REQUIRE(d != nullptr, HG_E_ARG, "tap descriptor is NULL");
  HG_REQUIRE(d->N >= 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->Ho >= 0 && d->Wo >= 0, HG_E_SHAPE, "bad shape");
  HG_REQUIRE(d->H < (1 << 24) && d->W < (1 << 24) && d->Ho < (1 << 24) && d->Wo < (1 << 24) && d->N < (1ll << 31) && d->C < 65536, HG_E_SHAPE,
             "shape too large");
  HG_REQUIRE(d->sy >= 1 && d->pad >= 0, HG_E_ARG, "bad row stride / padding");
  for (int s = 0; s < 2; ++s)
    HG_REQUIRE(d->set[s].T >= 1 && d->set[s].T <= 64 && d->set[s].sx >= 1, HG_E_ARG, "tap set %d: 1..64 taps, column stride >= 1", s);
  g.N = (int)d->N; g.C = (int)d->C; g.H = (int)d->H; g.W = (int)d->W; g.Ho = (int)d->Ho; g.Wo = (int)d->Wo;
  g.Hp = g.H + 2 * d->pad; g.Wp = g.W + 2 * d->pad;
  g.sy = d->sy; g.odd_limit = d->odd_limit; g.parity = d->parity & 1; g.pad = d->pad; g.pad_value = d->pad_value;
  g.set[0] = d->set[0]; g.set[1] = d->set[1];
  return HG_OK;
}

static unsigned dw_grid(long long total) {
  long long b = ceil_div(total, kDwThreads);
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_dwtaps_fwd(const hg_dwtaps_desc* d, const float* x, const float* w_even, const float* w_odd, float* y, hg_stream_t stream) {
  DwGeom g;
  int rc = dw_geom(d, g);
  if (rc) return rc;
  const long long total = (long long)g.N * g.C * g.Ho * g.Wo;
  if (total == 0) return HG_OK;
  HG_REQUIRE(x && y, HG_E_ARG, "NULL buffer");
  dwtaps_fwd_kernel<<<dw_grid(total), kDwThreads, 0, as_stream(stream)>>>(x, w_even, w_odd, y, g);
  return finish_launch("dwtaps_fwd");
}

int hg_dwtaps_dgrad(const hg_dwtaps_desc* d, const float* gy, const float* w_even, const float* w_odd, float* gx, hg_stream_t stream) {
  DwGeom g;
  int rc = dw_geom(d, g);
  if (rc) return rc;
  const long long total = (long long)g.N * g.C * g.Ho * g.Wo;
  if (total == 0) return HG_OK;
  HG_REQUIRE(gy && gx, HG_E_ARG, "NULL buffer");
  dwtaps_dgrad_kernel<<<dw_grid(total), kDwThreads, 0, as_stream(stream)>>>(gy, w_even, w_odd, gx, g);
  return finish_launch("dwtaps_dgrad");
}

int hg_dwtaps_wgrad(const hg_dwtaps_desc* d, const float* x, const float* gy, float* gw, int set, hg_stream_t stream) {
  DwGeom g;
  int rc = dw_geom(d, g);
  if (rc) return rc;
  HG_REQUIRE(set == 0 || set == 1, HG_E_ARG, "tap set must be 0 (even rows) or 1 (odd rows)");
  const long long per_c = (long long)g.N * g.Ho * g.Wo;
  if (per_c == 0) return HG_OK;
  HG_REQUIRE(x && gy && gw, HG_E_ARG, "NULL buffer");
  long long chunks = ceil_div(per_c, (long long)kDwThreads * 8);
  if (chunks > 64) chunks = 64;
  dim3 grid((unsigned)chunks, (unsigned)g.C);
  dwtaps_wgrad_kernel<<<grid, kDwThreads, 0, as_stream(stream)>>>(x, gy, gw, g, set);
  return finish_launch("dwtaps_wgrad");
}

}  // extern "C"
