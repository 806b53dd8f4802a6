// hg_hexgeom.cuh -- per-sample geometry of the hex-source resamplers (hex -> rect, hexresize, warp):
// cell lookup, triangle pick, barycentric weights.  ref: geometry_np.py:276-354, geometry_torch.py:278-356.
#pragma once
#include "hg_common.cuh"

namespace hg {

template <typename CT> struct Arith;
template <> struct Arith<double> {
  static __device__ __forceinline__ double add(double a, double b) { return dadd(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return dsub(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return dmul(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
};
template <> struct Arith<float> {
  static __device__ __forceinline__ float add(float a, float b) { return fadd(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return fsub(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return fmul(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
};

// The three lattice points fetched for one sample and their weights.
template <typename CT>
struct HexSample {
  int i_n, j_n;
  int off[3];   // linear offset i*w + j_off of P1, (P2 or P3), P4; -1 = outside (zero-filled)
  bool flag;    // up_down_flag = i_f > j_f
  CT wgt[3];    // barycentric weights alpha, beta, gamma (linear)
  int nearest;  // index (0..2) of the closest vertex (nearest)
};

template <typename CT, bool WANT_LINEAR, bool WANT_NEAREST>
__device__ __forceinline__ void hex_locate(CT x, CT y, int h, int w, CT hx, CT wy, CT ci, CT cj, HexSample<CT>& s) {
  using A = Arith<CT>;
  // ref geometry_np.py:276-285: i_ = x_ + (h-1)*0.5 ; j_ = 0.5*i_ + y_ + (w-0.5)*0.5
  const CT i_ = A::add(x, ci);
  const CT j_ = A::add(A::add(A::mul(CT(0.5), i_), y), cj);
  const int in = trunc_i32(i_), jn = trunc_i32(j_);
  const CT i_f = A::sub(i_, (CT)(float)in), j_f = A::sub(j_, (CT)(float)jn);
  const bool f = i_f > j_f;  // :298
  s.i_n = in; s.j_n = jn; s.flag = f;
  // :288-295 axial -> offset columns of the cell's lattice points
  const int j1 = jn - trunc_half(in + 1), j2 = jn - trunc_half(in + 2);
  const int iA = in, jA = j1;                       // P1
  const int iB = f ? in + 1 : in, jB = f ? j2 : j1 + 1;  // P2 (below) or P3 (right)
  const int iC = in + 1, jC = j2 + 1;               // P4
  s.off[0] = (iA >= 0 && iA < h && jA >= 0 && jA < w) ? iA * w + jA : -1;
  s.off[1] = (iB >= 0 && iB < h && jB >= 0 && jB < w) ? iB * w + jB : -1;
  s.off[2] = (iC >= 0 && iC < h && jC >= 0 && jC < w) ? iC * w + jC : -1;
  // :326-331 cartesian coordinates of the triangle vertices
  const CT fi = (CT)in, fj = (CT)jn, ff = f ? CT(1) : CT(0);
  const CT p1x = A::sub(fi, hx);
  const CT p1y = A::sub(A::sub(fj, A::mul(fi, CT(0.5))), wy);
  const CT p2x = A::sub(A::add(fi, ff), hx);
  const CT p2y = A::sub(A::sub(A::sub(A::add(fj, CT(1)), ff), A::mul(A::add(fi, ff), CT(0.5))), wy);
  const CT p3x = A::sub(A::add(fi, CT(1)), hx);
  const CT p3y = A::sub(A::sub(A::add(fj, CT(1)), A::mul(A::add(fi, CT(1)), CT(0.5))), wy);
  const CT ax = A::sub(x, p1x), ay = A::sub(y, p1y);
  const CT bx = A::sub(x, p2x), by = A::sub(y, p2y);
  const CT cx = A::sub(x, p3x), cy = A::sub(y, p3y);
  if (WANT_NEAREST) {  // :334-347, first minimum wins
    const CT d1 = A::add(A::mul(ax, ax), A::mul(ay, ay));
    const CT d2 = A::add(A::mul(bx, bx), A::mul(by, by));
    const CT d3 = A::add(A::mul(cx, cx), A::mul(cy, cy));
    int sel = 0; CT best = d1;
    if (d2 < best) { best = d2; sel = 1; }
    if (d3 < best) { sel = 2; }
    s.nearest = sel;
  }
  if (WANT_LINEAR) {   // :348-354 sub-triangle areas
    const CT S1 = A::mul(CT(0.5), A::abs(A::sub(A::mul(bx, cy), A::mul(by, cx))));
    const CT S2 = A::mul(CT(0.5), A::abs(A::sub(A::mul(ax, cy), A::mul(ay, cx))));
    const CT S3 = A::mul(CT(0.5), A::abs(A::sub(A::mul(ax, by), A::mul(ay, bx))));
    const CT tot = A::add(A::add(S1, S2), S3);
    s.wgt[0] = A::div(S1, tot);
    s.wgt[1] = A::div(S2, tot);
    s.wgt[2] = A::div(S3, tot);
  }
}

// Fast (fp32) variant: simplex interpolation in the axial unit cell (SURVEY 8a closed form):
// u = i_f, v = j_f;  u > v : (1-u, u-v, v)   else (1-v, v-u, u).  Index math stays exact.
__device__ __forceinline__ void hex_locate_fast(double x, double y, int h, int w, double ci, double cj,
                                                HexSample<float>& s) {
  const double i_ = dadd(x, ci);
  const double j_ = dadd(dadd(dmul(0.5, i_), y), cj);
  const int in = trunc_i32(i_), jn = trunc_i32(j_);
  const double i_f = dsub(i_, (double)in), j_f = dsub(j_, (double)jn);
  const bool f = i_f > j_f;
  s.i_n = in; s.j_n = jn; s.flag = f;
  const int j1 = jn - trunc_half(in + 1), j2 = jn - trunc_half(in + 2);
  const int iA = in, jA = j1;
  const int iB = f ? in + 1 : in, jB = f ? j2 : j1 + 1;
  const int iC = in + 1, jC = j2 + 1;
  s.off[0] = (iA >= 0 && iA < h && jA >= 0 && jA < w) ? iA * w + jA : -1;
  s.off[1] = (iB >= 0 && iB < h && jB >= 0 && jB < w) ? iB * w + jB : -1;
  s.off[2] = (iC >= 0 && iC < h && jC >= 0 && jC < w) ? iC * w + jC : -1;
  const float u = (float)i_f, v = (float)j_f;
  s.wgt[0] = f ? 1.f - u : 1.f - v;
  s.wgt[1] = f ? u - v : v - u;
  s.wgt[2] = f ? v : u;
}


}  // namespace hg
