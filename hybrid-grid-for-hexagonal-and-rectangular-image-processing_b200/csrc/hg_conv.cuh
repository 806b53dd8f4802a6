// hg_conv.cuh -- geometry shared by the hex-convolution kernels (direct stencil and tcgen05 paths).
//
// ref: HexFrames.py:96-169 HexConv2d.forward.  The reference materialises a 2x-wide "doubled" image and
// runs two dense strided F.conv2d with a zero-stuffed (2r-1)x(4r-3) window; in *offset* coordinates the
// same result is a sparse correlation whose tap columns depend only on the parity of the output row:
//
//   y[n,co,R,q] = bias[co] + sum_{ci,k} w[co,ci,0,k] * P[n, ci, s*R + ro[k], s*q + co[R&1][k]]
//
// with P the input framed by `pad` cells of pad_value (and literal zero one column beyond the frame,
// which is the doubled image's extra zero column), and for tap k = (a, t, m) in the reference's
// running-sum order (HexFrames.py:112-118):
//   ro[k]      = a*d
//   co[par][k] = floor((1 + par*s + t*d + 2*d*m - s_i) / 2),  s_i = ((s&1)*par + a*d + o) & 1,
//   o = (even_odd_offset + pad) & 1  ("parity" in hg_conv_desc).
#pragma once
#include "hg_common.cuh"
#include <stdlib.h>

namespace hg {

constexpr int kMaxTaps = 64;  // radius <= 5 (61 taps)

struct ConvTaps {
  int K;
  int ro[kMaxTaps];      // row offset in the padded frame
  int co[2][kMaxTaps];   // column offset in the padded frame, per output-row parity
};

struct ConvGeom {
  int N, Cin, Cout, H, W, Ho, Wo;
  int s, d, groups, pad, cin_g, cout_g;
  float pad_value;
  int relu;
  int pad_mode;          // 0 constant, 1 reflect, 2 replicate, 3 circular
};

// Source index of frame coordinate i (in [-pad, n + pad)) for the non-constant padding modes of F.pad.
__host__ __device__ __forceinline__ int conv_pad_remap(int i, int n, int mode) {
  if (mode == 1) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; }         // reflect (pad < n)
  else if (mode == 2) { i = i < 0 ? 0 : (i >= n ? n - 1 : i); }                   // replicate
  else { if (i < 0) i += n; else if (i >= n) i -= n; }                            // circular (pad <= n)
  return i;
}

static inline int conv_num_taps(int r) { return 3 * r * r - 3 * r + 1; }

static inline void conv_make_taps(int radius, int s, int d, int parity, ConvTaps& T) {
  int k = 0;
  for (int a = 0; a < 2 * radius - 1; ++a) {
    const int t = a - radius + 1 < 0 ? radius - 1 - a : a - radius + 1;
    const int ln = 2 * radius - 1 - t;
    for (int m = 0; m < ln; ++m, ++k) {
      T.ro[k] = a * d;
      for (int par = 0; par < 2; ++par) {
        const int si = ((s & 1) * par + a * d + parity) & 1;
        const int e = 1 + par * s + t * d + 2 * d * m - si;  // >= 0
        T.co[par][k] = e >> 1;
      }
    }
  }
  T.K = k;
}

// rows / cols of the interleaved output (HexFrames.py:127-162; k_h, k_w from :82-83)
static inline void conv_out_shape(int64_t Hp, int64_t Wp, int radius, int s, int d, int64_t& rows_e, int64_t& rows_o, int64_t& cols) {
  const int64_t k_h = (int64_t)(2 * radius - 2) * d + 1, k_w = (int64_t)2 * d * (2 * radius - 2) + 1;
  const int64_t wt = 2 * Wp - s;
  cols = wt >= k_w ? (wt - k_w) / (2 * s) + 1 : 0;
  rows_e = Hp >= k_h ? (Hp - k_h) / (2 * s) + 1 : 0;
  rows_o = Hp - s >= k_h ? (Hp - s - k_h) / (2 * s) + 1 : 0;
}

// Output rows per work item of the persistent tcgen05 kernels: `max_band` rows when the problem has plenty of items, shorter
// bands (down to 8 rows; each band re-stages its 2 halo rows) while a launch would otherwise leave most of the 148 SMs with
// one item or none -- the C5 layers (64 images of 64 x 64 or 32 x 32 cells) are 64-128 items at 32 rows per band.
// HG_CONV_BAND=<rows> forces a value (A/B runs).
static inline int conv_pick_band(int max_band, long long N, int Ho, int ctiles) {
  const char* e = getenv("HG_CONV_BAND");
  if (e && atoi(e) >= 1) return atoi(e) < max_band ? atoi(e) : max_band;
  int band = max_band;
  while (band > 8 && N * ((Ho + band - 1) / band) * ctiles < 4 * 148) band >>= 1;
  return band;
}

}  // namespace hg
