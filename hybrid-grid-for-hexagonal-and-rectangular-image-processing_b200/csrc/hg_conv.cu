// hg_conv.cu -- C ABI of the hex convolution: descriptor checks, tap tables, algorithm choice.
// ref: HexFrames.py:22-185 HexConv2d (forward :96-169); geometry in hg_conv.cuh.
#include "hg_conv.cuh"

namespace hg {
// hg_conv_direct.cu
int conv_fwd_direct(const ConvGeom&, const ConvTaps&, const void* x, int xdt, const float* w, const float* scale, const float* bias, void* y, int ydt, cudaStream_t);
int conv_dgrad_direct(const ConvGeom&, const ConvTaps&, const void* gy, int gdt, const float* w, void* gx, int xdt, cudaStream_t);
int conv_wgrad_direct(const ConvGeom&, const ConvTaps&, const void* x, int xdt, const void* gy, int gdt, float* gw, float* gbias, cudaStream_t);
// hg_conv_umma.cu
bool conv_umma_eligible(const hg_conv_desc* d, int op);
int conv_fwd_umma(const hg_conv_desc* d, const ConvGeom&, const ConvTaps&, const void* x, const float* w, const float* scale, const float* bias, void* y, cudaStream_t);
int conv_dgrad_umma(const hg_conv_desc* d, const ConvGeom&, const ConvTaps&, const void* gy, const float* w, void* gx, cudaStream_t);
int conv_wgrad_umma(const hg_conv_desc* d, const ConvGeom&, const ConvTaps&, const void* x, const void* gy, float* gw, float* gbias, cudaStream_t);

static int make_geom(const hg_conv_desc* d, ConvGeom& g, ConvTaps& tp) {
  HG_REQUIRE(d != nullptr, HG_E_ARG, "conv descriptor is NULL");
  HG_REQUIRE(d->N >= 0 && d->Cin > 0 && d->Cout > 0 && d->H > 0 && d->W > 0, HG_E_SHAPE, "bad conv shape N=%lld Cin=%lld Cout=%lld H=%lld W=%lld",
             (long long)d->N, (long long)d->Cin, (long long)d->Cout, (long long)d->H, (long long)d->W);
  HG_REQUIRE(d->radius >= 1 && conv_num_taps(d->radius) <= kMaxTaps, HG_E_UNSUPPORTED, "hexkernel_radius must be in 1..5 (got %d)", d->radius);
  HG_REQUIRE(d->stride >= 1 && d->dilation >= 1 && d->groups >= 1 && d->pad >= 0, HG_E_ARG, "bad stride/dilation/groups/padding");
  HG_REQUIRE(d->Cin % d->groups == 0, HG_E_ARG, "in_channels must be divisible by groups");
  HG_REQUIRE(d->Cout % d->groups == 0, HG_E_ARG, "out_channels must be divisible by groups");
  HG_REQUIRE((d->x_dtype == HG_F32 || d->x_dtype == HG_BF16) && (d->y_dtype == HG_F32 || d->y_dtype == HG_BF16), HG_E_DTYPE,
             "conv activations must be float32 or bfloat16");
  HG_REQUIRE(d->H < (1 << 24) && d->W < (1 << 24) && d->N < (1ll << 31) && d->Cin < (1 << 24) && d->Cout < (1 << 24), HG_E_SHAPE, "conv shape too large");
  int64_t re, ro, cols;
  conv_out_shape(d->H + 2 * d->pad, d->W + 2 * d->pad, d->radius, d->stride, d->dilation, re, ro, cols);
  HG_REQUIRE(re > 0 && ro >= 0 && cols > 0 && (re - ro == 0 || re - ro == 1), HG_E_SHAPE,
             "input %lldx%lld is too small for this hex kernel (even rows %lld, odd rows %lld, cols %lld)", (long long)d->H,
             (long long)d->W, (long long)re, (long long)ro, (long long)cols);
  HG_REQUIRE(d->Ho == re + ro && d->Wo == cols, HG_E_SHAPE, "output shape must be %lldx%lld (got %lldx%lld)", (long long)(re + ro),
             (long long)cols, (long long)d->Ho, (long long)d->Wo);
  g.N = (int)d->N; g.Cin = (int)d->Cin; g.Cout = (int)d->Cout; g.H = (int)d->H; g.W = (int)d->W; g.Ho = (int)d->Ho; g.Wo = (int)d->Wo;
  g.s = d->stride; g.d = d->dilation; g.groups = d->groups; g.pad = d->pad;
  g.cin_g = g.Cin / g.groups; g.cout_g = g.Cout / g.groups;
  g.pad_value = d->pad_value; g.relu = d->relu; g.pad_mode = d->pad_mode;
  HG_REQUIRE(d->pad_mode >= 0 && d->pad_mode <= 3, HG_E_ARG, "pad_mode must be 0 (constant), 1 (reflect), 2 (replicate) or 3 (circular)");
  HG_REQUIRE(d->pad_mode != 1 || (d->pad < d->H && d->pad < d->W), HG_E_SHAPE, "reflect padding must be smaller than the image");
  HG_REQUIRE(d->pad_mode != 3 || (d->pad <= d->H && d->pad <= d->W), HG_E_SHAPE, "circular padding must not exceed the image");
  conv_make_taps(d->radius, d->stride, d->dilation, d->parity & 1, tp);
  return HG_OK;
}

enum { OP_FWD = 0, OP_DGRAD = 1, OP_WGRAD = 2 };
static int pick_algo(const hg_conv_desc* d, int op, bool& umma) {
  HG_REQUIRE(d->algo >= 0 && d->algo <= 2, HG_E_ARG, "algo must be 0 (auto), 1 (direct) or 2 (tcgen05)");
  const bool ok = conv_umma_eligible(d, op);
  if (d->algo == 2) {
    HG_REQUIRE(ok, HG_E_UNSUPPORTED, "the tcgen05 implicit-GEMM path does not cover this configuration");
    umma = true;
  } else {
    umma = (d->algo == 0) && ok;
  }
  return HG_OK;
}
}  // namespace hg

using namespace hg;

extern "C" {

int hg_hexconv_out_shape(int64_t H, int64_t W, int radius, int stride, int dilation, int pad, int64_t* Ho, int64_t* Wo) {
  HG_REQUIRE(H > 0 && W > 0 && radius >= 1 && stride >= 1 && dilation >= 1 && pad >= 0 && Ho && Wo, HG_E_ARG, "bad arguments");
  int64_t re, ro, cols;
  conv_out_shape(H + 2 * pad, W + 2 * pad, radius, stride, dilation, re, ro, cols);
  HG_REQUIRE(re > 0 && ro >= 0 && cols > 0 && (re - ro == 0 || re - ro == 1), HG_E_SHAPE,
             "input %lldx%lld is too small for this hex kernel", (long long)H, (long long)W);
  *Ho = re + ro;
  *Wo = cols;
  return HG_OK;
}

int hg_hexconv_umma_eligible(const hg_conv_desc* d, int op) {
  ConvGeom g; ConvTaps tp;
  if (d == nullptr || op < 0 || op > 2 || make_geom(d, g, tp) != HG_OK) return 0;
  hg_conv_desc forced = *d;
  forced.algo = 2;
  return conv_umma_eligible(&forced, op) ? 1 : 0;
}

int hg_hexconv_fwd_affine(const hg_conv_desc* d, const void* x, const float* w, const float* scale, const float* shift, void* y,
                          hg_stream_t stream) {
  ConvGeom g; ConvTaps tp;
  int rc = make_geom(d, g, tp);
  if (rc) return rc;
  if (g.N == 0) return HG_OK;
  bool umma;
  if ((rc = pick_algo(d, OP_FWD, umma))) return rc;
  HG_REQUIRE(!d->accumulate || (umma && shift == nullptr && scale == nullptr), HG_E_UNSUPPORTED,
             "accumulate is a tcgen05-path option without bias / scale");
  if (umma) return conv_fwd_umma(d, g, tp, x, w, scale, shift, y, as_stream(stream));
  return conv_fwd_direct(g, tp, x, d->x_dtype, w, scale, shift, y, d->y_dtype, as_stream(stream));
}

int hg_hexconv_fwd(const hg_conv_desc* d, const void* x, const float* w, const float* bias, void* y, hg_stream_t stream) {
  return hg_hexconv_fwd_affine(d, x, w, nullptr, bias, y, stream);
}

int hg_hexconv_dgrad(const hg_conv_desc* d, const void* gy, const float* w, void* gx, hg_stream_t stream) {
  ConvGeom g; ConvTaps tp;
  int rc = make_geom(d, g, tp);
  if (rc) return rc;
  HG_REQUIRE(g.pad_mode == 0 || g.pad == 0, HG_E_UNSUPPORTED,
             "hg_hexconv_dgrad: reflect / replicate / circular frames are folded by the caller (pad = 0 on the padded geometry, then hg_pad2d_bwd)");
  if (g.N == 0) return HG_OK;
  bool umma;
  if ((rc = pick_algo(d, OP_DGRAD, umma))) return rc;
  HG_REQUIRE(!d->accumulate || umma, HG_E_UNSUPPORTED, "accumulate is a tcgen05-path option");
  if (umma) return conv_dgrad_umma(d, g, tp, gy, w, gx, as_stream(stream));
  return conv_dgrad_direct(g, tp, gy, d->y_dtype, w, gx, d->x_dtype, as_stream(stream));
}

int hg_hexconv_wgrad(const hg_conv_desc* d, const void* x, const void* gy, float* gw, float* gbias, hg_stream_t stream) {
  ConvGeom g; ConvTaps tp;
  int rc = make_geom(d, g, tp);
  if (rc) return rc;
  if (g.N == 0) return HG_OK;
  bool umma;
  if ((rc = pick_algo(d, OP_WGRAD, umma))) return rc;
  if (umma) return conv_wgrad_umma(d, g, tp, x, gy, gw, gbias, as_stream(stream));
  return conv_wgrad_direct(g, tp, x, d->x_dtype, gy, d->y_dtype, gw, gbias, as_stream(stream));
}

}  // extern "C"
