/* hygrid_b200.h -- C ABI of libhygrid_b200.so (sm_100a CUDA kernels for the HyGrid hot path).
 *
 * The reference (HyGrid, pure Python) has no FFI: the path sits behind plain Python call
 * signatures.  Every entry point below therefore names the reference
 * *function* it replaces as `ref:` (file:line under /root/reference/HyGrid/); INTEGRATION.md
 * shows the ctypes stub a HyGrid maintainer would add at that call site.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless the
 *     parameter is called host_*; tensors are dense, row-major, planes = N*C leading images.
 *   - the library never allocates, frees or keeps tensor memory; outputs are caller-allocated.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls only
 *     enqueue work; they do not synchronise.
 *   - return 0 on success, <0 invalid argument (HG_E_*), >0 a cudaError_t.  hg_last_error()
 *     returns a thread-local message for the last non-zero return.
 *   - re-entrant, no global mutable state besides the thread-local error string (and the per-device workspace
 *     of the hg_host_* entry points, guarded by one lock per device).
 */
#ifndef HYGRID_B200_H_
#define HYGRID_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_VERSION 100 /* 0.1.0 */

typedef void* hg_stream_t;

/* element types */
enum { HG_U8 = 0, HG_I16 = 1, HG_I32 = 2, HG_I64 = 3, HG_F32 = 4, HG_F64 = 5, HG_BF16 = 6, HG_U16 = 7 };
/* arithmetic of the interpolating kernels:
 *   HG_MATH_EXACT  float64, no FMA contraction, the reference's operation order.  With a
 *                  float64 destination the result is bit-identical to the reference; with a
 *                  float32 destination it is that value rounded once.
 *   HG_MATH_FAST   float32 weights and FMAs (|err| <= 1e-5 * max|x|).                          */
enum { HG_MATH_EXACT = 0, HG_MATH_FAST = 1 };
/* pooling reductions (ref: HexFrames.py:461-479) */
enum { HG_POOL_MAX = 0, HG_POOL_MIN = 1, HG_POOL_AVG = 2 };
/* error codes */
enum { HG_OK = 0, HG_E_ARG = -1, HG_E_DTYPE = -2, HG_E_SHAPE = -3, HG_E_UNSUPPORTED = -4 };

int hg_version(void);
const char* hg_last_error(void);
/* number of kernels launched by this thread since the last hg_reset_launch_count() (bench.py's
 * gpu_launches claim is read from here, not guessed). */
int64_t hg_launch_count(void);
/* name of the kernel family the calling thread launched last ("rect2hex_bilinear_stream", "hexconv_umma_tma", ...):
 * which of the library's kernels took a call is part of what the tests and the bench line state. */
const char* hg_last_launch(void);
void hg_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Lattice index kernels (integer; bit-exact with the reference)
 * ---------------------------------------------------------------------------------------- */

/* axial <-> offset column index of the "odd rows shifted right" hex lattice.
 * ref: geometry_np.py:288-295  j_off = j_ax - trunc((i+1)/2)  (true division, truncation). */
int hg_axial_to_offset_i32(const int32_t* i, const int32_t* j_ax, int32_t* j_off, int64_t n, hg_stream_t stream);
int hg_offset_to_axial_i32(const int32_t* i, const int32_t* j_off, int32_t* j_ax, int64_t n, hg_stream_t stream);

/* Row / column tables of rect->hex sampling.  ref: geometry_np.py:440-449.
 * xs[h1], ys[w1]: the reference's 1-D sample coordinates (host computes them with the very
 * same linspace call).  Outputs: i_n[h1], j_n[w1] truncated indices; i_f[h1], j_f[w1]
 * fractions (float64). */
int hg_rect2hex_index(const double* xs, const double* ys, int64_t h, int64_t w, int64_t h1, int64_t w1,
                      int32_t* i_n, double* i_f, int32_t* j_n, double* j_f, hg_stream_t stream);

/* Per-sample tables of sampling a hex-lattice image.  ref: geometry_np.py:276-316
 * (== geometry_torch.py:278-316).  Coordinates are either separable (coords_2d = 0: xs[h1],
 * ys[w1]) or full planes (coords_2d = 1: xs[h1*w1], ys[h1*w1]); coord_f32 = 1 evaluates the
 * index arithmetic in float32 like the torch warp (geometry_torch.py:99-118).
 * Outputs (each h1*w1): i_n, j_n (axial cell), tri (bit0 = up_down_flag i_f > j_f, bits 1..3 =
 * validity of the three fetched lattice points P1, P2|P3, P4), off[3*h1*w1] linear offsets
 * i*w + j_off of the three points (-1 where invalid). */
int hg_hexsrc_index(const void* xs, const void* ys, int coords_2d, int coord_f32,
                    int64_t h, int64_t w, int64_t h1, int64_t w1,
                    int32_t* i_n, int32_t* j_n, uint8_t* tri, int32_t* off, hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Resampling (memory-bound gathers).  src: [planes, h, w] -> dst: [planes, h1, w1]
 * ---------------------------------------------------------------------------------------- */

/* ref: geometry_np.py:358-519 rect_to_hex_resample(..., 'nearest'): literal 4-way argmin of
 * geometry_np.py:499-512.  dst has the src element type (elem_size 1, 2, 4 or 8 bytes). */
int hg_rect2hex_nearest(const void* src, void* dst, const double* xs, const double* ys,
                        int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                        int elem_size, hg_stream_t stream);
/* ref: geometry_np.py:358-519 rect_to_hex_resample(..., 'bilinear') (blend :514-517).
 * src_dtype in {HG_U8, HG_F32, HG_F64}, dst_dtype in {HG_F32, HG_F64}.
 * host_xs / host_ys (may be NULL): HOST copies of the same tables.  With them the library sizes the
 * source footprint of an output tile and, when the two lattices have similar pitch (float32 in/out),
 * runs the TMA-staged tile kernel instead of the direct gather; results are identical. */
int hg_rect2hex_bilinear(const void* src, void* dst, const double* xs, const double* ys,
                         const double* host_xs, const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                         int src_dtype, int dst_dtype, int math, hg_stream_t stream);

/* ref: geometry_torch.py:191-358 hex_to_square_resample / geometry_np.py:191-356
 * hex_to_rect_resample / geometry_np.py:520-681 hexresize -- they differ only in the 1-D
 * coordinate tables xs[h1], ys[w1], which the host computes with the reference's own
 * linspace call (float32 torch.linspace for the torch twin). */
int hg_hex2rect_nearest(const void* src, void* dst, const double* xs, const double* ys,
                        int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                        int elem_size, hg_stream_t stream);
/* host_xs / host_ys (may be NULL): host copies of the tables; with them float32 HG_MATH_FAST calls on
 * lattices of similar pitch run the TMA-staged tile kernel (identical indices, float32 weights). */
int hg_hex2rect_linear(const void* src, void* dst, const double* xs, const double* ys,
                       const double* host_xs, const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                       int src_dtype, int dst_dtype, int math, hg_stream_t stream);

/* ref: geometry_torch.py:7-189 image_geometric_transformation_gpu (coord_f32 = 1) /
 * geometry_np.py:6-189 image_geometric_transformation (coord_f32 = 0).  cx, cy: inverse-mapped
 * sample coordinate planes [h1*w1] (float32 or float64), shared by all planes. */
int hg_hexwarp_nearest(const void* src, void* dst, const void* cx, const void* cy, int coord_f32,
                       int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                       int elem_size, hg_stream_t stream);
int hg_hexwarp_linear(const void* src, void* dst, const void* cx, const void* cy, int coord_f32,
                      int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                      int src_dtype, int dst_dtype, hg_stream_t stream);
/* Same warp with the inverse affine map evaluated in-kernel (no coordinate planes in HBM):
 * sample (a,b) sits at X = row0 + a, Y = col0 + b + 0.5*(a odd); (x,y) = Hinv[0:2,:] * (X,Y,1)
 * in float64 (products summed left to right), then cast to float32 when coord_f32.
 * hinv: 6 doubles (HOST pointer, row-major first two rows of inv(H)). */
int hg_hexwarp_affine(const void* src, void* dst, const double* host_hinv, double row0, double col0,
                      int coord_f32, int interp, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                      int src_dtype, int dst_dtype, hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Doubled-raster layouts.  ref: HexImage.py:139-170 (encode), :106-111 (decode);
 * HexFrames.py:417-458 (torch twins).  hex [planes,H,W] <-> type1 [planes,H,2W+1] /
 * type2 [planes,2H,2W+1]; rows with (i + offset) odd carry the leading zero.
 * ---------------------------------------------------------------------------------------- */
int hg_hex_to_type1(const void* hex, void* t1, int64_t planes, int64_t H, int64_t W, int offset,
                    int src_dtype, int dst_dtype, hg_stream_t stream);
int hg_hex_to_type2(const void* hex, void* t2, int64_t planes, int64_t H, int64_t W, int offset,
                    int src_dtype, int dst_dtype, hg_stream_t stream);
/* rows_step = 1 decodes type1 ([..., 1::2]), 2 decodes type2 ([..., ::2, 1::2]); Wt = raster width */
int hg_type_to_hex(const void* t, void* hex, int64_t planes, int64_t Ht, int64_t Wt, int rows_step,
                   int src_dtype, int dst_dtype, hg_stream_t stream);

/* F.pad(x, (pl, pr, pt, pb), mode, value) on [planes,H,W] -> [planes,H+pt+pb,W+pl+pr].
 * ref: HexFrames.py:13-21 pad().  mode: 0 constant, 1 reflect, 2 replicate, 3 circular,
 * 4 symmetric (cv2.BORDER_REFLECT, geometry_np.py:720-725 heximpad).
 * dtype in {HG_U8, HG_F32, HG_F64, HG_BF16}; the backward ({HG_F32, HG_BF16}) writes gx fully. */
int hg_pad2d(const void* x, void* y, int64_t planes, int64_t H, int64_t W, int pl, int pr, int pt, int pb,
             int mode, double value, int dtype, hg_stream_t stream);
int hg_pad2d_bwd(const void* gy, void* gx, int64_t planes, int64_t H, int64_t W, int pl, int pr, int pt, int pb,
                 int mode, int dtype, hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Table-driven plane gather / scatter: rearrangements whose index rule depends on shapes only.
 * ref: the retired HexPixelShuffle, "HyGrid/codes in old versions.txt":68-126 -- 2*(3r^2-3r+1) strided
 * slice assignments into a doubled type1 canvas, then [..., 1::2] and a crop; and the viewer's fragment
 * shader, HexPixelArt/hexagon_mosaic_shader.py:25-81 (screen pixel -> hex cell).  Here one pass over a
 * host-built table of `cells` int64 source offsets (relative to the (batch, channel) base; < 0 = zero):
 *   gather : dst[b][c][e] = table[e] >= 0 ? src[b*batch_stride + c*chan_stride + table[e]] : 0
 *   scatter: gsrc[b*batch_stride + c*chan_stride + table[e]] = gdst[b][c][e]   (adjoint of the gather: the table
 *            must be injective and the caller zero-fills gsrc first)
 * dst / gdst are dense [batches, chans, cells].  Offsets that leave [0, batches*batch_stride) are skipped.
 * gather dtypes: {HG_F32, HG_BF16, HG_F64, HG_U8} -> HG_F32, HG_F64 -> HG_F64, HG_BF16 -> HG_BF16, HG_U8 -> HG_U8;
 * scatter: HG_F32, HG_F64, HG_BF16. */
int hg_plane_gather(const void* src, void* dst, const int64_t* table, int64_t batches, int64_t chans, int64_t cells,
                    int64_t batch_stride, int64_t chan_stride, int src_dtype, int dst_dtype, hg_stream_t stream);
int hg_plane_scatter(const void* gdst, void* gsrc, const int64_t* table, int64_t batches, int64_t chans, int64_t cells,
                     int64_t batch_stride, int64_t chan_stride, int dtype, hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hex pooling.  ref: HexFrames.py:255-341 HexPool2d, :344-401 HexAdaptivePool2d,
 * :402-414 HexGlobalPool2d, reductions :461-479 (NaN-aware).
 * The virtual input is x [planes,H,W] framed by `pad` cells of pad_value on every side and then
 * extended by tail_w columns / tail_h rows of tail_value (ceil_mode, value 0 or NaN).  Window of
 * out(I,J): rows sh*I + a, cols ((I%2)*shift)/2 + J*sw + b, a < kh, b < kw.
 * aux [planes,hn,wn] (may be NULL when no backward is needed; aux_bytes = 1 -> int8, 4 -> int32,
 * windows of more than 127 cells need int32): max/min -> winning window slot a*kw+b (-1 when the
 * winner is a masked NaN, which passes no gradient); avg -> number of non-NaN cells.
 * dtype in {HG_F32, HG_F64, HG_BF16}.  A window that leaves the virtual input returns HG_E_SHAPE
 * (the reference raises IndexError).
 * ---------------------------------------------------------------------------------------- */
int hg_hexpool_fwd(const void* x, void* y, void* aux, int aux_bytes, int64_t planes, int64_t H, int64_t W,
                   int64_t hn, int64_t wn, int kh, int kw, int sh, int sw, int shift,
                   int pad, double pad_value, int tail_h, int tail_w, double tail_value,
                   int method, int dtype, hg_stream_t stream);
/* gx [planes,H,W] is fully written (no pre-zeroing needed).  x (may be NULL) is only read by the
 * average method to keep NaN cells gradient-free. */
int hg_hexpool_bwd(const void* gy, const void* aux, int aux_bytes, const void* x, void* gx, int64_t planes,
                   int64_t H, int64_t W, int64_t hn, int64_t wn, int kh, int kw, int sh, int sw, int shift,
                   int pad, int method, int dtype, hg_stream_t stream);
/* x [planes, L] -> y [planes]; aux_idx [planes] int32 (argmax / non-NaN count), may be NULL */
int hg_hexglobalpool_fwd(const void* x, void* y, int32_t* aux_idx, int64_t planes, int64_t L,
                         int method, int dtype, hg_stream_t stream);
int hg_hexglobalpool_bwd(const void* gy, const void* x, const int32_t* aux_idx, void* gx, int64_t planes,
                         int64_t L, int method, int dtype, hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Hex convolution.  ref: HexFrames.py:22-185 HexConv2d.forward (closed form in DESIGN.md):
 *   y[n,co,R,q] = bias[co] + sum_{ci,tap} w[co,ci,0,tap] * D[s*R + a*d, 1 + (R%2)*s + 2*s*q + t*d + 2*d*m]
 * with D the doubled view of the (virtually) padded input, parity o = (even_odd_offset+pad)%2.
 * x [N,Cin,H,W], w [Cout,Cin/groups,1,K] (K = 3r^2-3r+1), y [N,Cout,Ho,Wo].  Padding is virtual in every
 * mode (pad_mode): the frame is pad_value, or the reflected / replicated / wrapped image, read in place.
 * io_dtype: element type of x / y / gradients in {HG_F32, HG_BF16}; weights and bias are
 * float32, accumulation is float32.  algo: 0 = auto, 1 = direct stencil, 2 = tcgen05 implicit GEMM.
 * ---------------------------------------------------------------------------------------- */
typedef struct hg_conv_desc {
  int64_t N, Cin, Cout, H, W, Ho, Wo;
  int radius, stride, dilation, groups, pad, parity;
  float pad_value;
  int x_dtype, y_dtype; /* HG_F32 / HG_BF16 */
  int algo;
  int relu;             /* fused epilogue of HexConvModule (conv -> act), 0/1 */
  int pad_mode;         /* how the `pad` frame is filled: 0 constant pad_value, 1 reflect, 2 replicate, 3 circular (F.pad modes,
                           ref HexFrames.py:13-21, :121; HexModules.py:185-190).  Modes 1..3 are resolved by the forward and
                           weight-gradient loaders by coordinate remapping -- no padded copy of x exists; hg_hexconv_dgrad
                           takes them on the padded geometry (pad = 0 on an [H+2p, W+2p] gradient) followed by hg_pad2d_bwd. */
  int accumulate;       /* forward / data gradient on the tcgen05 path only: add to what y / gx already hold (bias must be NULL).
                           Used by the float32 route that runs three bfloat16 tensor-core passes over split operands. */
} hg_conv_desc;

/* dst[0..n) = *scalar, both in device memory, dtype HG_F32 or HG_BF16 (dst 16-byte aligned): materialises the broadcast
 * gradient a `sum()` loss hands to the conv backward (torch's strided broadcast copy runs at a third of the HBM rate). */
int hg_broadcast_fill(void* dst, const void* scalar, int64_t n, int dtype, hg_stream_t stream);
/* x = hi + lo with hi = bf16(x), lo = bf16(x - hi): the operand split of the float32 tensor-core route
 * (y ~ x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, relative error ~2^-16).  hi / lo: bfloat16 [n]; either may be NULL. */
int hg_split_bf16(const float* x, void* hi, void* lo, int64_t n, hg_stream_t stream);

int hg_hexconv_out_shape(int64_t H, int64_t W, int radius, int stride, int dilation, int pad,
                         int64_t* Ho, int64_t* Wo);
/* 1 when the tcgen05 / TMEM implicit-GEMM kernel covers this configuration for op (0 forward, 1 data
 * gradient, 2 weight gradient): radius 2, stride 1, dilation 1, groups 1, channels in multiples of 16
 * (reduction channels beyond 64 run as passes over 64-channel slices; forward / data-gradient output channels
 * up to 256 as far as the weight image fits shared memory).  algo = 0 picks it by itself only for
 * bfloat16 activations (it rounds activations and weights to bfloat16, fp32 accumulation). */
int hg_hexconv_umma_eligible(const hg_conv_desc* d, int op);
int hg_hexconv_fwd(const hg_conv_desc* d, const void* x, const float* w, const float* bias, void* y,
                   hg_stream_t stream);
/* ref: HexModules.py:275-288 HexConvModule.forward in eval mode, order ('conv', 'norm', 'act'):
 *   y = act(conv(x, w) * scale[co] + shift[co])    with scale = gamma / sqrt(running_var + eps) and
 *   shift = beta - running_mean * scale (+ conv bias * scale); act = ReLU when d->relu, identity otherwise.
 * scale (may be NULL = 1) is folded into the weights on their way into shared memory, shift (may be NULL)
 * takes the bias slot: one pass over x and y instead of three.  Inference only (no backward counterpart). */
int hg_hexconv_fwd_affine(const hg_conv_desc* d, const void* x, const float* w, const float* scale,
                          const float* shift, void* y, hg_stream_t stream);
/* gx [N,Cin,H,W] fully written.  gy has y_dtype, gx has x_dtype. */
int hg_hexconv_dgrad(const hg_conv_desc* d, const void* gy, const float* w, void* gx, hg_stream_t stream);
/* gw [Cout,Cin/groups,1,K] float32 and gbias [Cout] float32 (may be NULL) are ACCUMULATED into
 * (caller zeroes them), so that per-layer partials can land directly in a flat all-reduce bucket. */
int hg_hexconv_wgrad(const hg_conv_desc* d, const void* x, const void* gy, float* gw, float* gbias,
                     hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise weighted tap gathers on doubled coordinates: the learned lattice resamplers the reference retired into
 * "HyGrid/codes in old versions.txt" -- Hex_to_Square_Conv2d_by_Double_Stride (:1-66), Square_to_Hex_Conv2d_by_Double_Stride
 * (:421-493), Hex_to_Square_original_resolution (:587-636).  One operator covers all three:
 *   y[n,c,R,J] = sum_t w_s[c,t] * A_s(sy*R + ry_s[t], sx_s*J + ex_s[t]),   s = (R odd and R < odd_limit) ? 1 : 0
 * A_s: set[s].doubled != 0 -> the doubled ("type1") view of the virtually padded hex lattice (column c of row i is cell
 * (c - s_i) >> 1 for s_i <= c < 2*Wp + s_i, s_i = (i + parity) & 1, literal 0 elsewhere); == 0 -> the plain padded image.
 * x [N,C,H,W] float32, y [N,C,Ho,Wo] float32, w_even / w_odd [C, set[s].T] float32 (NULL = all ones).
 *   hg_dwtaps_dgrad  gx += adjoint (caller zeroes gx);  hg_dwtaps_wgrad  gw[C, set[set].T] += sum gy * A  (caller zeroes).
 * ---------------------------------------------------------------------------------------- */
typedef struct hg_taps_set {
  int T, sx, doubled;          /* taps (1..64), column step per output, coordinate kind */
  int ry[64], ex[64];          /* row / column offset of every tap (may be negative) */
} hg_taps_set;
typedef struct hg_dwtaps_desc {
  int64_t N, C, H, W, Ho, Wo;
  int sy, odd_limit;           /* row step per output; odd output rows >= odd_limit use the even tap set */
  int parity, pad;             /* (even_odd_offset + pad) & 1; frame of pad_value cells around x */
  float pad_value;
  hg_taps_set set[2];          /* [0] even output rows, [1] odd output rows */
} hg_dwtaps_desc;
int hg_dwtaps_fwd(const hg_dwtaps_desc* d, const float* x, const float* w_even, const float* w_odd, float* y, hg_stream_t stream);
int hg_dwtaps_dgrad(const hg_dwtaps_desc* d, const float* gy, const float* w_even, const float* w_odd, float* gx,
                    hg_stream_t stream);
int hg_dwtaps_wgrad(const hg_dwtaps_desc* d, const float* x, const float* gy, float* gw, int set, hg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer entry points (what a numpy caller binds): pinned-staged, chunked over planes,
 * H2D / kernel / D2H overlapped on internal streams.  host_src / host_dst are HOST pointers
 * (pageable or pinned).  device: CUDA device ordinal.  Synchronous (returns when host_dst is
 * complete).  ref: the numpy-in / numpy-out contract of geometry_np.py:358 and :191.
 * ---------------------------------------------------------------------------------------- */
int hg_host_rect2hex(const void* host_src, void* host_dst, const double* host_xs, const double* host_ys,
                     int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                     int src_dtype, int dst_dtype, int interp, int math, int device);
int hg_host_hex2rect(const void* host_src, void* host_dst, const double* host_xs, const double* host_ys,
                     int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1,
                     int src_dtype, int dst_dtype, int interp, int math, int device);
/* Host-buffer writers / readers of the doubled rasters (the formats either side of the path): a host hex matrix
 * [planes,H,W] -> type1 [planes,H,2W+1] (rows_mul = 1) or type2 [planes,2H,2W+1] (rows_mul = 2) raster in host memory and
 * back, through the same pinned ring.  ref: HexImage.py:139-170 GenerateType1Image / GenerateType2Image (python loops over
 * bands x rows with np.insert / np.append), :106-111 (decode), :171-218 SaveHexImage (which writes those rasters).
 * Calls on different devices run concurrently (one ring and one lock per device). */
int hg_host_hex_to_type(const void* host_hex, void* host_raster, int64_t planes, int64_t H, int64_t W, int offset,
                        int src_dtype, int dst_dtype, int rows_mul, int device);
int hg_host_type_to_hex(const void* host_raster, void* host_hex, int64_t planes, int64_t H, int64_t W,
                        int src_dtype, int dst_dtype, int rows_mul, int device);
/* frees the streams, device ring and pinned bounce buffers the host entry points created lazily
 * (the only memory the library ever owns). */
void hg_host_release(void);

/* ----------------------------------------------------------------------------------------------------
 * Batch normalisation (+ fused ReLU) of HexConvModule: conv -> norm -> act (ref: HexModules.py:146-288; the norm
 * layer is torch.nn.BatchNorm2d built through mmcv's build_norm_layer, HexModules.py:57-76).  float32 NCHW,
 * x / y / dy / dx are [N, C, HW] contiguous; gamma / beta may be NULL (affine=False); HBM-bound streaming kernels.
 *   hg_bn_stats      sums[2c] += sum x, sums[2c+1] += sum x^2   (float64 [2*C], caller zeroes)
 *   hg_bn_apply      y = act(x * gamma*rstd + (beta - mean*gamma*rstd)); statistics from `sums` (training: mean,
 *                    biased variance and rstd are written to mean_out / var_out / rstd_out, each may be NULL) or from
 *                    mean_in / var_in (inference, sums == NULL); relu != 0 fuses the ReLU
 *   hg_bn_bwd_reduce dsums[2c] += sum dz, dsums[2c+1] += sum dz*xhat with dz = dy (masked by the fused ReLU's z > 0,
 *                    z recomputed from x) -- float64 [2*C], caller zeroes
 *   hg_bn_bwd_apply  dx = gamma*rstd*(dz - sum(dz)/n - xhat*sum(dz*xhat)/n) (training) or gamma*rstd*dz (inference);
 *                    dgamma / dbeta (may be NULL) = sum dz*xhat / sum dz
 * -------------------------------------------------------------------------------------------------- */
int hg_bn_stats(const float* x, double* sums, int64_t N, int64_t C, int64_t HW, hg_stream_t stream);
int hg_bn_apply(const float* x, float* y, const double* sums, const float* mean_in, const float* var_in,
                const float* gamma, const float* beta, float* mean_out, float* var_out, float* rstd_out,
                int64_t N, int64_t C, int64_t HW, float eps, int relu, hg_stream_t stream);
int hg_bn_bwd_reduce(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, double* dsums, int64_t N, int64_t C, int64_t HW, int relu,
                     hg_stream_t stream);
int hg_bn_bwd_apply(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, const double* dsums, float* dx, float* dgamma, float* dbeta, int64_t N,
                    int64_t C, int64_t HW, int relu, int training, hg_stream_t stream);

/* The same two passes per direction as ONE call each, with torch.nn.BatchNorm2d's training-mode bookkeeping done by the
 * kernels (the C5 training step is launch-bound: zeroing the sums, bumping num_batches_tracked and the mul_ / add_ pairs of
 * the running statistics were six extra launches per layer and step).
 *   hg_bn_train_fwd  zeroes `sums` (float64 [2*C] scratch), accumulates the batch statistics, writes y = act(bn(x)) and
 *                    mean / rstd for backward; when running_mean / running_var (float32 [C]) are given, increments
 *                    *num_batches_tracked (int64, may be NULL when momentum >= 0) and blends
 *                    running = (1 - f) * running + f * (mean | var * n / (n - 1)),  f = momentum, or
 *                    1 / num_batches_tracked when momentum < 0 (torch's momentum=None cumulative average).
 *                    ref: HexModules.py:57-76 (mmcv build_norm_layer -> torch.nn.BatchNorm2d), :275-288 forward order.
 *   hg_bn_bwd        zeroes `dsums`, reduces, writes dx; dgamma / dbeta (may be NULL) are overwritten, or added to when
 *                    accumulate_affine != 0 (slices of the all-reduce bucket, HyGrid/distributed.py). */
int hg_bn_train_fwd(const float* x, float* y, double* sums, const float* gamma, const float* beta, float* mean_out,
                    float* rstd_out, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    double momentum, int64_t N, int64_t C, int64_t HW, float eps, int relu, hg_stream_t stream);
int hg_bn_bwd(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma,
              const float* beta, double* dsums, float* dx, float* dgamma, float* dbeta, int accumulate_affine,
              int64_t N, int64_t C, int64_t HW, int relu, int training, hg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HYGRID_B200_H_ */
