// hg_layout.cu -- doubled-raster ("type1" / "type2") encode / decode and 2-D padding (sm_100a).
//
// Pure index shuffles, HBM-bound streaming copies.  A CTA owns a contiguous run of kChunk output
// elements; every thread resolves (row, column) of its first element with one division and then
// walks by blockDim with carries, so consecutive lanes always touch consecutive addresses.
//   ref: HexImage.py:139-170 (GenerateType1Image / GenerateType2Image), :106-111 (decode),
//        HexFrames.py:417-458 (heximage_to_type1/2, type1_to_heximage), HexFrames.py:13-21 (pad).
#include "hg_common.cuh"

namespace hg {

constexpr int kLayoutThreads = 256;
constexpr int kLayoutChunk = 256 * 16;

template <typename TS, typename TD> __device__ __forceinline__ TD convert(TS v) { return (TD)v; }
template <> __device__ __forceinline__ float convert<__nv_bfloat16, float>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ double convert<__nv_bfloat16, double>(__nv_bfloat16 v) { return (double)__bfloat162float(v); }
template <> __device__ __forceinline__ __nv_bfloat16 convert<__nv_bfloat16, __nv_bfloat16>(__nv_bfloat16 v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 convert<float, __nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ T zero_of() { return T(0); }
template <> __device__ __forceinline__ __nv_bfloat16 zero_of<__nv_bfloat16>() { return __float2bfloat16_rn(0.f); }

// out[r, c] (r over planes*Hout rows, c < Wt = 2W+1): source row i = (r % Hout) / rows_mul;
// s = (i + offset) & 1; cell (c - s) / 2 when s <= c < 2W + s, else zero.
template <typename TS, typename TD>
__global__ void __launch_bounds__(kLayoutThreads)
hex_to_type_kernel(const TS* __restrict__ hex, TD* __restrict__ out, int64_t total, int H, int W, int rows_mul, int offset) {
  const int Wt = 2 * W + 1, Hout = H * rows_mul;
  const int64_t base = (int64_t)blockIdx.x * kLayoutChunk;
  const int64_t end = min(base + (int64_t)kLayoutChunk, total);
  int64_t t = base + threadIdx.x;
  if (t >= end) return;
  int64_t r = t / Wt;
  int c = (int)(t - r * Wt);
  const int step_r = kLayoutThreads / Wt, step_c = kLayoutThreads % Wt;
  // rows are tracked as (plane, row-in-plane) with carries: no division inside the loop
  int64_t plane = r / Hout;
  int ro = (int)(r - plane * Hout);
  for (; t < end; t += kLayoutThreads) {
    const int i = rows_mul == 2 ? (ro >> 1) : ro;
    const int s = (i + offset) & 1;
    TD v = zero_of<TD>();
    if (c >= s && c < 2 * W + s) v = convert<TS, TD>(__ldg(hex + (plane * H + i) * (int64_t)W + ((c - s) >> 1)));
    out[t] = v;
    ro += step_r; c += step_c;
    if (c >= Wt) { c -= Wt; ++ro; }
    while (ro >= Hout) { ro -= Hout; ++plane; }
  }
}

// Vector variant: the output is treated as one flat array and every thread produces V = 16 / sizeof(TD)
// consecutive elements as ONE aligned 16-byte streaming store (rows of 2W+1 elements are never 16-byte aligned
// themselves, so vectors straddle row ends: each element resolves its own (row, column) by a carry).  Source
// cells are read with scalar loads (neighbouring lanes share lines through L1; each cell is needed twice).
template <typename TS, typename TD>
__global__ void __launch_bounds__(kLayoutThreads)
hex_to_type_vec_kernel(const TS* __restrict__ hex, TD* __restrict__ out, int64_t total, int H, int W, int rows_mul, int offset) {
  constexpr int V = 16 / (int)sizeof(TD);
  struct alignas(16) Vec { TD v[V]; };
  const int Wt = 2 * W + 1, Hout = H * rows_mul;
  const int64_t base = (int64_t)blockIdx.x * kLayoutChunk;
  const int64_t end = min(base + (int64_t)kLayoutChunk, total);      // kLayoutChunk % V == 0, total % V == 0
  int64_t t = base + (int64_t)threadIdx.x * V;
  if (t >= end) return;
  int64_t r = t / Wt;
  int c = (int)(t - r * Wt);
  constexpr int kStep = kLayoutThreads * V;
  const int step_r = kStep / Wt, step_c = kStep % Wt;
  int64_t plane = r / Hout;
  int ro = (int)(r - plane * Hout);
  for (; t < end; t += kStep) {
    Vec o;
    if (V == 4 && c >= 1 && c + V < 2 * W) {
      // interior vector (all four elements inside one row and inside the doubled cells): element e reads cell
      // (c + e - sft) >> 1 -- three consecutive cells cover it, no per-element bounds or carries
      const int i = rows_mul == 2 ? (ro >> 1) : ro;
      const int sft = (i + offset) & 1;
      const int q = c - sft;
      const TS* __restrict__ hp = hex + (plane * H + i) * (int64_t)W + (q >> 1);
      const TD v0 = convert<TS, TD>(__ldg(hp)), v1 = convert<TS, TD>(__ldg(hp + 1));
      if (q & 1) {
        const TD v2 = convert<TS, TD>(__ldg(hp + 2));
        o.v[0] = v0; o.v[1] = v1; o.v[V > 2 ? 2 : 0] = v1; o.v[V - 1] = v2;
      } else {
        o.v[0] = v0; o.v[1] = v0; o.v[V > 2 ? 2 : 0] = v1; o.v[V - 1] = v1;
      }
    } else {
    int cc = c, rr = ro;
    int64_t pp = plane;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int i = rows_mul == 2 ? (rr >> 1) : rr;
      const int sft = (i + offset) & 1;
      TD v = zero_of<TD>();
      if (cc >= sft && cc < 2 * W + sft) v = convert<TS, TD>(__ldg(hex + (pp * H + i) * (int64_t)W + ((cc - sft) >> 1)));
      o.v[e] = v;
      if (++cc == Wt) { cc = 0; if (++rr == Hout) { rr = 0; ++pp; } }
    }
    }
    __stcs(reinterpret_cast<uint4*>(out + t), *reinterpret_cast<const uint4*>(&o));
    ro += step_r; c += step_c;
    if (c >= Wt) { c -= Wt; ++ro; }
    while (ro >= Hout) { ro -= Hout; ++plane; }
  }
}

// the last (total % V) elements the vector kernel cannot cover
template <typename TS, typename TD>
__global__ void hex_to_type_tail_kernel(const TS* __restrict__ hex, TD* __restrict__ out, int64_t begin, int64_t total, int H, int W,
                                        int rows_mul, int offset) {
  const int64_t t = begin + threadIdx.x;
  if (t >= total) return;
  const int Wt = 2 * W + 1, Hout = H * rows_mul;
  const int64_t r = t / Wt;
  const int c = (int)(t - r * Wt);
  const int64_t plane = r / Hout;
  const int ro = (int)(r - plane * Hout);
  const int i = rows_mul == 2 ? (ro >> 1) : ro;
  const int sft = (i + offset) & 1;
  TD v = zero_of<TD>();
  if (c >= sft && c < 2 * W + sft) v = convert<TS, TD>(__ldg(hex + (plane * H + i) * (int64_t)W + ((c - sft) >> 1)));
  out[t] = v;
}

// hex[r, j] = t[plane, i * rows_step, 1 + 2 j]
template <typename TS, typename TD>
__global__ void __launch_bounds__(kLayoutThreads)
type_to_hex_kernel(const TS* __restrict__ tin, TD* __restrict__ hex, int64_t total, int Ht, int Wt, int H, int W, int rows_step) {
  const int64_t base = (int64_t)blockIdx.x * kLayoutChunk;
  const int64_t end = min(base + (int64_t)kLayoutChunk, total);
  int64_t t = base + threadIdx.x;
  if (t >= end) return;
  int64_t r = t / W;
  int c = (int)(t - r * W);
  const int step_r = kLayoutThreads / W, step_c = kLayoutThreads % W;
  int64_t plane = r / H;
  int i = (int)(r - plane * H);
  for (; t < end; t += kLayoutThreads) {
    hex[t] = convert<TS, TD>(__ldg(tin + (plane * Ht + (int64_t)i * rows_step) * Wt + 1 + 2 * c));
    i += step_r; c += step_c;
    if (c >= W) { c -= W; ++i; }
    while (i >= H) { i -= H; ++plane; }
  }
}

// F.pad(x, (pl, pr, pt, pb), mode, value) on [planes, H, W] -> [planes, H+pt+pb, W+pl+pr]
// mode: 0 constant, 1 reflect, 2 replicate, 3 circular, 4 symmetric (reflect repeating the edge)
__device__ __forceinline__ int pad_index(int i, int n, int mode, bool& inside) {
  inside = (i >= 0 && i < n);
  if (inside || mode == 0) return i;
  inside = true;
  if (mode == 1) {  // reflect without repeating the edge
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    int m = i % period; if (m < 0) m += period;
    return m < n ? m : period - m;
  }
  if (mode == 2) return i < 0 ? 0 : n - 1;
  if (mode == 4) {
    const int period = 2 * n;
    int m = i % period; if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
  }
  int m = i % n; if (m < 0) m += n;
  return m;
}

template <typename T>
__global__ void __launch_bounds__(kLayoutThreads)
pad2d_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total, int H, int W, int pl, int pr, int pt, int pb,
             int mode, T value) {
  const int Wo = W + pl + pr, Ho = H + pt + pb;
  const int64_t base = (int64_t)blockIdx.x * kLayoutChunk;
  const int64_t end = min(base + (int64_t)kLayoutChunk, total);
  int64_t t = base + threadIdx.x;
  if (t >= end) return;
  int64_t r = t / Wo;
  int c = (int)(t - r * Wo);
  const int step_r = kLayoutThreads / Wo, step_c = kLayoutThreads % Wo;
  int64_t plane = r / Ho;
  int i = (int)(r - plane * Ho);
  for (; t < end; t += kLayoutThreads) {
    bool in_r, in_c;
    const int si = pad_index(i - pt, H, mode, in_r), sj = pad_index(c - pl, W, mode, in_c);
    y[t] = (in_r && in_c) ? __ldg(x + (plane * H + si) * (int64_t)W + sj) : value;
    i += step_r; c += step_c;
    if (c >= Wo) { c -= Wo; ++i; }
    while (i >= Ho) { i -= Ho; ++plane; }
  }
}

// gradient of pad2d: gx[i, j] = sum of gy over all padded cells that read x[i, j]
template <typename T>
__global__ void __launch_bounds__(kLayoutThreads)
pad2d_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, int64_t total, int H, int W, int pl, int pr, int pt, int pb,
                 int mode) {
  const int Wo = W + pl + pr, Ho = H + pt + pb;
  const int64_t t = (int64_t)blockIdx.x * kLayoutThreads + threadIdx.x;
  if (t >= total) return;
  const int64_t r = t / W;
  const int j = (int)(t - r * W);
  const int64_t plane = r / H;
  const int i = (int)(r - plane * H);
  const T* g = gy + plane * (int64_t)Ho * Wo;
  if (mode == 0) { gx[t] = g[(int64_t)(i + pt) * Wo + j + pl]; return; }
  // candidates: the cell's own position plus the border rows / columns that may alias onto it
  float acc = 0.f;
  for (int ka = 0; ka < pt + pb + 1; ++ka) {
    const int a = ka < pt ? ka : (ka == pt ? i + pt : H + ka - 1);
    bool in_r; if (pad_index(a - pt, H, mode, in_r) != i) continue;
    for (int kb = 0; kb < pl + pr + 1; ++kb) {
      const int b = kb < pl ? kb : (kb == pl ? j + pl : W + kb - 1);
      bool in_c; if (pad_index(b - pl, W, mode, in_c) != j) continue;
      acc += to_f32(g[(int64_t)a * Wo + b]);
    }
  }
  gx[t] = from_f32<T>(acc);
}

static inline unsigned chunks(int64_t total) { return (unsigned)ceil_div(total, kLayoutChunk); }

template <typename TS, typename TD>
static int launch_to_type(const void* hex, void* out, int64_t planes, int64_t H, int64_t W, int rows_mul, int offset, cudaStream_t st) {
  const int64_t total = planes * H * rows_mul * (2 * W + 1);
  constexpr int V = 16 / (int)sizeof(TD);
  static_assert(kLayoutChunk % V == 0, "chunk must hold whole vectors");
  const int64_t vec_total = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && W >= 2 ? total / V * V : 0;
  if (vec_total > 0) {
    hex_to_type_vec_kernel<TS, TD><<<chunks(vec_total), kLayoutThreads, 0, st>>>((const TS*)hex, (TD*)out, vec_total, (int)H, (int)W, rows_mul, offset);
    int rc = finish_launch("hex_to_type_vec");
    if (rc || vec_total == total) return rc;
    // the last total % V elements (same plane, last row): one tiny scalar launch over the whole tail row range
  }
  if (vec_total == 0)
    hex_to_type_kernel<TS, TD><<<chunks(total), kLayoutThreads, 0, st>>>((const TS*)hex, (TD*)out, total, (int)H, (int)W, rows_mul, offset);
  else
    hex_to_type_tail_kernel<TS, TD><<<1, 32, 0, st>>>((const TS*)hex, (TD*)out, vec_total, total, (int)H, (int)W, rows_mul, offset);
  return finish_launch("hex_to_type");
}
template <typename TS, typename TD>
static int launch_from_type(const void* tin, void* hex, int64_t planes, int64_t Ht, int64_t Wt, int64_t H, int64_t W, int rows_step, cudaStream_t st) {
  const int64_t total = planes * H * W;
  type_to_hex_kernel<TS, TD><<<chunks(total), kLayoutThreads, 0, st>>>((const TS*)tin, (TD*)hex, total, (int)Ht, (int)Wt, (int)H, (int)W, rows_step);
  return finish_launch("type_to_hex");
}

// dispatch (src, dst): identical types move raw bits by element size; otherwise any supported
// source widens to float32 / float64 (the reference encoders return float64 numpy / float32 torch).
#define HG_LAYOUT_DISPATCH(FN, ...)                                                        \
  if (sdt == ddt) {                                                                        \
    switch (dtype_size(sdt)) {                                                             \
      case 1: return FN<uint8_t, uint8_t>(__VA_ARGS__);                                    \
      case 2: return FN<uint16_t, uint16_t>(__VA_ARGS__);                                  \
      case 4: return FN<uint32_t, uint32_t>(__VA_ARGS__);                                  \
      case 8: return FN<uint64_t, uint64_t>(__VA_ARGS__);                                  \
    }                                                                                      \
  }                                                                                        \
  if (ddt == HG_F32) {                                                                     \
    switch (sdt) {                                                                         \
      case HG_U8: return FN<uint8_t, float>(__VA_ARGS__);                                  \
      case HG_I16: return FN<int16_t, float>(__VA_ARGS__);                                 \
      case HG_U16: return FN<uint16_t, float>(__VA_ARGS__);                                \
      case HG_I32: return FN<int32_t, float>(__VA_ARGS__);                                 \
      case HG_I64: return FN<int64_t, float>(__VA_ARGS__);                                 \
      case HG_F64: return FN<double, float>(__VA_ARGS__);                                  \
      case HG_BF16: return FN<__nv_bfloat16, float>(__VA_ARGS__);                          \
    }                                                                                      \
  }                                                                                        \
  if (ddt == HG_F64) {                                                                     \
    switch (sdt) {                                                                         \
      case HG_U8: return FN<uint8_t, double>(__VA_ARGS__);                                 \
      case HG_I16: return FN<int16_t, double>(__VA_ARGS__);                                \
      case HG_U16: return FN<uint16_t, double>(__VA_ARGS__);                               \
      case HG_I32: return FN<int32_t, double>(__VA_ARGS__);                                \
      case HG_I64: return FN<int64_t, double>(__VA_ARGS__);                                \
      case HG_F32: return FN<float, double>(__VA_ARGS__);                                  \
      case HG_BF16: return FN<__nv_bfloat16, double>(__VA_ARGS__);                         \
    }                                                                                      \
  }                                                                                        \
  if (ddt == HG_BF16 && sdt == HG_F32) return FN<float, __nv_bfloat16>(__VA_ARGS__);

static int to_type(const void* hex, void* out, int64_t planes, int64_t H, int64_t W, int rows_mul, int offset, int sdt,
                   int ddt, hg_stream_t stream) {
  HG_REQUIRE(planes >= 0 && H >= 0 && W >= 0, HG_E_SHAPE, "bad shape planes=%lld H=%lld W=%lld", (long long)planes, (long long)H, (long long)W);
  HG_REQUIRE(2 * W + 1 < (1ll << 30) && H * rows_mul < (1ll << 30), HG_E_SHAPE, "raster too large");
  HG_REQUIRE(dtype_size(sdt) && dtype_size(ddt), HG_E_DTYPE, "unknown dtype");
  if (planes == 0 || H == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  offset = ((offset % 2) + 2) % 2;
  HG_LAYOUT_DISPATCH(launch_to_type, hex, out, planes, H, W, rows_mul, offset, st)
  set_error("hex_to_type: unsupported dtypes src=%d dst=%d", sdt, ddt);
  return HG_E_DTYPE;
}

// x = hi + lo, hi = bf16(x), lo = bf16(x - hi): four elements per thread, 16-byte loads, 8-byte stores
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < n; e += stride) {
    float v[4];
    if (e + 3 < n && (reinterpret_cast<uintptr_t>(x + e) & 15) == 0) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(x + e));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = e + k < n ? __ldg(x + e + k) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (e + k < n) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v[k]);
        if (hi) hi[e + k] = h;
        if (lo) lo[e + k] = __float2bfloat16_rn(v[k] - __bfloat162float(h));
      }
    }
  }
}

// dst[0..n) = *scalar (both in device memory): 16-byte streaming stores
template <typename T>
__global__ void __launch_bounds__(256) broadcast_fill_kernel(T* __restrict__ dst, const T* __restrict__ scalar, int64_t n) {
  constexpr int V = 16 / (int)sizeof(T);
  const T v = *scalar;
  union { uint4 q; T e[V]; } pack;
#pragma unroll
  for (int k = 0; k < V; ++k) pack.e[k] = v;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
  for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; e < n; e += stride) {
    if (e + V <= n) __stcs(reinterpret_cast<uint4*>(dst + e), pack.q);
    else for (int64_t k = e; k < n; ++k) dst[k] = v;
  }
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_broadcast_fill(void* dst, const void* scalar, int64_t n, int dtype, hg_stream_t stream) {
  HG_REQUIRE(n >= 0, HG_E_SHAPE, "bad length");
  if (n == 0) return HG_OK;
  HG_REQUIRE(dst && scalar && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, HG_E_ARG, "NULL or unaligned buffer");
  int64_t blocks = ceil_div(n, 256 * 8);
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = as_stream(stream);
  if (dtype == HG_F32) broadcast_fill_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)dst, (const float*)scalar, n);
  else if (dtype == HG_BF16) broadcast_fill_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)scalar, n);
  else { set_error("broadcast_fill: float32 or bfloat16"); return HG_E_DTYPE; }
  return finish_launch("broadcast_fill");
}

int hg_split_bf16(const float* x, void* hi, void* lo, int64_t n, hg_stream_t stream) {
  HG_REQUIRE(n >= 0, HG_E_SHAPE, "bad length");
  if (n == 0 || (!hi && !lo)) return HG_OK;
  HG_REQUIRE(x != nullptr, HG_E_ARG, "NULL buffer");
  int64_t blocks = ceil_div(n, 1024);
  if (blocks > 148 * 64) blocks = 148 * 64;
  split_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  return finish_launch("split_bf16");
}

int hg_hex_to_type1(const void* hex, void* t1, int64_t planes, int64_t H, int64_t W, int offset, int src_dtype,
                    int dst_dtype, hg_stream_t stream) {
  return to_type(hex, t1, planes, H, W, 1, offset, src_dtype, dst_dtype, stream);
}
int hg_hex_to_type2(const void* hex, void* t2, int64_t planes, int64_t H, int64_t W, int offset, int src_dtype,
                    int dst_dtype, hg_stream_t stream) {
  return to_type(hex, t2, planes, H, W, 2, offset, src_dtype, dst_dtype, stream);
}

int hg_type_to_hex(const void* t, void* hex, int64_t planes, int64_t Ht, int64_t Wt, int rows_step, int sdt, int ddt,
                   hg_stream_t stream) {
  HG_REQUIRE(planes >= 0 && Ht >= 0 && Wt >= 1, HG_E_SHAPE, "bad shape planes=%lld Ht=%lld Wt=%lld", (long long)planes, (long long)Ht, (long long)Wt);
  HG_REQUIRE(rows_step == 1 || rows_step == 2, HG_E_ARG, "rows_step must be 1 (type1) or 2 (type2)");
  HG_REQUIRE(Wt < (1ll << 30) && Ht < (1ll << 30), HG_E_SHAPE, "raster too large");
  HG_REQUIRE(dtype_size(sdt) && dtype_size(ddt), HG_E_DTYPE, "unknown dtype");
  const int64_t H = (Ht + rows_step - 1) / rows_step, W = (Wt - 1) / 2;
  if (planes == 0 || H == 0 || W == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  HG_LAYOUT_DISPATCH(launch_from_type, t, hex, planes, Ht, Wt, H, W, rows_step, st)
  set_error("type_to_hex: unsupported dtypes src=%d dst=%d", sdt, ddt);
  return HG_E_DTYPE;
}

int hg_pad2d(const void* x, void* y, int64_t planes, int64_t H, int64_t W, int pl, int pr, int pt, int pb, int mode,
             double value, int dtype, hg_stream_t stream) {
  HG_REQUIRE(planes >= 0 && H > 0 && W > 0, HG_E_SHAPE, "bad shape planes=%lld H=%lld W=%lld", (long long)planes, (long long)H, (long long)W);
  HG_REQUIRE(mode >= 0 && mode <= 4, HG_E_ARG, "pad mode must be 0..4");
  HG_REQUIRE(H + pt + pb > 0 && W + pl + pr > 0 && pl >= 0 && pr >= 0 && pt >= 0 && pb >= 0, HG_E_SHAPE, "negative padding");
  HG_REQUIRE(mode != 1 || (pl < W && pr < W && pt < H && pb < H), HG_E_SHAPE, "reflect padding must be smaller than the image");
  HG_REQUIRE(mode != 3 || (pl <= W && pr <= W && pt <= H && pb <= H), HG_E_SHAPE, "circular padding must not exceed the image");
  const int64_t total = planes * (H + pt + pb) * (W + pl + pr);
  if (total == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  switch (dtype) {
    case HG_F32:
      pad2d_kernel<float><<<chunks(total), kLayoutThreads, 0, st>>>((const float*)x, (float*)y, total, (int)H, (int)W, pl, pr, pt, pb, mode, (float)value);
      break;
    case HG_F64:
      pad2d_kernel<double><<<chunks(total), kLayoutThreads, 0, st>>>((const double*)x, (double*)y, total, (int)H, (int)W, pl, pr, pt, pb, mode, value);
      break;
    case HG_BF16:
      pad2d_kernel<__nv_bfloat16><<<chunks(total), kLayoutThreads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, total, (int)H, (int)W, pl, pr, pt, pb, mode, __float2bfloat16_rn((float)value));
      break;
    case HG_U8:
      pad2d_kernel<uint8_t><<<chunks(total), kLayoutThreads, 0, st>>>((const uint8_t*)x, (uint8_t*)y, total, (int)H, (int)W, pl, pr, pt, pb, mode, (uint8_t)value);
      break;
    default:
      set_error("pad2d: unsupported dtype %d", dtype);
      return HG_E_DTYPE;
  }
  return finish_launch("pad2d");
}

int hg_pad2d_bwd(const void* gy, void* gx, int64_t planes, int64_t H, int64_t W, int pl, int pr, int pt, int pb, int mode,
                 int dtype, hg_stream_t stream) {
  HG_REQUIRE(planes >= 0 && H > 0 && W > 0, HG_E_SHAPE, "bad shape");
  HG_REQUIRE(mode >= 0 && mode <= 4, HG_E_ARG, "pad mode must be 0..4");
  const int64_t total = planes * H * W;
  if (total == 0) return HG_OK;
  cudaStream_t st = as_stream(stream);
  const unsigned g = (unsigned)ceil_div(total, kLayoutThreads);
  switch (dtype) {
    case HG_F32: pad2d_bwd_kernel<float><<<g, kLayoutThreads, 0, st>>>((const float*)gy, (float*)gx, total, (int)H, (int)W, pl, pr, pt, pb, mode); break;
    case HG_BF16: pad2d_bwd_kernel<__nv_bfloat16><<<g, kLayoutThreads, 0, st>>>((const __nv_bfloat16*)gy, (__nv_bfloat16*)gx, total, (int)H, (int)W, pl, pr, pt, pb, mode); break;
    default: set_error("pad2d_bwd: unsupported dtype %d", dtype); return HG_E_DTYPE;
  }
  return finish_launch("pad2d_bwd");
}

}  // extern "C"
