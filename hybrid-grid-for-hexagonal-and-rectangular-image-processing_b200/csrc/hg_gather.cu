// hg_gather.cu -- table-driven plane gather / scatter.
//
// Some lattice rearrangements are pure index shuffles whose index rule depends on the shapes only (never on the
// data): the retired HexPixelShuffle of the reference builds its result with 2*(3r^2-3r+1) strided slice
// assignments into a doubled "type1" canvas, samples every second column and crops (reference
// "HyGrid/codes in old versions.txt":68-126).  Here the host evaluates the index rule once per shape into a table of
// one source offset per output cell (HyGrid/HexFrames.py: pixel_shuffle_table) and the rearrangement is ONE pass:
//
//   gather : dst[b][c][e] = table[e] >= 0 ? src[b*batch_stride + c*chan_stride + table[e]] : 0
//   scatter: src[b*batch_stride + c*chan_stride + table[e]] = dst[b][c][e]      (adjoint; table injective, src zeroed)
//
// HBM-bound: algorithmic bytes = sizeof(src elem) * gathered cells + sizeof(dst elem) * cells (+ 8-byte table entry
// per cell, amortised over the planes a block walks: the table value lives in a register across the plane loop).
#include "hg_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace hg {

constexpr int kGatherThreads = 256;
constexpr int kGatherPlanes = 8;   // planes walked per block with one table load

template <typename TS, typename TD> __device__ __forceinline__ TD convert(TS v) { return from_f32<TD>(to_f32<TS>(v)); }
template <> __device__ __forceinline__ double convert<double, double>(double v) { return v; }

template <typename TS, typename TD>
__global__ void __launch_bounds__(kGatherThreads)
plane_gather_kernel(const TS* __restrict__ src, TD* __restrict__ dst, const int64_t* __restrict__ table, int64_t planes,
                    int chans, int64_t cells, int64_t batch_stride, int64_t chan_stride, int64_t src_total, int planes_per_block) {
  const int64_t e = (int64_t)blockIdx.x * kGatherThreads + threadIdx.x;
  if (e >= cells) return;
  const int64_t off = table[e];
  const int64_t p0 = (int64_t)blockIdx.y * planes_per_block;
  const int64_t p1 = min(p0 + planes_per_block, planes);
  int64_t b = p0 / chans;
  int c = (int)(p0 - b * chans);
  for (int64_t p = p0; p < p1; ++p) {
    const int64_t idx = b * batch_stride + c * chan_stride + off;
    TD v = from_f32<TD>(0.f);
    if (off >= 0 && idx < src_total) v = convert<TS, TD>(src[idx]);
    st_stream(dst + p * cells + e, v);
    if (++c == chans) { c = 0; ++b; }
  }
}

// One-byte cells (the hex-mosaic preview of a uint8 image): EIGHT adjacent cells per thread -- four 16-byte table loads once,
// then per plane byte gathers and ONE 8-byte store -- and 32 planes per table load, so that the 8-byte table entry costs 0.25 B
// per output byte instead of 1 B.  History (96 x 1024^2 -> 4096^2 on B200): scalar kernel 0.13 of the HBM copy rate (half of the
// traffic was the table, every store a single byte); four cells per thread 0.27, instruction-bound: every output byte cost one
// predicated byte load with 64-bit index arithmetic and a bounds test.  Now (a) the offsets are checked once per thread, not per
// plane and byte: a plane whose largest offset stays inside the source takes a path without tests, through a plane pointer and
// 32-bit offsets; (b) neighbouring cells of a magnified image mostly read the SAME source texel: a group of four equal offsets
// is one load and a multiply by 0x01010101.
constexpr int kGatherPlanesU8 = 32;

template <int CELLS>     // cells per thread: 16 (one 16-byte store per plane) or 8
__global__ void __launch_bounds__(kGatherThreads)
plane_gather_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int64_t* __restrict__ table, int64_t planes,
                       int chans, int64_t cells, int64_t batch_stride, int64_t chan_stride, int64_t src_total, int planes_per_block) {
  constexpr int G = CELLS / 4;               // 4-byte groups
  const int64_t e = ((int64_t)blockIdx.x * kGatherThreads + threadIdx.x) * CELLS;
  if (e >= cells) return;
  int64_t off[CELLS];
#pragma unroll
  for (int k = 0; k < CELLS; k += 2) {
    const longlong2 t = __ldg(reinterpret_cast<const longlong2*>(table + e + k));
    off[k] = t.x; off[k + 1] = t.y;
  }
  int64_t omin = off[0], omax = off[0];
#pragma unroll
  for (int k = 1; k < CELLS; ++k) { omin = min(omin, off[k]); omax = max(omax, off[k]); }
  const bool small = omin >= 0 && omax < (1ll << 31);            // every cell has a source, offsets fit 32 bits
  uint32_t eq = 0;                                               // bit g: the four cells of group g read the same texel
#pragma unroll
  for (int g = 0; g < G; ++g)
    if (off[4 * g] == off[4 * g + 1] && off[4 * g + 1] == off[4 * g + 2] && off[4 * g + 2] == off[4 * g + 3]) eq |= 1u << g;
  const int64_t p0 = (int64_t)blockIdx.y * planes_per_block;
  const int64_t p1 = min(p0 + planes_per_block, planes);
  int64_t b = p0 / chans;
  int c = (int)(p0 - b * chans);
  uint8_t* __restrict__ out = dst + p0 * cells + e;
  // (unrolling this loop by four measured slower: 0.75 vs 0.67 ms)
#pragma unroll 1
  for (int64_t p = p0; p < p1; ++p, out += cells) {
    const int64_t base = b * batch_stride + c * chan_stride;
    uint32_t w[G];
    if (small && base + omax < src_total) {
      const uint8_t* __restrict__ sp = src + base;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if ((eq >> g) & 1u) w[g] = (uint32_t)__ldg(sp + (uint32_t)off[4 * g]) * 0x01010101u;
        else {
          w[g] = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) w[g] |= (uint32_t)__ldg(sp + (uint32_t)off[4 * g + k]) << (8 * k);
        }
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        w[g] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t idx = base + off[4 * g + k];
          const uint32_t v = (off[4 * g + k] >= 0 && idx < src_total) ? (uint32_t)__ldg(src + idx) : 0u;
          w[g] |= v << (8 * k);
        }
      }
    }
    if (CELLS == 16) __stcs(reinterpret_cast<uint4*>(out), make_uint4(w[0], w[1], w[2 % G], w[3 % G]));
    else __stcs(reinterpret_cast<uint2*>(out), make_uint2(w[0], w[1]));
    if (++c == chans) { c = 0; ++b; }
  }
}

template <typename T>
__global__ void __launch_bounds__(kGatherThreads)
plane_scatter_kernel(const T* __restrict__ gdst, T* __restrict__ gsrc, const int64_t* __restrict__ table, int64_t planes,
                     int chans, int64_t cells, int64_t batch_stride, int64_t chan_stride, int64_t src_total, int planes_per_block) {
  const int64_t e = (int64_t)blockIdx.x * kGatherThreads + threadIdx.x;
  if (e >= cells) return;
  const int64_t off = table[e];
  if (off < 0) return;
  const int64_t p0 = (int64_t)blockIdx.y * planes_per_block;
  const int64_t p1 = min(p0 + planes_per_block, planes);
  int64_t b = p0 / chans;
  int c = (int)(p0 - b * chans);
  for (int64_t p = p0; p < p1; ++p) {
    const int64_t idx = b * batch_stride + c * chan_stride + off;
    if (idx < src_total) gsrc[idx] = gdst[p * cells + e];
    if (++c == chans) { c = 0; ++b; }
  }
}

static int check_gather(int64_t batches, int64_t chans, int64_t cells, int64_t batch_stride, int64_t chan_stride) {
  HG_REQUIRE(batches >= 0 && chans >= 0 && cells >= 0, HG_E_SHAPE, "bad shape batches=%lld chans=%lld cells=%lld",
             (long long)batches, (long long)chans, (long long)cells);
  HG_REQUIRE(batch_stride >= 0 && chan_stride >= 0, HG_E_SHAPE, "negative stride");
  HG_REQUIRE(chans < (1ll << 31) && batches < (1ll << 31), HG_E_SHAPE, "too many planes");
  HG_REQUIRE(batches == 0 || batch_stride < (1ll << 62) / batches, HG_E_SHAPE, "source too large");
  return HG_OK;
}

struct GatherGrid {
  dim3 grid;
  int planes_per_block;
};

static GatherGrid gather_grid(int64_t planes, int64_t cells) {
  // planes walked per table load: 8, or 32 while the grid still has several waves of blocks (the 8-byte table entry is
  // 1 byte per moved cell at 8 planes -- an eighth of a float32 shuffle's traffic -- and a quarter of that at 32)
  static const int env_ppb = [] { const char* e = getenv("HG_GATHER_PPB"); return e ? atoi(e) : 0; }();
  int64_t ppb = kGatherPlanes;
  if (ceil_div(cells, kGatherThreads) * ceil_div(planes, 32) >= 148 * 16) ppb = 32;
  if (env_ppb > 0) ppb = env_ppb;
  if (ceil_div(planes, ppb) > 65535) ppb = ceil_div(planes, 65535);
  GatherGrid g;
  g.grid = dim3((unsigned)ceil_div(cells, kGatherThreads), (unsigned)ceil_div(planes, ppb), 1);
  g.planes_per_block = (int)ppb;
  return g;
}

template <typename TS, typename TD>
static int launch_gather(const void* src, void* dst, const int64_t* table, int64_t batches, int64_t chans, int64_t cells,
                         int64_t batch_stride, int64_t chan_stride, cudaStream_t st) {
  const int64_t planes = batches * chans;
  if (std::is_same<TS, uint8_t>::value && std::is_same<TD, uint8_t>::value && cells % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0) {
    static const int env_cells = [] { const char* e = getenv("HG_GATHER_CELLS"); return e ? atoi(e) : 0; }();
    static const int env_ppb = [] { const char* e = getenv("HG_GATHER_PPB"); return e ? atoi(e) : 0; }();
    const int per = (cells % 16 == 0 && env_cells == 16) ? 16 : 8;     // 16 measured slower (0.79 vs 0.67 ms on the 96-plane mosaic)
    const int64_t bx = ceil_div(cells / per, kGatherThreads);
    int64_t ppb = kGatherPlanesU8;           // (16 / 32 / 48 / 96 planes per table load: 0.70 / 0.67 / 0.67 / 0.69 ms)
    if (env_ppb > 0) ppb = env_ppb;
    if (ceil_div(planes, ppb) > 65535) ppb = ceil_div(planes, 65535);
    dim3 grid((unsigned)bx, (unsigned)ceil_div(planes, ppb), 1);
    auto kern = per == 16 ? plane_gather_u8_kernel<16> : plane_gather_u8_kernel<8>;
    kern<<<grid, kGatherThreads, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, table, planes, (int)chans, cells,
                                          batch_stride, chan_stride, batches * batch_stride, (int)ppb);
    return finish_launch("plane_gather_u8");
  }
  const GatherGrid g = gather_grid(planes, cells);
  plane_gather_kernel<TS, TD><<<g.grid, kGatherThreads, 0, st>>>((const TS*)src, (TD*)dst, table, planes, (int)chans, cells,
                                                                 batch_stride, chan_stride, batches * batch_stride,
                                                                 g.planes_per_block);
  return finish_launch("plane_gather");
}

template <typename T>
static int launch_scatter(const void* gdst, void* gsrc, const int64_t* table, int64_t batches, int64_t chans, int64_t cells,
                          int64_t batch_stride, int64_t chan_stride, cudaStream_t st) {
  const int64_t planes = batches * chans;
  const GatherGrid g = gather_grid(planes, cells);
  plane_scatter_kernel<T><<<g.grid, kGatherThreads, 0, st>>>((const T*)gdst, (T*)gsrc, table, planes, (int)chans, cells,
                                                             batch_stride, chan_stride, batches * batch_stride,
                                                             g.planes_per_block);
  return finish_launch("plane_scatter");
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_plane_gather(const void* src, void* dst, const int64_t* table, int64_t batches, int64_t chans, int64_t cells,
                    int64_t batch_stride, int64_t chan_stride, int src_dtype, int dst_dtype, hg_stream_t stream) {
  int rc = check_gather(batches, chans, cells, batch_stride, chan_stride);
  if (rc) return rc;
  HG_REQUIRE(cells < (1ll << 31) * kGatherThreads, HG_E_SHAPE, "plane too large");
  if (batches * chans == 0 || cells == 0) return HG_OK;
  HG_REQUIRE(src && dst && table, HG_E_ARG, "null pointer");
  cudaStream_t st = as_stream(stream);
#define HG_GATHER_CASE(S, TS, D, TD) \
  if (src_dtype == S && dst_dtype == D) return launch_gather<TS, TD>(src, dst, table, batches, chans, cells, batch_stride, chan_stride, st);
  HG_GATHER_CASE(HG_F32, float, HG_F32, float)
  HG_GATHER_CASE(HG_BF16, __nv_bfloat16, HG_F32, float)
  HG_GATHER_CASE(HG_F64, double, HG_F32, float)
  HG_GATHER_CASE(HG_U8, uint8_t, HG_F32, float)
  HG_GATHER_CASE(HG_F64, double, HG_F64, double)
  HG_GATHER_CASE(HG_BF16, __nv_bfloat16, HG_BF16, __nv_bfloat16)
  HG_GATHER_CASE(HG_U8, uint8_t, HG_U8, uint8_t)
#undef HG_GATHER_CASE
  set_error("plane_gather: unsupported dtypes src=%d dst=%d", src_dtype, dst_dtype);
  return HG_E_DTYPE;
}

int hg_plane_scatter(const void* gdst, void* gsrc, const int64_t* table, int64_t batches, int64_t chans, int64_t cells,
                     int64_t batch_stride, int64_t chan_stride, int dtype, hg_stream_t stream) {
  int rc = check_gather(batches, chans, cells, batch_stride, chan_stride);
  if (rc) return rc;
  HG_REQUIRE(cells < (1ll << 31) * kGatherThreads, HG_E_SHAPE, "plane too large");
  if (batches * chans == 0 || cells == 0) return HG_OK;
  HG_REQUIRE(gdst && gsrc && table, HG_E_ARG, "null pointer");
  cudaStream_t st = as_stream(stream);
  switch (dtype) {
    case HG_F32: return launch_scatter<float>(gdst, gsrc, table, batches, chans, cells, batch_stride, chan_stride, st);
    case HG_F64: return launch_scatter<double>(gdst, gsrc, table, batches, chans, cells, batch_stride, chan_stride, st);
    case HG_BF16: return launch_scatter<__nv_bfloat16>(gdst, gsrc, table, batches, chans, cells, batch_stride, chan_stride, st);
    default: set_error("plane_scatter: unsupported dtype %d", dtype); return HG_E_DTYPE;
  }
}

}  // extern "C"
