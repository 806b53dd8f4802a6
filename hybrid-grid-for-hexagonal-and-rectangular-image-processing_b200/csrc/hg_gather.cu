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

// One-byte cells (the hex-mosaic preview of a uint8 image): four adjacent cells per thread -- two 16-byte table loads, four
// byte gathers and ONE 4-byte store per plane -- and 32 planes per table load, so that the 8-byte table entry costs 0.25 B per
// output byte instead of 1 B (round-2 measurement of the scalar kernel on 96 x 1024^2 -> 4096^2: 0.13 of the HBM copy rate,
// half of the traffic was the table and every store was a single byte).
constexpr int kGatherPlanesU8 = 32;

__global__ void __launch_bounds__(kGatherThreads)
plane_gather_u8x4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int64_t* __restrict__ table, int64_t planes,
                         int chans, int64_t cells, int64_t batch_stride, int64_t chan_stride, int64_t src_total, int planes_per_block) {
  const int64_t e = ((int64_t)blockIdx.x * kGatherThreads + threadIdx.x) * 4;
  if (e >= cells) return;
  const longlong2 t01 = __ldg(reinterpret_cast<const longlong2*>(table + e));
  const longlong2 t23 = __ldg(reinterpret_cast<const longlong2*>(table + e + 2));
  const int64_t off[4] = {t01.x, t01.y, t23.x, t23.y};
  const int64_t p0 = (int64_t)blockIdx.y * planes_per_block;
  const int64_t p1 = min(p0 + planes_per_block, planes);
  int64_t b = p0 / chans;
  int c = (int)(p0 - b * chans);
  for (int64_t p = p0; p < p1; ++p) {
    const int64_t base = b * batch_stride + c * chan_stride;
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t idx = base + off[k];
      const uint32_t v = (off[k] >= 0 && idx < src_total) ? (uint32_t)__ldg(src + idx) : 0u;
      word |= v << (8 * k);
    }
    __stcs(reinterpret_cast<uint32_t*>(dst + p * cells + e), word);
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
  int64_t ppb = kGatherPlanes;
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
  if (std::is_same<TS, uint8_t>::value && std::is_same<TD, uint8_t>::value && cells % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 3) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0) {
    int64_t ppb = kGatherPlanesU8;
    if (ceil_div(planes, ppb) > 65535) ppb = ceil_div(planes, 65535);
    dim3 grid((unsigned)ceil_div(cells / 4, kGatherThreads), (unsigned)ceil_div(planes, ppb), 1);
    plane_gather_u8x4_kernel<<<grid, kGatherThreads, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, table, planes, (int)chans, cells,
                                                              batch_stride, chan_stride, batches * batch_stride, (int)ppb);
    return finish_launch("plane_gather_u8x4");
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
