// hg_host.cu -- host-buffer entry points: what a numpy caller of the reference binds.
//
// ref: the numpy-in / numpy-out contract of geometry_np.py:358 (rect_to_hex_resample), :191
// (hex_to_rect_resample), :520 (hexresize) and geometry_torch.py:191 (hex_to_square_resample), which
// moves the whole image H2D, computes, and moves the result D2H serially.
//
// Here the planes are cut into chunks that flow through a ring of kSlots device buffers, each slot on
// its own stream: H2D(chunk c+1), kernel(chunk c) and D2H(chunk c-1) run concurrently on the two copy
// engines and the SMs.  Pinned (page-locked / registered) host buffers are used in place; pageable
// ones are staged through pinned bounce buffers owned by the workspace.  The workspace (streams,
// device ring, bounce buffers) is created lazily per device, grows on demand and is released by
// hg_host_release(); it is the only memory this library ever owns.
#include "hg_common.cuh"
#include <mutex>
#include <string.h>

namespace hg {

constexpr int kSlots = 3;
constexpr int64_t kChunkBytes = 48ll << 20;  // target bytes (src + dst) per chunk
constexpr int kMaxDevices = 64;

struct Workspace {
  bool init = false;
  cudaStream_t st[kSlots] = {};
  void* dsrc[kSlots] = {};
  void* ddst[kSlots] = {};
  void* pin_in[kSlots] = {};
  void* pin_out[kSlots] = {};
  size_t dsrc_cap = 0, ddst_cap = 0, pin_in_cap = 0, pin_out_cap = 0;
  double* tables = nullptr;
  size_t tables_cap = 0;
};
static Workspace g_ws[kMaxDevices];
static std::mutex g_mu[kMaxDevices];  // one ring per device: calls on the same device are serialised (they saturate its
                                      // PCIe link anyway), calls on different devices -- one host thread per GPU -- run concurrently

#define HG_CUDA(call)                                                \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      set_error("%s: %s", #call, cudaGetErrorString(e_));            \
      return (int)e_;                                                \
    }                                                                \
  } while (0)

static int grow(void** bufs, size_t& cap, size_t need, bool pinned) {
  if (need <= cap) return HG_OK;
  for (int s = 0; s < kSlots; ++s) {
    if (bufs[s]) { if (pinned) cudaFreeHost(bufs[s]); else cudaFree(bufs[s]); bufs[s] = nullptr; }
  }
  cap = 0;
  for (int s = 0; s < kSlots; ++s) {
    if (pinned) HG_CUDA(cudaHostAlloc(&bufs[s], need, cudaHostAllocDefault));
    else HG_CUDA(cudaMalloc(&bufs[s], need));
  }
  cap = need;
  return HG_OK;
}

static void release(Workspace& w) {
  for (int s = 0; s < kSlots; ++s) {
    if (w.dsrc[s]) cudaFree(w.dsrc[s]);
    if (w.ddst[s]) cudaFree(w.ddst[s]);
    if (w.pin_in[s]) cudaFreeHost(w.pin_in[s]);
    if (w.pin_out[s]) cudaFreeHost(w.pin_out[s]);
    if (w.st[s]) cudaStreamDestroy(w.st[s]);
  }
  if (w.tables) cudaFree(w.tables);
  w = Workspace();
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Any early return (a CUDA error half-way through the ring) must not leave copies in flight that still read or
// write the caller's host buffers: drain the ring streams on the way out unless the run completed.
struct DrainOnExit {
  Workspace& w;
  bool armed = true;
  ~DrainOnExit() {
    if (!armed) return;
    for (int s = 0; s < kSlots; ++s) if (w.st[s]) cudaStreamSynchronize(w.st[s]);
    cudaGetLastError();
  }
};

// Streams `planes` independent planes through the device ring: H2D(chunk c+1) / launch(chunk c) / D2H(chunk c-1) overlap.
// launch(dsrc, ddst, n, stream) enqueues the kernel(s) for n planes and returns a status.  `prologue` (may be empty) runs once
// with the ring's first stream before any chunk (table uploads).
template <typename Prologue, typename Launch>
static int run_ring(const void* host_src, void* host_dst, int64_t planes, size_t src_plane, size_t dst_plane, int device,
                    Prologue prologue, Launch launch) {
  HG_REQUIRE(device >= 0 && device < kMaxDevices, HG_E_ARG, "bad device ordinal %d", device);
  std::lock_guard<std::mutex> lock(g_mu[device]);
  DeviceGuard guard(device);
  Workspace& ws = g_ws[device];
  if (!ws.init) {
    for (int s = 0; s < kSlots; ++s) HG_CUDA(cudaStreamCreateWithFlags(&ws.st[s], cudaStreamNonBlocking));
    ws.init = true;
  }
  int64_t per = kChunkBytes / (int64_t)(src_plane + dst_plane);
  if (per < 1) per = 1;
  if (per > planes) per = planes;
  // at least 2*kSlots chunks when there is enough work, so that the ring actually overlaps
  if (planes / per < 2 * kSlots && planes >= 2 * kSlots) per = planes / (2 * kSlots);
  const bool src_pinned = is_pinned(host_src), dst_pinned = is_pinned(host_dst);
  int rc;
  if ((rc = grow(ws.dsrc, ws.dsrc_cap, per * src_plane, false))) return rc;
  if ((rc = grow(ws.ddst, ws.ddst_cap, per * dst_plane, false))) return rc;
  if (!src_pinned && (rc = grow(ws.pin_in, ws.pin_in_cap, per * src_plane, true))) return rc;
  if (!dst_pinned && (rc = grow(ws.pin_out, ws.pin_out_cap, per * dst_plane, true))) return rc;
  if ((rc = prologue(ws))) return rc;

  const char* src = (const char*)host_src;
  char* dst = (char*)host_dst;
  DrainOnExit drain{ws};
  int64_t pending_p0[kSlots], pending_n[kSlots];
  for (int s = 0; s < kSlots; ++s) pending_n[s] = 0;
  int64_t c = 0;
  for (int64_t p0 = 0; p0 < planes; p0 += per, ++c) {
    const int s = (int)(c % kSlots);
    const int64_t n = planes - p0 < per ? planes - p0 : per;
    const void* hsrc = src + (size_t)p0 * src_plane;
    if (!src_pinned || !dst_pinned) {
      // the slot's previous chunk must have drained before its bounce buffers are touched
      HG_CUDA(cudaStreamSynchronize(ws.st[s]));
      if (!dst_pinned && pending_n[s]) {
        memcpy(dst + (size_t)pending_p0[s] * dst_plane, ws.pin_out[s], (size_t)pending_n[s] * dst_plane);
        pending_n[s] = 0;
      }
      if (!src_pinned) {
        memcpy(ws.pin_in[s], hsrc, (size_t)n * src_plane);
        hsrc = ws.pin_in[s];
      }
    }
    HG_CUDA(cudaMemcpyAsync(ws.dsrc[s], hsrc, (size_t)n * src_plane, cudaMemcpyHostToDevice, ws.st[s]));
    rc = launch(ws.dsrc[s], ws.ddst[s], n, ws.st[s]);
    if (rc) return rc;   // hg_last_error() already holds the kernel launcher's message; DrainOnExit waits for the ring
    void* hdst = dst_pinned ? (void*)(dst + (size_t)p0 * dst_plane) : ws.pin_out[s];
    HG_CUDA(cudaMemcpyAsync(hdst, ws.ddst[s], (size_t)n * dst_plane, cudaMemcpyDeviceToHost, ws.st[s]));
    if (!dst_pinned) { pending_p0[s] = p0; pending_n[s] = n; }
  }
  for (int k = 0; k < kSlots; ++k) {
    const int s = (int)((c + k) % kSlots);  // oldest first
    HG_CUDA(cudaStreamSynchronize(ws.st[s]));
    if (!dst_pinned && pending_n[s]) {
      memcpy(dst + (size_t)pending_p0[s] * dst_plane, ws.pin_out[s], (size_t)pending_n[s] * dst_plane);
      pending_n[s] = 0;
    }
  }
  drain.armed = false;
  return HG_OK;
}

// kind: 0 rect->hex, 1 hex-source (hex->rect / hexresize: they differ only in the coordinate tables)
static int run_host(int kind, const void* host_src, void* host_dst, const double* host_xs, const double* host_ys,
                    int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt, int ddt, int interp, int math,
                    int device) {
  HG_REQUIRE(planes >= 0 && h > 0 && w > 0 && h1 >= 0 && w1 >= 0, HG_E_SHAPE, "bad shape");
  HG_REQUIRE(interp == 0 || interp == 1, HG_E_ARG, "interp must be 0 (nearest) or 1 (bilinear / linear)");
  HG_REQUIRE(dtype_size(sdt) && dtype_size(ddt), HG_E_DTYPE, "unknown dtype");
  HG_REQUIRE(interp == 1 || dtype_size(sdt) == dtype_size(ddt), HG_E_DTYPE, "nearest keeps the element type");
  if (planes == 0 || h1 == 0 || w1 == 0) return HG_OK;
  HG_REQUIRE(host_src && host_dst && host_xs && host_ys, HG_E_ARG, "NULL buffer");
  const size_t src_plane = (size_t)h * w * dtype_size(sdt), dst_plane = (size_t)h1 * w1 * dtype_size(ddt);
  double* dxs = nullptr;
  double* dys = nullptr;
  auto prologue = [&](Workspace& ws) -> int {
    const size_t tab = (size_t)(h1 + w1) * sizeof(double);
    if (tab > ws.tables_cap) {
      if (ws.tables) cudaFree(ws.tables);
      ws.tables = nullptr; ws.tables_cap = 0;
      HG_CUDA(cudaMalloc((void**)&ws.tables, tab));
      ws.tables_cap = tab;
    }
    dxs = ws.tables;
    dys = ws.tables + h1;
    // tables are shared by every slot: upload on one ring stream and wait, so that no slot can run ahead
    HG_CUDA(cudaMemcpyAsync(dxs, host_xs, (size_t)h1 * sizeof(double), cudaMemcpyHostToDevice, ws.st[0]));
    HG_CUDA(cudaMemcpyAsync(dys, host_ys, (size_t)w1 * sizeof(double), cudaMemcpyHostToDevice, ws.st[0]));
    HG_CUDA(cudaStreamSynchronize(ws.st[0]));
    return HG_OK;
  };
  auto launch = [&](void* dsrc, void* ddst, int64_t n, cudaStream_t st) -> int {
    if (kind == 0)
      return interp == 0 ? hg_rect2hex_nearest(dsrc, ddst, dxs, dys, n, h, w, h1, w1, dtype_size(sdt), st)
                         : hg_rect2hex_bilinear(dsrc, ddst, dxs, dys, host_xs, host_ys, n, h, w, h1, w1, sdt, ddt, math, st);
    return interp == 0 ? hg_hex2rect_nearest(dsrc, ddst, dxs, dys, n, h, w, h1, w1, dtype_size(sdt), st)
                       : hg_hex2rect_linear(dsrc, ddst, dxs, dys, host_xs, host_ys, n, h, w, h1, w1, sdt, ddt, math, st);
  };
  return run_ring(host_src, host_dst, planes, src_plane, dst_plane, device, prologue, launch);
}

// rows_mul: 1 = type1 raster [planes, H, 2W+1], 2 = type2 raster [planes, 2H, 2W+1]; decode != 0 runs the inverse
static int run_host_layout(const void* host_src, void* host_dst, int64_t planes, int64_t H, int64_t W, int offset, int sdt,
                           int ddt, int rows_mul, int decode, int device) {
  HG_REQUIRE(planes >= 0 && H > 0 && W > 0, HG_E_SHAPE, "bad shape");
  HG_REQUIRE(rows_mul == 1 || rows_mul == 2, HG_E_ARG, "rows_mul must be 1 (type1) or 2 (type2)");
  HG_REQUIRE(dtype_size(sdt) && dtype_size(ddt), HG_E_DTYPE, "unknown dtype");
  if (planes == 0) return HG_OK;
  HG_REQUIRE(host_src && host_dst, HG_E_ARG, "NULL buffer");
  const size_t hex_plane = (size_t)H * W, ras_plane = (size_t)rows_mul * H * (2 * W + 1);
  const size_t src_plane = (decode ? ras_plane : hex_plane) * dtype_size(sdt);
  const size_t dst_plane = (decode ? hex_plane : ras_plane) * dtype_size(ddt);
  auto prologue = [](Workspace&) -> int { return HG_OK; };
  auto launch = [&](void* dsrc, void* ddst, int64_t n, cudaStream_t st) -> int {
    if (decode) return hg_type_to_hex(dsrc, ddst, n, rows_mul * H, 2 * W + 1, rows_mul, sdt, ddt, st);
    return rows_mul == 1 ? hg_hex_to_type1(dsrc, ddst, n, H, W, offset, sdt, ddt, st)
                         : hg_hex_to_type2(dsrc, ddst, n, H, W, offset, sdt, ddt, st);
  };
  return run_ring(host_src, host_dst, planes, src_plane, dst_plane, device, prologue, launch);
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_host_rect2hex(const void* host_src, void* host_dst, const double* host_xs, const double* host_ys, int64_t planes,
                     int64_t h, int64_t w, int64_t h1, int64_t w1, int src_dtype, int dst_dtype, int interp, int math,
                     int device) {
  return run_host(0, host_src, host_dst, host_xs, host_ys, planes, h, w, h1, w1, src_dtype, dst_dtype, interp, math, device);
}

int hg_host_hex2rect(const void* host_src, void* host_dst, const double* host_xs, const double* host_ys, int64_t planes,
                     int64_t h, int64_t w, int64_t h1, int64_t w1, int src_dtype, int dst_dtype, int interp, int math,
                     int device) {
  return run_host(1, host_src, host_dst, host_xs, host_ys, planes, h, w, h1, w1, src_dtype, dst_dtype, interp, math, device);
}

int hg_host_hex_to_type(const void* host_hex, void* host_raster, int64_t planes, int64_t H, int64_t W, int offset,
                        int src_dtype, int dst_dtype, int rows_mul, int device) {
  return run_host_layout(host_hex, host_raster, planes, H, W, offset, src_dtype, dst_dtype, rows_mul, 0, device);
}

int hg_host_type_to_hex(const void* host_raster, void* host_hex, int64_t planes, int64_t H, int64_t W, int src_dtype,
                        int dst_dtype, int rows_mul, int device) {
  return run_host_layout(host_raster, host_hex, planes, H, W, 0, src_dtype, dst_dtype, rows_mul, 1, device);
}

void hg_host_release(void) {
  for (int d = 0; d < kMaxDevices; ++d) {
    std::lock_guard<std::mutex> lock(g_mu[d]);
    if (!g_ws[d].init) continue;
    DeviceGuard guard(d);
    release(g_ws[d]);
  }
}

}  // extern "C"
