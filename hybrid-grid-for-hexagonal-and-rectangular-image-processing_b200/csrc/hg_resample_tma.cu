// hg_resample_tma.cu -- rect -> hex bilinear resampling with TMA-staged source tiles (sm_100a).
//
// ref: geometry_np.py:358-519 rect_to_hex_resample(..., 'bilinear').
//
// The direct gather kernel (hg_resample.cu) is issue-bound when the lattices have similar pitch: four
// predicated global loads with 64-bit addressing per output.  Here a persistent CTA walks a list of
// (tile position, plane) items; for each item ONE thread issues a 3-D TMA box load
// (cp.async.bulk.tensor) of the source footprint of the 32 x 128 (or 64 x 128) output tile -- rows
// i_n(a0) .. i_n(a1)+1, columns j_n(b0) .. j_n(b1)+1 -- into a ring of shared-memory stages.  Box elements
// outside the image are zero-filled by the TMA unit, which *is* the reference's zero-fill of
// out-of-range taps, so the compute loop has no bounds checks at all: per output two shared loads
// (the upper pair is carried in registers from the previous row whenever i_n advances by one), three
// FMAs and one coalesced streaming store.  Items are ordered position-major / plane-minor so that the
// per-thread row / column tables are computed once per position and reused for every plane.
#include "hg_common.cuh"
#include "hg_ptx.cuh"
#include <stdlib.h>

namespace hg {

constexpr int kTW = 128;       // output tile width  (4 columns per lane, 32 apart)
constexpr int kTmaThreads = 256;
constexpr int kTmaStages = 3;

// TMA needs the innermost box coordinate on a 16-byte boundary (measured on B200: a start column that is not
// a multiple of 4 floats faults with "illegal instruction"; negative multiples are fine and zero-fill).
template <typename TS> __device__ __forceinline__ int align_col(int c) {
  constexpr int A = 16 / (int)sizeof(TS);
  return c & ~(A - 1);                      // floor to a multiple of A, also for negative c (two's complement)
}

__device__ __forceinline__ void rect_axis_d(double coord, int n, int& idx, double& frac) {
  // i_ = x_ + (h-1)*0.5 ; i_n = trunc(i_) ; i_f = i_ - float32(i_n)      (geometry_np.py:440-449)
  const double c = dadd(coord, (double)(n - 1) * 0.5);
  idx = trunc_i32(c);
  frac = dsub(c, (double)(float)idx);
}

// Per-item lattice tables in shared memory (double-buffered): the 128 column entries and 8*RW row entries of
// the tile are computed by one thread each (instead of 4 + RW by every thread) while the previous tile is
// being blended.
template <typename WT, int RW>
struct TileTables {
  int cidx[kTW];        // j_n - col0  (column inside the staged box)
  WT cfrac[kTW];        // j_f
  int ridx[8 * RW];     // (i_n - row0) * BW  (row offset inside the staged box)
  WT rfrac[8 * RW];     // i_f
  int nrows;            // valid output rows of the tile
  int ncols;            // valid output columns of the tile
};

template <typename TS, typename TD, bool EXACT, int RW>  // RW = output rows per warp; tile height = 8*RW
__global__ void __launch_bounds__(kTmaThreads)
rect2hex_bilinear_tma_kernel(const __grid_constant__ CUtensorMap tmap, TD* __restrict__ dst, const double* __restrict__ xs,
                             const double* __restrict__ ys, int h, int w, int h1, int w1, int tiles_x, int tiles_y,
                             long long total_items, long long items_per_cta, int BW, int BH, int stage_bytes) {
  using WT = typename std::conditional<EXACT, double, float>::type;
  using Tab = TileTables<WT, RW>;
  constexpr int TH = 8 * RW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kTmaStages * stage_bytes);
  Tab* tabs = reinterpret_cast<Tab*>(smem_raw + (size_t)kTmaStages * stage_bytes + 64);

  const long long g_begin = (long long)blockIdx.x * items_per_cta;
  const long long g_end = min(total_items, g_begin + items_per_cta);
  if (g_begin >= g_end) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npos = tiles_x * tiles_y;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap);
    for (int s = 0; s < kTmaStages; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }

  // items run plane-major, tiles row-major inside a plane: horizontally and vertically adjacent tiles are
  // staged close in time by the same CTA, so their halo rows / columns come out of L2, not HBM.
  auto decode = [&](long long g, int& plane, int& tx, int& ty) {
    plane = (int)(g / npos);
    const int pos = (int)(g - (long long)plane * npos);
    ty = pos / tiles_x; tx = pos - ty * tiles_x;
  };
  auto origin = [&](int tx, int ty, int& row0, int& col0) {
    double f;
    rect_axis_d(xs[ty * TH], h, row0, f);
    rect_axis_d(ys[tx * kTW], w, col0, f);
    col0 = align_col<TS>(col0);
  };
  auto issue = [&](long long g, int s) {   // one thread
    int plane, tx, ty, row0, col0;
    decode(g, plane, tx, ty);
    origin(tx, ty, row0, col0);
    ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(BW * BH * (int)sizeof(TS)));
    ptx::tma_load_3d(smem_raw + (size_t)s * stage_bytes, &tmap, &full[s], col0, row0, plane);
  };
  auto build_tables = [&](long long g, Tab& T) {   // threads 0 .. kTW + TH - 1, one entry each
    int plane, tx, ty, row0, col0;
    decode(g, plane, tx, ty);
    origin(tx, ty, row0, col0);
    const int t = threadIdx.x;
    if (t < kTW) {
      const int b = tx * kTW + t;
      int jn = col0; double v = 0.0;
      if (b < w1) rect_axis_d(ys[b], w, jn, v);
      T.cidx[t] = jn - col0;
      T.cfrac[t] = (WT)v;
      if (t == 0) { T.ncols = min(kTW, w1 - tx * kTW); T.nrows = min(TH, h1 - ty * TH); }
    } else if (t < kTW + TH) {
      const int r = t - kTW, a = ty * TH + r;
      int in = row0; double u = 0.0;
      if (a < h1) rect_axis_d(xs[a], h, in, u);
      T.ridx[r] = (in - row0) * BW;
      T.rfrac[r] = (WT)u;
    }
  };

  build_tables(g_begin, tabs[0]);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages && g_begin + s < g_end; ++s) issue(g_begin + s, s);
  }

  for (long long k = 0; g_begin + k < g_end; ++k) {
    const int s = (int)(k % kTmaStages);
    const uint32_t parity = (uint32_t)((k / kTmaStages) & 1);
    const Tab& T = tabs[k & 1];
    if (g_begin + k + 1 < g_end) build_tables(g_begin + k + 1, tabs[(k + 1) & 1]);

    int plane, tx, ty;
    decode(g_begin + k, plane, tx, ty);
    int coff[4];
    WT jf[4];
    bool cok[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      coff[c] = T.cidx[lane + 32 * c];
      jf[c] = T.cfrac[lane + 32 * c];
      cok[c] = lane + 32 * c < T.ncols;
    }
    const int nrows = min(RW, T.nrows - warp * RW);
    TD* __restrict__ dp = dst + (size_t)plane * h1 * w1 + (size_t)(ty * TH + warp * RW) * w1 + (tx * kTW + lane);

    ptx::mbar_wait(&full[s], parity);
    const TS* __restrict__ t = reinterpret_cast<const TS*>(smem_raw + (size_t)s * stage_bytes);

    WT bl[4], br[4];   // lower pair of the previous row (= upper pair of this row when i_n advanced by one)
    int prev_roff = 0;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      if (r < nrows) {
        const int roff = T.ridx[warp * RW + r];
        const WT u = T.rfrac[warp * RW + r];
        const bool carry = (r > 0) && (roff == prev_roff + BW);   // warp-uniform
        prev_roff = roff;
        WT tl[4], tr[4];
        if (carry) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { tl[c] = bl[c]; tr[c] = br[c]; }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int o = roff + coff[c];
            tl[c] = (WT)t[o];
            tr[c] = (WT)t[o + 1];
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int o = roff + BW + coff[c];
          bl[c] = (WT)t[o];
          br[c] = (WT)t[o + 1];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          TD o;
          if (EXACT) {   // literal operation order of geometry_np.py:515-517, no contraction
            const double v = jf[c];
            const double u1 = dsub(1.0, u), v1 = dsub(1.0, v);
            const double t1 = dadd(dmul(u, bl[c]), dmul(u1, tl[c]));
            const double t2 = dadd(dmul(u, br[c]), dmul(u1, tr[c]));
            o = (TD)dadd(dmul(v, t2), dmul(v1, t1));
          } else {
            const float v = jf[c];
            const float t1 = fmaf(u, bl[c] - tl[c], tl[c]);
            const float t2 = fmaf(u, br[c] - tr[c], tr[c]);
            o = (TD)fmaf(v, t2 - t1, t1);
          }
          if (cok[c]) st_stream(dp + (size_t)r * w1 + 32 * c, o);
        }
      }
    }
    __syncthreads();                            // stage s fully read; next tile's tables complete
    if (threadIdx.x == 0 && g_begin + k + kTmaStages < g_end) issue(g_begin + k + kTmaStages, s);
  }
}

// shared-memory load from a 32-bit shared-window address (keeps tap addresses in one register each)
template <typename TS> __device__ __forceinline__ TS lds_at(uint32_t addr);
template <> __device__ __forceinline__ float lds_at<float>(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <> __device__ __forceinline__ double lds_at<double>(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <> __device__ __forceinline__ uint8_t lds_at<uint8_t>(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return (uint8_t)v;
}

// ---- warp-specialised variant ----------------------------------------------------------------------------
// Same tiles, same arithmetic, different choreography (ncu r1o/r1p of the kernel above: one __syncthreads per
// tile and the exposed global-load latency of the per-tile tables cost ~1/3 of the issue slots):
//   warps 0 .. NW-1  consumers: wait full[s] -> blend their RW rows of the tile -> arrive empty[s]
//   warp NW          producer : wait empty[s] -> one lane issues the TMA box load, all lanes compute the tile's
//                               lattice tables (global xs / ys reads, float64) into tabs[s] -> arrive full[s]
// so the consumers never touch global memory except for their output stores, never execute a CTA barrier, and
// the table / TMA latency is hidden behind kStages tiles of lookahead.  full[s] completes on two arrivals (TMA
// issue + tables written) plus the box's byte count; empty[s] on one arrival per consumer warp.
template <typename TS, typename TD, bool EXACT, int RW, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, EXACT ? 2 : (NW == 16 ? 2 : 3))
rect2hex_bilinear_ws_kernel(const __grid_constant__ CUtensorMap tmap, TD* __restrict__ dst, const double* __restrict__ xs,
                            const double* __restrict__ ys, int h, int w, int h1, int w1, int tiles_x, int tiles_y,
                            long long total_items, long long items_per_cta, int BW, int BH, int stage_bytes) {
  using WT = typename std::conditional<EXACT, double, float>::type;
  constexpr int TH = NW * RW;
  struct Tab {
    int cidx[kTW];
    WT cfrac[kTW];
    int ridx[TH];
    WT rfrac[TH];
    int nrows, ncols;
  };
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kTmaStages * stage_bytes);
  uint64_t* empty = full + kTmaStages;
  Tab* tabs = reinterpret_cast<Tab*>(smem_raw + (size_t)kTmaStages * stage_bytes + 64);

  const long long g_begin = (long long)blockIdx.x * items_per_cta;
  const long long g_end = min(total_items, g_begin + items_per_cta);
  if (g_begin >= g_end) return;
  const int n_items = (int)(g_end - g_begin);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npos = tiles_x * tiles_y;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap);
    for (int s = 0; s < kTmaStages; ++s) { ptx::mbar_init(&full[s], 2); ptx::mbar_init(&empty[s], NW); }
    ptx::fence_barrier_init();
  }
  __syncthreads();

  // first item of this CTA; afterwards (plane, ty, tx) advance by carries (items are plane-major, tiles row-major)
  int plane = (int)(g_begin / npos);
  int ty, tx;
  {
    const int pos = (int)(g_begin - (long long)plane * npos);
    ty = pos / tiles_x; tx = pos - ty * tiles_x;
  }
  int s = 0;
  uint32_t ph = 0;

  if (warp == NW) {
    // ===== producer ===========================================================================================
    for (int k = 0; k < n_items; ++k) {
      ptx::mbar_wait(&empty[s], ph ^ 1);                 // first pass: fresh barrier, returns at once
      int row0, col0;
      double f;
      rect_axis_d(xs[ty * TH], h, row0, f);
      rect_axis_d(ys[tx * kTW], w, col0, f);
      col0 = align_col<TS>(col0);
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(BW * BH * (int)sizeof(TS)));
        ptx::tma_load_3d(smem_raw + (size_t)s * stage_bytes, &tmap, &full[s], col0, row0, plane);
      }
      Tab& T = tabs[s];
#pragma unroll
      for (int i = 0; i < kTW / 32; ++i) {
        const int t = lane + 32 * i, b = tx * kTW + t;
        int jn = col0; double v = 0.0;
        if (b < w1) rect_axis_d(ys[b], w, jn, v);
        T.cidx[t] = (jn - col0) * (int)sizeof(TS);       // byte offset inside a staged row
        T.cfrac[t] = (WT)v;
      }
#pragma unroll
      for (int i = 0; i < (TH + 31) / 32; ++i) {
        const int r = lane + 32 * i, a = ty * TH + r;
        if (r < TH) {
          int in = row0; double u = 0.0;
          if (a < h1) rect_axis_d(xs[a], h, in, u);
          T.ridx[r] = (in - row0) * BW * (int)sizeof(TS);  // byte offset of the staged row
          T.rfrac[r] = (WT)u;
        }
      }
      if (lane == 0) { T.ncols = min(kTW, w1 - tx * kTW); T.nrows = min(TH, h1 - ty * TH); }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full[s]);
      if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++plane; } }
      if (++s == kTmaStages) { s = 0; ph ^= 1; }
    }
    return;
  }

  // ===== consumers ============================================================================================
  // Table entries are BYTE offsets into the staged box, so a tap address is one integer add; the row pairs
  // ping-pong between two register sets (the lower pair of row r is the upper pair of row r + 1 whenever i_n
  // advances by one -- no register moves); interior tiles take a path without any bounds predicate.
  const int bw_b = BW * (int)sizeof(TS);
  for (int k = 0; k < n_items; ++k) {
    ptx::mbar_wait(&full[s], ph);
    const Tab& T = tabs[s];
    const uint32_t tb = ptx::smem_u32(smem_raw + (size_t)s * stage_bytes);
    uint32_t cb[4];
    WT jf[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cb[c] = tb + T.cidx[lane + 32 * c];
      jf[c] = T.cfrac[lane + 32 * c];
    }
    const int ncols = T.ncols, nrows = T.nrows - warp * RW;
    TD* __restrict__ dp = dst + (size_t)plane * h1 * w1 + (size_t)(ty * TH + warp * RW) * w1 + (tx * kTW + lane);

    auto rows = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      WT v[2][4][2];     // [set][column][left / right tap]
      int prev_roff = 0;
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        if (FULL || r < nrows) {
          constexpr int kDummy = 0; (void)kDummy;
          const int A = r & 1, B = A ^ 1;                     // compile-time after unrolling
          const int roff = T.ridx[warp * RW + r];
          const WT u = T.rfrac[warp * RW + r];
          const bool carry = (r > 0) && (roff == prev_roff + bw_b);   // warp-uniform
          prev_roff = roff;
          if (!carry) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              v[A][c][0] = (WT)lds_at<TS>(cb[c] + roff);
              v[A][c][1] = (WT)lds_at<TS>(cb[c] + roff + (uint32_t)sizeof(TS));
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            v[B][c][0] = (WT)lds_at<TS>(cb[c] + roff + bw_b);
            v[B][c][1] = (WT)lds_at<TS>(cb[c] + roff + bw_b + (uint32_t)sizeof(TS));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const WT tl = v[A][c][0], tr = v[A][c][1], bl = v[B][c][0], br = v[B][c][1];
            TD o;
            if (EXACT) {   // literal operation order of geometry_np.py:515-517, no contraction
              const double vv = jf[c];
              const double u1 = dsub(1.0, u), v1 = dsub(1.0, vv);
              const double t1 = dadd(dmul(u, bl), dmul(u1, tl));
              const double t2 = dadd(dmul(u, br), dmul(u1, tr));
              o = (TD)dadd(dmul(vv, t2), dmul(v1, t1));
            } else {
              const float vv = jf[c];
              const float t1 = fmaf(u, bl - tl, tl);
              const float t2 = fmaf(u, br - tr, tr);
              o = (TD)fmaf(vv, t2 - t1, t1);
            }
            if (FULL || lane + 32 * c < ncols) st_stream(dp + 32 * c, o);
          }
          dp += w1;
        }
      }
    };
    if (ncols == kTW && nrows >= RW) rows(std::true_type{});
    else rows(std::false_type{});
    // every shared-memory value of the stage has been consumed by a store above; generic reads -> TMA refill
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[s]);
    if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++plane; } }
    if (++s == kTmaStages) { s = 0; ph ^= 1; }
  }
}

// ---- host side ---------------------------------------------------------------------------------------
static inline int host_axis_index(double coord, int64_t n) {
  const double c = coord + (double)(n - 1) * 0.5;
  return (int)c;  // truncation toward zero, like the device path
}

// Largest source footprint of any tile along one axis (+1 for the "index + 1" tap); -1 when the table is
// not monotone (then the tiled kernel does not apply).
static int axis_span(const double* tab, int64_t n_out, int64_t n_src, int tile) {
  int span = 0;
  for (int64_t t0 = 0; t0 < n_out; t0 += tile) {
    const int64_t t1 = (t0 + tile < n_out ? t0 + tile : n_out) - 1;
    const int first = host_axis_index(tab[t0], n_src);
    int prev = first;
    for (int64_t t = t0 + 1; t <= t1; ++t) {
      const int cur = host_axis_index(tab[t], n_src);
      if (cur < prev) return -1;
      prev = cur;
    }
    if (prev - first + 2 > span) span = prev - first + 2;
  }
  return span;
}

template <typename T> struct TmaType;
template <> struct TmaType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaType<double> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; };
template <> struct TmaType<uint8_t> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_UINT8; };

static int g_sm_count = 0;

template <typename TS, typename TD, bool EXACT, int RW>
static int launch_tma(const void* src, void* dst, const double* xs, const double* ys, int64_t planes, int64_t h, int64_t w,
                      int64_t h1, int64_t w1, int BW, int BH, cudaStream_t st) {
  constexpr int TH = 8 * RW;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return 1;
  alignas(64) CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)w * sizeof(TS), (cuuint64_t)w * h * sizeof(TS)};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tmap, TmaType<TS>::v, 3, const_cast<void*>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  const int stage_bytes = (int)ceil_div((int64_t)BW * BH * sizeof(TS), 128) * 128;
  using WT = typename std::conditional<EXACT, double, float>::type;
  const int smem = kTmaStages * stage_bytes + 64 + 2 * (int)sizeof(TileTables<WT, RW>);
  auto kern = rect2hex_bilinear_tma_kernel<TS, TD, EXACT, RW>;
  static SmemReservation reservation;
  if (reservation.reserve(kern, (size_t)smem) != cudaSuccess) return 1;
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTmaThreads, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return 1; }
  const int tiles_x = (int)ceil_div(w1, kTW), tiles_y = (int)ceil_div(h1, TH);
  const long long total = (long long)tiles_x * tiles_y * planes;
  if (kTW + TH > kTmaThreads) return 1;
  long long grid = (long long)g_sm_count * occ;
  if (grid > total) grid = total;
  const long long per = (total + grid - 1) / grid;
  grid = (total + per - 1) / per;
  kern<<<(unsigned)grid, kTmaThreads, smem, st>>>(tmap, (TD*)dst, xs, ys, (int)h, (int)w, (int)h1, (int)w1, tiles_x, tiles_y,
                                                   total, per, BW, BH, stage_bytes);
  return finish_launch("rect2hex_bilinear_tma");
}

template <typename TS, typename TD, bool EXACT, int RW, int NW>
static int launch_ws(const void* src, void* dst, const double* xs, const double* ys, int64_t planes, int64_t h, int64_t w,
                     int64_t h1, int64_t w1, int BW, int BH, cudaStream_t st) {
  constexpr int TH = NW * RW, kThreadsWs = (NW + 1) * 32;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return 1;
  alignas(64) CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)w * sizeof(TS), (cuuint64_t)w * h * sizeof(TS)};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tmap, TmaType<TS>::v, 3, const_cast<void*>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  const int stage_bytes = (int)ceil_div((int64_t)BW * BH * sizeof(TS), 128) * 128;
  using WT = typename std::conditional<EXACT, double, float>::type;
  const int tab_bytes = (int)((2 * kTW + 2 * TH) * (sizeof(int) + sizeof(WT)) / 2 + 16 + 15) / 16 * 16;   // >= sizeof(Tab)
  const int smem = kTmaStages * stage_bytes + 64 + kTmaStages * (tab_bytes + 16);
  auto kern = rect2hex_bilinear_ws_kernel<TS, TD, EXACT, RW, NW>;
  static SmemReservation reservation;
  if (reservation.reserve(kern, (size_t)smem) != cudaSuccess) return 1;
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreadsWs, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return 1; }
  const int tiles_x = (int)ceil_div(w1, kTW), tiles_y = (int)ceil_div(h1, TH);
  const long long total = (long long)tiles_x * tiles_y * planes;
  long long grid = (long long)g_sm_count * occ;
  if (grid > total) grid = total;
  const long long per = (total + grid - 1) / grid;
  grid = (total + per - 1) / per;
  kern<<<(unsigned)grid, kThreadsWs, smem, st>>>(tmap, (TD*)dst, xs, ys, (int)h, (int)w, (int)h1, (int)w1, tiles_x, tiles_y,
                                                  total, per, BW, BH, stage_bytes);
  return finish_launch("rect2hex_bilinear_ws");
}

static bool exact_sync_env() {
  static const bool v = [] { const char* e = getenv("HG_R2H_EXACT_SYNC"); return e && e[0] == '1'; }();
  return v;
}

// Returns HG_OK when the tiled kernel was launched, 1 when it does not apply (caller falls back to the
// direct gather), or an error code.
int try_rect2hex_bilinear_tma(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs,
                              const double* host_ys, int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt,
                              int ddt, int math, cudaStream_t st) {
  if (!host_xs || !host_ys || sdt != HG_F32 || ddt != HG_F32) return 1;
  if ((w * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 || planes >= (1ll << 31)) return 1;
  static const int rows_per_warp = [] { const char* e = getenv("HG_R2H_ROWS"); return e ? atoi(e) : 4; }();
  if (rows_per_warp == 0) return 1;              // HG_R2H_ROWS=0 forces the direct kernel (A/B measurements)
  // HG_R2H_WS: 0 = CTA-synchronous kernel, 8 / 16 = warp-specialised kernel with that many consumer warps (default 8)
  static const int ws_warps = [] { const char* e = getenv("HG_R2H_WS"); return e ? atoi(e) : 8; }();
  const int RW = rows_per_warp == 8 ? 8 : 4;
  const bool ws = ws_warps != 0 && (math != HG_MATH_EXACT || !exact_sync_env());
  const int TH = (ws && ws_warps == 16 ? 16 : 8) * (ws ? 4 : RW);
  const int span_r = axis_span(host_xs, h1, h, TH), span_c = axis_span(host_ys, w1, w, kTW);
  if (span_r < 0 || span_c < 0) return 1;
  const int BH = span_r, BW = (span_c + 3 + 3) / 4 * 4;   // + up to 3 columns for the 16-byte aligned box origin
  if (BH > 256 || BW > 256) return 1;
  // staging pays off while the footprint is close to the tile (every staged byte is used ~4 times);
  // for strong down-sampling the direct gather already runs at the HBM roofline.
  if ((int64_t)BH * BW > (int64_t)2 * TH * kTW) return 1;
  // warp-specialised kernel (measured: C2 0.88 / C4 0.85 of the HBM copy rate vs 0.82 / 0.83 for the CTA-synchronous
  // one with float32 math; HG_MATH_EXACT 0.74-0.81 vs 0.65).  HG_R2H_WS=0 / HG_R2H_EXACT_SYNC=1 select the
  // CTA-synchronous kernel for A/B runs.
  if (math != HG_MATH_EXACT && ws_warps == 16) return launch_ws<float, float, false, 4, 16>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st);
  if (math != HG_MATH_EXACT && ws_warps) return launch_ws<float, float, false, 4, 8>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st);
  if (math == HG_MATH_EXACT && ws_warps && !exact_sync_env()) return launch_ws<float, float, true, 4, 8>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st);
  if (math == HG_MATH_EXACT)
    return RW == 8 ? launch_tma<float, float, true, 8>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st)
                   : launch_tma<float, float, true, 4>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st);
  return RW == 8 ? launch_tma<float, float, false, 8>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st)
                 : launch_tma<float, float, false, 4>(src, dst, xs, ys, planes, h, w, h1, w1, BW, BH, st);
}

}  // namespace hg
