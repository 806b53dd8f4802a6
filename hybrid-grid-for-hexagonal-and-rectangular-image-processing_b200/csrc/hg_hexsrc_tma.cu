// hg_hexsrc_tma.cu -- hex -> rect / hex -> hex resampling ('linear') with TMA-staged source tiles (sm_100a).
//
// ref: geometry_np.py:191-356 hex_to_rect_resample, :520-681 hexresize, geometry_torch.py:191-358
// hex_to_square_resample -- they differ only in the 1-D coordinate tables.
//
// Same scheme as hg_resample_tma.cu (persistent CTAs, 3-D TMA box per item into a 3-stage shared ring, TMA
// zero-fill == the reference's zero-fill of out-of-range lattice points), with two differences:
//   * the per-sample geometry is not separable: the axial column j_ = (0.5*i_ + y_) + (w-0.5)/2 depends on
//     row and column together and must be truncated in float64 to stay bit-exact on the lattice indices
//     (geometry_np.py:276-298).  It is plane independent, so an item stages kG = 3 planes (an RGB image)
//     and every thread evaluates the cell / triangle / weights of a sample once and blends 3 planes;
//   * the three lattice points of a sample sit in two source rows at columns j_ax - (i+1)/2 (offset
//     storage, odd rows shifted right): relative to y_ + (w-0.5)/2 the source column moves by at most
//     [-1, +1.5], so the box origin depends only on the tile column.
// HG_MATH_FAST: float32 simplex weights (u > v ? (1-u, u-v, v) : (1-v, v-u, u)).
// HG_MATH_EXACT: the reference's own float64 sub-triangle-area weights and blend order (hg_hexgeom.cuh
// hex_locate), bit-identical to geometry_np when written as float64 -- ~250 float64 instructions per sample,
// evaluated once per item and shared by its three planes.
#include "hg_common.cuh"
#include "hg_hexgeom.cuh"
#include "hg_ptx.cuh"
#include <math.h>
#include <stdlib.h>

namespace hg {

constexpr int kHsTW = 128;
constexpr int kHsTH = 16;                // output rows per tile (2 or 1 per warp)
constexpr int kHsThreads = 256;
constexpr int kHsStages = 3;
constexpr int kHsG = 3;                  // planes per item

struct HsTables {
  double y[kHsTW];        // column coordinate
  double x[kHsTH];        // row coordinate (exact mode)
  double hrow[kHsTH];     // 0.5 * i_
  double u[kHsTH];        // i_f
  int in[kHsTH];          // i_n
  int nrows, ncols, row0, col0;
};

__device__ __forceinline__ int hs_col_origin(double y0, double cj) {
  // leftmost source column any sample of a tile starting at coordinate y0 can touch, minus slack, 16-byte aligned
  const int c = __double2int_rd(dadd(y0, cj)) - 2;
  return c & ~3;
}

template <typename TD, bool EXACT, int NWARPS>     // NWARPS = 8 (two output rows per warp) or 16 (one row per warp: twice
__global__ void __launch_bounds__(NWARPS * 32, 2)    //  the warps per SM for the latency-bound float64 variant)
hexsrc_linear_tma_kernel(const __grid_constant__ CUtensorMap tmap, TD* __restrict__ dst, const double* __restrict__ xs,
                         const double* __restrict__ ys, int h, int w, int h1, int w1, int planes, int tiles_x, int tiles_y,
                         long long total_items, long long items_per_cta, int BW, int BH, int stage_bytes, double ci, double cj,
                         double hx, double wy, int col_major, int R, int groups, int interleave) {
  using WT = typename std::conditional<EXACT, double, float>::type;
  constexpr int kHsRW = kHsTH / NWARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kHsStages * stage_bytes);
  HsTables* tabs = reinterpret_cast<HsTables*>(smem_raw + (size_t)kHsStages * stage_bytes + 64);

  const int npos = tiles_x * tiles_y;
  // Two ways to hand work to the persistent CTAs.  Contiguous: CTA b owns a consecutive range of items.  Interleaved: the
  // units (R plane groups of one tile position) are dealt round-robin, unit = block * npos + position, CTA b takes units
  // b, b + grid, ... -- at any moment the grid works on `grid` ADJACENT positions of the same plane groups, so the halo
  // rows / columns two neighbouring tiles share are requested within microseconds of each other and the second request hits
  // L2.  (Contiguous ranges with R shared plane groups put R * 3 planes * grid tiles = several hundred MB between the two
  // visits: every halo came from DRAM, 16.4 GB moved for 12.7 GB at R = 64, profiles/r3c_hexsrc_exact_R64_ncu_full.txt;
  // interleaved: 12.9 GB, profiles/r4b_hexsrc_exact_interleaved_ncu_full.txt.)
  const int full_blocks_ = groups / R, tail_ = groups - full_blocks_ * R;
  const long long n_units = (long long)npos * (full_blocks_ + (tail_ ? 1 : 0)), n_full_units = (long long)npos * full_blocks_;
  auto dealt = [&](long long n) { return n > (long long)blockIdx.x ? (n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ll; };
  const long long g_begin = interleave ? 0 : (long long)blockIdx.x * items_per_cta;
  const long long g_end = interleave ? dealt(n_full_units) * R + (dealt(n_units) - dealt(n_full_units)) * tail_
                                     : min(total_items, g_begin + items_per_cta);
  if (g_begin >= g_end) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int plane_elems = BW * BH;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmap);
    for (int s = 0; s < kHsStages; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }

  // plane-group major; inside a group the tiles run row-major.  (Column-major, which re-uses the 2-of-18-row halo
  // at once, was measured and rejected: C4 0.73 -> 0.71, C2 0.88 -> 0.69 -- DRAM page locality of consecutive row
  // segments matters more than the halo; HG_HEXSRC_ORDER=1 keeps it for A/B runs.)  A CTA's items are
  // consecutive, so positions advance by carries.
  struct Pos { int grp, tx, ty, blk, sub, pos; long long unit; };
  // Items run in blocks of R plane groups: inside a block the R groups of one tile position are consecutive, so a
  // thread evaluates the (plane independent) geometry of its samples once per position and re-uses it R times.
  const int full_blocks = groups / R, tail = groups - full_blocks * R;
  auto set_tile = [&](Pos& q) {
    if (col_major) { q.tx = q.pos / tiles_y; q.ty = q.pos - q.tx * tiles_y; }
    else { q.ty = q.pos / tiles_x; q.tx = q.pos - q.ty * tiles_x; }
  };
  auto decode = [&](long long g) {
    Pos q;
    if (interleave) {                        // only ever called with the CTA's first item
      q.unit = blockIdx.x;
      q.blk = (int)(q.unit / npos); q.pos = (int)(q.unit - (long long)q.blk * npos);
      q.sub = 0; q.grp = q.blk * R;
      set_tile(q);
      return q;
    }
    const long long per_block = (long long)R * npos;
    int rb, pos;
    if (g < (long long)full_blocks * per_block) {
      q.blk = (int)(g / per_block); rb = R;
      const int rem = (int)(g - (long long)q.blk * per_block);
      pos = rem / R; q.sub = rem - pos * R;
    } else {
      q.blk = full_blocks; rb = tail;
      const int rem = (int)(g - (long long)full_blocks * per_block);
      pos = rem / rb; q.sub = rem - pos * rb;
    }
    q.grp = q.blk * R + q.sub;
    q.pos = pos;
    set_tile(q);
    return q;
  };
  auto advance = [&](Pos& q) {
    const int rb = q.blk < full_blocks ? R : tail;
    if (++q.sub < rb) { ++q.grp; return; }
    q.sub = 0;
    if (interleave) {
      q.unit += gridDim.x;
      q.blk = (int)(q.unit / npos); q.pos = (int)(q.unit - (long long)q.blk * npos);
      set_tile(q);
      q.grp = q.blk * R;
      return;
    }
    bool wrapped = false;
    if (col_major) { if (++q.ty == tiles_y) { q.ty = 0; if (++q.tx == tiles_x) { q.tx = 0; wrapped = true; } } }
    else if (++q.tx == tiles_x) { q.tx = 0; if (++q.ty == tiles_y) { q.ty = 0; wrapped = true; } }
    if (wrapped) ++q.blk;
    q.grp = q.blk * R;
  };
  auto origin = [&](int tx, int ty, int& row0, int& col0) {
    row0 = trunc_i32(dadd(xs[ty * kHsTH], ci));
    col0 = hs_col_origin(ys[tx * kHsTW], cj);
  };
  auto issue = [&](const Pos& q, int s) {   // one thread
    int row0, col0;
    origin(q.tx, q.ty, row0, col0);
    ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(kHsG * plane_elems * 4));
    ptx::tma_load_3d(smem_raw + (size_t)s * stage_bytes, &tmap, &full[s], col0, row0, q.grp * kHsG);
  };
  auto build_tables = [&](const Pos& q, HsTables& T) {   // threads 0 .. kHsTW + kHsTH - 1
    int row0, col0;
    const int tx = q.tx, ty = q.ty;
    origin(tx, ty, row0, col0);
    const int t = threadIdx.x;
    if (t < kHsTW) {
      const int b = tx * kHsTW + t;
      T.y[t] = ys[b < w1 ? b : w1 - 1];
      if (t == 0) { T.ncols = min(kHsTW, w1 - tx * kHsTW); T.nrows = min(kHsTH, h1 - ty * kHsTH); T.row0 = row0; T.col0 = col0; }
    } else if (t < kHsTW + kHsTH) {
      const int r = t - kHsTW, a = min(ty * kHsTH + r, h1 - 1);
      T.x[r] = xs[a];
      const double i_ = dadd(xs[a], ci);                 // geometry_np.py:276
      const int in = trunc_i32(i_);
      T.in[r] = in;
      T.u[r] = dsub(i_, (double)(float)in);              // :284
      T.hrow[r] = dmul(0.5, i_);                         // first term of :277
    }
  };

  Pos cur = decode(g_begin), nxt = cur, iss = cur;
  advance(nxt);
  build_tables(cur, tabs[0]);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kHsStages && g_begin + s < g_end; ++s) { issue(iss, s); advance(iss); }
  }

  const int n_items = (int)(g_end - g_begin);
  int s = 0;
  uint32_t parity = 0;
  int o1[kHsRW][4], oB[kHsRW][4], o4[kHsRW][4];
  WT wa[kHsRW][4], wb[kHsRW][4], wc[kHsRW][4];
  for (int k = 0; k < n_items; ++k) {
    const HsTables& T = tabs[k & 1];
    if (k + 1 < n_items) build_tables(nxt, tabs[(k + 1) & 1]);
    const int grp = cur.grp, tx = cur.tx, ty = cur.ty;
    const int np = min(kHsG, planes - grp * kHsG);
    const int nrows = min(kHsRW, T.nrows - warp * kHsRW);
    const int row0 = T.row0, col0 = T.col0;

    // geometry of this thread's samples (plane independent): evaluated for the first group of a tile position
    if (k == 0 || cur.sub == 0) {
#pragma unroll
    for (int r = 0; r < kHsRW; ++r) {
      const int rl = warp * kHsRW + r;
      const int in = T.in[rl];
      const double u = T.u[rl], hrow = T.hrow[rl];
      const int roff = (in - row0) * BW - col0;
      const int kA = (in + 1) / 2, kB = (in + 2) / 2;   // true division then truncation (in >= 0 here)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int jn; bool f;
        if (EXACT) {   // the reference's own arithmetic, operation by operation
          HexSample<double> hs;
          hex_locate<double, true, false>(T.x[rl], T.y[lane + 32 * c], h, w, hx, wy, ci, cj, hs);
          jn = hs.j_n; f = hs.flag;
          wa[r][c] = (WT)hs.wgt[0]; wb[r][c] = (WT)hs.wgt[1]; wc[r][c] = (WT)hs.wgt[2];
        } else {
          const double j_ = dadd(dadd(hrow, T.y[lane + 32 * c]), cj);   // :277  (0.5*i_ + y_) + (w-0.5)*0.5
          jn = trunc_i32(j_);
          const double v = dsub(j_, (double)(float)jn);
          f = u > v;                                                    // :298 up_down_flag
          const float uf = (float)u, vf = (float)v;
          wa[r][c] = (WT)(f ? 1.f - uf : 1.f - vf);
          wb[r][c] = (WT)(f ? uf - vf : vf - uf);
          wc[r][c] = (WT)(f ? vf : uf);
        }
        const int p1 = roff + jn - kA;                                // P1 (i_n, j_n)
        const int p4 = roff + BW + jn - kB + 1;                       // P4 (i_n+1, j_n+1)
        o1[r][c] = p1;
        oB[r][c] = f ? p4 - 1 : p1 + 1;                               // P2 (i_n+1, j_n) or P3 (i_n, j_n+1)
        o4[r][c] = p4;
      }
    }
    }
    ptx::mbar_wait(&full[s], parity);
    const float* __restrict__ t = reinterpret_cast<const float*>(smem_raw + (size_t)s * stage_bytes);
    TD* __restrict__ dp = dst + ((size_t)grp * kHsG * h1 + (size_t)(ty * kHsTH + warp * kHsRW)) * w1 + (tx * kHsTW + lane);
    // Full tiles (all but the right / bottom edge): every load of a plane first, then the blends, then the stores -- straight-line
    // code the scheduler can interleave.  With the per-sample bounds test around each sample (the edge path below) the
    // compiler sinks load, conversion and blend into the predicated region and the 24 samples of a plane group run as 24
    // serial LDS -> F2F -> DMUL -> DADD -> DADD -> F2F -> STG chains (ncu r4b: stall_wait 3.0, short_scoreboard 2.3 per issue).
    const bool full_tile = nrows == kHsRW && T.ncols == kHsTW;
    if (full_tile) {
      for (int p = 0; p < np; ++p, t += plane_elems, dp += (size_t)h1 * w1) {
        float v1[kHsRW][4], vB[kHsRW][4], v4[kHsRW][4];
#pragma unroll
        for (int r = 0; r < kHsRW; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) { v1[r][c] = t[o1[r][c]]; vB[r][c] = t[oB[r][c]]; v4[r][c] = t[o4[r][c]]; }
#pragma unroll
        for (int r = 0; r < kHsRW; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            TD o;
            if (EXACT)
              o = (TD)dadd(dadd(dmul(wa[r][c], (double)v1[r][c]), dmul(wb[r][c], (double)vB[r][c])), dmul(wc[r][c], (double)v4[r][c]));
            else
              o = (TD)fmaf((float)wc[r][c], v4[r][c], fmaf((float)wb[r][c], vB[r][c], (float)wa[r][c] * v1[r][c]));
            st_stream(dp + (size_t)r * w1 + 32 * c, o);
          }
      }
    } else
    for (int p = 0; p < np; ++p, t += plane_elems, dp += (size_t)h1 * w1) {
#pragma unroll
      for (int r = 0; r < kHsRW; ++r) {
        if (r < nrows) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            TD o;
            if (EXACT)     // geometry_np.py:354  alpha * p1 + beta * p2 + gamma * p3, float64, no contraction
              o = (TD)dadd(dadd(dmul(wa[r][c], (double)t[o1[r][c]]), dmul(wb[r][c], (double)t[oB[r][c]])),
                           dmul(wc[r][c], (double)t[o4[r][c]]));
            else
              o = (TD)fmaf((float)wc[r][c], t[o4[r][c]], fmaf((float)wb[r][c], t[oB[r][c]], (float)wa[r][c] * t[o1[r][c]]));
            if (lane + 32 * c < T.ncols) st_stream(dp + (size_t)r * w1 + 32 * c, o);
          }
        }
      }
    }
    __syncthreads();                            // stage s fully read; next item's tables complete
    if (threadIdx.x == 0 && k + kHsStages < n_items) { issue(iss, s); advance(iss); }
    cur = nxt;
    advance(nxt);
    if (++s == kHsStages) { s = 0; parity ^= 1; }
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
static int g_hs_sms = 0;

// HG_OK launched, 1 not applicable (fall back to the direct gather), else an error code
int try_hexsrc_linear_tma(const void* src, void* dst, const double* xs, const double* ys, const double* host_xs, const double* host_ys,
                          int64_t planes, int64_t h, int64_t w, int64_t h1, int64_t w1, int sdt, int ddt, int math, cudaStream_t st) {
  if (!host_xs || !host_ys || sdt != HG_F32) return 1;
  if (!((ddt == HG_F32) || (ddt == HG_F64 && math == HG_MATH_EXACT))) return 1;
  if ((w * 4) % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 || planes >= (1ll << 31) || h1 < 1 || w1 < 1) return 1;
  static const bool off = [] { const char* e = getenv("HG_HEXSRC_NO_TMA"); return e && e[0] == '1'; }();
  if (off) return 1;
  const double ci = (h - 1) * 0.5, cj = (w - 0.5) * 0.5;
  // rows: monotone, non-negative cell rows; footprint of a tile = i_n(a0) .. i_n(a1) + 1
  int BH = 0;
  for (int64_t a0 = 0; a0 < h1; a0 += kHsTH) {
    const int64_t a1 = (a0 + kHsTH < h1 ? a0 + kHsTH : h1) - 1;
    int prev = (int)(host_xs[a0] + ci);
    if (host_xs[a0] + ci < 0.0) return 1;
    const int first = prev;
    for (int64_t a = a0 + 1; a <= a1; ++a) {
      const int cur = (int)(host_xs[a] + ci);
      if (cur < prev) return 1;
      prev = cur;
    }
    if (prev - first + 2 > BH) BH = prev - first + 2;
  }
  // columns: monotone; source columns of a tile lie in [floor(y0 + cj) - 2, floor(y1 + cj) + 3]
  int BW = 0;
  for (int64_t b0 = 0; b0 < w1; b0 += kHsTW) {
    const int64_t b1 = (b0 + kHsTW < w1 ? b0 + kHsTW : w1) - 1;
    for (int64_t b = b0 + 1; b <= b1; ++b)
      if (host_ys[b] < host_ys[b - 1]) return 1;
    const int lo = ((int)floor(host_ys[b0] + cj) - 2) & ~3;
    const int hi = (int)floor(host_ys[b1] + cj) + 3;
    if (hi - lo + 1 > BW) BW = hi - lo + 1;
  }
  BW = (BW + 3) / 4 * 4;
  if (BH > 256 || BW > 256) return 1;
  if ((int64_t)BH * BW > (int64_t)3 * kHsTH * kHsTW) return 1;   // strong down-sampling: the direct gather is already at the roofline
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return 1;
  alignas(64) CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
  const cuuint64_t gstr[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)kHsG};
  const cuuint32_t estr[3] = {1, 1, 1};
  // L2 promotion of the box rows (HG_HEXSRC_L2PROMO: 0 none, 1 64 B, 2 128 B, 3 256 B).  The exact variant at 64 shared plane
  // groups reads 10.0 GB for 6.4 GB of source (ncu, profiles/r3c_hexsrc_exact_R64_ncu_full.txt); the promotion width was the
  // suspect -- a 544-byte box row touches 5-6 128-byte lines -- but the sweep says it is not: 0.62-0.69 of the HBM copy rate for
  // every setting, float32 variants 0.81-0.84 (profiles/r3o_sweep_h2r_l2promo.jsonl).  128 B stays.
  const char* e_promo = getenv("HG_HEXSRC_L2PROMO");
  const int promo_sel = e_promo ? atoi(e_promo) : 2;
  const CUtensorMapL2promotion promo = promo_sel == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : promo_sel == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo_sel == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 1;
  const int stage_bytes = (int)ceil_div((int64_t)kHsG * BW * BH * 4, 128) * 128;
  const int smem = kHsStages * stage_bytes + 64 + 2 * (int)sizeof(HsTables);
  const char* e_warps = getenv("HG_HEXSRC_WARPS");
  const int nw16 = (e_warps ? atoi(e_warps) : 8) == 16 ? 1 : 0;   // 16 warps / CTA measured within +-2 % of 8 (profiles/r3d_sweep_kernels.jsonl)
  const int variant = (math == HG_MATH_FAST ? 0 : (ddt == HG_F32 ? 1 : 2)) + 3 * nw16;
  const int threads = nw16 ? 512 : 256;
  const void* kerns[6] = {(const void*)hexsrc_linear_tma_kernel<float, false, 8>, (const void*)hexsrc_linear_tma_kernel<float, true, 8>,
                          (const void*)hexsrc_linear_tma_kernel<double, true, 8>, (const void*)hexsrc_linear_tma_kernel<float, false, 16>,
                          (const void*)hexsrc_linear_tma_kernel<float, true, 16>, (const void*)hexsrc_linear_tma_kernel<double, true, 16>};
  const void* kern = kerns[variant];
  static SmemReservation reservation[6];
  if (reservation[variant].reserve(kern, (size_t)smem) != cudaSuccess) return 1;
  if (g_hs_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_hs_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return 1; }
  const int tiles_x = (int)ceil_div(w1, kHsTW), tiles_y = (int)ceil_div(h1, kHsTH);
  const long long groups = ceil_div(planes, kHsG);
  const long long total = (long long)tiles_x * tiles_y * groups;
  const char* e_dist = getenv("HG_HEXSRC_DIST");              // 0 contiguous item ranges per CTA, 1 (default) interleaved tile positions
  const int interleave = e_dist ? (atoi(e_dist) != 0) : 1;
  long long grid = (long long)g_hs_sms * occ;
  if (grid > total) grid = total;
  long long per = (total + grid - 1) / grid;
  if (!interleave) {
    grid = (total + per - 1) / per;
  }
  const double hx = (h - 1) / 2.0, wy = (w - 0.5) / 2.0;     // python-float expressions of geometry_np.py:326-331
  // A/B switches, read on every call so that one process can sweep them (tools/sweep_kernels.py)
  const char* e_order = getenv("HG_HEXSRC_ORDER");            // 0 row-, 1 column-major tile order
  const int order_env = e_order ? atoi(e_order) : -1;
  const int col_major = order_env > 0 ? 1 : 0;
  // plane groups that share one geometry evaluation (consecutive items of a CTA): more re-use of the float64 / index
  // arithmetic against a longer L2 re-use distance of the tile halos (measured)
  const char* e_share = getenv("HG_HEXSRC_SHARE");
  const int share_env = e_share ? atoi(e_share) : 0;
  // Measured on B200 (C4 / C2, fraction of the HBM copy rate; profiles/r3d_, r4a_, r4c_sweep*.jsonl).
  //  contiguous ranges: float32 weights R = 1 / 2 / 4: 0.84 / 0.81 / 0.77 and 0.85 / 0.82 / 0.78; float64 weights (exact)
  //    R = 8 / 16 / 64: 0.68 / 0.64 / 0.65 and 0.67 / 0.64 / 0.66 -- every halo re-read from DRAM (16.4 GB moved for 12.7 GB);
  //  interleaved units:  exact R = 8 / 16 / 64 / 256: 0.72 / 0.78 / 0.75 / 0.77 and 0.65 / 0.73 / 0.74 / 0.72 (12.9 GB moved).
  // Sharing amortises the ~250 float64 instructions of a sample's geometry; 32 groups are enough once the halos hit L2.
  const long long halo_footprint = (long long)tiles_x * stage_bytes * grid;        // bytes staged between two vertically adjacent tiles
  int R = share_env > 0 ? share_env
                        : (math == HG_MATH_EXACT ? (interleave ? 32 : 64) : (interleave ? 1 : (halo_footprint > (100ll << 20) ? 2 : 1)));
  if (R > (int)groups) R = (int)groups;
  if (interleave) {
    const long long n_units = (long long)tiles_x * tiles_y * ceil_div(groups, (long long)R);
    if (grid > n_units) grid = n_units;
  }
  void* args[] = {(void*)&tmap, (void*)&dst, (void*)&xs, (void*)&ys, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int hi = (int)h, wi = (int)w, h1i = (int)h1, w1i = (int)w1, pi = (int)planes, gi = (int)groups;
  long long total_ = total, per_ = per;
  int bw_ = BW, bh_ = BH, sb_ = stage_bytes, cm_ = col_major, r_ = R, il_ = interleave;
  double ci_ = ci, cj_ = cj, hx_ = hx, wy_ = wy;
  args[4] = &hi; args[5] = &wi; args[6] = &h1i; args[7] = &w1i; args[8] = &pi; args[9] = (void*)&tiles_x; args[10] = (void*)&tiles_y;
  args[11] = &total_; args[12] = &per_; args[13] = &bw_; args[14] = &bh_; args[15] = &sb_; args[16] = &ci_; args[17] = &cj_;
  args[18] = &hx_; args[19] = &wy_; args[20] = &cm_; args[21] = &r_; args[22] = &gi; args[23] = &il_;
  cudaLaunchKernel(kern, dim3((unsigned)grid), dim3((unsigned)threads), args, (size_t)smem, st);
  return finish_launch("hexsrc_linear_tma");
}

}  // namespace hg
