// hg_conv_wgrad_umma.cu -- weight / bias gradient of the 7-tap hex convolution on tcgen05 (sm_100a).
//
//   gw[co,ci,k] += sum_{n,R,q} gy[n,co,R,q] * P[n,ci, R + ro[k], q + co[R&1][k]]     (P = padded x, hg_conv.cuh)
//   gb[co]      += sum_{n,R,q} gy[n,co,R,q]
//
// GEMM view per tap: D_k[co, ci] += A[co, pixels] * B_k[ci, pixels]^T with the PIXELS as the reduction
// dimension.  Both operands are kept in shared memory exactly like the forward kernel's input rows --
// [channels/8][pixel][8 channels], one 16-byte unit per pixel -- but are handed to the tensor core as
// MN-major matrices (channels = M/N, contiguous 8-channel units; pixels = K, 16 bytes apart), so a tap's
// column shift is again a plain 16-byte offset of the descriptor start address and one staged x row
// serves all 7 taps of the three gy rows that touch it.  The 7 accumulators (7 x Cin fp32 columns) plus a
// 16-column bias accumulator (B = a block of ones) live in TMEM for the whole kernel; each persistent CTA
// reduces its share of (image, row band, column tile) items and adds its partial to gw / gb with one
// round of fp32 atomics at the end.  UMMA M is 64 for Cout <= 64 (the kernel is shared-memory-bandwidth bound and an
// M = 128 view would read the neighbouring ring slots as its upper half: 4 KB of A per MMA instead of 2 KB) and
// 128 for 64 < Cout <= 128; rows past Cout hold whatever the neighbouring ring memory holds and are never read.
//
// CTA = 15 warps: warps 0-11 converters (x rows and gy rows -> bf16 -> rings; warps 8-11 also run the final
// epilogue), warps 12 and 14 MMA issuers (taps 0-3 / taps 4-6 + bias: disjoint accumulators, so the two
// instruction streams need no ordering), warp 12 also allocates TMEM, warp 13 TMA producer.  Input path as in hg_conv_umma.cu:
//   TMA : 4-D boxes [C][1 row][px] of x and gy land in a shared raw staging ring (zero-fill of halos and of the
//         columns past Wo for free); needs pad_value == 0 and 16-byte aligned rows of both tensors.
//   LDG : coalesced global loads (any pad value / width); on lattices narrower than one tile the converter warps split into an
//         x group and a gy group so that both rows are in flight together.
//   CPA : cp.async rows into the raw ring for wide lattices whose rows TMA cannot address (float32, zero frame).
// Lattices narrower than a 128-pixel tile stage and reduce only ceil(Wo / 16) 16-pixel steps (template parameter FULL = whole
// tiles).  The final epilogue transposes the accumulators through the (then dead) ring memory so that every atomic adds 32
// consecutive floats of gw, and every CTA starts at a different place of gw.
#include "hg_conv.cuh"
#include "hg_ptx.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace hg {

constexpr int kWuTile = 128;
constexpr int kWuPW = 144;
constexpr int kWuConv = 384;                 // converter threads (warps 0-11)
constexpr int kWuConvWarps = kWuConv / 32;
constexpr int kWuThreads = 480;               // 12 converter warps, MMA issuer A (12), TMA producer (13), MMA issuer B (14)
constexpr int kWuBand = 32;
constexpr int kWuMaxQ = 3;                   // ceil(8 * 144 / 384)
constexpr int kWuTaps = 7;
constexpr int kWuStagePitch = kWuTaps * 32 + 1;   // final epilogue: one gw row chunk ([32 ci][7 taps]) per accumulator row, + 1 word
constexpr int kWuStageBytes = 4 * 32 * kWuStagePitch * 4;   // four epilogue warps x 32 rows (reuses the ring memory)

struct WgParams {
  int N, Cin, Cout, H, W, Ho, Wo;
  int row0, col0;                // x row / col of (gy 0,0) for row slot 0 / shift 0
  int ra[kWuTaps];
  int sh[2][kWuTaps];
  int pad;
  float pad_value;
  int pad_mode;                  // 1 reflect / 2 replicate / 3 circular frame, resolved by the x loader (LDG variant)
  int xslots, gslots, bands, ctiles, has_bias;
  int band;                      // output rows per work item
  int rstages, raw_bytes;        // TMA / cp.async variants
  int epi_direct;
  int split;                     // narrow lattices, LDG variant: x rows and gy rows loaded by separate warp groups
  int xpitch, gpitch, ksteps;    // pixels of an x / gy row that are staged and 16-pixel reduction steps per row: a lattice narrower
                                 // than one 128-pixel tile stages and multiplies only ceil(Wo / 16) steps (gy is zero beyond Wo)
  int m64;                       // Cout <= 64: UMMA M = 64 (half the A-operand shared-memory reads of an M = 128 view)
  int cin_total, ci_off, cout_total, co_off;   // this launch covers x channels [ci_off, ci_off + Cin) and gy channels [co_off, co_off + Cout)
  long long items;
};

template <typename T> __device__ __forceinline__ float wu_ld(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ float wu_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

__device__ __forceinline__ uint32_t wu_pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// MN-major, no swizzle: element (mn, k) at start + (mn/8)*SBO + (mn%8)*2 + (k/8)*LBO + (k%8)*16   (bf16)
__device__ __forceinline__ uint64_t wu_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return ptx::umma_desc_kmajor_noswizzle(addr, lbo, sbo);   // same bit fields; the major-ness is in the instruction descriptor
}
__host__ __device__ constexpr uint32_t wu_idesc(int M, int N) {   // bf16 x bf16 -> fp32, A and B MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// SRC: 0 = LDG (loads straight into registers), 1 = TMA boxes, 2 = cp.async rows ("software TMA") into the raw ring
// FULL: rows of whole 128-pixel tiles -- staged widths, reduction steps and the task -> (channel group, pixel) map are
// compile-time constants (the run-time versions cost the C3 layer 3 %)
template <typename TX, typename TG, int SRC, bool FULL>
__global__ void __launch_bounds__(kWuThreads, 1)
hexconv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap gmap,
                          const TX* __restrict__ x, const TG* __restrict__ gy, float* __restrict__ gw, float* __restrict__ gb,
                          WgParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr bool TMA = SRC == 1, CPA = SRC == 2, RAW = SRC != 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xpitch = FULL ? kWuPW : P.xpitch, gpitch = FULL ? kWuTile : P.gpitch, ksteps = FULL ? kWuTile / 16 : P.ksteps;
  const int gslot_bytes = P.Cout * kWuTile * 2;         // [Cout/8][128 px][16 B]
  const int xslot_bytes = P.Cin * kWuPW * 2;            // [Cin/8][PW px][16 B]
  unsigned char* gring = smem;                          // gy ring first: its M = 128 view may run into the x ring
  unsigned char* xring = gring + P.gslots * gslot_bytes;
  unsigned char* raw = xring + P.xslots * xslot_bytes;  // [rstage] raw rows (TMA variant)
  const int front = P.gslots * gslot_bytes + P.xslots * xslot_bytes + P.rstages * P.raw_bytes;
  unsigned char* ones = smem + max(front, kWuStageBytes);        // 512 B of bf16 1.0 (bias accumulator operand); the front
                                                                 // region doubles as the final epilogue's staging space
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 512);
  uint64_t* xfull = bars;                   // [xslots]   converters -> MMA   (one arrival per converter warp)
  uint64_t* xempty = xfull + P.xslots;      // [xslots]   MMA commit -> converters
  uint64_t* gfull = xempty + P.xslots;      // [gslots]
  uint64_t* gempty = gfull + P.gslots;      // [gslots]
  uint64_t* rfull = gempty + P.gslots;      // [rstages]  TMA bytes landed
  uint64_t* rempty = rfull + P.rstages;     // [rstages]  converters done
  uint64_t* done = rempty + P.rstages;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  for (int e = tid; e < 256; e += kWuThreads) reinterpret_cast<__nv_bfloat16*>(ones)[e] = __float2bfloat16_rn(1.f);
  if (tid == 0) {
    const int fill_warps = (!FULL && !RAW && P.split) ? kWuConvWarps / 2 : kWuConvWarps;     // arrivals per filled slot
    for (int s = 0; s < P.xslots; ++s) { ptx::mbar_init(&xfull[s], fill_warps); ptx::mbar_init(&xempty[s], 2); }   // one commit per issuer
    for (int s = 0; s < P.gslots; ++s) { ptx::mbar_init(&gfull[s], fill_warps); ptx::mbar_init(&gempty[s], 2); }
    for (int s = 0; s < P.rstages; ++s) { ptx::mbar_init(&rfull[s], CPA ? kWuConv : 1); ptx::mbar_init(&rempty[s], kWuConvWarps); }
    ptx::mbar_init(done, 2);
    if (TMA) { ptx::prefetch_tensormap(&xmap); ptx::prefetch_tensormap(&gmap); }
    ptx::fence_barrier_init();
  }
  if (warp == 12) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the issue loop
  const int per_n = P.bands * P.ctiles;
  const int bias_col = kWuTaps * P.Cin;

  if (warp < kWuConvWarps) {
    // ===== converters ========================================================================================
    const size_t xplane = (size_t)P.H * P.W, gplane = (size_t)P.Ho * P.Wo;
    const int xtasks = (P.Cin >> 3) * xpitch, gtasks = (P.Cout >> 3) * gpitch;
    uint32_t xs = 0, xph = 0, gs = 0, gph = 0, rs = 0, rph = 0;      // ring positions / parities
    auto publish = [&](uint64_t* bar) {     // all of this warp's writes fenced, then one arrival per warp
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar);
    };
    // The stage may be refilled by TMA (async proxy) as soon as rempty completes.  The refill of a stage that held a
    // gy row is an x row and vice versa, so the zero-fill of the new box's out-of-range columns (written without
    // any memory latency) lands on *valid* data of the old layout: the shared-memory reads above must have been
    // performed before the arrival.  The values are packed (register dependency on every LDS) and a
    // generic -> async proxy fence is issued first -- the CUTLASS consumer_release pattern for TMA-fed stages.
    // (Measured on B200 without it: intermittent wrong gw rows, only on the TMA path.)
    auto raw_done = [&]() {                 // this warp has the raw row in registers
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&rempty[rs]);
      if (++rs == (uint32_t)P.rstages) { rs = 0; rph ^= 1; }
    };
    auto pack8 = [&](const float (&v)[8]) {
      uint4 pk;
      pk.x = wu_pack(v[0], v[1]); pk.y = wu_pack(v[2], v[3]); pk.z = wu_pack(v[4], v[5]); pk.w = wu_pack(v[6], v[7]);
      return pk;
    };
    // ---- cp.async variant: every converter thread copies its share of the raw rows rstages - 1 rows ahead of the row
    // being converted.  The raw rows of an item come in the order x0, x1, (x[rr + 2], gy[rr]) for rr = 0 .. rows - 1.
    struct RawCur { long long item; int step, rows, r0, c0; const TX* xn; const TG* gn; };
    auto cur_set = [&](RawCur& c, long long item) {
      c.item = item; c.step = 0;
      if (item < P.items) {
        const int n = (int)(item / per_n);
        const int rem = (int)(item - (long long)n * per_n);
        const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
        c.r0 = band * P.band; c.rows = min(P.band, P.Ho - c.r0); c.c0 = ct * kWuTile;
        c.xn = x + ((size_t)n * P.cin_total + P.ci_off) * xplane;
        c.gn = gy + ((size_t)n * P.cout_total + P.co_off) * gplane;
      }
    };
    uint32_t prs = 0, prph = 0;              // producer side position in the raw ring
    auto cpa_issue = [&](RawCur& c) {
      ptx::mbar_wait(&rempty[prs], prph ^ 1);                 // every converter warp has the stage's old row in registers
      const uint32_t dst0 = ptx::smem_u32(raw + (size_t)prs * P.raw_bytes);
      const bool is_g = c.step >= 2 && ((c.step - 2) & 1);
      if (!is_g) {
        int i = c.r0 + P.row0 + (c.step < 2 ? c.step : 2 + ((c.step - 2) >> 1));
        const bool row_frame = i >= -P.pad && i < P.H + P.pad;
        if (P.pad_mode && row_frame) i = conv_pad_remap(i, P.H, P.pad_mode);
        const bool row_in = i >= 0 && i < P.H;
        for (int p = lane; p < xpitch; p += 32) {           // lanes along the row (coalesced), warps over the channels
          int j = c.c0 + P.col0 + p;
          const bool col_frame = j >= -P.pad && j < P.W + P.pad;
          if (P.pad_mode && col_frame) j = conv_pad_remap(j, P.W, P.pad_mode);
          const bool ok = row_in && j >= 0 && j < P.W;
          const TX* __restrict__ src = ok ? c.xn + (size_t)i * P.W + j : x;
          const int nch = ok ? min(P.Cin, P.cin_total - P.ci_off) : 0;       // channels that exist; the rest is zero fill
          for (int ch = warp; ch < P.Cin; ch += kWuConvWarps)
            ptx::cp_async_4(dst0 + (uint32_t)(ch * xpitch + p) * 4u, ch < nch ? src + (size_t)ch * xplane : x, ch < nch ? 4u : 0u);
        }
      } else {
        const int R = c.r0 + ((c.step - 2) >> 1);
        for (int p = lane; p < gpitch; p += 32) {
          const int cc = c.c0 + p;
          const bool ok = cc < P.Wo;
          const TG* __restrict__ src = ok ? c.gn + (size_t)R * P.Wo + cc : gy;
          for (int ch = warp; ch < P.Cout; ch += kWuConvWarps)
            ptx::cp_async_4(dst0 + (uint32_t)(ch * gpitch + p) * 4u, ok ? src + (size_t)ch * gplane : gy, ok ? 4u : 0u);
        }
      }
      ptx::cp_async_mbar_arrive_noinc(&rfull[prs]);
      if (++prs == (uint32_t)P.rstages) { prs = 0; prph ^= 1; }
      if (++c.step == 2 + 2 * c.rows) cur_set(c, c.item + gridDim.x);
    };
    RawCur pc;
    auto cpa_ahead = [&]() {
      if constexpr (CPA && sizeof(TX) == 4 && sizeof(TG) == 4) { if (pc.item < P.items) cpa_issue(pc); }
    };
    if constexpr (CPA && sizeof(TX) == 4 && sizeof(TG) == 4) {
      cur_set(pc, blockIdx.x);
      for (int d = 0; d < P.rstages - 1 && pc.item < P.items; ++d) cpa_issue(pc);
    }
    // the (channel group, pixel) of this thread's tasks never changes: narrow lattices work it out once (run-time widths),
    // full tiles divide by constants on the spot
    // Narrow lattices, LDG variant: the twelve converter warps split into an x group (warps 0-5) and a gy group (6-11),
    // so an x row and a gy row are in flight together (one after the other -- gy in two passes at 128 channels -- the
    // converters waited for global loads 40 % of the time on the C5 layers)
    const bool split = !FULL && !RAW && P.split;
    const bool grp_g = split && warp >= kWuConvWarps / 2;
    const int cthreads = split ? kWuConv / 2 : kWuConv, ctid = grp_g ? tid - kWuConv / 2 : tid;
    int xkc_[kWuMaxQ], xp_[kWuMaxQ], gkc_[2 * kWuMaxQ], gp_[2 * kWuMaxQ];      // < 0: no task
    if (!FULL) {
#pragma unroll
      for (int q = 0; q < kWuMaxQ; ++q) {
        const int task = ctid + q * cthreads;
        xkc_[q] = task / xpitch; xp_[q] = task - xkc_[q] * xpitch;
        if (task >= xtasks) xkc_[q] = -1;
      }
#pragma unroll
      for (int q = 0; q < 2 * kWuMaxQ; ++q) {
        const int task = ctid + q * cthreads;
        gkc_[q] = task / gpitch; gp_[q] = task - gkc_[q] * gpitch;
        if (task >= gtasks) gkc_[q] = -1;
      }
    }
    auto xkc = [&](int q) { return FULL ? (tid + q * kWuConv) / kWuPW : xkc_[q]; };
    auto xp = [&](int q) { return FULL ? (tid + q * kWuConv) % kWuPW : xp_[q]; };
    auto gkc = [&](int q) { return FULL ? (tid + q * kWuConv) / kWuTile : gkc_[q]; };
    auto gp = [&](int q) { return FULL ? (tid + q * kWuConv) % kWuTile : gp_[q]; };
    auto xok = [&](int q) { return FULL ? (tid + q * kWuConv) < xtasks : xkc_[q] >= 0; };
    auto gok = [&](int q) { return FULL ? (tid + q * kWuConv) < gtasks : gkc_[q] >= 0; };
    // LDG variant: what a task reads does not depend on the row -- per item, the pointer to the first of its 8 channels at
    // its (frame-remapped) column, or nullptr outside the image; a row adds i * W (see hg_conv_umma.cu)
    const TX* xsrc[kWuMaxQ];
    const TG* gsrc[2 * kWuMaxQ];
    uint32_t xframe = 0;                     // bit q: the x task's column lies in the pad_value frame
    auto item_tasks = [&](const TX* __restrict__ xn, const TG* __restrict__ gn, int c0) {
      xframe = 0;
#pragma unroll
      for (int q = 0; q < kWuMaxQ; ++q) {
        xsrc[q] = nullptr;
        if (xok(q)) {
          int j = c0 + P.col0 + xp(q);
          const bool col_frame = j >= -P.pad && j < P.W + P.pad;
          if (P.pad_mode && col_frame) j = conv_pad_remap(j, P.W, P.pad_mode);
          if (col_frame) xframe |= 1u << q;
          if (j >= 0 && j < P.W) xsrc[q] = xn + (size_t)(xkc(q) * 8) * xplane + j;
        }
      }
#pragma unroll
      for (int q = 0; q < 2 * kWuMaxQ; ++q) {
        gsrc[q] = nullptr;
        if (gok(q) && c0 + gp(q) < P.Wo) gsrc[q] = gn + (size_t)(gkc(q) * 8) * gplane + c0 + gp(q);
      }
    };
    auto load_x = [&](const TX* __restrict__ xn, int i, int c0) {    // i: frame row, remapped below for the non-constant modes
      unsigned char* sb = xring + (size_t)xs * xslot_bytes;
      uint4 pk[kWuMaxQ];
      if (RAW) {
        cpa_ahead();
        ptx::mbar_wait(&rfull[rs], rph);
        const TX* __restrict__ rp = reinterpret_cast<const TX*>(raw + (size_t)rs * P.raw_bytes);
#pragma unroll
        for (int q = 0; q < kWuMaxQ; ++q) {
          if (xok(q)) {
            const int kc = xkc(q), p = xp(q);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = to_f32(rp[(kc * 8 + e) * xpitch + p]);
            pk[q] = pack8(v);
          }
        }
        raw_done();
        ptx::mbar_wait(&xempty[xs], xph ^ 1);
      } else {
        ptx::mbar_wait(&xempty[xs], xph ^ 1);
        const bool row_frame = i >= -P.pad && i < P.H + P.pad;
        if (P.pad_mode && row_frame) i = conv_pad_remap(i, P.H, P.pad_mode);          // frame rows read the image in place
        const bool row_in = i >= 0 && i < P.H;
#pragma unroll
        for (int q = 0; q < kWuMaxQ; ++q) {
          if (xok(q)) {
            const float fill = (row_frame && ((xframe >> q) & 1u)) ? P.pad_value : 0.f;
            const bool in_img = row_in && xsrc[q] != nullptr;
            const TX* __restrict__ src = xsrc[q] + (size_t)i * P.W;
            const int c_left = P.cin_total - P.ci_off - xkc(q) * 8;  // channels of x that exist from this group on (RGB: 3 of 16)
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[e] = e < c_left ? (in_img ? wu_ld(src) : fill) : 0.f; src += xplane; }
            pk[q] = pack8(v);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kWuMaxQ; ++q) {
        if (xok(q)) *reinterpret_cast<uint4*>(sb + (size_t)(xkc(q) * kWuPW + xp(q)) * 16) = pk[q];
      }
      publish(&xfull[xs]);
      if (++xs == (uint32_t)P.xslots) { xs = 0; xph ^= 1; }
    };
    auto load_g = [&](const TG* __restrict__ gn, int R, int c0) {
      unsigned char* sb = gring + (size_t)gs * gslot_bytes;
      bool waited = false;
      if (RAW) cpa_ahead();
#pragma unroll
      for (int half = 0; half < 2; ++half) {                           // Cout = 128 needs two passes
        const int base = half * kWuMaxQ * cthreads;
        if (base >= gtasks) break;
        uint4 pk[kWuMaxQ];
        if (RAW) {
          if (base == 0) ptx::mbar_wait(&rfull[rs], rph);
          const TG* __restrict__ rp = reinterpret_cast<const TG*>(raw + (size_t)rs * P.raw_bytes);
#pragma unroll
          for (int q = 0; q < kWuMaxQ; ++q) {
            if (gok(half * kWuMaxQ + q)) {
              const int kc = gkc(half * kWuMaxQ + q), p = gp(half * kWuMaxQ + q);
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = to_f32(rp[(kc * 8 + e) * gpitch + p]);
              pk[q] = pack8(v);
            }
          }
          if (base + kWuMaxQ * cthreads >= gtasks) raw_done();
        } else {
#pragma unroll
          for (int q = 0; q < kWuMaxQ; ++q) {
            if (gok(half * kWuMaxQ + q)) {
              const TG* __restrict__ src = gsrc[half * kWuMaxQ + q] + (size_t)R * P.Wo;
              const bool in_img = gsrc[half * kWuMaxQ + q] != nullptr;
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { v[e] = in_img ? wu_ld(src) : 0.f; src += gplane; }
              pk[q] = pack8(v);
            }
          }
        }
        if (!waited) { ptx::mbar_wait(&gempty[gs], gph ^ 1); waited = true; }
#pragma unroll
        for (int q = 0; q < kWuMaxQ; ++q) {
          if (gok(half * kWuMaxQ + q)) *reinterpret_cast<uint4*>(sb + (size_t)(gkc(half * kWuMaxQ + q) * kWuTile + gp(half * kWuMaxQ + q)) * 16) = pk[q];
        }
      }
      publish(&gfull[gs]);
      if (++gs == (uint32_t)P.gslots) { gs = 0; gph ^= 1; }
    };
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int n = (int)(item / per_n);
      const int rem = (int)(item - (long long)n * per_n);
      const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
      const int r0 = band * P.band, rows = min(P.band, P.Ho - r0), c0 = ct * kWuTile;
      const TX* __restrict__ xn = x + ((size_t)n * P.cin_total + P.ci_off) * xplane;
      const TG* __restrict__ gn = gy + ((size_t)n * P.cout_total + P.co_off) * gplane;
      if (!RAW) item_tasks(xn, gn, c0);
      if (split) {
        // (LDG variant only.  With TMA the raw ring carries x and gy rows in ONE sequence: each group has to wait for and
        //  release the other group's stages as well -- a parity wait only works for a waiter that has observed every earlier
        //  phase; skipping blindly failed both ways on B200 -- and then the split gains nothing: 0.080 -> 0.086 ms on the
        //  C5 first layer, with more registers.)
        if (grp_g) { for (int rr = 0; rr < rows; ++rr) load_g(gn, r0 + rr, c0); }
        else { for (int t = 0; t < rows + 2; ++t) load_x(xn, r0 + P.row0 + t, c0); }
      } else {
        load_x(xn, r0 + P.row0 + 0, c0);
        load_x(xn, r0 + P.row0 + 1, c0);
        for (int rr = 0; rr < rows; ++rr) {
          load_x(xn, r0 + P.row0 + rr + 2, c0);
          load_g(gn, r0 + rr, c0);
        }
      }
    }
    if (warp >= 8) {
      // ===== final epilogue (warps 8-11): TMEM partials -> fp32 atomics ======================================
      // A TMEM lane is an output channel, so a warp-wide atomic straight from the registers would touch 32 rows of gw
      // (32 sectors per instruction; measured: ~230 cycles each, 50 us for the 448 of a 64 -> 128 layer -- the whole
      // kernel on a small lattice).  Each warp transposes its rows through shared memory instead (the rings are dead
      // once `done` has completed): a row of gw holds [ci][tap] contiguously, so the warp then adds 32 consecutive
      // floats per instruction (4 sectors).  Row pitch 7 * 32 + 1 words: conflict-free both ways.
      const int q4 = warp & 3;
      // accumulator row (= output channel) held by this thread's TMEM lane.  M = 128: row == lane.  M = 64
      // (cta_group::1): row r sits in lane 32 * (r / 16) + r % 16, i.e. 16 rows in the lower half of every
      // 32-lane quadrant (measured on B200 against the oracle: the "first 64 lanes" reading is wrong).
      const int rows_q = P.m64 ? 16 : 32;                       // accumulator rows per TMEM lane quadrant
      const int co0 = q4 * rows_q;
      float* stg = reinterpret_cast<float*>(smem) + (size_t)q4 * 32 * kWuStagePitch;
      ptx::mbar_wait(done, 0);
      ptx::tc_fence_after_sync();
      if (P.epi_direct) {                   // A/B switch (HG_WU_EPI_DIRECT=1): adds straight from the registers, 32 gw rows per instruction
        const int co = lane < rows_q ? co0 + lane : (1 << 30);
        for (int k = 0; k < kWuTaps; ++k)
          for (int cb = 0; cb < P.Cin; cb += 32) {
            uint32_t v[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(k * P.Cin + cb), v);
            ptx::tmem_ld_wait();
            if (co < P.Cout) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cb + j < P.Cin && P.ci_off + cb + j < P.cin_total)
                  atomicAdd(gw + ((size_t)(P.co_off + co) * P.cin_total + P.ci_off + cb + j) * kWuTaps + k, __uint_as_float(v[j]));
            }
          }
      } else
      for (int cb = 0; cb < P.Cin; cb += 32) {
        const int cwv = max(0, min(min(32, P.Cin - cb), P.cin_total - P.ci_off - cb));   // channels of this chunk that exist
        for (int k = 0; k < kWuTaps; ++k) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(k * P.Cin + cb), v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * kWuStagePitch + j * kWuTaps + k] = __uint_as_float(v[j]);
        }
        __syncwarp();
        const int run = kWuTaps * cwv;                          // contiguous floats of one gw row in this chunk
        // every CTA adds to the same gw: each CTA starts at a different 32-float unit of the block and wraps around, or all
        // of them would queue on the same L2 lines at the same time (measured at the C3 shape, every CTA walking the rows in
        // the same order: 0.79 -> 0.92 ms)
        const int nrows = min(rows_q, P.Cout - co0);
        const int upr = (run + 31) >> 5;                        // units per row
        const int units = nrows * upr;
        int c, m;                                               // row / unit in the row, advanced without divisions
        { const int u = ((int)blockIdx.x * 5) % units; c = u / upr; m = u - c * upr; }
        float* __restrict__ gbase = gw + ((size_t)(P.co_off + co0) * P.cin_total + P.ci_off + cb) * kWuTaps;
        const size_t grow = (size_t)P.cin_total * kWuTaps;
#pragma unroll 4
        for (int u0 = 0; u0 < units; ++u0) {
          const int t = m * 32 + lane;
          if (t < run) atomicAdd(gbase + c * grow + t, stg[c * kWuStagePitch + t]);
          if (++m == upr) { m = 0; if (++c == nrows) c = 0; }
        }
        __syncwarp();
      }
      if (P.has_bias) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)bias_col, v);
        ptx::tmem_ld_wait();
        const int co = lane < rows_q ? co0 + lane : (1 << 30);
        if (co < P.Cout) atomicAdd(gb + P.co_off + co, __uint_as_float(v[0]));
      }
    }
  } else if (warp == 12 || warp == 14) {
    // ===== MMA issuers =====================================================================================
    const int k_begin = warp == 12 ? 0 : 4, k_end = warp == 12 ? 4 : kWuTaps;
    const bool do_bias = warp == 14 && P.has_bias;
    const int umma_m = P.m64 ? 64 : 128;
    const uint32_t idesc = wu_idesc(umma_m, P.Cin), idesc_b = wu_idesc(umma_m, 16);
    const uint32_t g_addr = ptx::smem_u32(gring), x_addr = ptx::smem_u32(xring), o_addr = ptx::smem_u32(ones);
    const uint32_t sbo_g = kWuTile * 16, sbo_x = kWuPW * 16;
    // descriptor = constant high word (SBO, version) | low word (start >> 4, LBO = 128 B between 8-pixel groups)
    const uint32_t g_hi = (sbo_g >> 4) | (1u << 14), x_hi = (sbo_x >> 4) | (1u << 14), o_hi = (256u >> 4) | (1u << 14);
    const uint32_t lo_const = (128u >> 4) << 16;
    const uint64_t od = ((uint64_t)o_hi << 32) | ((o_addr >> 4) + lo_const);
    uint32_t slot0 = 0, phase0 = 0, gs = 0, gphase = 0;
    auto next_slot = [&](uint32_t& sl, uint32_t& ph) { if (++sl == (uint32_t)P.xslots) { sl = 0; ph ^= 1; } };
    uint32_t started = 0;
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int rem = (int)(item % per_n);
      const int band = rem / P.ctiles;
      const int r0 = band * P.band, rows = min(P.band, P.Ho - r0);
      for (int rr = 0; rr < rows; ++rr) {
        uint32_t s1 = slot0, p1 = phase0; next_slot(s1, p1);
        uint32_t s2 = s1, p2 = p1; next_slot(s2, p2);
        ptx::mbar_wait(&xfull[slot0], phase0);
        ptx::mbar_wait(&xfull[s1], p1);
        ptx::mbar_wait(&xfull[s2], p2);
        ptx::mbar_wait(&gfull[gs], gphase);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {                // single-thread region: descriptors travel R -> UR once per MMA
          const int par = (r0 + rr) & 1;
          const uint32_t a_lo0 = ((g_addr + gs * (uint32_t)gslot_bytes) >> 4) + lo_const;
          const uint32_t xb0 = (x_addr + slot0 * (uint32_t)xslot_bytes) >> 4;
          const uint32_t xb1 = (x_addr + s1 * (uint32_t)xslot_bytes) >> 4;
          const uint32_t xb2 = (x_addr + s2 * (uint32_t)xslot_bytes) >> 4;
#pragma unroll
          for (int k = 0; k < kWuTaps; ++k) {
            if (k < k_begin || k >= k_end) continue;             // warp-uniform
            const int ra = P.ra[k];
            uint32_t b_lo = (ra == 0 ? xb0 : (ra == 1 ? xb1 : xb2)) + (uint32_t)P.sh[par][k] + lo_const;
            uint32_t a_lo = a_lo0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(k * P.Cin);
            if (FULL) {                                       // full tiles: straight-line issue
#pragma unroll
              for (int j = 0; j < kWuTile / 16; ++j) {
                ptx::umma_bf16(d_tmem, ((uint64_t)g_hi << 32) | a_lo, ((uint64_t)x_hi << 32) | b_lo, idesc, started | (uint32_t)j);
                a_lo += 16; b_lo += 16;                       // 16 pixels = 256 bytes
              }
            } else {
              for (int j = 0; j < ksteps; ++j) {
                ptx::umma_bf16(d_tmem, ((uint64_t)g_hi << 32) | a_lo, ((uint64_t)x_hi << 32) | b_lo, idesc, started | (uint32_t)j);
                a_lo += 16; b_lo += 16;
              }
            }
          }
          if (do_bias) {
            uint32_t a_lo = a_lo0;
#pragma unroll 8
            for (int j = 0; j < ksteps; ++j) {
              ptx::umma_bf16(tmem_base + (uint32_t)bias_col, ((uint64_t)g_hi << 32) | a_lo, od, idesc_b, started | (uint32_t)j);
              a_lo += 16;
            }
          }
          started = 1;
          ptx::umma_commit(&gempty[gs]);
          ptx::umma_commit(&xempty[slot0]);
          if (rr == rows - 1) {
            ptx::umma_commit(&xempty[s1]);
            ptx::umma_commit(&xempty[s2]);
          }
        }
        __syncwarp();
        next_slot(slot0, phase0);
        if (++gs == (uint32_t)P.gslots) { gs = 0; gphase ^= 1; }
      }
      next_slot(slot0, phase0);
      next_slot(slot0, phase0);
    }
    if (ptx::elect_one()) ptx::umma_commit(done);
    __syncwarp();
  } else if (TMA) {
    // ===== TMA producer (warp 13, one lane): same row order as the converters =================================
    if (lane == 0) {
      uint32_t rs = 0, rph = 0;
      auto push = [&](const CUtensorMap* m, uint32_t bytes, int c, int r, int ch, int n) {
        ptx::mbar_wait(&rempty[rs], rph ^ 1);
        ptx::mbar_arrive_expect_tx(&rfull[rs], bytes);
        ptx::tma_load_4d(raw + (size_t)rs * P.raw_bytes, m, &rfull[rs], c, r, ch, n);
        if (++rs == (uint32_t)P.rstages) { rs = 0; rph ^= 1; }
      };
      const uint32_t xbytes = (uint32_t)(P.Cin * xpitch * (int)sizeof(TX)), gbytes = (uint32_t)(P.Cout * gpitch * (int)sizeof(TG));
      for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
        const int n = (int)(item / per_n);
        const int rem = (int)(item - (long long)n * per_n);
        const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
        const int r0 = band * P.band, rows = min(P.band, P.Ho - r0), c0 = ct * kWuTile;
        push(&xmap, xbytes, c0 + P.col0, r0 + P.row0 + 0, P.ci_off, n);
        push(&xmap, xbytes, c0 + P.col0, r0 + P.row0 + 1, P.ci_off, n);
        for (int rr = 0; rr < rows; ++rr) {
          push(&xmap, xbytes, c0 + P.col0, r0 + P.row0 + rr + 2, P.ci_off, n);
          push(&gmap, gbytes, c0, r0 + rr, P.co_off, n);
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side --------------------------------------------------------------------------------------------
static int g_wu_sms = 0, g_wu_smem_max = 0;
static bool g_wu_no_tma = [] { const char* e = getenv("HG_CONV_NO_TMA"); return e && e[0] == '1'; }();

static size_t wu_smem_bytes(int Cin, int Cout, int xslots, int gslots, int rstages, int raw_bytes) {
  const size_t front = (size_t)gslots * Cout * kWuTile * 2 + (size_t)xslots * Cin * kWuPW * 2 + (size_t)rstages * raw_bytes;
  return (front > (size_t)kWuStageBytes ? front : (size_t)kWuStageBytes) + 512 +
         (size_t)(2 * xslots + 2 * gslots + 2 * rstages + 1) * 8 + 16;
}

static bool wu_limits() {
  if (g_wu_smem_max == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_wu_smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&g_wu_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaGetLastError() != cudaSuccess || g_wu_smem_max <= 0) { g_wu_smem_max = 0; return false; }
  }
  return true;
}

// the M = 128 view of the last gy slot runs (16 - Cout/8) channel groups past it: that must stay inside the x ring
static bool wu_view_fits(int Cin, int Cout, int xslots) {
  return (size_t)(16 - Cout / 8) * kWuTile * 16 <= (size_t)xslots * Cin * kWuPW * 2;
}

// src: 0 LDG, 1 TMA, 2 cp.async; xpitch / gpitch: staged pixels per x / gy row
static bool wu_pick(int Cin, int Cout, int xes, int ges, int src, int xpitch, int gpitch, int& xslots, int& gslots, int& rstages, int& raw_bytes) {
  xslots = gslots = rstages = raw_bytes = 0;
  if (!wu_limits()) return false;
  if (src) {
    const int64_t xb = (int64_t)Cin * xpitch * xes, gb = (int64_t)Cout * gpitch * ges;
    raw_bytes = (int)ceil_div(xb > gb ? xb : gb, 128) * 128;
  }
  // small layers are bound by the latency of the rows in flight: up to six raw stages when they fit next to full rings
  for (int r = 6; src && r >= 4; --r)
    if (wu_smem_bytes(Cin, Cout, 6, 3, r, raw_bytes) <= (size_t)g_wu_smem_max && wu_view_fits(Cin, Cout, 6)) {
      xslots = 6; gslots = 3; rstages = r;
      return true;
    }
  for (int r = src ? 3 : 0; r >= (src ? 2 : 0); --r)
    for (int g = 3; g >= 2; --g)
      for (int xsl = 6; xsl >= 4; --xsl)
        if (wu_smem_bytes(Cin, Cout, xsl, g, r, raw_bytes) <= (size_t)g_wu_smem_max && wu_view_fits(Cin, Cout, xsl)) {
          xslots = xsl; gslots = g; rstages = r;
          return true;
        }
  raw_bytes = 0;
  return false;
}

bool conv_wgrad_umma_eligible(const hg_conv_desc* d) {
  if (d->radius != 2 || d->stride != 1 || d->dilation != 1 || d->groups != 1) return false;
  // one launch covers <= 64 input channels (7*Cin + 16 TMEM columns <= 512) x <= 128 output channels (UMMA M);
  // larger layers run as a grid of launches over channel slices (independent blocks of gw)
  // (input channels that are not a multiple of 16 -- an RGB first layer -- run with the slice rounded up: the loader
  //  supplies zeros for the channels that do not exist and the epilogue drops their gradients)
  if (d->Cin < 1 || d->Cin > 1024) return false;
  const int64_t cin16 = (d->Cin + 15) / 16 * 16;
  if (d->Cout % 8 != 0 || d->Cout < 8 || d->Cout > 1024) return false;
  int a, b, c, e;
  if (!wu_pick((int)(cin16 > 64 ? 64 : cin16), (int)(d->Cout > 128 ? 128 : d->Cout), 4, 4, 0, kWuPW, kWuTile, a, b, c, e)) return false;
  if (cin16 > 64 && cin16 % 64 != 0 && !wu_pick((int)(cin16 % 64), (int)(d->Cout > 128 ? 128 : d->Cout), 4, 4, 0, kWuPW, kWuTile, a, b, c, e)) return false;
  if (d->Cout > 128 && d->Cout % 128 != 0 && !wu_pick((int)(cin16 > 64 ? 64 : cin16), (int)(d->Cout % 128), 4, 4, 0, kWuPW, kWuTile, a, b, c, e)) return false;
  if (d->algo == 0 && (d->x_dtype != HG_BF16 || d->Cin * d->Cout < 32 * 32)) return false;
  return true;
}

template <typename TX, typename TG, int SRC, bool FULL>
static int launch_wu_full(const CUtensorMap& xmap, const CUtensorMap& gmap, const void* x, const void* gy, float* gw, float* gb,
                     const WgParams& P, cudaStream_t st) {
  const size_t smem = wu_smem_bytes(P.Cin, P.Cout, P.xslots, P.gslots, P.rstages, P.raw_bytes);
  auto kern = hexconv_wgrad_umma_kernel<TX, TG, SRC, FULL>;
  static SmemReservation reservation;
  cudaError_t e = reservation.reserve(kern, smem);
  if (e != cudaSuccess) { set_error("hexconv_wgrad_umma: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
  long long grid = g_wu_sms > 0 ? g_wu_sms : 148;
  if (grid > P.items) grid = P.items;
  kern<<<(unsigned)grid, kWuThreads, smem, st>>>(xmap, gmap, (const TX*)x, (const TG*)gy, gw, gb, P);
  return finish_launch(SRC == 1 ? "hexconv_wgrad_umma_tma" : SRC == 2 ? "hexconv_wgrad_umma_cpasync" : "hexconv_wgrad_umma");
}

template <typename TX, typename TG, int SRC>
static int launch_wu(const CUtensorMap& xmap, const CUtensorMap& gmap, const void* x, const void* gy, float* gw, float* gb,
                     const WgParams& P, cudaStream_t st) {
  return (P.ksteps == kWuTile / 16 && !P.split) ? launch_wu_full<TX, TG, SRC, true>(xmap, gmap, x, gy, gw, gb, P, st)
                                  : launch_wu_full<TX, TG, SRC, false>(xmap, gmap, x, gy, gw, gb, P, st);
}

template <typename T>
static bool wu_encode(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int Wd, int Hd, int Cd, int Nd, int box_w, int box_c) {
  constexpr int es = (int)sizeof(T);
  const cuuint64_t gdim[4] = {(cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Cd, (cuuint64_t)Nd};
  const cuuint64_t gstr[3] = {(cuuint64_t)Wd * es, (cuuint64_t)Wd * Hd * es, (cuuint64_t)Wd * Hd * Cd * es};
  const cuuint32_t box[4] = {(cuuint32_t)box_w, 1, (cuuint32_t)box_c, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  return enc(m, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// x rows and gy rows converted by separate warp groups (two rows in flight instead of one after the other): whenever a group
// of six warps covers a row's tasks -- narrow lattices, or few channels (an RGB first layer)
static int wu_can_split(const WgParams& P) {
  static const bool no_split = [] { const char* e = getenv("HG_WU_NO_SPLIT"); return e && e[0] == '1'; }();
  return (!no_split && (P.Cin / 8) * P.xpitch <= kWuMaxQ * (kWuConv / 2) && (P.Cout / 8) * P.gpitch <= 2 * kWuMaxQ * (kWuConv / 2)) ? 1 : 0;
}

static bool g_wu_no_cpa = [] { const char* e = getenv("HG_CONV_NO_CPASYNC"); return e && e[0] == '1'; }();

template <typename TX, typename TG>
static int launch_wu_any(const void* x, const void* gy, float* gw, float* gb, WgParams P, cudaStream_t st) {
  alignas(64) CUtensorMap xmap, gmap;
  memset(&xmap, 0, sizeof(xmap));
  memset(&gmap, 0, sizeof(gmap));
  constexpr int xes = (int)sizeof(TX), ges = (int)sizeof(TG), A = 16 / xes;
  // a lattice narrower than one tile: reduce over ceil(Wo / 16) 16-pixel steps only; gy is staged (zero beyond Wo) for exactly
  // those pixels and x for the same plus the tap shifts and the alignment slack -- every staged pixel is written every row
  P.ksteps = P.ctiles == 1 ? (int)ceil_div(P.Wo < kWuTile ? P.Wo : kWuTile, 16) : kWuTile / 16;
  P.gpitch = 16 * P.ksteps;
  P.xpitch = P.ksteps == kWuTile / 16 ? kWuPW : P.gpitch + 16;
  PFN_encodeTiled enc = get_encode_tiled();
  bool tma = !g_wu_no_tma && enc != nullptr && P.pad_value == 0.f && P.pad_mode == 0 &&      // (channels past cin_total -- RGB rounded up to 16 -- are TMA zero fill)
             ((int64_t)P.W * xes) % 16 == 0 && ((int64_t)P.Wo * ges) % 16 == 0 &&
             (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(gy) & 15) == 0;
  if (tma) {
    int xs_, gs_, rst, rb;
    tma = wu_pick(P.Cin, P.Cout, xes, ges, 1, P.xpitch, P.gpitch, xs_, gs_, rst, rb);
    if (tma) {
      const int col0a = (int)(floor((double)P.col0 / A)) * A, e0 = P.col0 - col0a;
      int smax = 0;
      for (int par = 0; par < 2; ++par) for (int k = 0; k < kWuTaps; ++k) smax = max(smax, P.sh[par][k] + e0);
      if (P.gpitch + smax > P.xpitch || !wu_encode<TX>(enc, &xmap, x, P.W, P.H, P.cin_total, P.N, P.xpitch, P.Cin) ||
          !wu_encode<TG>(enc, &gmap, gy, P.Wo, P.Ho, P.cout_total, P.N, P.gpitch, P.Cout))
        tma = false;
      else {
        P.col0 = col0a;
        for (int par = 0; par < 2; ++par) for (int k = 0; k < kWuTaps; ++k) P.sh[par][k] += e0;
        P.xslots = xs_; P.gslots = gs_; P.rstages = rst; P.raw_bytes = rb;
        P.split = 0;
        return launch_wu<TX, TG, 1>(xmap, gmap, x, gy, gw, gb, P, st);
      }
    }
  }
  if constexpr (xes == 4 && ges == 4) {
    if (!g_wu_no_cpa && P.ctiles > 1 && (P.pad_value == 0.f || P.pad_mode != 0)) {     // see hg_conv_umma.cu: wide rows only
      int rst, rb;
      if (wu_pick(P.Cin, P.Cout, xes, ges, 2, P.xpitch, P.gpitch, P.xslots, P.gslots, rst, rb)) {
        P.rstages = rst; P.raw_bytes = rb; P.split = 0;
        return launch_wu<TX, TG, 2>(xmap, gmap, x, gy, gw, gb, P, st);
      }
    }
  }
  int rst, rb;
  HG_REQUIRE(wu_pick(P.Cin, P.Cout, xes, ges, 0, P.xpitch, P.gpitch, P.xslots, P.gslots, rst, rb), HG_E_UNSUPPORTED, "hexconv_wgrad_umma: shared memory does not fit");
  P.rstages = 0; P.raw_bytes = 0;
  P.split = wu_can_split(P);
  return launch_wu<TX, TG, 0>(xmap, gmap, x, gy, gw, gb, P, st);
}

int conv_wgrad_umma(const hg_conv_desc* d, const ConvGeom& g, const ConvTaps& tp, const void* x, const void* gy, float* gw,
                    float* gbias, cudaStream_t st) {
  WgParams P{};
  P.N = g.N; P.Cin = g.Cin; P.Cout = g.Cout; P.H = g.H; P.W = g.W; P.Ho = g.Ho; P.Wo = g.Wo;
  int cmin = 1 << 30, cmax = -(1 << 30);
  for (int par = 0; par < 2; ++par)
    for (int k = 0; k < kWuTaps; ++k) { cmin = min(cmin, tp.co[par][k]); cmax = max(cmax, tp.co[par][k]); }
  HG_REQUIRE(tp.K == kWuTaps && cmax - cmin <= 3, HG_E_UNSUPPORTED, "hexconv_wgrad_umma: unexpected tap geometry");
  for (int k = 0; k < kWuTaps; ++k) {
    P.ra[k] = tp.ro[k];
    for (int par = 0; par < 2; ++par) P.sh[par][k] = tp.co[par][k] - cmin;
  }
  P.row0 = -g.pad; P.col0 = cmin - g.pad;
  P.pad = g.pad; P.pad_value = g.pad_value; P.pad_mode = g.pad ? g.pad_mode : 0;
  P.has_bias = gbias != nullptr;
  P.ctiles = (int)ceil_div(g.Wo, kWuTile);
  P.band = conv_pick_band(kWuBand, g.N, g.Ho, P.ctiles);
  P.bands = (int)ceil_div(g.Ho, P.band);
  P.items = (long long)g.N * P.bands * P.ctiles;
  const int xdt = d->x_dtype, gdt = d->y_dtype;
  P.cin_total = g.Cin; P.cout_total = g.Cout;
  for (int co0 = 0; co0 < g.Cout; co0 += 128) {
    for (int ci0 = 0; ci0 < g.Cin; ci0 += 64) {
      P.co_off = co0; P.Cout = g.Cout - co0 < 128 ? g.Cout - co0 : 128;
      P.ci_off = ci0; P.Cin = g.Cin - ci0 < 64 ? (g.Cin - ci0 + 15) / 16 * 16 : 64;      // rounded up: see the eligibility note
      P.has_bias = gbias != nullptr && ci0 == 0;
      // The kernel is shared-memory-bandwidth bound (per gy row: 64 MMAs x 6 KB of operand reads + ~175 KB of
      // staging / conversion traffic at 128 B/clk): with Cout <= 64 an M = 64 tile reads 2 KB of A per MMA
      // instead of 4 KB (half of an M = 128 view would be the neighbouring ring slots): 1.11 -> 0.81 ms on C3.
      static const bool m128_only = [] { const char* e = getenv("HG_WU_M128"); return e && e[0] == '1'; }();
      P.m64 = (P.Cout <= 64 && !m128_only) ? 1 : 0;
      static const bool epi_direct = [] { const char* e = getenv("HG_WU_EPI_DIRECT"); return e && e[0] == '1'; }();
      P.epi_direct = epi_direct ? 1 : 0;
      float* gb = P.has_bias ? gbias : nullptr;
      int rc;
      if (xdt == HG_F32 && gdt == HG_F32) rc = launch_wu_any<float, float>(x, gy, gw, gb, P, st);
      else if (xdt == HG_BF16 && gdt == HG_F32) rc = launch_wu_any<__nv_bfloat16, float>(x, gy, gw, gb, P, st);
      else if (xdt == HG_F32 && gdt == HG_BF16) rc = launch_wu_any<float, __nv_bfloat16>(x, gy, gw, gb, P, st);
      else if (xdt == HG_BF16 && gdt == HG_BF16) rc = launch_wu_any<__nv_bfloat16, __nv_bfloat16>(x, gy, gw, gb, P, st);
      else { set_error("hexconv_wgrad_umma: unsupported dtypes x=%d gy=%d", xdt, gdt); return HG_E_DTYPE; }
      if (rc) return rc;
    }
  }
  return HG_OK;
}

}  // namespace hg
