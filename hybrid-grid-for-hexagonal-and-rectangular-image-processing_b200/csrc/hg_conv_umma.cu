// hg_conv_umma.cu -- 7-tap hex convolution (radius 2, stride 1, dilation 1) as an implicit GEMM on the
// 5th-generation tensor cores: tcgen05.mma with the accumulator in TMEM (sm_100a).  Forward and data
// gradient share one kernel (the data gradient is the same 3-row stencil with transposed weights and
// mirrored tap offsets).  ref: HexFrames.py:96-169; geometry in hg_conv.cuh.
//
// GEMM view, per output row segment of 128 pixels:   D[128 px, Nout] = sum over 7 taps
//        A_tap[128 px, Cred] * B_tap[Cred, Nout]      (bf16 x bf16 -> fp32 in TMEM)
// A_tap is a *shifted view* of an input row kept in shared memory: input rows are stored K-major without
// swizzle as [Cred/8][PW pixels][8 channels] (one 16-byte core-matrix row per pixel), so that moving the
// UMMA descriptor start address by 16 bytes moves the view by exactly one pixel.  One input row is
// therefore fetched from HBM once, converted fp32 -> bf16 once, and serves all 7 taps of the three output
// rows that touch it.  NCHW activations are read with plain coalesced loads (lanes along the row) -- the
// [pixel][channel] transposition TMA cannot do happens in registers on the way to shared memory.
//
// Persistent CTA = 19 warps:  warps 0-7 loaders / converters (-> bf16 -> smem ring of input rows),
//                             warps 8-15 epilogue (TMEM -> registers -> bias/ReLU -> coalesced NCHW stores),
//                             warps 16 and 18 MMA issuers (one lane each, alternate output rows; 16 also allocates TMEM),
//                             warp 17 TMA producer (one lane; TMA variant only).
// Three variants of the input path:
//   TMA  : a 4-D tensor map over NCHW; one cp.async.bulk.tensor box [Cred][1 row][PW px] per input row lands
//          in a raw staging ring (deep hardware prefetch, zero-fill of halo rows / columns for free); the
//          converter warps read it conflict-free (lanes along pixels), pack 8 channels to 16 bytes, and
//          write the K-major ring.  Needs pad_value == 0 and 16-byte aligned rows.
//   LDG  : the same conversion straight from global memory with coalesced loads (any pad value / width).
//   CPA  : cp.async rows into the raw ring for wide lattices whose rows TMA cannot address (float32, zero frame).
// Pipelines: raw full/empty (TMA <-> converters), full/empty per ring slot (converters <-> MMA, slots released
//            by tcgen05.commit), tmem_full/tmem_empty per accumulator stage (MMA <-> epilogue).
#include "hg_conv.cuh"
#include "hg_ptx.cuh"
#include <algorithm>
#include <type_traits>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace hg {

constexpr int kUmTile = 128;                 // output pixels per MMA (UMMA M)
constexpr int kUmPW = 144;                   // pixels per ring slot: tile + tap shifts (<= 3) + alignment slack (<= 7)
constexpr int kUmLoaders = 256;
constexpr int kUmEpiWarps = 8;               // two warps per TMEM lane quadrant, alternating 32-channel chunks
constexpr int kUmMmaWarp = kUmLoaders / 32 + kUmEpiWarps;
constexpr int kUmTmaWarp = kUmMmaWarp + 1;
constexpr int kUmMmaWarpB = kUmMmaWarp + 2;    // second MMA issuer (alternate output rows; P.dual)
constexpr int kUmThreads = kUmLoaders + kUmEpiWarps * 32 + 32 + 32 + 32;
constexpr int kUmBand = 32;                  // output rows per work item (upper bound; small problems take shorter bands, umma_common)
constexpr int kUmMaxQ = 5;                   // ceil(8 * 144 / 256): (chunk, pixel) tasks per loader thread, Cred <= 64
constexpr int kTaps = 7;
constexpr int kUmMaxAcc = 8;                 // accumulator stages in TMEM (barrier arrays are sized for this)
constexpr int kUmMaxCred = 512;              // reduction channels per call (passes of <= 64)

struct UmmaParams {
  int N, Cred, Nout, Hi, Wi, Ho, Wo;
  int row0, col0;                // input row / col of (output 0,0) for row slot 0 / shift 0
  int ra[kTaps];                 // row slot 0..2 of each tap
  int sh[2][kTaps];              // column shift 0..3 of each tap, per output-row parity
  int pad;                       // frame of pad_value around the input; literal zero beyond
  float pad_value;
  int pad_mode;                  // 1 reflect / 2 replicate / 3 circular frame: the loader reads the image in place (LDG variant)
  int relu, has_bias, transpose_w;   // transpose_w: weights indexed [red][out] (dgrad)
  int cred_total, c_off, accumulate; // reduction channels > 64 run as passes of <= 64: this pass covers channels
                                     // [c_off, c_off + Cred) of cred_total and (accumulate) adds to the output of the previous pass
  int slots, bands, ctiles;
  int band;                      // output rows per work item
  int rstages, raw_bytes;        // TMA / cp.async variants: raw staging ring
  int nacc;                      // accumulator stages in TMEM (2 .. kUmMaxAcc)
  int dual;                      // two MMA-issuing warps take alternate output rows (small layers are paced by the issuing thread)
  int wpr;                       // LDG variant: loader warps per row (8, or fewer on narrow lattices: several rows in flight)
  int rpitch;                    // pixels of a row that are staged (<= kUmPW; narrow lattices stage only what their outputs read)
  long long items;
};

template <typename T> __device__ __forceinline__ float ld_in(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_in<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
template <typename T> __device__ __forceinline__ float ld_out(const T* p) { return *p; }
template <> __device__ __forceinline__ float ld_out<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_out(T* p, float v) { __stcs(p, v); }
template <> __device__ __forceinline__ void st_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// SRC: 0 = LDG (loads straight into registers), 1 = TMA boxes, 2 = cp.async rows ("software TMA") into the raw ring
template <typename TIN, typename TOUT, int SRC, bool ACC>   // ACC: add to the output of the previous channel-slice pass
__global__ void __launch_bounds__(kUmThreads, 1)
hexconv_umma_kernel(const __grid_constant__ CUtensorMap tmap, const TIN* __restrict__ in, const float* __restrict__ w,
                    const float* __restrict__ scale, const float* __restrict__ bias, TOUT* __restrict__ out, UmmaParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr bool TMA = SRC == 1, CPA = SRC == 2, RAW = SRC != 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkc = P.Cred >> 3;
  const int slot_bytes = P.Cred * P.rpitch * 2;            // ring slot: [Cred/8][rpitch][8] bf16
  const int wtap_bytes = P.Cred * P.Nout * 2;
  unsigned char* w_smem = smem;                                   // [tap][Cred/8][Nout][8] bf16
  unsigned char* ring = smem + kTaps * wtap_bytes;                // [slot][Cred/8][PW][8] bf16
  unsigned char* raw = ring + P.slots * slot_bytes;               // [rstage][Cred][PW] TIN (TMA variant)
  float* bias_s = reinterpret_cast<float*>(raw + (size_t)P.rstages * P.raw_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + ((P.Nout + 31) & ~31));
  uint64_t* full = bars;                    // [slots]  loaders -> MMA        (one arrival per loader warp)
  uint64_t* empty = bars + P.slots;         // [slots]  MMA commit -> loaders (count 1)
  uint64_t* tfull = empty + P.slots;        // [kUmMaxAcc] MMA commit -> epilogue   (P.nacc accumulator stages in use)
  uint64_t* tempty = tfull + kUmMaxAcc;     // [kUmMaxAcc] epilogue -> MMA          (count: all epilogue threads)
  uint64_t* rfull = tempty + kUmMaxAcc;     // [rstages] TMA bytes landed
  uint64_t* rempty = rfull + P.rstages;     // [rstages] converters done      (one arrival per warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + P.rstages);

  // ---- one-time setup ---------------------------------------------------------------------------------
  // weights: fp32 [Cout][Cin][1][7] in global -> bf16 UMMA B image in shared memory
  // A thread owns (reduction channel, output channel) pairs and moves the pair's 7 taps -- 28 contiguous bytes in global
  // memory -- with independent loads, four pairs in flight.  (The first version walked the image element by element:
  // one dependent L2 round trip and two integer divisions per element, 100 elements per thread at 64 -> 128 channels =
  // 55 us before the first MMA -- more than the whole layer on a small lattice; measured with one-image launches.)
  {
    const int per_tap = P.Cred * P.Nout;
    const int per_kc = P.Nout * 8;
    __nv_bfloat16* wimg = reinterpret_cast<__nv_bfloat16*>(w_smem);
#pragma unroll 4
    for (int r = tid; r < per_tap; r += kUmThreads) {          // r = offset of the pair inside a tap image: [kc][n][j]
      const int kc = r / per_kc, r2 = r - kc * per_kc;
      const int n = r2 >> 3, j = r2 & 7;
      const int red = P.c_off + kc * 8 + j;
      // (forward with Cin rounded up to 16 -- RGB first layers: the channels that do not exist carry zero weights)
      const bool real = red < P.cred_total;
      const float* __restrict__ wp = w + (P.transpose_w ? ((size_t)red * P.Nout + n)           // w[co=red][ci=n][k]
                                                        : ((size_t)n * P.cred_total + red)) * kTaps;   // w[co=n][ci=red][k]
      float v[kTaps];
#pragma unroll
      for (int k = 0; k < kTaps; ++k) v[k] = real ? __ldg(wp + k) : 0.f;
      // per-output-channel scale (BN-inference affine of HexConvModule) folded into the weight image: free at run time
      const float sc = scale ? __ldg(scale + n) : 1.f;
#pragma unroll
      for (int k = 0; k < kTaps; ++k) wimg[k * per_tap + r] = __float2bfloat16_rn(v[k] * sc);
    }
    for (int e = tid; e < ((P.Nout + 31) & ~31); e += kUmThreads) bias_s[e] = (P.has_bias && e < P.Nout) ? __ldg(bias + e) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < P.slots; ++s) { ptx::mbar_init(&full[s], CPA ? kUmLoaders / 32 : P.wpr); ptx::mbar_init(&empty[s], P.dual ? 2 : 1); }
    for (int s = 0; s < kUmMaxAcc; ++s) { ptx::mbar_init(&tfull[s], 1); ptx::mbar_init(&tempty[s], kUmEpiWarps * 32); }
    for (int s = 0; s < P.rstages; ++s) { ptx::mbar_init(&rfull[s], CPA ? kUmLoaders : 1); ptx::mbar_init(&rempty[s], kUmLoaders / 32); }
    if (TMA) ptx::prefetch_tensormap(&tmap);
    ptx::fence_barrier_init();
  }
  // two accumulator stages of Nout columns; the epilogue always loads 32-column chunks, so the last chunk of stage 1
  // may reach up to Nout + roundup32(Nout): the allocation covers that (Nout = 16: 64 columns, not 32)
  // P.nacc accumulator stages of Nout columns; the epilogue always loads 32-column chunks, so the last chunk of the last stage
  // may reach up to roundup32(Nout) past its start: the allocation covers that (Nout = 16, two stages: 64 columns, not 32)
  const uint32_t tmem_need = (uint32_t)(P.nacc - 1) * (uint32_t)P.Nout + (((uint32_t)P.Nout + 31u) & ~31u);
  const uint32_t tmem_cols = tmem_need <= 32 ? 32 : tmem_need <= 64 ? 64 : tmem_need <= 128 ? 128 : tmem_need <= 256 ? 256 : 512;
  if (warp == kUmMmaWarp) { ptx::tmem_alloc(tmem_slot, tmem_cols); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();            // weight image written by the generic proxy, read by tcgen05.mma
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the issue loop

  const int per_n = P.bands * P.ctiles;

  if (warp < 8) {
    // ===== loaders ========================================================================================
    // tasks = (8-channel group, pixel) pairs over the staged width P.rpitch; ring address of a task: [kc][p] 16-byte units
    const int ntasks = nkc * P.rpitch;
    const size_t plane = (size_t)P.Hi * P.Wi;
    // cp.async variant: this thread's own producer cursor runs rstages - 1 rows ahead of the row being converted
    struct RowCur { long long item; int t, rows, r0, c0; const TIN* in_n; };
    auto cur_set = [&](RowCur& c, long long item) {
      c.item = item; c.t = 0;
      if (item < P.items) {
        const int n = (int)(item / per_n);
        const int rem = (int)(item - (long long)n * per_n);
        const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
        c.r0 = band * P.band; c.rows = min(P.band, P.Ho - c.r0); c.c0 = ct * kUmTile;
        c.in_n = in + ((size_t)n * P.cred_total + P.c_off) * plane;
      }
    };
    auto cur_next = [&](RowCur& c) { if (++c.t == c.rows + 2) cur_set(c, c.item + gridDim.x); };
    uint32_t prs = 0, prph = 0;              // producer position in the raw ring / parity (advanced without divisions)
    auto cpa_issue = [&](const RowCur& c) {
      const int rs = (int)prs;
      ptx::mbar_wait(&rempty[rs], prph ^ 1);    // every converter warp has the old row in registers
      int i = c.r0 + P.row0 + c.t;
      const bool row_frame = i >= -P.pad && i < P.Hi + P.pad;
      if (P.pad_mode && row_frame) i = conv_pad_remap(i, P.Hi, P.pad_mode);
      const bool row_in = i >= 0 && i < P.Hi;
      const uint32_t dst0 = ptx::smem_u32(raw + (size_t)rs * P.raw_bytes);
      for (int p = lane; p < P.rpitch; p += 32) {             // lanes along the row (coalesced), warps over the channels
        int j = c.c0 + P.col0 + p;
        const bool col_frame = j >= -P.pad && j < P.Wi + P.pad;
        if (P.pad_mode && col_frame) j = conv_pad_remap(j, P.Wi, P.pad_mode);
        const bool ok = row_in && j >= 0 && j < P.Wi;
        const TIN* __restrict__ src = ok ? c.in_n + (size_t)i * P.Wi + j : in;
        const int nch = ok ? min(P.Cred, P.cred_total - P.c_off) : 0;        // channels that exist (RGB: 3 of 16); the rest is zero fill
        for (int ch = warp; ch < P.Cred; ch += kUmLoaders / 32)
          ptx::cp_async_4(dst0 + (uint32_t)(ch * P.rpitch + p) * 4u, ch < nch ? src + (size_t)ch * plane : in, ch < nch ? 4u : 0u);
      }
      ptx::cp_async_mbar_arrive_noinc(&rfull[rs]);
      if (++prs == (uint32_t)P.rstages) { prs = 0; prph ^= 1; }
    };
    RowCur pc;
    if (CPA) {
      cur_set(pc, blockIdx.x);
      for (int d = 0; d < P.rstages - 1 && pc.item < P.items; ++d) { cpa_issue(pc); cur_next(pc); }
    }
    // the (channel group, pixel) of this thread's tasks never changes: worked out once (a division by the run-time
    // width inside the row loop cost the C3 forward 10 %)
    // LDG variant on narrow lattices: a row needs only P.wpr of the eight loader warps, so the warps form 8 / wpr groups
    // that load consecutive rows at the same time (one row in flight per CTA made these layers latency-bound)
    const int wpr = CPA ? kUmLoaders / 32 : P.wpr, ngroups = (kUmLoaders / 32) / wpr;     // (the cp.async rows are copied by all eight warps)
    const int group = warp / wpr, gthreads = wpr * 32, tid_g = tid - group * gthreads;
    int tkc[kUmMaxQ], tp[kUmMaxQ];           // tkc < 0: no task
#pragma unroll
    for (int q = 0; q < kUmMaxQ; ++q) {
      const int task = tid_g + q * gthreads;
      tkc[q] = task / P.rpitch; tp[q] = task - tkc[q] * P.rpitch;
      if (task >= ntasks) tkc[q] = -1;
    }
    // ring positions advance by counters: the 64-bit % and / by run-time stage counts this loop started with were a fifth of
    // the kernel's instructions on a small layer, and in the serial chain of every row
    uint32_t lslot = 0, lph = 0, lrs = 0, lrph = 0, lgrp = 0;   // ring slot / parity, raw stage / parity, row's loader group
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int n = (int)(item / per_n);
      const int rem = (int)(item - (long long)n * per_n);
      const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
      const int r0 = band * P.band, rows = min(P.band, P.Ho - r0), c0 = ct * kUmTile;
      const TIN* __restrict__ in_n = in + ((size_t)n * P.cred_total + P.c_off) * plane;
      // LDG variant: what a task reads does not depend on the row -- column (frame columns remapped), frame flag and
      // the pointer to the first of its 8 channels are worked out per item; a row adds i * Wi.  (ncu on the C5 layers: a
      // third of all instructions were this address arithmetic, repeated per row, in warps that run one row at a time.)
      const TIN* tsrc[kUmMaxQ];              // nullptr: column outside the image
      uint32_t cframe = 0;                   // bit q: the task's column lies in the pad_value frame
      if (!RAW) {
#pragma unroll
        for (int q = 0; q < kUmMaxQ; ++q) {
          tsrc[q] = nullptr;
          if (tkc[q] >= 0) {
            int j = c0 + P.col0 + tp[q];
            const bool col_frame = j >= -P.pad && j < P.Wi + P.pad;
            if (P.pad_mode && col_frame) j = conv_pad_remap(j, P.Wi, P.pad_mode);
            if (col_frame) cframe |= 1u << q;
            if (j >= 0 && j < P.Wi) tsrc[q] = in_n + (size_t)(tkc[q] * 8) * plane + j;
          }
        }
      }
      for (int t = 0; t < rows + 2; ++t) {
        const int slot = (int)lslot, rs = (int)lrs;
        const uint32_t eph = lph ^ 1, rph = lrph;                // parities to wait for: slot empty, raw stage full
        const bool mine = (int)lgrp == group;
        if (++lslot == (uint32_t)P.slots) { lslot = 0; lph ^= 1; }
        if (RAW) { if (++lrs == (uint32_t)P.rstages) { lrs = 0; lrph ^= 1; } }
        if (++lgrp == (uint32_t)ngroups) lgrp = 0;
        if (!mine) {                                             // another group's row
          // TMA variant: wait for it and release it all the same.  A parity wait only works for a waiter that has observed
          // every earlier phase of the barrier, and a stage must not be recycled past a group that has not seen it yet
          // (the weight-gradient kernel's x / gy groups failed both ways before they did this).
          if (RAW) {
            ptx::mbar_wait(&rfull[rs], rph);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&rempty[rs]);
          }
          continue;
        }
        unsigned char* sb = ring + (size_t)slot * slot_bytes;
        uint4 pk[kUmMaxQ];                   // 8 channels of one pixel, packed to bf16 as soon as they are loaded
        auto pack8 = [](const float (&v)[8]) {
          uint4 r;
          r.x = pack_bf16(v[0], v[1]); r.y = pack_bf16(v[2], v[3]); r.z = pack_bf16(v[4], v[5]); r.w = pack_bf16(v[6], v[7]);
          return r;
        };
        if (RAW) {
          if (CPA) {
            if constexpr (sizeof(TIN) == 4) { if (pc.item < P.items) { cpa_issue(pc); cur_next(pc); } }
          }
          // raw row [Cred][rpitch] landed by TMA / cp.async (halo rows / columns already zero-filled) -> registers
          ptx::mbar_wait(&rfull[rs], rph);
          const TIN* __restrict__ rp = reinterpret_cast<const TIN*>(raw + (size_t)rs * P.raw_bytes);
#pragma unroll
          for (int q = 0; q < kUmMaxQ; ++q) {
            if (tkc[q] >= 0) {
              const int kc = tkc[q], p = tp[q];
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = to_f32(rp[(kc * 8 + e) * P.rpitch + p]);
              pk[q] = pack8(v);
            }
          }
          // generic-proxy reads of the stage -> TMA (async proxy) refill: proxy fence before the release (the
          // CUTLASS consumer_release pattern; see hg_conv_wgrad_umma.cu for the failure it prevents)
          ptx::fence_proxy_async_smem();
          __syncwarp();                      // the whole warp holds its values in registers:
          if (lane == 0) ptx::mbar_arrive(&rempty[rs]);   // one arrival per warp, the stage can be refilled
          ptx::mbar_wait(&empty[slot], eph);
        } else {
          ptx::mbar_wait(&empty[slot], eph);
          int i = r0 + P.row0 + t;
          const bool row_frame = i >= -P.pad && i < P.Hi + P.pad;
          if (P.pad_mode && row_frame) i = conv_pad_remap(i, P.Hi, P.pad_mode);      // frame rows read the image in place
          const bool row_in = i >= 0 && i < P.Hi;
#pragma unroll
          for (int q = 0; q < kUmMaxQ; ++q) {
            if (tkc[q] >= 0) {
              const float fill = (row_frame && ((cframe >> q) & 1u)) ? P.pad_value : 0.f;
              const bool in_img = row_in && tsrc[q] != nullptr;
              const TIN* __restrict__ src = tsrc[q] + (size_t)i * P.Wi;
              const int c_left = P.cred_total - P.c_off - tkc[q] * 8;  // channels that exist from this group on (RGB: 3 of 16)
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { v[e] = e < c_left ? (in_img ? ld_in(src) : fill) : 0.f; src += plane; }
              pk[q] = pack8(v);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < kUmMaxQ; ++q) {
          if (tkc[q] >= 0) *reinterpret_cast<uint4*>(sb + (size_t)(tkc[q] * P.rpitch + tp[q]) * 16) = pk[q];
        }
        ptx::fence_proxy_async_smem();       // generic-proxy smem writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&full[slot]);    // one arrival per converter warp
      }
    }
  } else if (warp < kUmMmaWarp) {
    // ===== epilogue ========================================================================================
    // ncu r1t: with 4 epilogue warps the MMA warp spent ~39 polls per row waiting for a free accumulator --
    // the epilogue (64 strided channel-plane stores per thread and row) paced the whole pipeline.  8 warps:
    // quadrant q4 is shared by warps q4 and q4 + 4, which take the even / odd 32-channel chunks.
    const int q4 = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - kUmLoaders / 32) >> 2;
    const int px = q4 * 32 + lane;
    const size_t cstride = (size_t)P.Ho * P.Wo;
    uint32_t acc = 0, acc_phase = 0;
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int n = (int)(item / per_n);
      const int rem = (int)(item - (long long)n * per_n);
      const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
      const int r0 = band * P.band, rows = min(P.band, P.Ho - r0), c = ct * kUmTile + px;
      TOUT* __restrict__ orow = out + (((size_t)n * P.Nout) * P.Ho + r0) * P.Wo + c;
      for (int rr = 0; rr < rows; ++rr, orow += P.Wo) {
        ptx::mbar_wait(&tfull[acc], acc_phase);
        ptx::tc_fence_after_sync();
        const int last_cb = ((P.Nout - 1) >> 5) << 5;                      // first channel of the last chunk
        int my_last = ((last_cb >> 5) & 1) == half ? last_cb : last_cb - 32;   // last chunk this warp owns (< 0: none)
        int cb0 = half * 32;
        if (P.Nout <= 32) {                  // a single chunk: the two warps of a quadrant take alternate rows (= accumulator stages),
          const bool own = (int)(acc & 1u) == half; //  so two rows are being stored at any time (C5 first layer: the stores paced the kernel)
          my_last = own ? 0 : -1;
          cb0 = own ? 0 : P.Nout;
        }
        if (my_last < 0) {                   // no chunk in this row: only release the accumulator
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(&tempty[acc]);
        }
        for (int cb = cb0; cb < P.Nout; cb += 64) {
          TOUT* __restrict__ op = orow + (size_t)cb * cstride;
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * (uint32_t)P.Nout + (uint32_t)cb, v);
          // bias of this chunk in registers (8 vector LDS in flight together with the TMEM load) -- a scalar
          // LDS in front of every FADD serialises the epilogue on shared-memory latency (ncu r1i).  Accumulating
          // passes (ACC) carry no bias: the same registers take the 32 old output values instead, all loads in
          // flight at once (one load in front of every add made the pass 4x slower than a first pass, ncu r2h).
          float4 bb[8];
          if (ACC) {
            float* bo = reinterpret_cast<float*>(bb);
            const int nv_ = min(32, P.Nout - cb);
#pragma unroll
            for (int j = 0; j < 32; ++j) bo[j] = (c < P.Wo && j < nv_) ? ld_out(op + (size_t)j * cstride) : 0.f;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) bb[i] = *reinterpret_cast<const float4*>(bias_s + cb + 4 * i);
          }
          ptx::tmem_ld_wait();
          const bool last = cb == my_last;
          if (last) {                        // this warp's part of the accumulator is in registers: hand it back
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(&tempty[acc]);
          }
          if (c < P.Wo) {
            const int nvalid = min(32, P.Nout - cb);      // warp-uniform; 32 except for a trailing 16-channel chunk
            const float* bf = reinterpret_cast<const float*>(bb);
            if (nvalid == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float f = __uint_as_float(v[j]) + bf[j];
                if (P.relu) f = fmaxf(f, 0.f);
                st_out(op, f);
                op += cstride;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float f = __uint_as_float(v[j]) + bf[j];
                if (P.relu) f = fmaxf(f, 0.f);
                st_out(op, f);
                op += cstride;
              }
            }
          }
        }
        if (++acc == (uint32_t)P.nacc) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == kUmMmaWarp || (warp == kUmMmaWarpB && P.dual)) {
    // ===== MMA issuer(s) ======================================================================================
    // Single-lane issue loop: everything that does not change per instruction is hoisted -- ring slots and
    // mbarrier parities advance incrementally (no 64-bit division), descriptors are built from a constant
    // high word and a low word that only gets an address increment added.
    const uint32_t idesc = ptx::umma_idesc_bf16(kUmTile, P.Nout);
    const uint32_t ring_addr = ptx::smem_u32(ring), w_addr = ptx::smem_u32(w_smem);
    const uint32_t lbo_a = (uint32_t)P.rpitch * 16, lbo_b = (uint32_t)P.Nout * 16;
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
    const uint32_t a_lo_const = ((lbo_a >> 4) << 16), b_lo_const = ((lbo_b >> 4) << 16);
    const uint32_t a_step = (2u * lbo_a) >> 4, b_step = (2u * lbo_b) >> 4;  // one K = 16 step, in 16-byte units
    const int ksteps = P.Cred >> 4;
    uint32_t slot0 = 0, phase0 = 0;          // ring slot / parity of the current output row's first input row
    uint32_t acc = 0, acc_phase = 0;         // accumulator stage / parity
    // Two issuers (P.dual): consecutive output rows alternate between them, each row into its own accumulator stage.  A ring
    // slot is read by three consecutive output rows, i.e. by BOTH issuers, and a tcgen05.commit only tracks the MMAs of the
    // thread that issues it: empty[] counts two arrivals and each issuer releases a slot after ITS last row that reads it
    // (input slot t of a band is read by output rows t-2, t-1, t: the owner of row q releases slots q and q+1, and q+2 when it
    // has no later row in the band; a slot that only one issuer reads -- the band's first and last -- gets both arrivals from it).
    const uint32_t me = warp == kUmMmaWarp ? 0u : 1u;
    uint32_t rowc = 0;                       // rows seen so far (parity = the row's issuer when dual)
    auto next_slot = [&](uint32_t& sl, uint32_t& ph) { if (++sl == (uint32_t)P.slots) { sl = 0; ph ^= 1; } };
    for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
      const int rem = (int)(item % per_n);
      const int band = rem / P.ctiles;
      const int r0 = band * P.band, rows = min(P.band, P.Ho - r0);
      for (int rr = 0; rr < rows; ++rr, ++rowc) {
        uint32_t s1 = slot0, p1 = phase0; next_slot(s1, p1);
        uint32_t s2 = s1, p2 = p1; next_slot(s2, p2);
        if (!P.dual || (rowc & 1u) == me) {
          ptx::mbar_wait(&full[slot0], phase0);
          ptx::mbar_wait(&full[s1], p1);
          ptx::mbar_wait(&full[s2], p2);
          ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
          ptx::tc_fence_after_sync();
          if (ptx::elect_one()) {                // single-thread region: descriptors travel R -> UR once per MMA
            const int par = (r0 + rr) & 1;
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)P.Nout;
            const uint32_t rb0 = (ring_addr + slot0 * (uint32_t)slot_bytes) >> 4;
            const uint32_t rb1 = (ring_addr + s1 * (uint32_t)slot_bytes) >> 4;
            const uint32_t rb2 = (ring_addr + s2 * (uint32_t)slot_bytes) >> 4;
            uint32_t b_lo = (w_addr >> 4) + b_lo_const;            // taps are contiguous: one running weight descriptor
            // the K loop of a tap is straight-line code for the usual channel counts (16 / 32 / 64 reduction channels): its
            // run-time trip count cost the issuing thread a loop dispatch per tap
            auto issue_row = [&](auto ks_tag) {
              constexpr int KS = decltype(ks_tag)::value;          // 0: run-time trip count
              uint32_t accum = 0;
#pragma unroll
              for (int k = 0; k < kTaps; ++k) {
                const int ra = P.ra[k];
                uint32_t a_lo = (ra == 0 ? rb0 : (ra == 1 ? rb1 : rb2)) + (uint32_t)P.sh[par][k] + a_lo_const;
                if constexpr (KS > 0) {
#pragma unroll
                  for (int j = 0; j < KS; ++j) {
                    ptx::umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, accum);
                    accum = 1;
                    a_lo += a_step; b_lo += b_step;
                  }
                } else {
                  for (int j = 0; j < ksteps; ++j) {
                    ptx::umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, accum);
                    accum = 1;
                    a_lo += a_step; b_lo += b_step;
                  }
                }
              }
            };
            if (ksteps == 4) issue_row(std::integral_constant<int, 4>{});
            else if (ksteps == 1) issue_row(std::integral_constant<int, 1>{});
            else if (ksteps == 2) issue_row(std::integral_constant<int, 2>{});
            else issue_row(std::integral_constant<int, 0>{});
            ptx::umma_commit(&tfull[acc]);                         // accumulator ready for the epilogue
            const bool last = rr == rows - 1;
            if (!P.dual) {
              ptx::umma_commit(&empty[slot0]);                     // input row rr is not needed any more
              if (last) {                                          // band done: release its two trailing rows too
                ptx::umma_commit(&empty[s1]);
                ptx::umma_commit(&empty[s2]);
              }
            } else {
              ptx::umma_commit(&empty[slot0]);                     // slot q: this issuer's last read
              if (rr == 0) ptx::umma_commit(&empty[slot0]);        //   (the band's first slot has no other reader)
              ptx::umma_commit(&empty[s1]);                        // slot q + 1: this issuer's next row (q + 2) does not read it
              if (rows == 1) ptx::umma_commit(&empty[s1]);
              if (rr + 2 > rows - 1) {                             // slot q + 2: no later row of this issuer in the band
                ptx::umma_commit(&empty[s2]);
                if (last) ptx::umma_commit(&empty[s2]);            //   (the band's last slot: read by the last row only)
              }
            }
          }
          __syncwarp();
        }
        next_slot(slot0, phase0);
        if (++acc == (uint32_t)P.nacc) { acc = 0; acc_phase ^= 1; }
      }
      next_slot(slot0, phase0);              // the band's two trailing input rows
      next_slot(slot0, phase0);
    }
  } else if (TMA && warp == kUmTmaWarp) {
    // ===== TMA producer (one lane) ===========================================================================
    if (lane == 0) {
      uint32_t trs = 0, tph = 0;
      for (long long item = blockIdx.x; item < P.items; item += gridDim.x) {
        const int n = (int)(item / per_n);
        const int rem = (int)(item - (long long)n * per_n);
        const int band = rem / P.ctiles, ct = rem - band * P.ctiles;
        const int r0 = band * P.band, rows = min(P.band, P.Ho - r0), c0 = ct * kUmTile;
        for (int t = 0; t < rows + 2; ++t) {
          const int rs = (int)trs;
          ptx::mbar_wait(&rempty[rs], tph ^ 1);
          if (++trs == (uint32_t)P.rstages) { trs = 0; tph ^= 1; }
          ptx::mbar_arrive_expect_tx(&rfull[rs], (uint32_t)(P.Cred * P.rpitch * (int)sizeof(TIN)));
          ptx::tma_load_4d(raw + (size_t)rs * P.raw_bytes, &tmap, &rfull[rs], c0 + P.col0, r0 + P.row0 + t, P.c_off, n);
        }
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------------------
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == kUmMmaWarp) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- host side --------------------------------------------------------------------------------------------
static int g_um_sms = 0, g_um_smem_max = 0;

static size_t umma_smem_bytes(int Cred, int Nout, int slots, int rstages, int raw_bytes, int rpitch) {
  // the UMMA A view is always 128 pixels (+ tap shift) long: with a slot pitch below that it runs past its channel group --
  // harmless, those accumulator rows are never stored -- and, for the last group of the last slot, past the ring: slack
  const size_t slack = rpitch < kUmPW ? (size_t)(kUmPW - rpitch) * 16 : 0;
  return slack + (size_t)kTaps * Cred * Nout * 2 + (size_t)slots * Cred * rpitch * 2 + (size_t)rstages * raw_bytes + (size_t)((Nout + 31) & ~31) * 4 +
         (size_t)(2 * slots + 2 * kUmMaxAcc + 2 * rstages) * 8 + 16;
}

static bool umma_device_limits() {
  if (g_um_smem_max == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_um_smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&g_um_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaGetLastError() != cudaSuccess || g_um_smem_max <= 0) { g_um_smem_max = 0; return false; }
  }
  return true;
}

// ring slots (and raw stages for the TMA / cp.async variants) that fit; 0 slots = does not fit.  src: 0 LDG, 1 TMA, 2 cp.async
static void umma_pick_stages(int Cred, int Nout, int in_elem, int src, int rpitch, int& slots, int& rstages, int& raw_bytes) {
  slots = rstages = raw_bytes = 0;
  if (!umma_device_limits()) return;
  if (src) {
    raw_bytes = (int)ceil_div((int64_t)Cred * rpitch * in_elem, 128) * 128;
    // as many ring slots as fit first (what the large layers had: 6 slots + 2 raw stages fill the shared memory at 64
    // channels), then as many raw stages as still fit, up to six: on small layers a TMA row is ~1 us of latency and two
    // stages in flight bounded the kernel at 0.5 us per row (C5 first layer)
    for (int sl = 6; sl >= 4 && !slots; --sl)
      for (int r = 6; r >= 2; --r)
        if (umma_smem_bytes(Cred, Nout, sl, r, raw_bytes, rpitch) <= (size_t)g_um_smem_max) { slots = sl; rstages = r; break; }
    if (!slots) raw_bytes = 0;
    return;
  }
  for (int sl = 6; sl >= 4; --sl)
    if (umma_smem_bytes(Cred, Nout, sl, 0, 0, rpitch) <= (size_t)g_um_smem_max) { slots = sl; return; }
}

// op: 0 forward (reduce over Cin), 1 dgrad (reduce over Cout), 2 wgrad (reduce over pixels)
bool conv_wgrad_umma_eligible(const hg_conv_desc* d);   // hg_conv_wgrad_umma.cu

bool conv_umma_eligible(const hg_conv_desc* d, int op) {
  if (op == 2) return conv_wgrad_umma_eligible(d);
  if (d->radius != 2 || d->stride != 1 || d->dilation != 1 || d->groups != 1) return false;
  int64_t Cred = op == 0 ? d->Cin : d->Cout;
  const int64_t Nout = op == 0 ? d->Cout : d->Cin;
  if (op == 0 && Cred >= 1) Cred = (Cred + 15) / 16 * 16;                  // forward: input channels are rounded up to 16 in the loader
  if (Cred % 16 != 0 || Cred < 16 || Cred > kUmMaxCred) return false;      // > 64: passes of <= 64 channels
  if (Nout % 16 != 0 || Nout < 16 || Nout > 256) return false;
  if (op == 1 && d->relu) return false;
  // auto: bf16 activations only (fp32 callers keep fp32 accuracy on the direct stencil) and only when the
  // channel contraction is dense enough to feed a 128 x Nout x Cred tile
  if (d->algo == 0 && ((op == 0 ? d->x_dtype : d->y_dtype) != HG_BF16 || Cred * Nout < 32 * 32)) return false;
  int slots, rst, rb;
  umma_pick_stages((int)(Cred > 64 ? 64 : Cred), (int)Nout, 4, 0, kUmPW, slots, rst, rb);
  return slots > 0;
}

static bool g_um_no_tma = [] { const char* e = getenv("HG_CONV_NO_TMA"); return e && e[0] == '1'; }();

// Loader warps per row: the fewest that still cover the row's tasks (kUmMaxQ per thread), so that the eight warps form
// several groups working on consecutive rows at once -- at most max_rows of them (ring slots / raw stages that can be in flight).
static void umma_pick_wpr(UmmaParams& P, int max_rows) {
  static const int wpr_env = [] { const char* e = getenv("HG_CONV_WPR"); return e ? atoi(e) : 0; }();
  P.wpr = kUmLoaders / 32;
  for (int wv = 2; wv < kUmLoaders / 32; wv *= 2)
    if ((P.Cred / 8) * P.rpitch <= kUmMaxQ * 32 * wv && (kUmLoaders / 32) / wv <= max_rows) { P.wpr = wv; break; }
  if (wpr_env == 1 || wpr_env == 2 || wpr_env == 4 || wpr_env == 8) P.wpr = wpr_env;      // A/B switch (the task bound still has to hold)
  if ((P.Cred / 8) * P.rpitch > kUmMaxQ * 32 * P.wpr) P.wpr = kUmLoaders / 32;
}

static bool g_um_no_cpa = [] { const char* e = getenv("HG_CONV_NO_CPASYNC"); return e && e[0] == '1'; }();

template <typename TIN, typename TOUT, int SRC, bool ACC>
static int launch_umma_acc(const CUtensorMap& tmap, const void* in, const float* w, const float* scale, const float* bias, void* out,
                           const UmmaParams& P, cudaStream_t st) {
  const size_t smem = umma_smem_bytes(P.Cred, P.Nout, P.slots, P.rstages, P.raw_bytes, P.rpitch);
  auto kern = hexconv_umma_kernel<TIN, TOUT, SRC, ACC>;
  static SmemReservation reservation;
  cudaError_t e = reservation.reserve(kern, smem);
  if (e != cudaSuccess) { set_error("hexconv_umma: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
  long long grid = g_um_sms > 0 ? g_um_sms : 148;
  if (grid > P.items) grid = P.items;
  kern<<<(unsigned)grid, kUmThreads, smem, st>>>(tmap, (const TIN*)in, w, scale, bias, (TOUT*)out, P);
  return finish_launch(SRC == 1 ? "hexconv_umma_tma" : SRC == 2 ? "hexconv_umma_cpasync" : "hexconv_umma");
}

template <typename TIN, typename TOUT, int SRC>
static int launch_umma(const CUtensorMap& tmap, const void* in, const float* w, const float* scale, const float* bias, void* out,
                       const UmmaParams& P, cudaStream_t st) {
  return P.accumulate ? launch_umma_acc<TIN, TOUT, SRC, true>(tmap, in, w, scale, bias, out, P, st)
                      : launch_umma_acc<TIN, TOUT, SRC, false>(tmap, in, w, scale, bias, out, P, st);
}

// Completes P (stage counts, column alignment, staged width) and launches the variant the input qualifies for: TMA boxes
// (16-byte aligned rows, zero frame), else cp.async rows (float32, zero frame or a frame read from the image), else plain loads.
template <typename TIN, typename TOUT>
static int launch_umma_any(const void* in, const float* w, const float* scale, const float* bias, void* out, UmmaParams P, cudaStream_t st) {
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  constexpr int es = (int)sizeof(TIN), A = 16 / es;
  // a lattice narrower than one 128-pixel tile stages only the pixels its outputs read (tile + tap shifts + alignment slack)
  P.rpitch = P.ctiles == 1 ? (int)std::min<int64_t>(kUmPW, ((int64_t)std::min(P.Wo, kUmTile) + 16 + 7) / 8 * 8) : kUmPW;
  PFN_encodeTiled enc = get_encode_tiled();
  bool tma = !g_um_no_tma && enc != nullptr && P.pad_value == 0.f && P.pad_mode == 0 &&      // (channels past cred_total -- RGB rounded up to 16 -- are TMA zero fill)
             ((int64_t)P.Wi * es) % 16 == 0 &&
             (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (tma) {
    int slots, rst, rb;
    umma_pick_stages(P.Cred, P.Nout, es, 1, P.rpitch, slots, rst, rb);
    tma = slots > 0;
    if (tma) {
      // 16-byte aligned box origin: move the ring origin left by e0 pixels and the tap views right by e0
      const int col0a = (int)(floor((double)P.col0 / A)) * A, e0 = P.col0 - col0a;
      int smax = 0;
      for (int par = 0; par < 2; ++par) for (int k = 0; k < kTaps; ++k) smax = max(smax, P.sh[par][k] + e0);
      if (std::min(P.Wo, kUmTile) + smax > P.rpitch) tma = false;
      else {
        const cuuint64_t gdim[4] = {(cuuint64_t)P.Wi, (cuuint64_t)P.Hi, (cuuint64_t)P.cred_total, (cuuint64_t)P.N};
        const cuuint64_t gstr[3] = {(cuuint64_t)P.Wi * es, (cuuint64_t)P.Wi * P.Hi * es, (cuuint64_t)P.Wi * P.Hi * P.cred_total * es};
        const cuuint32_t box[4] = {(cuuint32_t)P.rpitch, 1, (cuuint32_t)P.Cred, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        if (enc(&tmap, dt, 4, const_cast<void*>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
          tma = false;
        else {
          P.col0 = col0a;
          for (int par = 0; par < 2; ++par) for (int k = 0; k < kTaps; ++k) P.sh[par][k] += e0;
          P.slots = slots; P.rstages = rst; P.raw_bytes = rb;
          umma_pick_wpr(P, std::min(slots - 2, rst - 1));
          return launch_umma<TIN, TOUT, 1>(tmap, in, w, scale, bias, out, P, st);
        }
      }
    }
  }
  if constexpr (es == 4) {
    // measured: rows of several tiles gain from the deeper prefetch (C3 at width 255: 2.48 -> 1.89 ms); on a lattice narrower
    // than one tile the plain loads of the few staged pixels are quicker than the extra hop through the raw ring
    if (!g_um_no_cpa && P.ctiles > 1 && (P.pad_value == 0.f || P.pad_mode != 0)) {
      int slots, rst, rb;
      umma_pick_stages(P.Cred, P.Nout, es, 2, P.rpitch, slots, rst, rb);
      if (slots > 0) {
        P.slots = slots; P.rstages = rst; P.raw_bytes = rb; P.wpr = kUmLoaders / 32;
        return launch_umma<TIN, TOUT, 2>(tmap, in, w, scale, bias, out, P, st);
      }
    }
  }
  int rst, rb;
  if (P.ctiles == 1) P.rpitch = (int)std::min<int64_t>(kUmPW, ((int64_t)std::min(P.Wo, kUmTile) + 3 + 7) / 8 * 8);   // no alignment slack needed
  umma_pick_stages(P.Cred, P.Nout, es, 0, P.rpitch, P.slots, rst, rb);
  P.rstages = 0; P.raw_bytes = 0;
  HG_REQUIRE(P.slots > 0, HG_E_UNSUPPORTED, "hexconv_umma: shared memory does not fit");
  umma_pick_wpr(P, P.slots - 2);
  return launch_umma<TIN, TOUT, 0>(tmap, in, w, scale, bias, out, P, st);
}

static int dispatch_umma_pass(int in_dt, int out_dt, const void* in, const float* w, const float* scale, const float* bias, void* out,
                              const UmmaParams& P, cudaStream_t st) {
  if (in_dt == HG_F32 && out_dt == HG_F32) return launch_umma_any<float, float>(in, w, scale, bias, out, P, st);
  if (in_dt == HG_BF16 && out_dt == HG_F32) return launch_umma_any<__nv_bfloat16, float>(in, w, scale, bias, out, P, st);
  if (in_dt == HG_F32 && out_dt == HG_BF16) return launch_umma_any<float, __nv_bfloat16>(in, w, scale, bias, out, P, st);
  if (in_dt == HG_BF16 && out_dt == HG_BF16) return launch_umma_any<__nv_bfloat16, __nv_bfloat16>(in, w, scale, bias, out, P, st);
  set_error("hexconv_umma: unsupported dtypes in=%d out=%d", in_dt, out_dt);
  return HG_E_DTYPE;
}

// Reduction channels beyond 64 do not fit the shared-memory budget (7 weight taps + the input-row ring) of one
// CTA: they run as successive passes over channel slices of <= 64, every pass after the first adding to the
// output of the previous one in the epilogue (bias in the first pass, ReLU in the last).
static int dispatch_umma(int in_dt, int out_dt, const void* in, const float* w, const float* scale, const float* bias, void* out, UmmaParams P,
                         cudaStream_t st) {
  const int total = P.Cred, relu = P.relu, has_bias = P.has_bias, base_acc = P.accumulate;
  if (P.cred_total <= 0) P.cred_total = total;           // the forward sets the real channel count when it rounded Cred up
  // a narrow lattice has short ring slots: 128 reduction channels fit in one pass (no second launch, no read-back of the output)
  int step = 64;
  if (total > 64 && P.ctiles == 1) {
    const int rp = (int)std::min<int64_t>(kUmPW, ((int64_t)std::min(P.Wo, kUmTile) + 16 + 7) / 8 * 8);
    int sl, rst, rb;
    umma_pick_stages(128, P.Nout, 4, 1, rp, sl, rst, rb);
    if (sl > 0 && (128 / 8) * rp <= kUmMaxQ * kUmLoaders) step = 128;
  }
  for (int c0 = 0; c0 < total; c0 += step) {
    P.c_off = c0;
    P.Cred = total - c0 < step ? total - c0 : step;
    P.accumulate = c0 > 0 || base_acc;
    P.has_bias = has_bias && c0 == 0;
    P.relu = relu && c0 + step >= total;
    int rc = dispatch_umma_pass(in_dt, out_dt, in, w, scale, P.has_bias ? bias : nullptr, out, P, st);
    if (rc) return rc;
  }
  return HG_OK;
}

static void umma_common(UmmaParams& P, int Ho, int Wo, int N) {
  P.N = N; P.Ho = Ho; P.Wo = Wo;
  P.ctiles = (int)ceil_div(Wo, kUmTile);
  P.band = conv_pick_band(kUmBand, N, Ho, P.ctiles);
  P.bands = (int)ceil_div(Ho, P.band);
  P.items = (long long)N * P.bands * P.ctiles;
  // accumulator stages: two.  (HG_CONV_NACC = 4 / 8 takes more of the 512 TMEM columns; measured on the C5 and C3 layers it
  // changes nothing -- the issuer never waits for a free accumulator.  What paces a small layer is the single issuing thread:
  // with one tap's MMAs per row instead of seven the RGB layer ran in 46 instead of 78 us, ~190 cycles per tcgen05.mma of
  // address selection, descriptor moves and issue, against 16 cycles of tensor work; profiles/r5j.)
  static const int nacc_env = [] { const char* e = getenv("HG_CONV_NACC"); return e ? atoi(e) : 0; }();
  const int nround = (P.Nout + 31) & ~31;
  int nacc = 2;
  if (nacc_env >= 2 && nacc_env <= kUmMaxAcc && (nacc_env & 1) == 0 && (nacc_env - 1) * P.Nout + nround <= 512) nacc = nacc_env;
  P.nacc = nacc;
  static const int dual_env = [] { const char* e = getenv("HG_CONV_DUAL"); return e ? atoi(e) : -1; }();
  P.dual = dual_env >= 0 ? (dual_env ? 1 : 0) : 1;
  if (P.dual) P.nacc = 2;                    // a row's accumulator stage = its issuer
}

int conv_fwd_umma(const hg_conv_desc* d, const ConvGeom& g, const ConvTaps& tp, const void* x, const float* w, const float* scale,
                  const float* bias, void* y, cudaStream_t st) {
  UmmaParams P{};
  P.Cred = (g.Cin + 15) / 16 * 16; P.cred_total = g.Cin; P.Nout = g.Cout; P.Hi = g.H; P.Wi = g.W;
  int cmin = 1 << 30, cmax = -(1 << 30);
  for (int par = 0; par < 2; ++par)
    for (int k = 0; k < kTaps; ++k) { cmin = min(cmin, tp.co[par][k]); cmax = max(cmax, tp.co[par][k]); }
  HG_REQUIRE(tp.K == kTaps && cmax - cmin <= 3, HG_E_UNSUPPORTED, "hexconv_umma: unexpected tap geometry");
  for (int k = 0; k < kTaps; ++k) {
    P.ra[k] = tp.ro[k];
    for (int par = 0; par < 2; ++par) P.sh[par][k] = tp.co[par][k] - cmin;
  }
  P.row0 = -g.pad; P.col0 = cmin - g.pad;
  P.pad = g.pad; P.pad_value = g.pad_value; P.pad_mode = g.pad ? g.pad_mode : 0;
  P.relu = g.relu; P.has_bias = bias != nullptr; P.transpose_w = 0;
  umma_common(P, g.Ho, g.Wo, g.N);
  P.accumulate = d->accumulate ? 1 : 0;
  return dispatch_umma(d->x_dtype, d->y_dtype, x, w, scale, bias, y, P, st);
}

// gx[n,ci,i,j] = sum_{co,k} w[co,ci,k] * gy[n,co, i + pad - ro[k], j + pad - co[(i + pad - ro[k]) & 1][k]]   (zero outside)
int conv_dgrad_umma(const hg_conv_desc* d, const ConvGeom& g, const ConvTaps& tp, const void* gy, const float* w, void* gx,
                    cudaStream_t st) {
  UmmaParams P{};
  P.Cred = g.Cout; P.Nout = g.Cin; P.Hi = g.Ho; P.Wi = g.Wo;
  int cs[2][kTaps];
  int cmin = 1 << 30, cmax = -(1 << 30);
  HG_REQUIRE(tp.K == kTaps, HG_E_UNSUPPORTED, "hexconv_umma: unexpected tap geometry");
  for (int ipar = 0; ipar < 2; ++ipar)
    for (int k = 0; k < kTaps; ++k) {
      const int rpar = (ipar + g.pad - tp.ro[k]) & 1;           // parity of the gy row this tap reads
      cs[ipar][k] = g.pad - tp.co[rpar][k];
      cmin = min(cmin, cs[ipar][k]); cmax = max(cmax, cs[ipar][k]);
    }
  HG_REQUIRE(cmax - cmin <= 3, HG_E_UNSUPPORTED, "hexconv_umma: unexpected tap geometry");
  for (int k = 0; k < kTaps; ++k) {
    P.ra[k] = 2 - tp.ro[k];
    for (int ipar = 0; ipar < 2; ++ipar) P.sh[ipar][k] = cs[ipar][k] - cmin;
  }
  P.row0 = g.pad - 2; P.col0 = cmin;
  P.pad = 0; P.pad_value = 0.f; P.pad_mode = 0;
  P.relu = 0; P.has_bias = 0; P.transpose_w = 1;
  umma_common(P, g.H, g.W, g.N);
  P.accumulate = d->accumulate ? 1 : 0;
  return dispatch_umma(d->y_dtype, d->x_dtype, gy, w, nullptr, nullptr, gx, P, st);
}

}  // namespace hg
