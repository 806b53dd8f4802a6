// hg_norm.cu -- batch normalisation (+ fused ReLU) of HexConvModule, forward and backward (sm_100a).  HBM-bound.
//
// ref: HexModules.py:146-288 HexConvModule (conv -> norm -> act; the norm layer is torch.nn.BatchNorm2d built by
// mmcv's build_norm_layer, HexModules.py:57-76).  The reference runs three library kernels per module and step
// (cuDNN batch-norm forward, ReLU, and their backward twins); measured on B200 for the C5 network, cuDNN's
// bn_fw_tr / bn_bw kernels run at 0.45 / 0.24 TB/s -- half of the whole training step.  Here:
//   forward : stats (one read of x)  -> apply (+ReLU) (one read of x, one write of y)
//   backward: reduce (reads x, dy)   -> apply (reads x, dy, writes dx)
// Every kernel streams whole float4 rows of one (image, channel) plane per CTA chunk; per-channel sums are
// accumulated in float inside a thread (16 values), in double across threads and CTAs (atomicAdd on double),
// so the statistics do not depend on the launch geometry beyond float64 rounding.
//   training:  mean = S1 / n ; var = S2 / n - mean^2 (biased, as torch normalises) ; rstd = 1 / sqrt(var + eps)
//   y  = act(x * (gamma * rstd) + (beta - mean * gamma * rstd))
//   dz = dy * [z > 0]  (ReLU fused; z recomputed from x, nothing but x is kept for backward)
//   dbeta = sum dz ; dgamma = sum dz * xhat ; dx = gamma * rstd * (dz - dbeta / n - xhat * dgamma / n)
#include "hg_common.cuh"

namespace hg {

constexpr int kBnThreads = 256;
constexpr int kBnChunk = 4096;      // elements of one plane per CTA

__device__ __forceinline__ void bn_block_sum2(float a, float b, double* dst) {   // dst[0] += sum a, dst[1] += sum b
  __shared__ double red[2][kBnThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = (double)a; red[1][warp] = (double)b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int w = 0; w < kBnThreads / 32; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
    atomicAdd(dst, s0);
    atomicAdd(dst + 1, s1);
  }
}

struct BnGeom {
  int C;
  long long HW;
  int chunks;          // CTAs per plane
  double inv_n;        // 1 / (N * HW)
};

// torch.nn.BatchNorm2d's training-mode bookkeeping, done by the kernels that have the numbers anyway (hg_bn_train_fwd):
// the stats kernel bumps num_batches_tracked, the apply kernel blends mean / unbiased variance into the running buffers.
struct BnRunning {
  float* mean;                 // running_mean [C] or NULL
  float* var;                  // running_var  [C]
  long long* batches;          // num_batches_tracked (int64 scalar) or NULL
  float momentum;              // < 0: cumulative moving average, factor = 1 / num_batches_tracked (after the increment)
  float unbias;                // n / (n - 1)
};

__device__ __forceinline__ void bn_locate(const BnGeom& g, long long& base, int& n, int& c) {
  const long long plane = blockIdx.x / g.chunks;
  const int ck = (int)(blockIdx.x - plane * g.chunks);
  c = (int)(plane % g.C);
  const long long off = (long long)ck * kBnChunk;
  base = plane * g.HW + off;
  n = (int)min((long long)kBnChunk, g.HW - off);
}

// sums[2c] += sum x, sums[2c+1] += sum x^2 over this CTA's chunk
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, BnGeom g, long long* __restrict__ batches) {
  long long base; int n, c;
  bn_locate(g, base, n, c);
  if (batches && blockIdx.x == 0 && threadIdx.x == 0) *batches += 1;      // read by the apply kernel that follows in the stream
  const float* __restrict__ p = x + base;
  float s = 0.f, ss = 0.f;
  if (VEC) {
    for (int i = threadIdx.x * 4; i < n; i += kBnThreads * 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
      s += (v.x + v.y) + (v.z + v.w);
      ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kBnThreads) { const float v = __ldg(p + i); s += v; ss += v * v; }
  }
  bn_block_sum2(s, ss, sums + 2 * c);
}

// per-channel scale / shift from the sums (training) or from given mean / var (inference), every CTA for itself
__device__ __forceinline__ void bn_channel(const double* sums, const float* mean_in, const float* var_in, const float* gamma,
                                           const float* beta, int c, double inv_n, float eps, float& mean, float& rstd,
                                           float& scale, float& shift) {
  double m, v;
  if (sums) { m = sums[2 * c] * inv_n; v = sums[2 * c + 1] * inv_n - m * m; if (v < 0.0) v = 0.0; }
  else { m = (double)mean_in[c]; v = (double)var_in[c]; }
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(v + (double)eps));
  // float from here on, in exactly the form the backward kernels repeat (same z, same ReLU mask)
  scale = (gamma ? gamma[c] : 1.f) * rstd;
  shift = fmaf(-mean, scale, beta ? beta[c] : 0.f);
}

template <bool VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const double* __restrict__ sums, const float* __restrict__ mean_in,
                const float* __restrict__ var_in, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ mean_out, float* __restrict__ var_out, float* __restrict__ rstd_out, BnGeom g, float eps, int relu,
                BnRunning run) {
  long long base; int n, c;
  bn_locate(g, base, n, c);
  float mean, rstd, scale, shift;
  bn_channel(sums, mean_in, var_in, gamma, beta, c, g.inv_n, eps, mean, rstd, scale, shift);
  if (blockIdx.x < (unsigned)(g.C * g.chunks) && blockIdx.x % g.chunks == 0 && threadIdx.x == 0) {
    // first image's CTAs publish the statistics: mean / biased variance for the running averages, mean / rstd for backward
    if (rstd_out) rstd_out[c] = rstd;
    if (sums && mean_out) {
      const double m = sums[2 * c] * g.inv_n;
      const double v = sums[2 * c + 1] * g.inv_n - m * m;
      mean_out[c] = (float)m;
      if (var_out) var_out[c] = (float)(v < 0.0 ? 0.0 : v);
    } else if (mean_out) {
      mean_out[c] = mean;
    }
    if (sums && run.mean) {      // running = (1 - f) * running + f * batch statistic, float32 as torch does it; unbiased variance
      const double m = sums[2 * c] * g.inv_n;
      double v = sums[2 * c + 1] * g.inv_n - m * m;
      if (v < 0.0) v = 0.0;
      const float f = run.momentum >= 0.f ? run.momentum : 1.f / (float)(*run.batches);
      run.mean[c] = run.mean[c] * (1.f - f) + f * (float)m;
      run.var[c] = run.var[c] * (1.f - f) + (f * run.unbias) * (float)v;
    }
  }
  const float* __restrict__ p = x + base;
  float* __restrict__ q = y + base;
  if (VEC) {
    for (int i = threadIdx.x * 4; i < n; i += kBnThreads * 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
      v.x = fmaf(v.x, scale, shift); v.y = fmaf(v.y, scale, shift); v.z = fmaf(v.z, scale, shift); v.w = fmaf(v.w, scale, shift);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      *reinterpret_cast<float4*>(q + i) = v;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kBnThreads) {
      float v = fmaf(__ldg(p + i), scale, shift);
      if (relu) v = fmaxf(v, 0.f);
      q[i] = v;
    }
  }
}

// dsums[2c] += sum dz, dsums[2c+1] += sum dz * xhat      (dz = dy, or dy where the fused ReLU was open)
template <bool VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, double* __restrict__ dsums, BnGeom g, int relu) {
  long long base; int n, c;
  bn_locate(g, base, n, c);
  const float m = mean[c], r = rstd[c];
  const float scale = (gamma ? gamma[c] : 1.f) * r, shift = fmaf(-m, scale, beta ? beta[c] : 0.f);
  const float* __restrict__ p = x + base;
  const float* __restrict__ d = dy + base;
  float s = 0.f, sx = 0.f;
  auto one = [&](float xv, float dv) {
    if (relu && !(fmaf(xv, scale, shift) > 0.f)) dv = 0.f;       // the forward's own z
    s += dv; sx = fmaf(dv, (xv - m) * r, sx);
  };
  if (VEC) {
    for (int i = threadIdx.x * 4; i < n; i += kBnThreads * 4) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(p + i)), dv = __ldg(reinterpret_cast<const float4*>(d + i));
      one(xv.x, dv.x); one(xv.y, dv.y); one(xv.z, dv.z); one(xv.w, dv.w);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kBnThreads) one(__ldg(p + i), __ldg(d + i));
  }
  bn_block_sum2(s, sx, dsums + 2 * c);
}

template <bool VEC>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const double* __restrict__ dsums,
                    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, BnGeom g, int relu, int training,
                    int accumulate) {
  long long base; int n, c;
  bn_locate(g, base, n, c);
  const float m = mean[c], r = rstd[c];
  const float sc = (gamma ? gamma[c] : 1.f) * r, shift = fmaf(-m, sc, beta ? beta[c] : 0.f);
  const float sdz = (float)dsums[2 * c], sdzx = (float)dsums[2 * c + 1];
  if (blockIdx.x < (unsigned)(g.C * g.chunks) && blockIdx.x % g.chunks == 0 && threadIdx.x == 0) {
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + sdz : sdz;          // accumulate: the slot is a slice of the all-reduce bucket
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + sdzx : sdzx;
  }
  // training: statistics depend on x; inference (running statistics): dx = gamma * rstd * dz
  const float k0 = training ? (float)((double)sdz * g.inv_n) : 0.f, k1 = training ? (float)((double)sdzx * g.inv_n) : 0.f;
  const float* __restrict__ p = x + base;
  const float* __restrict__ d = dy + base;
  float* __restrict__ q = dx + base;
  auto one = [&](float xv, float dv) {
    if (relu && !(fmaf(xv, sc, shift) > 0.f)) dv = 0.f;
    return sc * (dv - k0 - (xv - m) * r * k1);
  };
  if (VEC) {
    for (int i = threadIdx.x * 4; i < n; i += kBnThreads * 4) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(p + i)), dv = __ldg(reinterpret_cast<const float4*>(d + i));
      float4 o;
      o.x = one(xv.x, dv.x); o.y = one(xv.y, dv.y); o.z = one(xv.z, dv.z); o.w = one(xv.w, dv.w);
      *reinterpret_cast<float4*>(q + i) = o;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kBnThreads) q[i] = one(__ldg(p + i), __ldg(d + i));
  }
}

static int bn_geom(int64_t N, int64_t C, int64_t HW, BnGeom& g, unsigned& grid, bool& vec, const void* a, const void* b, const void* c3) {
  HG_REQUIRE(N > 0 && C > 0 && HW > 0 && C < (1 << 24), HG_E_SHAPE, "bad batch-norm shape N=%lld C=%lld HW=%lld", (long long)N, (long long)C, (long long)HW);
  g.C = (int)C; g.HW = HW;
  g.chunks = (int)ceil_div(HW, kBnChunk);
  g.inv_n = 1.0 / ((double)N * (double)HW);
  const int64_t blocks = N * C * g.chunks;
  HG_REQUIRE(blocks < (1ll << 31), HG_E_SHAPE, "batch-norm tensor too large for one launch");
  grid = (unsigned)blocks;
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  vec = HW % 4 == 0 && al(a) && al(b) && al(c3);
  return HG_OK;
}

}  // namespace hg

using namespace hg;

extern "C" {

int hg_bn_stats(const float* x, double* sums, int64_t N, int64_t C, int64_t HW, hg_stream_t stream) {
  BnGeom g; unsigned grid; bool vec;
  int rc = bn_geom(N, C, HW, g, grid, vec, x, nullptr, nullptr);
  if (rc) return rc;
  HG_REQUIRE(x && sums, HG_E_ARG, "NULL buffer");
  cudaStream_t st = as_stream(stream);
  if (vec) bn_stats_kernel<true><<<grid, kBnThreads, 0, st>>>(x, sums, g, nullptr);
  else bn_stats_kernel<false><<<grid, kBnThreads, 0, st>>>(x, sums, g, nullptr);
  return finish_launch("bn_stats");
}

int hg_bn_apply(const float* x, float* y, const double* sums, const float* mean_in, const float* var_in, const float* gamma,
                const float* beta, float* mean_out, float* var_out, float* rstd_out, int64_t N, int64_t C, int64_t HW, float eps,
                int relu, hg_stream_t stream) {
  BnGeom g; unsigned grid; bool vec;
  int rc = bn_geom(N, C, HW, g, grid, vec, x, y, nullptr);
  if (rc) return rc;
  HG_REQUIRE(x && y && (sums || (mean_in && var_in)), HG_E_ARG, "batch-norm needs either the batch sums or mean / var");
  cudaStream_t st = as_stream(stream);
  const BnRunning none{nullptr, nullptr, nullptr, 0.f, 1.f};
  if (vec) bn_apply_kernel<true><<<grid, kBnThreads, 0, st>>>(x, y, sums, mean_in, var_in, gamma, beta, mean_out, var_out, rstd_out, g, eps, relu, none);
  else bn_apply_kernel<false><<<grid, kBnThreads, 0, st>>>(x, y, sums, mean_in, var_in, gamma, beta, mean_out, var_out, rstd_out, g, eps, relu, none);
  return finish_launch("bn_apply");
}

int hg_bn_bwd_reduce(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     double* dsums, int64_t N, int64_t C, int64_t HW, int relu, hg_stream_t stream) {
  BnGeom g; unsigned grid; bool vec;
  int rc = bn_geom(N, C, HW, g, grid, vec, x, dy, nullptr);
  if (rc) return rc;
  HG_REQUIRE(x && dy && mean && rstd && dsums, HG_E_ARG, "NULL buffer");
  cudaStream_t st = as_stream(stream);
  if (vec) bn_bwd_reduce_kernel<true><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, g, relu);
  else bn_bwd_reduce_kernel<false><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, g, relu);
  return finish_launch("bn_bwd_reduce");
}

int hg_bn_bwd_apply(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma, const float* beta,
                    const double* dsums, float* dx, float* dgamma, float* dbeta, int64_t N, int64_t C, int64_t HW, int relu,
                    int training, hg_stream_t stream) {
  BnGeom g; unsigned grid; bool vec;
  int rc = bn_geom(N, C, HW, g, grid, vec, x, dy, dx);
  if (rc) return rc;
  HG_REQUIRE(x && dy && mean && rstd && dsums && dx, HG_E_ARG, "NULL buffer");
  cudaStream_t st = as_stream(stream);
  if (vec) bn_bwd_apply_kernel<true><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, g, relu, training, 0);
  else bn_bwd_apply_kernel<false><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, g, relu, training, 0);
  return finish_launch("bn_bwd_apply");
}

// One call per training-mode forward: the C5 step is launch-bound (about 85 launches in 2 ms), and the torch-side bookkeeping
// of a BatchNorm2d -- zero the sums, bump the counter, two mul_ / add_ pairs for the running statistics -- was six launches
// per layer next to the two kernels that do the work.
int hg_bn_train_fwd(const float* x, float* y, double* sums, const float* gamma, const float* beta, float* mean_out, float* rstd_out,
                    float* running_mean, float* running_var, int64_t* num_batches_tracked, double momentum, int64_t N, int64_t C,
                    int64_t HW, float eps, int relu, hg_stream_t stream) {
  BnGeom g; unsigned grid; bool vec;
  int rc = bn_geom(N, C, HW, g, grid, vec, x, y, nullptr);
  if (rc) return rc;
  HG_REQUIRE(x && y && sums && mean_out && rstd_out, HG_E_ARG, "NULL buffer");
  HG_REQUIRE((running_mean == nullptr) == (running_var == nullptr), HG_E_ARG, "running_mean and running_var come together");
  HG_REQUIRE(momentum >= 0.0 || running_mean == nullptr || num_batches_tracked != nullptr, HG_E_ARG,
             "a cumulative moving average (momentum < 0) needs num_batches_tracked");
  HG_REQUIRE(N * HW > 1, HG_E_SHAPE, "training-mode batch norm needs more than one value per channel");
  cudaStream_t st = as_stream(stream);
  cudaError_t me = cudaMemsetAsync(sums, 0, (size_t)(2 * C) * sizeof(double), st);
  HG_REQUIRE(me == cudaSuccess, (int)me, "hg_bn_train_fwd: cudaMemsetAsync: %s", cudaGetErrorString(me));
  long long* nbt = reinterpret_cast<long long*>(num_batches_tracked);
  if (vec) bn_stats_kernel<true><<<grid, kBnThreads, 0, st>>>(x, sums, g, nbt);
  else bn_stats_kernel<false><<<grid, kBnThreads, 0, st>>>(x, sums, g, nbt);
  rc = finish_launch("bn_stats");
  if (rc) return rc;
  const double n = (double)N * (double)HW;
  const BnRunning run{running_mean, running_var, nbt, (float)momentum, (float)(n / (n - 1.0))};
  if (vec) bn_apply_kernel<true><<<grid, kBnThreads, 0, st>>>(x, y, sums, nullptr, nullptr, gamma, beta, mean_out, nullptr, rstd_out, g, eps, relu, run);
  else bn_apply_kernel<false><<<grid, kBnThreads, 0, st>>>(x, y, sums, nullptr, nullptr, gamma, beta, mean_out, nullptr, rstd_out, g, eps, relu, run);
  return finish_launch("bn_apply");
}

// Backward in one call: zeroes `dsums`, reduces, applies.  accumulate_affine != 0 adds dgamma / dbeta to their destinations
// (slices of HyGrid.distributed.FlatGradBucket, zeroed at the start of the step) instead of overwriting them.
int hg_bn_bwd(const float* x, const float* dy, const float* mean, const float* rstd, const float* gamma, const float* beta,
              double* dsums, float* dx, float* dgamma, float* dbeta, int accumulate_affine, int64_t N, int64_t C, int64_t HW,
              int relu, int training, hg_stream_t stream) {
  BnGeom g; unsigned grid; bool vec;
  int rc = bn_geom(N, C, HW, g, grid, vec, x, dy, dx);
  if (rc) return rc;
  HG_REQUIRE(x && dy && mean && rstd && dsums && dx, HG_E_ARG, "NULL buffer");
  cudaStream_t st = as_stream(stream);
  cudaError_t me = cudaMemsetAsync(dsums, 0, (size_t)(2 * C) * sizeof(double), st);
  HG_REQUIRE(me == cudaSuccess, (int)me, "hg_bn_bwd: cudaMemsetAsync: %s", cudaGetErrorString(me));
  if (vec) bn_bwd_reduce_kernel<true><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, g, relu);
  else bn_bwd_reduce_kernel<false><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, g, relu);
  rc = finish_launch("bn_bwd_reduce");
  if (rc) return rc;
  const int acc = accumulate_affine ? 1 : 0;
  if (vec) bn_bwd_apply_kernel<true><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, g, relu, training, acc);
  else bn_bwd_apply_kernel<false><<<grid, kBnThreads, 0, st>>>(x, dy, mean, rstd, gamma, beta, dsums, dx, dgamma, dbeta, g, relu, training, acc);
  return finish_launch("bn_bwd_apply");
}

}  // extern "C"
