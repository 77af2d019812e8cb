// improved_ops.cu — the non-GEMM operators of arch_unet.ImprovedUNet (reference arch_unet.py:420-531):
// GroupNorm forward / backward (norm2d('gn', c, 32), arch_unet.py:7-15; optional fused LeakyReLU or residual add, the two
// forms ResBlock uses, :421-432), LeakyReLU / Sigmoid forward and backward from the OUTPUT (the reference runs them in
// place), PixelShuffle(2) and its inverse (UpBlock, :456-466) and the residual add of RDB (:449).  All HBM-bound passes over
// fp32 NCHW tensors; reductions are fixed-order (double accumulators) and therefore deterministic.
#include "common.cuh"

namespace n2n {

constexpr int kGnThreads = 256;

__device__ __forceinline__ void block_reduce2(double& a, double& b, double* smem /* [16] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smem[warp] = a; smem[8 + warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0, sb = 0;
    for (int w = 0; w < kGnThreads / 32; ++w) { sa += smem[w]; sb += smem[8 + w]; }
    a = sa; b = sb;
  }
}

// V = 4 (float4 accesses, H*W % 4 == 0) or 1 (any H*W, e.g. the 2 x 3 bottleneck of a 32 x 48 input)
template <int V> __device__ __forceinline__ void ldv(const float* p, long long i, float (&v)[V]) {
  if constexpr (V == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else v[0] = __ldg(p + i);
}
template <int V> __device__ __forceinline__ void stv(float* p, long long i, const float (&v)[V]) {
  if constexpr (V == 4) reinterpret_cast<float4*>(p)[i] = make_float4(v[0], v[1], v[2], v[3]);
  else p[i] = v[0];
}

// ---- GroupNorm forward ------------------------------------------------------------------------------------------------
// (n, g) owns the contiguous span x[(n*C + g*cg) * hw, +cg*hw).  Stage 1: `splits` blocks per span write (sum, sum of
// squares) partials; stage 2 (the apply kernel) folds them — so a single 704 x 704 image still fills the machine.
template <int V>
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const float* __restrict__ x, long long span, int splits, double* __restrict__ partial) {
  pdl_enter();
  __shared__ double red[16];
  const long long ng = blockIdx.x / splits;
  const int s = blockIdx.x % splits;
  const long long chunk = (span / V + splits - 1) / splits * V;        // span % V == 0
  const long long lo = (long long)s * chunk, hi = lo + chunk < span ? lo + chunk : span;
  const float* p = x + ng * span;
  double a = 0.0, b = 0.0;
  for (long long i = lo / V + threadIdx.x; i < hi / V; i += kGnThreads) {
    float v[V];
    ldv<V>(p, i, v);
#pragma unroll
    for (int j = 0; j < V; ++j) { a += (double)v[j]; b += (double)v[j] * v[j]; }
  }
  block_reduce2(a, b, red);
  if (threadIdx.x == 0) { partial[2 * (long long)blockIdx.x] = a; partial[2 * (long long)blockIdx.x + 1] = b; }
}

__device__ __forceinline__ void gn_fold(const double* partial, long long ng, int splits, double inv_count, float eps,
                                        float& mean, float& rstd) {
  double a = 0.0, b = 0.0;
  for (int s = 0; s < splits; ++s) { a += partial[2 * (ng * splits + s)]; b += partial[2 * (ng * splits + s) + 1]; }
  const double m = a * inv_count;
  double var = b * inv_count - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// grid = (blocks per plane, n * C); y = (x - mean) * rstd * gamma[c] + beta[c]; LeakyReLU when slope >= 0; + res when given
template <int V>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ res, float* __restrict__ y, const double* __restrict__ partial, int splits,
                float* __restrict__ mean_rstd, int C, int cg, int hw, float eps, float slope) {
  pdl_enter();
  const int plane = blockIdx.y;
  const int n = plane / C, c = plane % C;
  const long long ng = (long long)n * (C / cg) + c / cg;
  float mean, rstd;
  gn_fold(partial, ng, splits, 1.0 / ((double)cg * hw), eps, mean, rstd);
  if (mean_rstd && blockIdx.x == 0 && threadIdx.x == 0 && c % cg == 0) { mean_rstd[2 * ng] = mean; mean_rstd[2 * ng + 1] = rstd; }
  const float sc = rstd * gamma[c], sh = beta[c] - mean * sc;
  const float* px = x + (long long)plane * hw;
  const float* pr = res ? res + (long long)plane * hw : nullptr;
  float* py = y + (long long)plane * hw;
  for (int i = blockIdx.x * kGnThreads + threadIdx.x; i < hw / V; i += gridDim.x * kGnThreads) {
    float v[V];
    ldv<V>(px, i, v);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = fmaf(v[j], sc, sh);
      if (slope >= 0.f) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
    }
    if (pr) {
      float r[V];
      ldv<V>(pr, i, r);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] += r[j];
    }
    stv<V>(py, i, v);
  }
}

// ---- GroupNorm backward -----------------------------------------------------------------------------------------------
// dy_eff = dy * LeakyReLU'(y) (sign of the OUTPUT) when the activation was fused.  Per plane (n, c):
// a = sum dy_eff * xhat, b = sum dy_eff.  Then dgamma[c] = sum_n a, dbeta[c] = sum_n b and, with s1 = sum_{c in g} gamma a,
// s2 = sum_{c in g} gamma b, M = cg * hw:  dx = rstd * (dy_eff * gamma - (xhat * s1 + s2) / M).
template <int V>
__global__ void __launch_bounds__(kGnThreads)
gn_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                   const float* __restrict__ mean_rstd, double* __restrict__ ab, int C, int cg, int hw, float slope) {
  pdl_enter();
  __shared__ double red[16];
  const int plane = blockIdx.x;
  const int n = plane / C, c = plane % C;
  const long long ng = (long long)n * (C / cg) + c / cg;
  const float mean = mean_rstd[2 * ng], rstd = mean_rstd[2 * ng + 1];
  const float* px = x + (long long)plane * hw;
  const float* py = (slope >= 0.f) ? y + (long long)plane * hw : nullptr;
  const float* pd = dy + (long long)plane * hw;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < hw / V; i += kGnThreads) {
    float xv[V], d[V];
    ldv<V>(px, i, xv);
    ldv<V>(pd, i, d);
    if (py) {
      float o[V];
      ldv<V>(py, i, o);
#pragma unroll
      for (int j = 0; j < V; ++j) d[j] *= o[j] > 0.f ? 1.f : slope;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { a += (double)(d[j] * ((xv[j] - mean) * rstd)); b += (double)d[j]; }
  }
  block_reduce2(a, b, red);
  if (threadIdx.x == 0) { ab[2 * (long long)plane] = a; ab[2 * (long long)plane + 1] = b; }
}

__global__ void gn_bwd_params_kernel(const double* __restrict__ ab, float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int C) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int n = 0; n < N; ++n) { a += ab[2 * ((long long)n * C + c)]; b += ab[2 * ((long long)n * C + c) + 1]; }
  dgamma[c] = (float)a; dbeta[c] = (float)b;
}

template <int V>
__global__ void __launch_bounds__(kGnThreads)
gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                    const float* __restrict__ gamma, const float* __restrict__ mean_rstd, const double* __restrict__ ab,
                    float* __restrict__ dx, int C, int cg, int hw, float slope) {
  pdl_enter();
  const int plane = blockIdx.y;
  const int n = plane / C, c = plane % C;
  const int g = c / cg;
  const long long ng = (long long)n * (C / cg) + g;
  const float mean = mean_rstd[2 * ng], rstd = mean_rstd[2 * ng + 1];
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < cg; ++k) {
    const int cc = g * cg + k;
    s1 += (double)gamma[cc] * ab[2 * ((long long)n * C + cc)];
    s2 += (double)gamma[cc] * ab[2 * ((long long)n * C + cc) + 1];
  }
  const float invM = 1.0f / ((float)cg * (float)hw);
  const float k1 = (float)s1 * invM, k2 = (float)s2 * invM, gm = gamma[c];
  const float* px = x + (long long)plane * hw;
  const float* py = (slope >= 0.f) ? y + (long long)plane * hw : nullptr;
  const float* pd = dy + (long long)plane * hw;
  float* po = dx + (long long)plane * hw;
  for (int i = blockIdx.x * kGnThreads + threadIdx.x; i < hw / V; i += gridDim.x * kGnThreads) {
    float xv[V], d[V], r[V];
    ldv<V>(px, i, xv);
    ldv<V>(pd, i, d);
    if (py) {
      float o[V];
      ldv<V>(py, i, o);
#pragma unroll
      for (int j = 0; j < V; ++j) d[j] *= o[j] > 0.f ? 1.f : slope;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) r[j] = rstd * (d[j] * gm - ((xv[j] - mean) * rstd * k1 + k2));
    stv<V>(po, i, r);
  }
}

// ---- activations (forward; backward from the output), add, PixelShuffle(2) -----------------------------------------------
// kind 1: LeakyReLU(slope), kind 2: Sigmoid
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long count, int kind, float slope) {
  pdl_enter();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < count; i += gridDim.x * 256LL) {
    const float v = x[i];
    y[i] = kind == 1 ? (v > 0.f ? v : v * slope) : 1.0f / (1.0f + expf(-v));
  }
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx,
                                                      long long count, int kind, float slope) {
  pdl_enter();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < count; i += gridDim.x * 256LL) {
    const float o = y[i], d = dy[i];
    dx[i] = kind == 1 ? d * (o > 0.f ? 1.f : slope) : d * o * (1.0f - o);
  }
}
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long count) {
  pdl_enter();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < count; i += gridDim.x * 256LL) out[i] = a[i] + b[i];
}
// forward: y[n, c, 2h+dy, 2w+dx] = x[n, 4c + 2dy + dx, h, w]; inverse: the same index map read the other way.
// One thread per (input pixel pair row): reads the four source planes, writes two float2 — coalesced on both sides.
__global__ void __launch_bounds__(256) pixel_shuffle2_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W,
                                                             long long total, int inverse) {
  pdl_enter();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int w = (int)(i % W);
    long long t = i / W;
    const int h = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const long long n = t / C;
    const long long lo = ((n * 4 * C + 4 * c) * H + h) * W + w;          // plane 4c of the low-resolution tensor
    const long long hi = ((n * C + c) * 2 * H + 2 * h) * (2LL * W) + 2 * w;
    const long long ps = (long long)H * W;
    if (!inverse) {
      *reinterpret_cast<float2*>(dst + hi) = make_float2(src[lo], src[lo + ps]);
      *reinterpret_cast<float2*>(dst + hi + 2 * W) = make_float2(src[lo + 2 * ps], src[lo + 3 * ps]);
    } else {
      const float2 r0 = *reinterpret_cast<const float2*>(src + hi), r1 = *reinterpret_cast<const float2*>(src + hi + 2 * W);
      dst[lo] = r0.x; dst[lo + ps] = r0.y; dst[lo + 2 * ps] = r1.x; dst[lo + 3 * ps] = r1.y;
    }
  }
}

static int gn_splits(long long groups_total, long long span) {
  int s = (int)((kSMs * 4 + groups_total - 1) / groups_total);
  if (s > 32) s = 32;
  while (s > 1 && span / s < 4096) --s;
  return s < 1 ? 1 : s;
}

}  // namespace n2n

using namespace n2n;

extern "C" int n2n_groupnorm_groups(int channels, int groups) {
  int g = groups < channels ? groups : channels;
  while (g > 1 && channels % g != 0) --g;                 // arch_unet.py:12-14
  return g < 1 ? 1 : g;
}

extern "C" size_t n2n_groupnorm_workspace_bytes(int n, int c) {
  return ((size_t)n * c * 2 + (size_t)n * c * 2 * 32) * sizeof(double);   // per-plane (a, b) + forward split partials
}

extern "C" int n2n_groupnorm_fwd(const float* x, const float* gamma, const float* beta, const float* residual, float* y,
                                 float* mean_rstd, int n, int c, int hw, int groups, float eps, float act_slope,
                                 void* workspace, void* stream) {
  N2N_CHECK_ARG(x && gamma && beta && y && workspace && n > 0 && c > 0 && hw > 0 && groups > 0 && c % groups == 0,
                "groupnorm_fwd: bad arguments (n=%d c=%d hw=%d groups=%d)", n, c, hw, groups);
  N2N_CHECK_ARG(!(residual && act_slope >= 0.f), "groupnorm_fwd: fused activation and residual are exclusive");
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = c / groups;
  const long long span = (long long)cg * hw, ngt = (long long)n * groups;
  const int splits = gn_splits(ngt, span);
  double* partial = (double*)workspace + (size_t)n * c * 2;
  const bool v4 = hw % 4 == 0;
  (void)launch_pdl_v(v4 ? gn_stats_kernel<4> : gn_stats_kernel<1>, dim3((unsigned)(ngt * splits)), dim3(kGnThreads), 0, st, x, span, splits, partial);
  N2N_LAUNCH_CHECK();
  int bpp = (hw / 4 + kGnThreads - 1) / kGnThreads;
  if (bpp < 1) bpp = 1;
  const long long planes = (long long)n * c;
  N2N_CHECK_ARG(planes < 65536, "groupnorm_fwd: n * c = %lld planes exceed the grid", planes);
  while (bpp > 1 && planes * bpp > (long long)kSMs * 16) bpp = (bpp + 1) / 2;
  (void)launch_pdl_v(v4 ? gn_apply_kernel<4> : gn_apply_kernel<1>, dim3(bpp, (unsigned)planes), dim3(kGnThreads), 0, st, x, gamma, beta, residual, y,
                     (const double*)partial, splits, mean_rstd, c, cg, hw, eps, act_slope);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_groupnorm_bwd(const float* x, const float* gamma, const float* y, const float* dy, const float* mean_rstd,
                                 float* dx, float* dgamma, float* dbeta, int n, int c, int hw, int groups, float act_slope,
                                 void* workspace, void* stream) {
  N2N_CHECK_ARG(x && gamma && dy && mean_rstd && dx && dgamma && dbeta && workspace && n > 0 && c > 0 && hw > 0 && groups > 0 &&
                c % groups == 0, "groupnorm_bwd: bad arguments (n=%d c=%d hw=%d groups=%d)", n, c, hw, groups);
  N2N_CHECK_ARG(act_slope < 0.f || y, "groupnorm_bwd: the fused activation needs the forward output");
  cudaStream_t st = (cudaStream_t)stream;
  const int cg = c / groups;
  const long long planes = (long long)n * c;
  N2N_CHECK_ARG(planes < 65536, "groupnorm_bwd: n * c = %lld planes exceed the grid", planes);
  double* ab = (double*)workspace;
  const bool v4 = hw % 4 == 0;
  (void)launch_pdl_v(v4 ? gn_bwd_sums_kernel<4> : gn_bwd_sums_kernel<1>, dim3((unsigned)planes), dim3(kGnThreads), 0, st, x, y, dy, mean_rstd, ab, c, cg, hw, act_slope);
  N2N_LAUNCH_CHECK();
  (void)launch_pdl_v(gn_bwd_params_kernel, dim3((c + 127) / 128), dim3(128), 0, st, (const double*)ab, dgamma, dbeta, n, c);
  N2N_LAUNCH_CHECK();
  int bpp = (hw / 4 + kGnThreads - 1) / kGnThreads;
  if (bpp < 1) bpp = 1;
  while (bpp > 1 && planes * bpp > (long long)kSMs * 16) bpp = (bpp + 1) / 2;
  (void)launch_pdl_v(v4 ? gn_bwd_apply_kernel<4> : gn_bwd_apply_kernel<1>, dim3(bpp, (unsigned)planes), dim3(kGnThreads), 0, st, x, y, dy, gamma, mean_rstd,
                     (const double*)ab, dx, c, cg, hw, act_slope);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_act_fwd(const float* x, float* y, int64_t count, int kind, float slope, void* stream) {
  N2N_CHECK_ARG(x && y && count > 0 && (kind == 1 || kind == 2), "act_fwd: bad arguments");
  (void)launch_pdl_v(act_fwd_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, (long long)count, kind, slope);
  N2N_LAUNCH_CHECK();
  return 0;
}
extern "C" int n2n_act_bwd(const float* y, const float* dy, float* dx, int64_t count, int kind, float slope, void* stream) {
  N2N_CHECK_ARG(y && dy && dx && count > 0 && (kind == 1 || kind == 2), "act_bwd: bad arguments");
  (void)launch_pdl_v(act_bwd_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, y, dy, dx, (long long)count, kind, slope);
  N2N_LAUNCH_CHECK();
  return 0;
}
extern "C" int n2n_add_f32(const float* a, const float* b, float* out, int64_t count, void* stream) {
  N2N_CHECK_ARG(a && b && out && count > 0, "add_f32: bad arguments");
  (void)launch_pdl_v(add_kernel, dim3(grid_for(count, 256)), dim3(256), 0, (cudaStream_t)stream, a, b, out, (long long)count);
  N2N_LAUNCH_CHECK();
  return 0;
}
extern "C" int n2n_pixel_shuffle2(const float* src, float* dst, int n, int c_out, int h, int w, int inverse, void* stream) {
  N2N_CHECK_ARG(src && dst && n > 0 && c_out > 0 && h > 0 && w > 0, "pixel_shuffle2: bad arguments");
  const long long total = (long long)n * c_out * h * w;
  (void)launch_pdl_v(pixel_shuffle2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, dst, c_out, h, w, total, inverse);
  N2N_LAUNCH_CHECK();
  return 0;
}
