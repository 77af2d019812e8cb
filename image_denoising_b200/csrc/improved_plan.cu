// improved_plan.cu — native no-grad executor of arch_unet.ImprovedUNet (reference arch_unet.py:420-531).
//
// One C call runs the whole forward on the tensor-core engines with every activation resident in the blocked C16 layout
// (bf16 in "bf16" mode, fp32 on the parity engine): the dense concats of the RDBs (arch_unet.py:443-449) are block ranges of
// ONE buffer that the four growth convolutions fill in place, the skip concat of an UpBlock (:462) is written in place by
// the encoder's last GroupNorm and by the PixelShuffle kernel, the RDB / ResBlock residuals ride in the GEMM epilogue
// (addend) and in the GroupNorm apply pass.  Layers wider than one launch allows (N > 256 accumulator columns) are issued
// as column chunks.  Used by image_denoising_b200.improved.ImprovedUNet for no-grad calls; the training path composes the
// per-layer C-ABI calls under autograd (improved.py).
#include <vector>

#include "common.cuh"
#include "layers.cuh"

namespace n2n {

// ---- GroupNorm on C16 ------------------------------------------------------------------------------------------------
// stats: per (image, channel) sum / sum of squares over the pixels, `splits` partial rows per channel block
template <typename T>
__global__ void __launch_bounds__(256)
c16_gn_stats_kernel(View x, int splits, double* __restrict__ partial /* [N][Cb][splits][32] */) {
  pdl_enter();
  __shared__ float red[8][32];
  const int cb = blockIdx.y, n = blockIdx.z, s = blockIdx.x;
  const long long hw = (long long)x.H * x.W;
  const long long chunk = (hw + splits - 1) / splits, lo = s * chunk, hi = lo + chunk < hw ? lo + chunk : hw;
  float a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = b[i] = 0.f;
  const T* base = (const T*)x.ptr + n * x.sN + cb * x.sCb;
  for (long long p = lo + threadIdx.x; p < hi; p += 256) {
    const int yy = (int)(p / x.W), xx = (int)(p % x.W);
    float v[16];
    Block16<T>::load(base + yy * x.sY + xx * x.sX, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] += v[i]; b[i] = fmaf(v[i], v[i], b[i]); }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a[i] += __shfl_xor_sync(0xffffffffu, a[i], o); b[i] += __shfl_xor_sync(0xffffffffu, b[i], o); }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { red[warp][i] = a[i]; red[warp][16 + i] = b[i]; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double sum = 0.0;
    for (int w = 0; w < 8; ++w) sum += (double)red[w][threadIdx.x];
    partial[(((long long)n * gridDim.y + cb) * splits + s) * 32 + threadIdx.x] = sum;
  }
}
// fold: per (image, channel) scale = rstd * gamma, shift = beta - mean * scale of the channel's group (0 for pad channels)
__global__ void c16_gn_fold_kernel(const double* __restrict__ partial, int splits, int N, int C, int Cb, int cg, long long hw, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scale_shift /* [N][Cb*16][2] */,
                                   float* __restrict__ mean_rstd /* training plans: [N][Cb*16][2], else NULL */) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Cb * 16) return;
  const int n = i / (Cb * 16), c = i % (Cb * 16);
  float sc = 0.f, sh = 0.f, mu = 0.f, rs = 0.f;
  if (c < C) {
    const int g0 = (c / cg) * cg;
    double a = 0.0, b = 0.0;
    for (int k = g0; k < g0 + cg; ++k)
      for (int s = 0; s < splits; ++s) {
        const double* p = partial + (((long long)n * Cb + (k >> 4)) * splits + s) * 32;
        a += p[k & 15]; b += p[16 + (k & 15)];
      }
    const double inv = 1.0 / ((double)cg * (double)hw), mean = a * inv;
    double var = b * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    sc = rstd * gamma[c];
    sh = beta[c] - (float)mean * sc;
    mu = (float)mean; rs = rstd;
  }
  scale_shift[2 * i] = sc; scale_shift[2 * i + 1] = sh;
  if (mean_rstd) { mean_rstd[2 * i] = mu; mean_rstd[2 * i + 1] = rs; }
}
template <typename T>
__global__ void __launch_bounds__(256)
c16_gn_apply_kernel(View x, View res, int has_res, View y, const float* __restrict__ scale_shift, float slope, long long items) {
  pdl_enter();
  const int W = x.W, H = x.H, Cb = x.Cb;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
    const int xx = (int)(i % W);
    long long t = i / W;
    const int yy = (int)(t % H); t /= H;
    const int cb = (int)(t % Cb);
    const long long n = t / Cb;
    float v[16];
    Block16<T>::load((const T*)x.ptr + n * x.sN + cb * x.sCb + yy * x.sY + xx * x.sX, v);
    const float* ss = scale_shift + ((n * Cb + cb) * 16) * 2;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      v[e] = fmaf(v[e], ss[2 * e], ss[2 * e + 1]);
      if (slope >= 0.f) v[e] = v[e] > 0.f ? v[e] : v[e] * slope;
    }
    if (has_res) {
      float r[16];
      Block16<T>::load((const T*)res.ptr + n * res.sN + cb * res.sCb + yy * res.sY + xx * res.sX, r);
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] += r[e];
    }
    Block16<T>::store((T*)y.ptr + n * y.sN + cb * y.sCb + yy * y.sY + xx * y.sX, v);
  }
}

// ---- PixelShuffle(2) on C16: out[n, co, 2y+i, 2x+j] = in[n, 4 co + 2i + j, y, x]; pad lanes of the last output block = 0 ---------
template <typename T>
__global__ void __launch_bounds__(256)
c16_pixel_shuffle2_kernel(View src, View dst, int c_out, long long items) {
  pdl_enter();
  const int W = src.W, H = src.H, Cbo = dst.Cb;
  for (long long it = blockIdx.x * 256LL + threadIdx.x; it < items; it += gridDim.x * 256LL) {
    const int xx = (int)(it % W);
    long long t = it / W;
    const int yy = (int)(t % H); t /= H;
    const int cbo = (int)(t % Cbo);
    const long long n = t / Cbo;
    float in[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (4 * cbo + q < src.Cb) Block16<T>::load((const T*)src.ptr + n * src.sN + (4 * cbo + q) * src.sCb + yy * src.sY + xx * src.sX, in[q]);
      else {
#pragma unroll
        for (int e = 0; e < 16; ++e) in[q][e] = 0.f;
      }
    }
#pragma unroll
    for (int ij = 0; ij < 4; ++ij) {
      float o[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) o[e] = (16 * cbo + e < c_out) ? in[e >> 2][4 * (e & 3) + ij] : 0.f;
      Block16<T>::store((T*)dst.ptr + n * dst.sN + cbo * dst.sCb + (2 * yy + (ij >> 1)) * dst.sY + (2 * xx + (ij & 1)) * dst.sX, o);
    }
  }
}
// write a 1-channel fp32 NCHW map into lane `lane` of block 0 of a C16 view (the sigma map of the noise estimator, :516-517)
template <typename T>
__global__ void c16_set_lane_kernel(const float* __restrict__ src, View dst, int lane, long long items) {
  pdl_enter();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
    const int xx = (int)(i % dst.W);
    long long t = i / dst.W;
    const int yy = (int)(t % dst.H);
    const long long n = t / dst.H;
    ((T*)dst.ptr)[n * dst.sN + yy * dst.sY + xx * dst.sX + lane] = from_f32<T>(src[i]);
  }
}

// ---- backward kernels on C16 (training plans) ----------------------------------------------------------------------------
// g *= LeakyReLU'(act)  (the activation was applied in place by the producer, so its sign is the mask)
template <typename T>
__global__ void __launch_bounds__(256) c16_lrelu_mask_kernel(View act, View g, float slope, long long items) {
  pdl_enter();
  const int W = g.W, H = g.H, Cb = g.Cb;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
    const int xx = (int)(i % W);
    long long t = i / W;
    const int yy = (int)(t % H); t /= H;
    const int cb = (int)(t % Cb);
    const long long n = t / Cb;
    float a[16], v[16];
    Block16<T>::load((const T*)act.ptr + n * act.sN + cb * act.sCb + yy * act.sY + xx * act.sX, a);
    T* gp = (T*)g.ptr + n * g.sN + cb * g.sCb + yy * g.sY + xx * g.sX;
    Block16<T>::load(gp, v);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] *= a[e] > 0.f ? 1.f : slope;
    Block16<T>::store(gp, v);
  }
}
// dst += src
template <typename T>
__global__ void __launch_bounds__(256) c16_add_into_kernel(View src, View dst, long long items) {
  pdl_enter();
  const int W = dst.W, H = dst.H, Cb = dst.Cb;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
    const int xx = (int)(i % W);
    long long t = i / W;
    const int yy = (int)(t % H); t /= H;
    const int cb = (int)(t % Cb);
    const long long n = t / Cb;
    float a[16], v[16];
    Block16<T>::load((const T*)src.ptr + n * src.sN + cb * src.sCb + yy * src.sY + xx * src.sX, a);
    T* dp = (T*)dst.ptr + n * dst.sN + cb * dst.sCb + yy * dst.sY + xx * dst.sX;
    Block16<T>::load(dp, v);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] += a[e];
    Block16<T>::store(dp, v);
  }
}
// MaxPool2d(2) backward, accumulating: gsrc[first maximum of the 2x2 cell] += gpool  (ATen's tie rule: first in row-major order)
template <typename T>
__global__ void __launch_bounds__(128) c16_unpool_acc_kernel(View act, View gpool, View gsrc, long long items) {
  pdl_enter();
  const int Wo = gpool.W, Ho = gpool.H, Cb = gpool.Cb;
  for (long long i = blockIdx.x * 128LL + threadIdx.x; i < items; i += gridDim.x * 128LL) {
    long long r = i;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho); r /= Ho;
    const int cb = (int)(r % Cb);
    const long long n = r / Cb;
    const T* s = (const T*)act.ptr + n * act.sN + cb * act.sCb + (2 * y) * act.sY + (2 * x) * act.sX;
    float v[4][16], g[16];
    Block16<T>::load(s, v[0]);
    Block16<T>::load(s + act.sX, v[1]);
    Block16<T>::load(s + act.sY, v[2]);
    Block16<T>::load(s + act.sY + act.sX, v[3]);
    Block16<T>::load((const T*)gpool.ptr + n * gpool.sN + cb * gpool.sCb + y * gpool.sY + x * gpool.sX, g);
    T* d = (T*)gsrc.ptr + n * gsrc.sN + cb * gsrc.sCb + (2 * y) * gsrc.sY + (2 * x) * gsrc.sX;
    const long long offs[4] = {0, gsrc.sX, gsrc.sY, gsrc.sY + gsrc.sX};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float o[16];
      Block16<T>::load(d + offs[q], o);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        int best = 0; float m = v[0][k];
        if (v[1][k] > m) { m = v[1][k]; best = 1; }
        if (v[2][k] > m) { m = v[2][k]; best = 2; }
        if (v[3][k] > m) { m = v[3][k]; best = 3; }
        if (best == q) o[k] += g[k];
      }
      Block16<T>::store(d + offs[q], o);
    }
  }
}
// PixelShuffle(2) backward, accumulating: gsrc[n, 4 co + 2i + j, y, x] += gdst[n, co, 2y+i, 2x+j]
template <typename T>
__global__ void __launch_bounds__(256) c16_pixel_unshuffle2_acc_kernel(View gdst, View gsrc, int c_out, long long items) {
  pdl_enter();
  const int W = gsrc.W, H = gsrc.H, Cbo = gdst.Cb;
  for (long long it = blockIdx.x * 256LL + threadIdx.x; it < items; it += gridDim.x * 256LL) {
    const int xx = (int)(it % W);
    long long t = it / W;
    const int yy = (int)(t % H); t /= H;
    const int cbo = (int)(t % Cbo);
    const long long n = t / Cbo;
    float hi[4][16];
#pragma unroll
    for (int ij = 0; ij < 4; ++ij)
      Block16<T>::load((const T*)gdst.ptr + n * gdst.sN + cbo * gdst.sCb + (2 * yy + (ij >> 1)) * gdst.sY + (2 * xx + (ij & 1)) * gdst.sX, hi[ij]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (4 * cbo + q >= gsrc.Cb) continue;
      T* sp = (T*)gsrc.ptr + n * gsrc.sN + (4 * cbo + q) * gsrc.sCb + yy * gsrc.sY + xx * gsrc.sX;
      float o[16];
      Block16<T>::load(sp, o);
#pragma unroll
      for (int l = 0; l < 16; ++l) {                 // lane l of input block 4 cbo + q = channel 4 (4 q + l / 4) + l % 4 relative to 64 cbo
        const int e = 4 * q + (l >> 2), ij = l & 3;
        if (16 * cbo + e < c_out) o[l] += hi[ij][e];
      }
      Block16<T>::store(sp, o);
    }
  }
}
// lane `lane` of block 0 of a C16 view -> a 1-channel fp32 NCHW map
template <typename T>
__global__ void c16_get_lane_kernel(View src, int lane, float* __restrict__ dst, long long items) {
  pdl_enter();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
    const int xx = (int)(i % src.W);
    long long t = i / src.W;
    const int yy = (int)(t % src.H);
    const long long n = t / src.H;
    dst[i] = to_f32<T>(((const T*)src.ptr)[n * src.sN + yy * src.sY + xx * src.sX + lane]);
  }
}

// GroupNorm backward.  stats: per (image, channel) a = sum dy_eff * xhat, b = sum dy_eff (dy_eff = dy * LeakyReLU'(y) when fused)
template <typename T>
__global__ void __launch_bounds__(256)
c16_gn_bwd_stats_kernel(View x, View y, View gy, const float* __restrict__ mean_rstd /* [N][Cb*16][2] */, float slope, int splits,
                        double* __restrict__ partial /* [N][Cb][splits][32] */) {
  pdl_enter();
  __shared__ float red[8][32];
  const int cb = blockIdx.y, n = blockIdx.z, s = blockIdx.x;
  const long long hw = (long long)x.H * x.W;
  const long long chunk = (hw + splits - 1) / splits, lo = s * chunk, hi = lo + chunk < hw ? lo + chunk : hw;
  float a[16], b[16], mu[16], rs[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = b[i] = 0.f;
    mu[i] = mean_rstd[(((long long)n * gridDim.y + cb) * 16 + i) * 2]; rs[i] = mean_rstd[(((long long)n * gridDim.y + cb) * 16 + i) * 2 + 1];
  }
  for (long long p = lo + threadIdx.x; p < hi; p += 256) {
    const int yy = (int)(p / x.W), xx = (int)(p % x.W);
    float v[16], d[16];
    Block16<T>::load((const T*)x.ptr + n * x.sN + cb * x.sCb + yy * x.sY + xx * x.sX, v);
    Block16<T>::load((const T*)gy.ptr + n * gy.sN + cb * gy.sCb + yy * gy.sY + xx * gy.sX, d);
    if (slope >= 0.f) {
      float o[16];
      Block16<T>::load((const T*)y.ptr + n * y.sN + cb * y.sCb + yy * y.sY + xx * y.sX, o);
#pragma unroll
      for (int i = 0; i < 16; ++i) d[i] *= o[i] > 0.f ? 1.f : slope;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = fmaf(d[i], (v[i] - mu[i]) * rs[i], a[i]); b[i] += d[i]; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a[i] += __shfl_xor_sync(0xffffffffu, a[i], o); b[i] += __shfl_xor_sync(0xffffffffu, b[i], o); }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { red[warp][i] = a[i]; red[warp][16 + i] = b[i]; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double sum = 0.0;
    for (int w = 0; w < 8; ++w) sum += (double)red[w][threadIdx.x];
    partial[(((long long)n * gridDim.y + cb) * splits + s) * 32 + threadIdx.x] = sum;
  }
}
// fold: per (image, channel) k1 = sum_{c' in group} gamma a / M, k2 = sum gamma b / M; per channel dgamma = sum_n a, dbeta = sum_n b
__global__ void c16_gn_bwd_fold_kernel(const double* __restrict__ partial, int splits, int N, int C, int Cb, int cg, long long hw,
                                       const float* __restrict__ gamma, float* __restrict__ coef /* [N][Cb*16][2] */,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * Cb * 16) {
    const int n = i / (Cb * 16), c = i % (Cb * 16);
    float k1 = 0.f, k2 = 0.f;
    if (c < C) {
      const int g0 = (c / cg) * cg;
      double s1 = 0.0, s2 = 0.0;
      for (int k = g0; k < g0 + cg; ++k) {
        double a = 0.0, b = 0.0;
        for (int s = 0; s < splits; ++s) {
          const double* p = partial + (((long long)n * Cb + (k >> 4)) * splits + s) * 32;
          a += p[k & 15]; b += p[16 + (k & 15)];
        }
        s1 += (double)gamma[k] * a; s2 += (double)gamma[k] * b;
      }
      const double inv = 1.0 / ((double)cg * (double)hw);
      k1 = (float)(s1 * inv); k2 = (float)(s2 * inv);
    }
    coef[2 * i] = k1; coef[2 * i + 1] = k2;
  }
  if (i < C) {
    double a = 0.0, b = 0.0;
    for (int n = 0; n < N; ++n)
      for (int s = 0; s < splits; ++s) {
        const double* p = partial + (((long long)n * Cb + (i >> 4)) * splits + s) * 32;
        a += p[i & 15]; b += p[16 + (i & 15)];
      }
    dgamma[i] = (float)a; dbeta[i] = (float)b;
  }
}
// gx += rstd * (dy_eff * gamma - (xhat * k1 + k2));  gres += dy (the residual branch, un-masked)
template <typename T>
__global__ void __launch_bounds__(256)
c16_gn_bwd_apply_kernel(View x, View y, View gy, View gx, View gres, int has_res, const float* __restrict__ mean_rstd,
                        const float* __restrict__ coef, const float* __restrict__ gamma, int C, float slope, long long items) {
  pdl_enter();
  const int W = x.W, H = x.H, Cb = x.Cb;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < items; i += gridDim.x * 256LL) {
    const int xx = (int)(i % W);
    long long t = i / W;
    const int yy = (int)(t % H); t /= H;
    const int cb = (int)(t % Cb);
    const long long n = t / Cb;
    float v[16], d[16], o[16];
    Block16<T>::load((const T*)x.ptr + n * x.sN + cb * x.sCb + yy * x.sY + xx * x.sX, v);
    Block16<T>::load((const T*)gy.ptr + n * gy.sN + cb * gy.sCb + yy * gy.sY + xx * gy.sX, d);
    if (has_res) {
      T* rp = (T*)gres.ptr + n * gres.sN + cb * gres.sCb + yy * gres.sY + xx * gres.sX;
      float r[16];
      Block16<T>::load(rp, r);
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] += d[e];
      Block16<T>::store(rp, r);
    }
    if (slope >= 0.f) {
      Block16<T>::load((const T*)y.ptr + n * y.sN + cb * y.sCb + yy * y.sY + xx * y.sX, o);
#pragma unroll
      for (int e = 0; e < 16; ++e) d[e] *= o[e] > 0.f ? 1.f : slope;
    }
    T* gp = (T*)gx.ptr + n * gx.sN + cb * gx.sCb + yy * gx.sY + xx * gx.sX;
    Block16<T>::load(gp, o);
    const float* mr = mean_rstd + ((n * Cb + cb) * 16) * 2;
    const float* kc = coef + ((n * Cb + cb) * 16) * 2;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int c = cb * 16 + e;
      if (c < C) {
        const float rs = mr[2 * e + 1], xh = (v[e] - mr[2 * e]) * rs;
        o[e] += rs * (d[e] * gamma[c] - (xh * kc[2 * e] + kc[2 * e + 1]));
      }
    }
    Block16<T>::store(gp, o);
  }
}

template <typename F32K, typename BF16K, typename... Args>
static cudaError_t launch_by_dtype(int dtype, F32K kf, BF16K kb, dim3 grid, int threads, cudaStream_t st, Args... args) {
  return dtype == N2N_BF16 ? launch_pdl_v(kb, grid, dim3(threads), 0, st, args...) : launch_pdl_v(kf, grid, dim3(threads), 0, st, args...);
}

}  // namespace n2n

using namespace n2n;

// ------------------------------------------------------------------------------------------------------------------------
struct ImpBuf { size_t off = 0; int cb = 0, h = 0, w = 0; };
struct Win { int buf = -1, cb0 = 0, cb = 0; };         // blocks [cb0, cb0 + cb) of buffer `buf`

enum { OP_CONV = 0, OP_GN, OP_POOL, OP_PSHUF, OP_COPY, OP_SIGMA, OP_FINAL };
struct ImpOp {                                          // one forward operator, as the backward needs it
  int kind = OP_CONV;
  Win x, y, res;                                        // input / output / residual (conv addend, GroupNorm residual)
  bool has_res = false, has_bias = false, want_dgrad = true;
  int c0 = 0, c1 = 0, cout = 0, k = 3, pidx = 0;        // conv: input segments, outputs, kernel, first parameter
  float slope = -1.f;
  int C = 0, stat = -1;                                 // GroupNorm: channels, slot of the saved (mean, rstd)
  int wg0 = 0, dg0 = 0;                                 // conv: first weight-gradient / input-gradient chunk in the plan's tables
};

struct n2n_improved_plan {
  int in_nc, out_nc, nf, depth, noise, N, H, W, dtype;
  bool train = false;
  std::vector<ImpBuf> bufs;
  std::vector<size_t> conv_wp, conv_bias;     // per convolution chunk: offsets of its packed weights / padded bias
  std::vector<size_t> gn_stat;                // training: per GroupNorm, offset of its saved (mean, rstd) [N][Cb*16][2]
  std::vector<ImpOp> tape;                    // training: the forward operators in execution order
  // training: per weight-gradient chunk (dY <= 8 blocks x X <= 9 blocks) its partial / bias-partial offsets (relative to
  // off_partial / off_bpartial; chunks of ONE conv do not overlap, convs reuse the region) and pixel splits; per
  // input-gradient chunk (<= 16 blocks of Cin) the offset of its packed transposed weights (all packed up front)
  std::vector<size_t> wg_partial, wg_bpartial, dg_wd;
  std::vector<int> wg_splits;
  // executions captured as CUDA graphs, keyed by every pointer the launch sequence touches (a caller that keeps its
  // buffers in place pays the ~170 / ~540 launches and their tensor-map encodes once)
  struct Captured { unsigned long long key; cudaGraphExec_t exec; int launches; };
  std::vector<Captured> graphs;
  cudaStream_t cap_stream = nullptr;          // private stream the sequences are captured on (the caller's may be the legacy stream)
  ~n2n_improved_plan() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
  }
  size_t act_bytes = 0;                       // activations occupy [0, act_bytes); training: gradients [off_grad, off_grad + act_bytes)
  size_t off_wp = 0, off_bias = 0, off_gn_partial = 0, off_gn_ss = 0, off_f32 = 0, off_grad = 0, off_stat = 0, off_wd = 0,
         off_partial = 0, off_bpartial = 0, off_f32b = 0, total = 0;
  int launches = 0, bwd_launches = 0;
  int nparams = 0;
};

namespace {

// channel segments of a conv input restricted to padded blocks [blk0, blk0 + nblk): padded index -> real channel
Segs chunk_segs(int c0, int c1, int blk0, int nblk) {
  Segs s; s.n = 0;
  const int lo = 16 * blk0, hi = 16 * (blk0 + nblk);
  const int ps[2] = {0, cblocks(c0) * 16}, rs[2] = {0, c0}, rc[2] = {c0, c1};
  for (int i = 0; i < 2; ++i) {
    if (rc[i] <= 0) continue;
    const int a = ps[i] > lo ? ps[i] : lo, b = ps[i] + rc[i] < hi ? ps[i] + rc[i] : hi;
    if (a >= b) continue;
    s.src0[s.n] = rs[i] + (a - ps[i]); s.cnt[s.n] = b - a; s.dst0[s.n] = a - lo; s.off[s.n] = 0;
    ++s.n;
  }
  if (s.n == 0) { s.n = 1; s.src0[0] = 0; s.cnt[0] = 0; s.dst0[0] = 0; }
  return s;
}

enum { IMP_SIZE = 0, IMP_PACK = 1, IMP_EXEC = 2 };
struct Builder {      // walks the network: sizing (plan creation), then per forward once to collect every layer's weight
                      // pack job (a handful of batched launches up front) and once to launch the layers
  n2n_improved_plan* p;
  const float* const* prm;
  char* ws;
  cudaStream_t st;
  int mode = IMP_SIZE;
  int pi = 0;                       // next parameter (state_dict order, arch_unet.py:476-513)
  int ci = 0;                       // next convolution chunk
  int gi = 0;                       // next GroupNorm
  int nbuf_run = 0;
  size_t off = 0, wp_total = 0, bias_total = 0, stat_total = 0, gn_partial_max = 0, gn_ss_max = 0;
  std::vector<PackJob> pack_jobs;
  std::vector<BiasPadJob> bias_jobs;
  bool run() const { return mode == IMP_EXEC; }
  bool taping() const { return p->train && mode == IMP_SIZE; }

  int new_buf(int channels_blocks, int h, int w) {
    if (mode == IMP_SIZE) {
      ImpBuf b; b.off = off; b.cb = channels_blocks; b.h = h; b.w = w;
      off += align_up((size_t)p->N * channels_blocks * h * w * 16 * dtype_size(p->dtype), 1024);
      p->bufs.push_back(b);
      return (int)p->bufs.size() - 1;
    }
    return nbuf_run++;
  }
  static Win win(int b, int cb0, int cb) { Win w; w.buf = b; w.cb0 = cb0; w.cb = cb; return w; }
  View view(const Win& w) const {
    const ImpBuf& B = p->bufs[w.buf];
    return make_view(ws ? ws + B.off : nullptr, p->dtype, p->N, B.h, B.w, B.cb, w.cb0, w.cb);
  }
  const float* param(int i) const { return mode == IMP_SIZE ? nullptr : prm[i]; }

  // conv (k = 3 / 1) over a one- or two-segment input window, optional bias / LeakyReLU / residual addend
  int conv(const Win& xw, int c0, int c1, int cout, int k, bool has_bias, const Win& yw, float slope, const Win* addend,
           bool want_dgrad = true) {
    if (taping()) {
      ImpOp op; op.kind = OP_CONV; op.x = xw; op.y = yw; op.has_res = addend != nullptr; if (addend) op.res = *addend;
      op.has_bias = has_bias; op.want_dgrad = want_dgrad; op.c0 = c0; op.c1 = c1; op.cout = cout; op.k = k; op.pidx = pi; op.slope = slope;
      p->tape.push_back(op);
    }
    const float* w = param(pi);
    const float* b = has_bias ? param(pi + 1) : nullptr;
    pi += has_bias ? 2 : 1;
    LayerGeom L;
    L.kind = k == 3 ? L_CONV3 : L_CONV1;
    L.cin = c1 > 0 ? chan2(c0, c1) : chan1(c0);
    const int cin_real = c0 + c1;
    for (int n0 = 0; n0 < cout; n0 += 256) {
      const int nc = cout - n0 < 256 ? cout - n0 : 256;
      L.cout = nc;
      if (mode == IMP_SIZE) {
        p->conv_wp.push_back(wp_total); p->conv_bias.push_back(bias_total);
        wp_total += align_up(L.fwd_pack_bytes(p->dtype), 1024);
        bias_total += align_up((size_t)L.cout_blocks() * 16 * sizeof(float), 256);
        continue;
      }
      void* wp = ws + p->off_wp + p->conv_wp[ci];
      float* bias = (float*)(ws + p->off_bias + p->conv_bias[ci]);
      ++ci;
      if (mode == IMP_PACK) {
        pack_jobs.push_back(make_fwd_pack(L, w + (size_t)n0 * cin_real * k * k, wp));
        bias_jobs.push_back(BiasPadJob{b ? b + n0 : nullptr, bias, nc, L.cout_blocks() * 16});
        continue;
      }
      const View x = view(xw), y = view(yw);
      TapGemm g = make_conv_fwd(L, p->dtype, x, sub_blocks(y, p->dtype, n0 / 16, L.cout_blocks()), wp, bias);
      if (slope >= 0.f) { g.act = 1; g.slope = slope; }
      if (addend) { g.has_addend = true; g.addend = sub_blocks(view(*addend), p->dtype, n0 / 16, L.cout_blocks()); }
      N2N_TRY(launch_tapgemm(g, st));
    }
    return 0;
  }

  static int gn_groups(int C) { int g = C < 32 ? C : 32; while (g > 1 && C % g) --g; return g; }   // arch_unet.py:12-14
  int gn_splits(const ImpBuf& B, int cb) const {
    const long long hw = (long long)B.h * B.w;
    int splits = (int)((kSMs * 2 + (long long)p->N * cb - 1) / ((long long)p->N * cb));
    if (splits > 64) splits = 64;
    while (splits > 1 && hw / splits < 512) --splits;
    return splits;
  }
  int groupnorm(const Win& xw, int C, const Win& yw, float slope, const Win* res) {
    const int slot = gi++;
    if (taping()) {
      ImpOp op; op.kind = OP_GN; op.x = xw; op.y = yw; op.has_res = res != nullptr; if (res) op.res = *res;
      op.pidx = pi; op.slope = slope; op.C = C; op.stat = slot;
      p->tape.push_back(op);
    }
    const float* gamma = param(pi);
    const float* beta = param(pi + 1);
    pi += 2;
    const int cg = C / gn_groups(C);
    const ImpBuf& B = p->bufs[xw.buf];
    const long long hw = (long long)B.h * B.w;
    const int splits = gn_splits(B, xw.cb);
    const size_t pb = (size_t)p->N * xw.cb * splits * 32 * sizeof(double), sb = (size_t)p->N * xw.cb * 16 * 2 * sizeof(float);
    if (pb > gn_partial_max) gn_partial_max = pb;
    if (sb > gn_ss_max) gn_ss_max = sb;
    if (mode == IMP_SIZE && p->train) { p->gn_stat.push_back(stat_total); stat_total += align_up(sb, 256); }
    if (!run()) return 0;
    const View x = view(xw), y = view(yw);
    double* partial = (double*)(ws + p->off_gn_partial);
    float* ss = (float*)(ws + p->off_gn_ss);
    float* mr = p->train ? (float*)(ws + p->off_stat + p->gn_stat[slot]) : nullptr;
    (void)launch_by_dtype(p->dtype, c16_gn_stats_kernel<float>, c16_gn_stats_kernel<__nv_bfloat16>, dim3(splits, x.Cb, x.N), 256, st, x, splits, partial);
    N2N_LAUNCH_CHECK();
    (void)launch_pdl_v(c16_gn_fold_kernel, dim3((x.N * x.Cb * 16 + 127) / 128), dim3(128), 0, st, (const double*)partial, splits, x.N, C,
                       x.Cb, cg, hw, 1e-5f, gamma, beta, ss, mr);
    N2N_LAUNCH_CHECK();
    const long long items = (long long)x.N * x.Cb * hw;
    (void)launch_by_dtype(p->dtype, c16_gn_apply_kernel<float>, c16_gn_apply_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st, x,
                          res ? view(*res) : x, res ? 1 : 0, y, (const float*)ss, slope, items);
    N2N_LAUNCH_CHECK();
    return 0;
  }

  // RDB (arch_unet.py:435-449): `R` holds [x | o1 | o2 | o3 | o4]; x is already in blocks [0, xb); result -> `out`
  int rdb(int R, int C, const Win& out) {
    const int xb = cblocks(C);
    for (int j = 0; j < 4; ++j)
      N2N_TRY(conv(win(R, 0, xb + 2 * j), C, 32 * j, 32, 3, true, win(R, xb + 2 * j, 2), 0.2f, nullptr));
    const Win xv = win(R, 0, xb);
    return conv(win(R, 0, xb + 8), C, 128, C, 1, true, out, -1.f, &xv);                   // x + lff(cat)
  }
  // ResBlock (:420-432): x (`xin`) -> `out`.  No-grad plans reuse two scratch buffers; training plans keep all three tensors.
  int resblock(const Win& xin, int C, int h, int w, const Win& out) {
    const int xb = cblocks(C);
    const int t1 = new_buf(xb, h, w), t2 = new_buf(xb, h, w), t3 = p->train ? new_buf(xb, h, w) : t1;
    N2N_TRY(conv(xin, C, 0, C, 3, false, win(t1, 0, xb), -1.f, nullptr));
    N2N_TRY(groupnorm(win(t1, 0, xb), C, win(t2, 0, xb), 0.2f, nullptr));
    N2N_TRY(conv(win(t2, 0, xb), C, 0, C, 3, false, win(t3, 0, xb), -1.f, nullptr));
    return groupnorm(win(t3, 0, xb), C, out, -1.f, &xin);
  }
  void tape_simple(int kind, const Win& x, const Win& y, int C = 0) {
    if (!taping()) return;
    ImpOp op; op.kind = kind; op.x = x; op.y = y; op.C = C;
    p->tape.push_back(op);
  }

  int network(const float* x, float* y) {
    const int N = p->N, H = p->H, W = p->W, dt = p->dtype, in_nc = p->in_nc, nf0 = p->nf, D = p->depth;
    // input block: [x (in_nc) | sigma (1)]
    const int IN = new_buf(1, H, W);
    if (run()) N2N_TRY(launch_nchw_to_c16(x, in_nc, view(win(IN, 0, 1)), dt, st));
    if (p->noise) {
      const int NE1 = new_buf(cblocks(nf0), H, W), NE2 = new_buf(1, H, W);
      N2N_TRY(conv(win(IN, 0, 1), in_nc, 0, nf0, 3, true, win(NE1, 0, cblocks(nf0)), 0.2f, nullptr, false));
      N2N_TRY(conv(win(NE1, 0, cblocks(nf0)), nf0, 0, 1, 3, true, win(NE2, 0, 1), -1.f, nullptr));
      tape_simple(OP_SIGMA, win(NE2, 0, 1), win(IN, 0, 1));
      if (run()) {
        float* sig = (float*)(ws + p->off_f32);        // kept for the backward (sigmoid')
        N2N_TRY(launch_c16_to_nchw(view(win(NE2, 0, 1)), dt, sig, 1, st));
        N2N_TRY(n2n_act_fwd(sig, sig, (int64_t)N * H * W, 2, 0.f, st));
        const long long items = (long long)N * H * W;
        (void)launch_by_dtype(dt, c16_set_lane_kernel<float>, c16_set_lane_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st,
                              (const float*)sig, view(win(IN, 0, 1)), in_nc, items);
        N2N_LAUNCH_CHECK();
      }
    }
    // the fuse-concat buffers of the decoder, [PixelShuffle output | encoder skip], exist before the encoder writes the skips
    std::vector<int> F(D), skip_c(D);
    int nf = nf0;
    for (int i = 0; i < D; ++i) { F[i] = new_buf(cblocks(nf / 2) + cblocks(nf), H >> i, W >> i); skip_c[i] = nf; nf *= 2; }
    // encoder
    int cur = IN, cur_c = p->noise ? in_nc + 1 : 1;
    nf = nf0;
    for (int i = 0; i < D; ++i) {
      const int h = H >> i, w = W >> i, xb = cblocks(nf);
      const int R = new_buf(xb + 8, h, w), A = new_buf(xb, h, w);
      N2N_TRY(conv(win(cur, 0, cblocks(cur_c)), cur_c, 0, nf, 3, true, win(R, 0, xb), 0.2f, nullptr, i > 0 || p->noise != 0));
      N2N_TRY(rdb(R, nf, win(A, 0, xb)));
      const Win skip = win(F[i], cblocks(nf / 2), xb);
      N2N_TRY(resblock(win(A, 0, xb), nf, h, w, skip));
      const int P = new_buf(xb, h / 2, w / 2);
      tape_simple(OP_POOL, skip, win(P, 0, xb));
      if (run()) N2N_TRY(launch_maxpool(view(skip), view(win(P, 0, xb)), dt, st));
      cur = P; cur_c = nf;
      nf *= 2;
    }
    nf /= 2;       // bottleneck width
    {
      const int h = H >> D, w = W >> D, xb = cblocks(nf);
      const int R = new_buf(xb + 8, h, w), A = new_buf(xb, h, w), B = new_buf(xb, h, w);
      tape_simple(OP_COPY, win(cur, 0, xb), win(R, 0, xb));
      if (run()) {   // RDB input must sit in R[0:xb): copy the pooled tensor (a D2D copy of the smallest level)
        const View src = view(win(cur, 0, xb)), dst = view(win(R, 0, xb));
        const size_t es = dtype_size(dt), row = (size_t)xb * h * w * 16 * es;
        N2N_CUDA(cudaMemcpy2DAsync(dst.ptr, (size_t)dst.sN * es, src.ptr, (size_t)src.sN * es, row, N, cudaMemcpyDeviceToDevice, st));
      }
      N2N_TRY(rdb(R, nf, win(A, 0, xb)));
      N2N_TRY(resblock(win(A, 0, xb), nf, h, w, win(B, 0, xb)));
      cur = B; cur_c = nf;
    }
    // decoder
    const int FIN = new_buf(cblocks(nf0 / 2) + 1, H, W);
    for (int j = 0; j < D; ++j) {
      const int lvl = D - 1 - j, h = H >> lvl, w = W >> lvl, oc = cur_c / 2, ob = cblocks(oc);
      const int PS = new_buf(cblocks(4 * oc), h / 2, w / 2);
      N2N_TRY(conv(win(cur, 0, cblocks(cur_c)), cur_c, 0, 4 * oc, 3, true, win(PS, 0, cblocks(4 * oc)), -1.f, nullptr));
      tape_simple(OP_PSHUF, win(PS, 0, cblocks(4 * oc)), win(F[lvl], 0, ob), oc);
      if (run()) {
        const long long items = (long long)N * ob * (h / 2) * (w / 2);
        (void)launch_by_dtype(dt, c16_pixel_shuffle2_kernel<float>, c16_pixel_shuffle2_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st,
                              view(win(PS, 0, cblocks(4 * oc))), view(win(F[lvl], 0, ob)), oc, items);
        N2N_LAUNCH_CHECK();
      }
      const int R = new_buf(ob + 8, h, w), A = new_buf(ob, h, w);
      N2N_TRY(conv(win(F[lvl], 0, ob + cblocks(skip_c[lvl])), oc, skip_c[lvl], oc, 3, true, win(R, 0, ob), 0.2f, nullptr));
      N2N_TRY(rdb(R, oc, win(A, 0, ob)));
      int O = -1;
      Win out;
      if (j == D - 1) out = win(FIN, 0, ob);
      else { O = new_buf(ob, h, w); out = win(O, 0, ob); }
      N2N_TRY(resblock(win(A, 0, ob), oc, h, w, out));
      cur = O; cur_c = oc;
    }
    // final: sigmoid(conv(cat[x, orig]))
    const int fb = cblocks(nf0 / 2);
    if (run()) N2N_TRY(launch_nchw_to_c16(x, in_nc, view(win(FIN, fb, 1)), dt, st));
    const int OUT = new_buf(cblocks(p->out_nc), H, W);
    N2N_TRY(conv(win(FIN, 0, fb + 1), nf0 / 2, in_nc, p->out_nc, 3, true, win(OUT, 0, cblocks(p->out_nc)), -1.f, nullptr));
    tape_simple(OP_FINAL, win(OUT, 0, cblocks(p->out_nc)), win(OUT, 0, cblocks(p->out_nc)));
    if (run()) {
      N2N_TRY(launch_c16_to_nchw(view(win(OUT, 0, cblocks(p->out_nc))), dt, y, p->out_nc, st));
      N2N_TRY(n2n_act_fwd(y, y, (int64_t)N * p->out_nc * H * W, 2, 0.f, st));
    }
    return 0;
  }
};

// engine limits of one weight-gradient launch: dY (common operand) <= 8 blocks, X (variant) <= 9 blocks (the widest geometry the
// UNet plans exercise); one input-gradient launch: N = Cin <= 16 blocks
constexpr int kWgDyBlocks = 8, kWgXBlocks = 9, kDgBlocks = 16;

struct Backward {
  n2n_improved_plan* p;
  const float* const* prm;
  float* const* grads;
  char* ws;
  cudaStream_t st;
  View act(const Win& w) const {
    const ImpBuf& B = p->bufs[w.buf];
    return make_view(ws + B.off, p->dtype, p->N, B.h, B.w, B.cb, w.cb0, w.cb);
  }
  View grd(const Win& w) const {
    const ImpBuf& B = p->bufs[w.buf];
    return make_view(ws + p->off_grad + B.off, p->dtype, p->N, B.h, B.w, B.cb, w.cb0, w.cb);
  }
  static long long items_of(const View& v) { return (long long)v.N * v.Cb * v.H * v.W; }
  int add_into(const View& src, const View& dst) {
    const long long items = items_of(dst);
    (void)launch_by_dtype(p->dtype, c16_add_into_kernel<float>, c16_add_into_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st, src, dst, items);
    N2N_LAUNCH_CHECK();
    return 0;
  }

  int conv(const ImpOp& op) {
    const int dt = p->dtype, k = op.k, cin_real = op.c0 + op.c1, xb = op.x.cb, yb = cblocks(op.cout);
    const View x = act(op.x), y = act(op.y), gy = grd(op.y), gx = grd(op.x);
    if (op.slope >= 0.f) {        // every consumer has added its share: apply this layer's LeakyReLU' once, in place
      const long long items = items_of(gy);
      (void)launch_by_dtype(dt, c16_lrelu_mask_kernel<float>, c16_lrelu_mask_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st, y, gy,
                            op.slope, items);
      N2N_LAUNCH_CHECK();
    }
    if (op.has_res) N2N_TRY(add_into(gy, grd(op.res)));
    float* dw = grads[op.pidx];
    float* db = op.has_bias ? grads[op.pidx + 1] : nullptr;
    // weight gradient: engine-sized chunks into disjoint partial regions, then ONE fixed-order reduction launch for the layer
    std::vector<UnpackJob> jobs;
    int wi = op.wg0;
    for (int o0 = 0; o0 < yb; o0 += kWgDyBlocks) {
      const int ob = yb - o0 < kWgDyBlocks ? yb - o0 : kWgDyBlocks;
      for (int c0 = 0; c0 < xb; c0 += kWgXBlocks, ++wi) {
        const int cb = xb - c0 < kWgXBlocks ? xb - c0 : kWgXBlocks;
        LayerGeom L; L.kind = k == 3 ? L_CONV3 : L_CONV1; L.cin = chan1(cb * 16); L.cout = ob * 16;
        const int splits = p->wg_splits[wi];
        float* partial = (float*)(ws + p->off_partial + p->wg_partial[wi]);
        float* bpartial = (float*)(ws + p->off_bpartial + p->wg_bpartial[wi]);
        const bool bias_here = db != nullptr && c0 == 0;
        TapWgrad g = make_conv_wgrad(L, dt, sub_blocks(x, dt, c0, cb), sub_blocks(gy, dt, o0, ob), partial, bias_here ? bpartial : nullptr, splits);
        N2N_TRY(launch_tapwgrad(g, st));
        UnpackJob j;
        j.partial = partial; j.bias_partial = bias_here ? bpartial : nullptr; j.dst_w = dw; j.dst_b = bias_here ? db : nullptr;
        j.splits = splits; j.ntaps = L.ntaps(); j.npad = ob * 16; j.cpad = cb * 16; j.bias_rows = splits;
        if (k == 3) { j.s_t = 1; j.s_n = (long long)cin_real * 9; j.s_c = 9; } else { j.s_t = 0; j.s_n = cin_real; j.s_c = 1; }
        j.nseg.n = 1; j.nseg.src0[0] = 16 * o0; j.nseg.cnt[0] = op.cout - 16 * o0 < 16 * ob ? op.cout - 16 * o0 : 16 * ob; j.nseg.dst0[0] = 0;
        j.cseg = chunk_segs(op.c0, op.c1, c0, cb);
        jobs.push_back(j);
      }
    }
    N2N_TRY(launch_unpack(jobs.data(), (int)jobs.size(), st));
    if (!op.want_dgrad) return 0;
    // input gradient, accumulated into the (zero-initialised / partly filled) gradient window of the input
    LayerGeom L; L.kind = k == 3 ? L_CONV3 : L_CONV1; L.cin = op.c1 > 0 ? chan2(op.c0, op.c1) : chan1(op.c0); L.cout = op.cout;
    int di = op.dg0;
    for (int b0 = 0; b0 < xb; b0 += kDgBlocks, ++di) {
      const int nb = xb - b0 < kDgBlocks ? xb - b0 : kDgBlocks;
      const View dx = sub_blocks(gx, dt, b0, nb);
      TapGemm g = make_conv_dgrad(L, dt, gy, dx, ws + p->off_wd + p->dg_wd[di], nb);
      g.has_addend = true; g.addend = dx;
      N2N_TRY(launch_tapgemm(g, st));
    }
    return 0;
  }
  // all transposed weights of the input-gradient GEMMs -> engine layout, a few batched launches before the walk
  int pack_dgrad_weights() {
    std::vector<PackJob> jobs;
    for (const ImpOp& op : p->tape) {
      if (op.kind != OP_CONV || !op.want_dgrad) continue;
      LayerGeom L; L.kind = op.k == 3 ? L_CONV3 : L_CONV1; L.cin = op.c1 > 0 ? chan2(op.c0, op.c1) : chan1(op.c0); L.cout = op.cout;
      int di = op.dg0;
      for (int b0 = 0; b0 < op.x.cb; b0 += kDgBlocks, ++di) {
        const int nb = op.x.cb - b0 < kDgBlocks ? op.x.cb - b0 : kDgBlocks;
        PackJob pj = make_dgrad_pack(L, prm[op.pidx], ws + p->off_wd + p->dg_wd[di], nb);
        pj.nseg = chunk_segs(op.c0, op.c1, b0, nb);
        jobs.push_back(pj);
      }
    }
    return launch_pack(jobs.data(), (int)jobs.size(), p->dtype, st);
  }

  int groupnorm(const ImpOp& op) {
    const int dt = p->dtype, C = op.C;
    const View x = act(op.x), y = act(op.y), gy = grd(op.y), gx = grd(op.x);
    const ImpBuf& B = p->bufs[op.x.buf];
    const long long hw = (long long)B.h * B.w;
    const int cg = C / Builder::gn_groups(C);
    Builder tmp{p, nullptr, nullptr, nullptr};
    const int splits = tmp.gn_splits(B, op.x.cb);
    double* partial = (double*)(ws + p->off_gn_partial);
    float* coef = (float*)(ws + p->off_gn_ss);
    const float* mr = (const float*)(ws + p->off_stat + p->gn_stat[op.stat]);
    const float* gamma = prm[op.pidx];
    (void)launch_by_dtype(dt, c16_gn_bwd_stats_kernel<float>, c16_gn_bwd_stats_kernel<__nv_bfloat16>, dim3(splits, x.Cb, x.N), 256, st, x, y, gy, mr,
                          op.slope, splits, partial);
    N2N_LAUNCH_CHECK();
    const int nthreads = x.N * x.Cb * 16 > C ? x.N * x.Cb * 16 : C;
    (void)launch_pdl_v(c16_gn_bwd_fold_kernel, dim3((nthreads + 127) / 128), dim3(128), 0, st, (const double*)partial, splits, x.N, C, x.Cb, cg, hw,
                       gamma, coef, grads[op.pidx], grads[op.pidx + 1]);
    N2N_LAUNCH_CHECK();
    const long long items = items_of(x);
    (void)launch_by_dtype(dt, c16_gn_bwd_apply_kernel<float>, c16_gn_bwd_apply_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st, x, y, gy,
                          gx, op.has_res ? grd(op.res) : gx, op.has_res ? 1 : 0, mr, (const float*)coef, gamma, C, op.slope, items);
    N2N_LAUNCH_CHECK();
    return 0;
  }

  int run(const float* dy, const float* y_out) {
    const int dt = p->dtype, N = p->N, H = p->H, W = p->W;
    N2N_CUDA(cudaMemsetAsync(ws + p->off_grad, 0, p->act_bytes, st));
    N2N_TRY(pack_dgrad_weights());
    for (int i = (int)p->tape.size() - 1; i >= 0; --i) {
      const ImpOp& op = p->tape[i];
      switch (op.kind) {
        case OP_FINAL: {        // y = sigmoid(out): d out = dy * y * (1 - y)
          float* tmp = (float*)(ws + p->off_f32b);
          N2N_TRY(n2n_act_bwd(y_out, dy, tmp, (int64_t)N * p->out_nc * H * W, 2, 0.f, st));
          N2N_TRY(launch_nchw_to_c16(tmp, p->out_nc, grd(op.y), dt, st));
          break;
        }
        case OP_CONV: N2N_TRY(conv(op)); break;
        case OP_GN: N2N_TRY(groupnorm(op)); break;
        case OP_POOL: {
          const View a = act(op.x), gp = grd(op.y), gs = grd(op.x);
          const long long items = items_of(gp);
          (void)launch_by_dtype(dt, c16_unpool_acc_kernel<float>, c16_unpool_acc_kernel<__nv_bfloat16>, dim3(grid_for(items, 128)), 128, st, a, gp, gs, items);
          N2N_LAUNCH_CHECK();
          break;
        }
        case OP_PSHUF: {
          const View gd = grd(op.y), gs = grd(op.x);
          const long long items = (long long)gs.N * gd.Cb * gs.H * gs.W;
          (void)launch_by_dtype(dt, c16_pixel_unshuffle2_acc_kernel<float>, c16_pixel_unshuffle2_acc_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)),
                                256, st, gd, gs, op.C, items);
          N2N_LAUNCH_CHECK();
          break;
        }
        case OP_COPY: N2N_TRY(add_into(grd(op.y), grd(op.x))); break;
        case OP_SIGMA: {        // sigma = sigmoid(NE2) sits in lane in_nc of the input block
          float* tmp = (float*)(ws + p->off_f32b);
          const float* sig = (const float*)(ws + p->off_f32);
          const long long items = (long long)N * H * W;
          (void)launch_by_dtype(dt, c16_get_lane_kernel<float>, c16_get_lane_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), 256, st, grd(op.y),
                                p->in_nc, tmp, items);
          N2N_LAUNCH_CHECK();
          N2N_TRY(n2n_act_bwd(sig, tmp, tmp, (int64_t)items, 2, 0.f, st));
          N2N_TRY(launch_nchw_to_c16(tmp, 1, grd(op.x), dt, st));
          break;
        }
      }
    }
    return 0;
  }
};

}  // namespace

extern "C" int n2n_improved_plan_create(n2n_improved_plan** plan, int in_nc, int out_nc, int n_feature, int depth, int noise, int n,
                                        int h, int w, int dtype, int with_backward) {
  N2N_CHECK_ARG(plan && in_nc >= 1 && in_nc <= 15 && out_nc >= 1 && out_nc <= 16 && n_feature >= 2 && n_feature % 2 == 0 && depth >= 1 &&
                depth <= 6 && n >= 1 && h >= 1 && w >= 1, "improved_plan_create: bad arguments");
  N2N_CHECK_ARG(h % (1 << depth) == 0 && w % (1 << depth) == 0, "improved_plan_create: H and W must be multiples of 2^depth");
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, "improved_plan_create: bad dtype");
  N2N_CHECK_ARG(noise || in_nc == 1, "improved_plan_create: noise = 0 expects one input channel (arch_unet.py:496)");
  n2n_improved_plan* p = new n2n_improved_plan();
  p->in_nc = in_nc; p->out_nc = out_nc; p->nf = n_feature; p->depth = depth; p->noise = noise; p->N = n; p->H = h; p->W = w; p->dtype = dtype;
  p->train = with_backward != 0;
  Builder b{p, nullptr, nullptr, nullptr};
  const int r = b.network(nullptr, nullptr);
  if (r != 0) { delete p; return r; }
  size_t off = b.off;
  p->act_bytes = b.off;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  p->off_wp = take(b.wp_total); p->off_bias = take(b.bias_total);
  p->off_gn_partial = take(b.gn_partial_max); p->off_gn_ss = take(b.gn_ss_max);
  p->off_f32 = take((size_t)n * h * w * sizeof(float));
  if (p->train) {
    p->off_grad = take(p->act_bytes);
    p->off_stat = take(b.stat_total);
    p->off_f32b = take((size_t)n * (out_nc > 1 ? out_nc : 1) * h * w * sizeof(float));
    // weight-gradient chunks of one conv get disjoint partial regions (one reduction launch per conv), the region is reused by
    // the next conv; every input-gradient chunk of the network keeps its own packed weights (packed in one batch up front)
    size_t partial = 0, bpartial = 0, wd = 0;
    for (ImpOp& op : p->tape) {
      if (op.kind != OP_CONV) continue;
      const ImpBuf& B = p->bufs[op.x.buf];
      const int xb = op.x.cb, yb = cblocks(op.cout);
      op.wg0 = (int)p->wg_splits.size();
      size_t po = 0, bo = 0;
      for (int o0 = 0; o0 < yb; o0 += kWgDyBlocks)           // the same chunking as Backward::conv
        for (int c0 = 0; c0 < xb; c0 += kWgXBlocks) {
          LayerGeom L; L.kind = op.k == 3 ? L_CONV3 : L_CONV1;
          L.cin = chan1((xb - c0 < kWgXBlocks ? xb - c0 : kWgXBlocks) * 16); L.cout = (yb - o0 < kWgDyBlocks ? yb - o0 : kWgDyBlocks) * 16;
          const int splits = layer_wgrad_splits(L, dtype, n, B.h, B.w);
          p->wg_splits.push_back(splits); p->wg_partial.push_back(po); p->wg_bpartial.push_back(bo);
          po += align_up(L.partial_bytes(splits), 1024); bo += align_up(L.bias_partial_bytes(splits), 1024);
        }
      if (po > partial) partial = po;
      if (bo > bpartial) bpartial = bo;
      op.dg0 = (int)p->dg_wd.size();
      if (op.want_dgrad) {
        LayerGeom G; G.kind = op.k == 3 ? L_CONV3 : L_CONV1; G.cin = chan1(op.c0 + op.c1); G.cout = op.cout;
        for (int b0 = 0; b0 < xb; b0 += kDgBlocks) {
          p->dg_wd.push_back(wd);
          wd += align_up(G.dgrad_pack_bytes(dtype, xb - b0 < kDgBlocks ? xb - b0 : kDgBlocks), 1024);
        }
      }
    }
    p->off_partial = take(partial); p->off_bpartial = take(bpartial); p->off_wd = take(wd);
  }
  p->total = off;
  p->nparams = b.pi;
  *plan = p;
  return 0;
}
extern "C" void n2n_improved_plan_destroy(n2n_improved_plan* plan) { delete plan; }
extern "C" size_t n2n_improved_workspace_bytes(const n2n_improved_plan* plan) { return plan ? plan->total : 0; }
extern "C" int n2n_improved_num_params(const n2n_improved_plan* plan) { return plan ? plan->nparams : 0; }
extern "C" int n2n_improved_launches(const n2n_improved_plan* plan, int backward) { return plan ? (backward ? plan->bwd_launches : plan->launches) : 0; }

static int improved_forward_body(n2n_improved_plan* p, const float* const* params, const float* x, float* y, void* ws, cudaStream_t st);
namespace {
unsigned long long mix(unsigned long long h, const void* ptr) {
  h ^= (unsigned long long)(uintptr_t)ptr + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  return h;
}
// Run `body` (a launch sequence on `st`) through a CUDA graph cached under `key`; falls back to plain launches when the
// stream is already being captured, per-launch profiling is armed or N2N_IMPROVED_NO_GRAPH=1.
template <typename Body>
int run_captured(n2n_improved_plan* p, unsigned long long key, cudaStream_t st, int* launches, Body body) {
  static const char* const off = getenv("N2N_IMPROVED_NO_GRAPH");
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  N2N_CUDA(cudaStreamIsCapturing(st, &cs));
  if ((off && atoi(off)) || cs != cudaStreamCaptureStatusNone || profiling_active()) {
    const long long l0 = g_launch_count;
    N2N_TRY(body(st));
    *launches = (int)(g_launch_count - l0);
    return 0;
  }
  for (auto& g : p->graphs)
    if (g.key == key) {
      N2N_CUDA(cudaGraphLaunch(g.exec, st));
      *launches = g.launches;
      g_launch_count += g.launches;
      return 0;
    }
  const long long l0 = g_launch_count;
  if (!p->cap_stream) N2N_CUDA(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
  N2N_CUDA(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = body(p->cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(p->cap_stream, &graph);
  if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
  N2N_CUDA(e);
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  N2N_CUDA(ei);
  if (p->graphs.size() >= 8) { cudaGraphExecDestroy(p->graphs.front().exec); p->graphs.erase(p->graphs.begin()); }
  *launches = (int)(g_launch_count - l0);
  p->graphs.push_back({key, exec, *launches});
  N2N_CUDA(cudaGraphLaunch(exec, st));
  return 0;
}
}  // namespace

extern "C" int n2n_improved_forward(n2n_improved_plan* p, const float* const* params, const float* x, float* y, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && x && y && ws, "improved_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long key = mix(mix(mix(1, ws), x), y);
  for (int i = 0; i < p->nparams; ++i) key = mix(key, params[i]);
  return run_captured(p, key, st, &p->launches, [&](cudaStream_t s) -> int { return improved_forward_body(p, params, x, y, ws, s); });
}

static int improved_forward_body(n2n_improved_plan* p, const float* const* params, const float* x, float* y, void* ws, cudaStream_t st) {
  {   // every layer's weights -> engine layout, biases -> padded rows: a few batched launches for the whole network
    Builder pk{p, params, (char*)ws, st};
    pk.mode = IMP_PACK;
    N2N_TRY(pk.network(x, y));
    N2N_TRY(launch_pack(pk.pack_jobs.data(), (int)pk.pack_jobs.size(), p->dtype, st));
    N2N_TRY(launch_bias_pad(pk.bias_jobs.data(), (int)pk.bias_jobs.size(), st));
  }
  Builder b{p, params, (char*)ws, st};
  b.mode = IMP_EXEC;
  N2N_TRY(b.network(x, y));
  N2N_CHECK_ARG(b.pi == p->nparams, "improved_forward: walked %d parameters, plan has %d", b.pi, p->nparams);
  return 0;
}

// Debug / layer-level parity hook: buffer `buf` (activation, or its gradient mirror when grad != 0) as fp32 NCHW over all of its
// 16-channel blocks.  out == NULL: returns the element count and fills dims = {blocks * 16, h, w}.
extern "C" long long n2n_improved_read_buffer(const n2n_improved_plan* p, const void* ws, int buf, int grad, float* out, int* dims,
                                              void* stream) {
  if (!p || buf < 0 || buf >= (int)p->bufs.size() || (grad && !p->train)) return -1;
  const ImpBuf& B = p->bufs[buf];
  if (dims) { dims[0] = B.cb * 16; dims[1] = B.h; dims[2] = B.w; }
  const long long count = (long long)p->N * B.cb * 16 * B.h * B.w;
  if (!out) return count;
  if (!ws) return -1;
  const View v = make_view((char*)ws + (grad ? p->off_grad : 0) + B.off, p->dtype, p->N, B.h, B.w, B.cb, 0, B.cb);
  const int rc = launch_c16_to_nchw(v, p->dtype, out, B.cb * 16, (cudaStream_t)stream);
  return rc < 0 ? rc : count;
}

extern "C" int n2n_improved_backward(n2n_improved_plan* p, const float* const* params, const float* dy, const float* y,
                                     float* const* grads, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && dy && y && grads && ws, "improved_backward: null argument");
  N2N_CHECK_ARG(p->train, "improved_backward: plan was created without with_backward");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long key = mix(mix(mix(2, ws), dy), y);
  for (int i = 0; i < p->nparams; ++i) key = mix(mix(key, params[i]), grads[i]);
  return run_captured(p, key, st, &p->bwd_launches, [&](cudaStream_t s) -> int {
    Backward b{p, params, grads, (char*)ws, s};
    return b.run(dy, y);
  });
}
