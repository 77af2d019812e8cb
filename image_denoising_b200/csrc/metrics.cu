// metrics.cu — PSNR / SSIM reduction kernel (utils_eval.py:19-53) for batches of uint8 images.
// SSIM: separable 11-tap Gaussian (sigma 1.5) over the "valid" interior, five moments,
// float64 accumulation like the reference; PSNR from the exact integer sum of squared
// differences.  Each CTA owns a 32x8 tile: it stages the (8+10)x(32+10) input window of both
// images in shared memory once (HBM traffic = the two images), runs the horizontal pass into
// shared memory and the vertical pass from it, and publishes one partial per CTA; a second tiny
// kernel sums the partials in a fixed order (deterministic).
#include "common.cuh"

namespace n2n {

constexpr int TX = 32, TY = 8, R = 5, KS = 11;
__constant__ double c_gauss[KS];

__global__ void __launch_bounds__(256)
psnr_ssim_tile_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, int H, int W, int C,
                      double* __restrict__ partial /* [batch*C][tiles][2] */) {
  __shared__ double sa[TY + 2 * R][TX + 2 * R];
  __shared__ double sb[TY + 2 * R][TX + 2 * R];
  __shared__ double hp[5][TY + 2 * R][TX];
  __shared__ double red[2][8];
  const int plane = blockIdx.z;            // image * C + channel
  const int img = plane / C, ch = plane - img * C;
  const uint8_t* a = A + (long long)img * H * W * C + ch;
  const uint8_t* b = B + (long long)img * H * W * C + ch;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tid = threadIdx.x;
  // stage the input window (zero outside the image; those entries never reach a valid output)
  for (int i = tid; i < (TY + 2 * R) * (TX + 2 * R); i += 256) {
    const int yy = i / (TX + 2 * R), xx = i - yy * (TX + 2 * R);
    const int y = y0 + yy, x = x0 + xx;
    double va = 0.0, vb = 0.0;
    if (y < H && x < W) {
      va = (double)a[((long long)y * W + x) * C];
      vb = (double)b[((long long)y * W + x) * C];
    }
    sa[yy][xx] = va; sb[yy][xx] = vb;
  }
  __syncthreads();
  // squared error over the tile's own pixels (exact: integers < 2^53)
  double sq = 0.0;
  {
    const int yy = tid / TX, xx = tid - yy * TX;      // 256 threads == TY*TX pixels
    if (y0 + yy < H && x0 + xx < W) { const double d = sa[yy][xx] - sb[yy][xx]; sq = d * d; }
  }
  // horizontal pass
  for (int i = tid; i < (TY + 2 * R) * TX; i += 256) {
    const int yy = i / TX, xx = i - yy * TX;
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const double g = c_gauss[k], p = sa[yy][xx + k], q = sb[yy][xx + k];
      m1 += g * p; m2 += g * q; s11 += g * (p * p); s22 += g * (q * q); s12 += g * (p * q);
    }
    hp[0][yy][xx] = m1; hp[1][yy][xx] = m2; hp[2][yy][xx] = s11; hp[3][yy][xx] = s22; hp[4][yy][xx] = s12;
  }
  __syncthreads();
  double ss = 0.0;
  {
    const int yy = tid / TX, xx = tid - yy * TX;
    if (y0 + yy < H - 2 * R && x0 + xx < W - 2 * R) {
      double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const double g = c_gauss[k];
        m1 += g * hp[0][yy + k][xx]; m2 += g * hp[1][yy + k][xx];
        s11 += g * hp[2][yy + k][xx]; s22 += g * hp[3][yy + k][xx]; s12 += g * hp[4][yy + k][xx];
      }
      const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
      const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
      ss = ((2 * m12 + C1) * (2 * (s12 - m12) + C2)) / ((m11 + m22 + C1) * ((s11 - m11) + (s22 - m22) + C2));
    }
  }
  // block reduce (fixed order)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = sq; red[1][tid >> 5] = ss; }
  __syncthreads();
  if (tid == 0) {
    double tq = 0, ts = 0;
    for (int w = 0; w < 8; ++w) { tq += red[0][w]; ts += red[1][w]; }
    const long long tiles = (long long)gridDim.x * gridDim.y;
    double* o = partial + ((long long)plane * tiles + (long long)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    o[0] = tq; o[1] = ts;
  }
}

__global__ void psnr_ssim_final_kernel(const double* __restrict__ partial, long long tiles, int H, int W, int C,
                                       double* __restrict__ result) {
  // one warp per image
  const int img = blockIdx.x, lane = threadIdx.x;
  double sq = 0.0, ssim_mean = 0.0;
  for (int ch = 0; ch < C; ++ch) {
    const double* p = partial + ((long long)(img * C + ch)) * tiles * 2;
    double q = 0, s = 0;
    for (long long t = lane; t < tiles; t += 32) { q += p[2 * t]; s += p[2 * t + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { q += __shfl_xor_sync(0xffffffffu, q, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
    sq += q;
    ssim_mean += s / ((double)(H - 2 * R) * (double)(W - 2 * R));
  }
  if (lane == 0) {
    const double mse = sq / ((double)H * W * C);
    result[2 * img] = 10.0 * log10(255.0 * 255.0 / mse);      // +inf when identical (utils_eval.py:52)
    result[2 * img + 1] = ssim_mean / C;
  }
}

static int upload_gauss(cudaStream_t st) {
  static bool done[64] = {false};
  int dev = 0;
  N2N_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return 0;
  double g[KS], s = 0;
  for (int i = 0; i < KS; ++i) { const double d = i - (KS - 1) / 2.0; g[i] = exp(-(d * d) / (2.0 * 1.5 * 1.5)); s += g[i]; }
  for (int i = 0; i < KS; ++i) g[i] /= s;     // cv2.getGaussianKernel(11, 1.5)
  N2N_CUDA(cudaMemcpyToSymbolAsync(c_gauss, g, sizeof(g), 0, cudaMemcpyHostToDevice, st));
  N2N_CUDA(cudaStreamSynchronize(st));
  if (dev < 64) done[dev] = true;
  return 0;
}

}  // namespace n2n

using namespace n2n;

static inline long long ps_tiles(int h, int w) { return (long long)((w + TX - 1) / TX) * ((h + TY - 1) / TY); }

extern "C" size_t n2n_psnr_ssim_workspace_bytes(int batch, int h, int w, int channels) {
  return (size_t)batch * channels * ps_tiles(h, w) * 2 * sizeof(double);
}

extern "C" int n2n_psnr_ssim_u8(const uint8_t* a, const uint8_t* b, int batch, int h, int w, int channels,
                                double* result, void* workspace, void* stream) {
  N2N_CHECK_ARG(a && b && result && workspace, "psnr_ssim: null pointer");
  N2N_CHECK_ARG(batch > 0 && h > 2 * R && w > 2 * R && (channels == 1 || channels == 3),
                "psnr_ssim: need batch>0, h,w>10, channels in {1,3} (got %d,%d,%d,%d)", batch, h, w, channels);
  N2N_CHECK_ARG((long long)batch * channels <= 65535, "psnr_ssim: batch*channels too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  N2N_TRY(upload_gauss(st));
  dim3 grid((w + TX - 1) / TX, (h + TY - 1) / TY, batch * channels);
  psnr_ssim_tile_kernel<<<grid, 256, 0, st>>>(a, b, h, w, channels, (double*)workspace);
  N2N_LAUNCH_CHECK();
  psnr_ssim_final_kernel<<<batch, 32, 0, st>>>((const double*)workspace, ps_tiles(h, w), h, w, channels, result);
  N2N_LAUNCH_CHECK();
  return 0;
}
