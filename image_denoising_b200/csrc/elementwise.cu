// elementwise.cu — HBM-bound layout / pooling / evaluation kernels on the C16 layout.
#include "common.cuh"
#include "layers.cuh"

namespace n2n {

// ------------------------------------------------------------------------------------------
// NCHW fp32  <->  one C16 block (C <= 16 channels, zero padded)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_c16_kernel(const float* __restrict__ src, int C, View dst, long long items) {
  pdl_enter();
  const long long hw = (long long)dst.H * dst.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % hw;
    const int cb = (int)((i / hw) % dst.Cb);
    const int n = (int)(i / (hw * dst.Cb));
    const int y = (int)(r / dst.W), x = (int)(r - (long long)y * dst.W);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int ch = cb * 16 + c;
      v[c] = (ch < C) ? src[((long long)n * C + ch) * hw + r] : 0.f;
    }
    T* p = (T*)dst.ptr + n * dst.sN + cb * dst.sCb + y * dst.sY + x * dst.sX;
    Block16<T>::store(p, v);
  }
}

template <typename T>
__global__ void c16_to_nchw_kernel(View src, float* __restrict__ dst, int C, long long items) {
  pdl_enter();
  const long long hw = (long long)src.H * src.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % hw;
    const int cb = (int)((i / hw) % src.Cb);
    const int n = (int)(i / (hw * src.Cb));
    const int y = (int)(r / src.W), x = (int)(r - (long long)y * src.W);
    float v[16];
    const T* p = (const T*)src.ptr + n * src.sN + cb * src.sCb + y * src.sY + x * src.sX;
    Block16<T>::load(p, v);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int ch = cb * 16 + c;
      if (ch < C) dst[((long long)n * C + ch) * hw + r] = v[c];
    }
  }
}

template <typename T, int CT>
__global__ void nchw_to_im2col9_kernel(const float* __restrict__ src, int Crt, View dst, long long items) {
  pdl_enter();
  const int C = CT > 0 ? CT : Crt;        // compile-time channel count (1, 3) keeps the tap decode division-free
  const int H = dst.H, W = dst.W;
  const long long hw = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % hw;
    const int cb = (int)((i / hw) % dst.Cb);
    const int n = (int)(i / (hw * dst.Cb));
    const int y = (int)(r / W), x = (int)(r - (long long)y * W);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int k = cb * 16 + c;
      const int tap = k / C, ch = k - tap * C;
      float val = 0.f;
      if (tap < 9) {
        const int sy = y + tap / 3 - 1, sx = x + tap % 3 - 1;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) val = src[((long long)n * C + ch) * hw + (long long)sy * W + sx];
      }
      v[c] = val;
    }
    T* p = (T*)dst.ptr + n * dst.sN + cb * dst.sCb + y * dst.sY + x * dst.sX;
    Block16<T>::store(p, v);
  }
}

int launch_nchw_to_im2col9(const float* src, int C, const View& dst, int dtype, cudaStream_t st) {
  N2N_CHECK_ARG(C >= 1 && 9 * C <= 16 * dst.Cb, "nchw_to_im2col9: 9*%d channels do not fit %d blocks", C, dst.Cb);
  long long items = (long long)dst.N * dst.Cb * dst.H * dst.W;
  int grid = grid_for(items, 256);
  if (dtype == N2N_BF16) {
    if (C == 1) (void)launch_pdl_v(nchw_to_im2col9_kernel<__nv_bfloat16, 1>, dim3(grid), dim3(256), 0, st, src, C, dst, items);
    else if (C == 3) (void)launch_pdl_v(nchw_to_im2col9_kernel<__nv_bfloat16, 3>, dim3(grid), dim3(256), 0, st, src, C, dst, items);
    else (void)launch_pdl_v(nchw_to_im2col9_kernel<__nv_bfloat16, 0>, dim3(grid), dim3(256), 0, st, src, C, dst, items);
  } else {
    (void)launch_pdl_v(nchw_to_im2col9_kernel<float, 0>, dim3(grid), dim3(256), 0, st, src, C, dst, items);
  }
  N2N_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Input stage of the bf16 UNet plan: one pass over the fp32 NCHW input writes BOTH the 3x3 im2col
// block(s) (dec_conv1a's raw-input operand, enc_conv0's weight-gradient operand) and enc_conv0's
// activated output (conv3x3 in_nc -> Cout + bias + LeakyReLU, arch_unet.py:114-116 / :201) as C16
// bf16.  The conv has K = 9*in_nc <= 27: it is HBM-bound (write 32 B per 16 output channels per
// pixel), so it runs on the CUDA cores in fp32 straight from the fp32 weights instead of going
// through a K=16-padded tensor-core GEMM and a second pass over the im2col block.
// ------------------------------------------------------------------------------------------
// PX horizontally adjacent pixels per thread: every weight quad fetched from shared memory (one LDS.128) feeds 4 * PX FMAs.
// With one pixel per thread the RGB form (K = 27) was bound by those loads (4 LDS.128 per 16 FMAs), not by the HBM writes.
template <int C, int PX>
__global__ void __launch_bounds__(128)
input_stage_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                   View col, View e0, float slope, long long groups) {
  pdl_enter();
  constexpr int K = 9 * C;
  __shared__ __align__(16) float s_w[K * 64];      // [k = tap*C + c][co], co padded to 64
  __shared__ float s_b[64];
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) {
    const int k = i >> 6, co = i & 63;
    const int tap = k / C, c = k - tap * C;
    s_w[i] = co < Cout ? w[((long long)co * C + c) * 9 + tap] : 0.f;
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_b[i] = i < Cout ? bias[i] : 0.f;
  __syncthreads();
  const int H = e0.H, W = e0.W, WG = W / PX;
  const long long hw = (long long)H * W, ghw = (long long)H * WG;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(g / ghw);
    const int r = (int)(g - (long long)n * ghw);
    const int y = r / WG, x0 = (r - y * WG) * PX;
    float win[C][3][PX + 2];                          // the (PX + 2) x 3 input window of the group, zero outside the image
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int sy = y + ky - 1;
        const float* row = x + ((long long)n * C + c) * hw + (long long)sy * W;
#pragma unroll
        for (int j = 0; j < PX + 2; ++j) {
          const int sx = x0 + j - 1;
          win[c][ky][j] = (sy >= 0 && sy < H && sx >= 0 && sx < W) ? __ldg(row + sx) : 0.f;
        }
      }
    // im2col block(s): element k = tap * C + c of pixel px is win[c][tap / 3][px + tap % 3]
#pragma unroll
    for (int px = 0; px < PX; ++px)
#pragma unroll
      for (int cb = 0; cb < (K + 15) / 16; ++cb) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int k = cb * 16 + q, kk = k < K ? k : 0;
          v[q] = k < K ? win[kk % C][(kk / C) / 3][px + (kk / C) % 3] : 0.f;
        }
        Block16<__nv_bfloat16>::store((__nv_bfloat16*)col.ptr + n * col.sN + cb * col.sCb + y * col.sY + (x0 + px) * col.sX, v);
      }
    // enc_conv0 + bias + LeakyReLU
    for (int cb = 0; cb < e0.Cb; ++cb) {
      float acc[PX][16];
#pragma unroll
      for (int px = 0; px < PX; ++px)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[px][q] = s_b[cb * 16 + q];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4* wr = reinterpret_cast<const float4*>(&s_w[k * 64 + cb * 16]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
#pragma unroll
          for (int px = 0; px < PX; ++px) {
            const float a = win[k % C][(k / C) / 3][px + (k / C) % 3];
            acc[px][4 * q] = fmaf(a, wv.x, acc[px][4 * q]); acc[px][4 * q + 1] = fmaf(a, wv.y, acc[px][4 * q + 1]);
            acc[px][4 * q + 2] = fmaf(a, wv.z, acc[px][4 * q + 2]); acc[px][4 * q + 3] = fmaf(a, wv.w, acc[px][4 * q + 3]);
          }
        }
      }
#pragma unroll
      for (int px = 0; px < PX; ++px) {
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[px][q] = acc[px][q] > 0.f ? acc[px][q] : acc[px][q] * slope;
        Block16<__nv_bfloat16>::store((__nv_bfloat16*)e0.ptr + n * e0.sN + cb * e0.sCb + y * e0.sY + (x0 + px) * e0.sX, acc[px]);
      }
    }
  }
}

// Returns kSgNotEligible when the shape is not covered (caller uses im2col + GEMM).
int launch_input_stage(const float* x, int C, const float* w, const float* bias, int Cout, const View& col,
                       const View& e0, float slope, cudaStream_t st) {
  if ((C != 1 && C != 3) || Cout > 64 || e0.Cb * 16 < Cout || col.Cb != (9 * C + 15) / 16) return kSgNotEligible;
  { const char* e = getenv("N2N_NO_INPUT_STAGE"); if (e && atoi(e)) return kSgNotEligible; }
  const long long pixels = (long long)e0.N * e0.H * e0.W;
  static int px_knob = -1;
  if (px_knob < 0) { const char* e = getenv("N2N_INPUT_STAGE_PX"); px_knob = e ? atoi(e) : 0; }
  // measured (64 x 1 x 256 x 256 / 32 x 3 x 256 x 256): C = 1 is bound by its HBM writes at any PX (PX = 4 is slower: registers);
  // C = 3 (27 taps) was bound by the weight loads: PX = 2 takes 40 us off the launch
  const int want = px_knob > 0 ? px_knob : (C == 1 ? 1 : 2);
  const int px = (want >= 4 && C == 1 && e0.W % 4 == 0) ? 4 : ((want >= 2 && e0.W % 2 == 0) ? 2 : 1);
  const long long groups = pixels / px;
  const int grid = grid_for(groups, 128, 16);
#define N2N_IS_LAUNCH(CC, PP) (void)launch_pdl_v(input_stage_kernel<CC, PP>, dim3(grid), dim3(128), 0, st, x, w, bias, Cout, col, e0, slope, groups)
  if (C == 1) { if (px == 4) N2N_IS_LAUNCH(1, 4); else if (px == 2) N2N_IS_LAUNCH(1, 2); else N2N_IS_LAUNCH(1, 1); }
  else { if (px == 2) N2N_IS_LAUNCH(3, 2); else N2N_IS_LAUNCH(3, 1); }
#undef N2N_IS_LAUNCH
  N2N_LAUNCH_CHECK();
  return 0;
}

int launch_nchw_to_c16(const float* src, int C, const View& dst, int dtype, cudaStream_t st) {
  N2N_CHECK_ARG(C >= 1 && C <= 16 * dst.Cb, "nchw_to_c16: C=%d does not fit %d blocks", C, dst.Cb);
  long long items = (long long)dst.N * dst.Cb * dst.H * dst.W;
  int grid = grid_for(items, 256);
  if (dtype == N2N_BF16) (void)launch_pdl_v(nchw_to_c16_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, src, C, dst, items);
  else (void)launch_pdl_v(nchw_to_c16_kernel<float>, dim3(grid), dim3(256), 0, st, src, C, dst, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

int launch_c16_to_nchw(const View& src, int dtype, float* dst, int C, cudaStream_t st) {
  N2N_CHECK_ARG(C >= 1 && C <= 16 * src.Cb, "c16_to_nchw: C=%d does not fit %d blocks", C, src.Cb);
  long long items = (long long)src.N * src.Cb * src.H * src.W;
  int grid = grid_for(items, 256);
  if (dtype == N2N_BF16) (void)launch_pdl_v(c16_to_nchw_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, src, dst, C, items);
  else (void)launch_pdl_v(c16_to_nchw_kernel<float>, dim3(grid), dim3(256), 0, st, src, dst, C, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

struct BiasPadBatch { BiasPadJob j[32]; int n; };
__global__ void bias_pad_kernel(const __grid_constant__ BiasPadBatch b) {
  pdl_enter_no_release();   // its output is prefetched by the next GEMM's prologue
  const BiasPadJob& J = b.j[blockIdx.x];
  for (int i = threadIdx.x; i < J.npad; i += blockDim.x) J.dst[i] = (i < J.n && J.src) ? J.src[i] : 0.f;
}
int launch_bias_pad(const BiasPadJob* jobs, int njobs, cudaStream_t st) {
  for (int base = 0; base < njobs; base += 32) {
    BiasPadBatch b;
    b.n = njobs - base < 32 ? njobs - base : 32;
    for (int i = 0; i < b.n; ++i) b.j[i] = jobs[base + i];
    (void)launch_pdl_v(bias_pad_kernel, dim3(b.n), dim3(128), 0, st, b);
    N2N_LAUNCH_CHECK();
  }
  return 0;
}
int launch_nchw_to_c16_multi(const float* src, int C, const View& dst, int dtype, cudaStream_t st) {
  return launch_nchw_to_c16(src, C, dst, dtype, st);
}
int launch_c16_to_nchw_multi(const View& src, int dtype, float* dst, int C, cudaStream_t st) {
  return launch_c16_to_nchw(src, dtype, dst, C, st);
}

// ------------------------------------------------------------------------------------------
// MaxPool2d(2) forward (arch_unet.py:120-136) and fused max-pool + LeakyReLU backward.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_kernel(View src, View dst, long long items) {
  pdl_enter();
  const int Wo = dst.W, Ho = dst.H, Cb = dst.Cb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho); r /= Ho;
    const int cb = (int)(r % Cb);
    const int n = (int)(r / Cb);
    const T* s = (const T*)src.ptr + n * src.sN + cb * src.sCb + (2 * y) * src.sY + (2 * x) * src.sX;
    float a[16], b[16], c[16], d[16], o[16];
    Block16<T>::load(s, a);
    Block16<T>::load(s + src.sX, b);
    Block16<T>::load(s + src.sY, c);
    Block16<T>::load(s + src.sY + src.sX, d);
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    Block16<T>::store((T*)dst.ptr + n * dst.sN + cb * dst.sCb + y * dst.sY + x * dst.sX, o);
  }
}

int launch_maxpool(const View& src, const View& dst, int dtype, cudaStream_t st) {
  N2N_CHECK_ARG(dst.H == src.H / 2 && dst.W == src.W / 2 && dst.Cb == src.Cb && dst.N == src.N,
                "maxpool: shape mismatch");
  long long items = (long long)dst.N * dst.Cb * dst.H * dst.W;
  int grid = grid_for(items, 256);
  if (dtype == N2N_BF16) (void)launch_pdl_v(maxpool_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, src, dst, items);
  else (void)launch_pdl_v(maxpool_kernel<float>, dim3(grid), dim3(256), 0, st, src, dst, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

// gact[2y+a, 2x+b] = (first max of the window at (a,b) ? gpool[y,x] : 0) * (act > 0 ? 1 : slope)
// Tie rule: ATen keeps the first maximum in row-major window order (update only on >).
template <typename T>
__global__ void unpool_lrelu_kernel(View act, View gpool, View gact, float slope, long long items) {
  pdl_enter();
  const int Wo = gpool.W, Ho = gpool.H, Cb = gpool.Cb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho); r /= Ho;
    const int cb = (int)(r % Cb);
    const int n = (int)(r / Cb);
    const T* s = (const T*)act.ptr + n * act.sN + cb * act.sCb + (2 * y) * act.sY + (2 * x) * act.sX;
    float v[4][16], g[16], o[4][16];
    Block16<T>::load(s, v[0]);
    Block16<T>::load(s + act.sX, v[1]);
    Block16<T>::load(s + act.sY, v[2]);
    Block16<T>::load(s + act.sY + act.sX, v[3]);
    Block16<T>::load((const T*)gpool.ptr + n * gpool.sN + cb * gpool.sCb + y * gpool.sY + x * gpool.sX, g);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      int best = 0; float m = v[0][k];
      if (v[1][k] > m) { m = v[1][k]; best = 1; }
      if (v[2][k] > m) { m = v[2][k]; best = 2; }
      if (v[3][k] > m) { m = v[3][k]; best = 3; }
      const float gg = g[k] * (m > 0.f ? 1.f : slope);
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][k] = (q == best) ? gg : 0.f;
    }
    T* d = (T*)gact.ptr + n * gact.sN + cb * gact.sCb + (2 * y) * gact.sY + (2 * x) * gact.sX;
    Block16<T>::store(d, o[0]);
    Block16<T>::store(d + gact.sX, o[1]);
    Block16<T>::store(d + gact.sY, o[2]);
    Block16<T>::store(d + gact.sY + gact.sX, o[3]);
  }
}

int launch_unpool_lrelu(const View& act, const View& gpool, const View& gact, float slope, int dtype,
                        cudaStream_t st) {
  N2N_CHECK_ARG(gpool.H == act.H / 2 && gpool.W == act.W / 2 && gpool.Cb == act.Cb && gact.H == act.H &&
                    gact.W == act.W && gact.Cb == act.Cb,
                "unpool: shape mismatch");
  long long items = (long long)gpool.N * gpool.Cb * gpool.H * gpool.W;
  int grid = grid_for(items, 128);
  if (dtype == N2N_BF16) (void)launch_pdl_v(unpool_lrelu_kernel<__nv_bfloat16>, dim3(grid), dim3(128), 0, st, act, gpool, gact, slope, items);
  else (void)launch_pdl_v(unpool_lrelu_kernel<float>, dim3(grid), dim3(128), 0, st, act, gpool, gact, slope, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

// Device-side patch cropper (the data path of train.py:208-228 / finetune.py:94-150 moved off the host): the training
// images stay resident in HBM as the reference holds them (float32 H x W x C, values 0..255); one launch cuts a batch
// of patches out of them at (image, top, left) and writes the [B, C, ps, ps] network input scaled by `scale`
// (1/255, train.py:358, finetune.py:146-147).  Clean and noisy tables share the crop coordinates, so the pair stays
// registered.  sel = int32 [B][3] (image index, top, left); imgs = device pointer table; dims = int32 [nimg][2] (H, W).
__global__ void crop_patches_kernel(const float* const* __restrict__ imgs, const int* __restrict__ dims,
                                    const int* __restrict__ sel, int C, int ps, float scale, float* __restrict__ out,
                                    long long items) {
  pdl_enter();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < items; t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int x = (int)(r % ps); r /= ps;
    const int y = (int)(r % ps); r /= ps;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const int im = sel[3 * b], top = sel[3 * b + 1], left = sel[3 * b + 2];
    const int W = dims[2 * im + 1];
    out[t] = imgs[im][((long long)(top + y) * W + (left + x)) * C + c] * scale;
  }
}
int launch_crop_patches(const float* const* imgs, const int* dims, const int* sel, int batch, int C, int ps, float scale,
                        float* out, cudaStream_t st) {
  const long long items = (long long)batch * C * ps * ps;
  if (items <= 0) return 0;
  (void)launch_pdl_v(crop_patches_kernel, dim3(grid_for(items, 256)), dim3(256), 0, st, imgs, dims, sel, C, ps, scale, out, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

// y[i] += x[i] (fp32): the global residual of arch_unet.RESNET (arch_unet.py:409) and its input gradient
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
int launch_add_inplace(float* y, const float* x, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  (void)launch_pdl_v(add_inplace_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, y, x, n);
  N2N_LAUNCH_CHECK();
  return 0;
}

int launch_fill_zero(void* p, size_t bytes, cudaStream_t st) {
  N2N_CUDA(cudaMemsetAsync(p, 0, bytes, st));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Evaluation post-processing
// ------------------------------------------------------------------------------------------
__global__ void quantize_u8_kernel(const float* __restrict__ p, uint8_t* __restrict__ o, long long n, float bias) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = fminf(fmaxf(p[i], 0.f), 1.f);
    // evaluation.py:83 / evaluation_704.py:120: fp32 multiply then add, clip, truncate
    v = __fadd_rn(__fmul_rn(v, 255.0f), bias);
    v = fminf(fmaxf(v, 0.f), 255.f);
    o[i] = (uint8_t)v;
  }
}

__global__ void tile_accumulate_kernel(const float* __restrict__ tile, int ps, const float* __restrict__ wm,
                                       float* __restrict__ acc, float* __restrict__ cnt, int W, int r0, int c0,
                                       int th, int tw) {
  const long long n = (long long)th * tw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / tw), x = (int)(i - (long long)y * tw);
    const float p = fminf(fmaxf(tile[(long long)y * ps + x], 0.f), 1.f);
    const float w = wm[(long long)y * ps + x];
    const long long o = (long long)(r0 + y) * W + (c0 + x);
    acc[o] = __fadd_rn(acc[o], __fmul_rn(p, w));
    cnt[o] = __fadd_rn(cnt[o], w);
  }
}

__global__ void tile_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ cnt,
                                     uint8_t* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float c = cnt[i];
    if (c == 0.f) c = 1.f;
    float v = __fmul_rn(__fdiv_rn(acc[i], c), 255.0f);
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = (uint8_t)v;
  }
}

// evaluation_704.py:82-101 on the device: tile t = (ty, tx) of image b starts at (ty * stride, tx * stride), is cut to the
// image, scaled by 1/255 and extended to ps x ps exactly as np.pad(patch, ((0, ps-th), (0, ps-tw)), mode='reflect') does
// (numpy's reflect continues periodically, period 2 (n - 1), when the pad exceeds the patch — the 128-pixel edge tiles of
// a 704 x 704 image are padded by 224).  tiles: fp32 [B * T][1][ps][ps], T = tiles_y * tiles_x, row-major tile order.
__device__ __forceinline__ int reflect_index(int i, int n) {
  if (i < n) return i;
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  const int m = i % period;
  return m < n ? m : period - m;
}
__global__ void tile_gather_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int ps, int stride, int tiles_y, int tiles_x,
                                      float* __restrict__ tiles, long long items) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int x = (int)(r % ps); r /= ps;
    const int y = (int)(r % ps); r /= ps;
    const int tx = (int)(r % tiles_x); r /= tiles_x;
    const int ty = (int)(r % tiles_y);
    const long long b = r / tiles_y;
    const int r0 = ty * stride, c0 = tx * stride;
    const int th = min(ps, H - r0), tw = min(ps, W - c0);
    const int sy = r0 + reflect_index(y, th), sx = c0 + reflect_index(x, tw);
    tiles[i] = __fdiv_rn((float)img[(b * H + sy) * W + sx], 255.0f);         // evaluation_704.py:89: astype(float32) / 255.0
  }
}
// evaluation_704.py:103-120: out = clip(acc / cnt * 255) with acc = sum_t clamp(pred_t, 0, 1) * w, cnt = sum_t w over the
// tiles covering the pixel, accumulated in the reference's tile order (fp32, same rounding as its sequential loop).
__global__ void tile_blend_u8_kernel(const float* __restrict__ pred, const float* __restrict__ wm, int H, int W, int ps, int stride,
                                     int tiles_y, int tiles_x, uint8_t* __restrict__ out, long long items) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    float acc = 0.f, cnt = 0.f;
    for (int ty = 0; ty < tiles_y; ++ty) {
      const int r0 = ty * stride;
      if (y < r0 || y >= min(r0 + ps, H)) continue;
      for (int tx = 0; tx < tiles_x; ++tx) {
        const int c0 = tx * stride;
        if (x < c0 || x >= min(c0 + ps, W)) continue;
        const long long t = (b * tiles_y + ty) * tiles_x + tx;
        const long long o = (long long)(y - r0) * ps + (x - c0);
        const float p = fminf(fmaxf(pred[t * ps * ps + o], 0.f), 1.f);
        const float w = wm[o];
        acc = __fadd_rn(acc, __fmul_rn(p, w));
        cnt = __fadd_rn(cnt, w);
      }
    }
    if (cnt == 0.f) cnt = 1.f;
    float v = __fmul_rn(__fdiv_rn(acc, cnt), 255.0f);
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = (uint8_t)v;
  }
}

}  // namespace n2n

using namespace n2n;

extern "C" int n2n_tile_gather_u8(const uint8_t* images, int batch, int h, int w, int ps, int stride, float* tiles, void* stream) {
  N2N_CHECK_ARG(images && tiles && batch >= 1 && h >= 1 && w >= 1 && ps >= 1 && stride >= 1, "tile_gather_u8: bad arguments");
  const int ty = (h + stride - 1) / stride, tx = (w + stride - 1) / stride;
  const long long items = (long long)batch * ty * tx * ps * ps;
  tile_gather_u8_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(images, h, w, ps, stride, ty, tx, tiles, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_tile_blend_u8(const float* pred_tiles, const float* weight_mask, int batch, int h, int w, int ps, int stride,
                                 uint8_t* out, void* stream) {
  N2N_CHECK_ARG(pred_tiles && weight_mask && out && batch >= 1 && h >= 1 && w >= 1 && ps >= 1 && stride >= 1, "tile_blend_u8: bad arguments");
  const int ty = (h + stride - 1) / stride, tx = (w + stride - 1) / stride;
  const long long items = (long long)batch * h * w;
  tile_blend_u8_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(pred_tiles, weight_mask, h, w, ps, stride, ty, tx, out, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_quantize_u8(const float* pred, uint8_t* out, int64_t count, float bias, void* stream) {
  N2N_CHECK_ARG(pred && out && count >= 0, "quantize_u8: bad arguments");
  if (count == 0) return 0;
  quantize_u8_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(pred, out, count, bias);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_tile_accumulate(const float* pred_tile, int ps, const float* weight_mask, float* acc,
                                   float* cnt, int H, int W, int r0, int c0, int th, int tw, void* stream) {
  N2N_CHECK_ARG(pred_tile && weight_mask && acc && cnt, "tile_accumulate: null pointer");
  N2N_CHECK_ARG(th >= 0 && tw >= 0 && th <= ps && tw <= ps && r0 >= 0 && c0 >= 0 && r0 + th <= H && c0 + tw <= W,
                "tile_accumulate: tile out of range");
  if (th == 0 || tw == 0) return 0;
  tile_accumulate_kernel<<<grid_for((long long)th * tw, 256), 256, 0, (cudaStream_t)stream>>>(
      pred_tile, ps, weight_mask, acc, cnt, W, r0, c0, th, tw);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_tile_finalize_u8(const float* acc, const float* cnt, uint8_t* out, int64_t count, void* stream) {
  N2N_CHECK_ARG(acc && cnt && out && count >= 0, "tile_finalize: bad arguments");
  if (count == 0) return 0;
  tile_finalize_kernel<<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(acc, cnt, out, count);
  N2N_LAUNCH_CHECK();
  return 0;
}
