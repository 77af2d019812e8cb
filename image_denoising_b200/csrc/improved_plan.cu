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
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scale_shift /* [N][Cb*16][2] */) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Cb * 16) return;
  const int n = i / (Cb * 16), c = i % (Cb * 16);
  float sc = 0.f, sh = 0.f;
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
  }
  scale_shift[2 * i] = sc; scale_shift[2 * i + 1] = sh;
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

template <typename F32K, typename BF16K, typename... Args>
static cudaError_t launch_by_dtype(int dtype, F32K kf, BF16K kb, dim3 grid, cudaStream_t st, Args... args) {
  return dtype == N2N_BF16 ? launch_pdl_v(kb, grid, dim3(256), 0, st, args...) : launch_pdl_v(kf, grid, dim3(256), 0, st, args...);
}

}  // namespace n2n

using namespace n2n;

// ------------------------------------------------------------------------------------------------------------------------
struct ImpBuf { size_t off = 0; int cb = 0, h = 0, w = 0; };

struct n2n_improved_plan {
  int in_nc, out_nc, nf, depth, noise, N, H, W, dtype;
  std::vector<ImpBuf> bufs;
  std::vector<size_t> conv_wp, conv_bias;     // per convolution chunk: offsets of its packed weights / padded bias
  size_t off_wp = 0, off_bias = 0, off_gn_partial = 0, off_gn_ss = 0, off_f32 = 0, total = 0;
  int launches = 0;
  int nparams = 0;
};

namespace {

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
  size_t off = 0, wp_total = 0, bias_total = 0, gn_partial_max = 0, gn_ss_max = 0;
  std::vector<PackJob> pack_jobs;
  std::vector<BiasPadJob> bias_jobs;
  bool run() const { return mode == IMP_EXEC; }

  int new_buf(int channels_blocks, int h, int w) {
    if (mode == IMP_SIZE) {
      ImpBuf b; b.off = off; b.cb = channels_blocks; b.h = h; b.w = w;
      off += align_up((size_t)p->N * channels_blocks * h * w * 16 * dtype_size(p->dtype), 1024);
      p->bufs.push_back(b);
      return (int)p->bufs.size() - 1;
    }
    return nbuf_run++;
  }
  int nbuf_run = 0;
  View view(int b, int cb0, int cb) const {
    const ImpBuf& B = p->bufs[b];
    return make_view(ws ? ws + B.off : nullptr, p->dtype, p->N, B.h, B.w, B.cb, cb0, cb);
  }
  const float* param(int i) const { return mode == IMP_SIZE ? nullptr : prm[i]; }

  // conv (k = 3 / 1) over a one- or two-segment input window, optional bias / LeakyReLU / residual addend
  int conv(const View& x, int c0, int c1, int cout, int k, bool has_bias, const View& y, float slope, const View* addend) {
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
      TapGemm g = make_conv_fwd(L, p->dtype, x, sub_blocks(y, p->dtype, n0 / 16, L.cout_blocks()), wp, bias);
      if (slope >= 0.f) { g.act = 1; g.slope = slope; }
      if (addend) { g.has_addend = true; g.addend = sub_blocks(*addend, p->dtype, n0 / 16, L.cout_blocks()); }
      N2N_TRY(launch_tapgemm(g, st));
    }
    return 0;
  }

  int groupnorm(const View& x, int C, const View& y, float slope, const View* res) {
    const float* gamma = param(pi);
    const float* beta = param(pi + 1);
    pi += 2;
    int groups = C < 32 ? C : 32;
    while (groups > 1 && C % groups) --groups;                  // arch_unet.py:12-14
    const int cg = C / groups;
    const long long hw = (long long)x.H * x.W;
    int splits = (int)((kSMs * 2 + (long long)x.N * x.Cb - 1) / ((long long)x.N * x.Cb));
    if (splits > 64) splits = 64;
    while (splits > 1 && hw / splits < 512) --splits;
    const size_t pb = (size_t)x.N * x.Cb * splits * 32 * sizeof(double), sb = (size_t)x.N * x.Cb * 16 * 2 * sizeof(float);
    if (pb > gn_partial_max) gn_partial_max = pb;
    if (sb > gn_ss_max) gn_ss_max = sb;
    if (!run()) return 0;
    double* partial = (double*)(ws + p->off_gn_partial);
    float* ss = (float*)(ws + p->off_gn_ss);
    (void)launch_by_dtype(p->dtype, c16_gn_stats_kernel<float>, c16_gn_stats_kernel<__nv_bfloat16>, dim3(splits, x.Cb, x.N), st, x, splits, partial);
    N2N_LAUNCH_CHECK();
    (void)launch_pdl_v(c16_gn_fold_kernel, dim3((x.N * x.Cb * 16 + 127) / 128), dim3(128), 0, st, (const double*)partial, splits, x.N, C,
                       x.Cb, cg, hw, 1e-5f, gamma, beta, ss);
    N2N_LAUNCH_CHECK();
    const long long items = (long long)x.N * x.Cb * hw;
    (void)launch_by_dtype(p->dtype, c16_gn_apply_kernel<float>, c16_gn_apply_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), st, x,
                          res ? *res : x, res ? 1 : 0, y, (const float*)ss, slope, items);
    N2N_LAUNCH_CHECK();
    return 0;
  }

  // RDB (arch_unet.py:435-449): `R` holds [x | o1 | o2 | o3 | o4]; x is already in blocks [0, xb); result -> `out`
  int rdb(int R, int C, const View& out) {
    const int xb = cblocks(C);
    for (int j = 0; j < 4; ++j)
      N2N_TRY(conv(view(R, 0, xb + 2 * j), C, 32 * j, 32, 3, true, view(R, xb + 2 * j, 2), 0.2f, nullptr));
    const View xv = view(R, 0, xb);
    return conv(view(R, 0, xb + 8), C, 128, C, 1, true, out, -1.f, &xv);                   // x + lff(cat)
  }
  // ResBlock (:420-432): x (view `xin`) -> `out`; t1 / t2 are scratch buffers of the same shape
  int resblock(const View& xin, int C, int t1, int t2, const View& out) {
    const int xb = cblocks(C);
    N2N_TRY(conv(xin, C, 0, C, 3, false, view(t1, 0, xb), -1.f, nullptr));
    N2N_TRY(groupnorm(view(t1, 0, xb), C, view(t2, 0, xb), 0.2f, nullptr));
    N2N_TRY(conv(view(t2, 0, xb), C, 0, C, 3, false, view(t1, 0, xb), -1.f, nullptr));
    return groupnorm(view(t1, 0, xb), C, out, -1.f, &xin);
  }

  int network(const float* x, float* y) {
    const int N = p->N, H = p->H, W = p->W, dt = p->dtype, in_nc = p->in_nc, nf0 = p->nf, D = p->depth;
    // input block: [x (in_nc) | sigma (1)]
    const int IN = new_buf(1, H, W);
    if (run()) N2N_TRY(launch_nchw_to_c16(x, in_nc, view(IN, 0, 1), dt, st));
    if (p->noise) {
      const int NE1 = new_buf(cblocks(nf0), H, W), NE2 = new_buf(1, H, W);
      N2N_TRY(conv(view(IN, 0, 1), in_nc, 0, nf0, 3, true, view(NE1, 0, cblocks(nf0)), 0.2f, nullptr));
      N2N_TRY(conv(view(NE1, 0, cblocks(nf0)), nf0, 0, 1, 3, true, view(NE2, 0, 1), -1.f, nullptr));
      if (run()) {
        float* sig = (float*)(ws + p->off_f32);
        N2N_TRY(launch_c16_to_nchw(view(NE2, 0, 1), dt, sig, 1, st));
        N2N_TRY(n2n_act_fwd(sig, sig, (int64_t)N * H * W, 2, 0.f, st));
        const long long items = (long long)N * H * W;
        const View iv = view(IN, 0, 1);
        (void)launch_by_dtype(dt, c16_set_lane_kernel<float>, c16_set_lane_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), st,
                              (const float*)sig, iv, in_nc, items);
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
      const int R = new_buf(xb + 8, h, w), A = new_buf(xb, h, w), T1 = new_buf(xb, h, w), T2 = new_buf(xb, h, w);
      N2N_TRY(conv(view(cur, 0, cblocks(cur_c)), cur_c, 0, nf, 3, true, view(R, 0, xb), 0.2f, nullptr));
      N2N_TRY(rdb(R, nf, view(A, 0, xb)));
      const View skip = view(F[i], cblocks(nf / 2), xb);
      N2N_TRY(resblock(view(A, 0, xb), nf, T1, T2, skip));
      // pool -> the x blocks of the next level's RDB input producer; the next conv reads it from a plain buffer
      const int P = new_buf(xb, h / 2, w / 2);
      if (run()) N2N_TRY(launch_maxpool(skip, view(P, 0, xb), dt, st));
      cur = P; cur_c = nf;
      nf *= 2;
    }
    nf /= 2;       // bottleneck width
    {
      const int h = H >> D, w = W >> D, xb = cblocks(nf);
      const int R = new_buf(xb + 8, h, w), A = new_buf(xb, h, w), T1 = new_buf(xb, h, w), T2 = new_buf(xb, h, w), B = new_buf(xb, h, w);
      if (run()) {   // RDB input must sit in R[0:xb): copy the pooled tensor (a D2D copy of the smallest level)
        const View src = view(cur, 0, xb), dst = view(R, 0, xb);
        const size_t es = dtype_size(dt), row = (size_t)xb * h * w * 16 * es;
        N2N_CUDA(cudaMemcpy2DAsync(dst.ptr, (size_t)dst.sN * es, src.ptr, (size_t)src.sN * es, row, N, cudaMemcpyDeviceToDevice, st));
      }
      N2N_TRY(rdb(R, nf, view(A, 0, xb)));
      N2N_TRY(resblock(view(A, 0, xb), nf, T1, T2, view(B, 0, xb)));
      cur = B; cur_c = nf;
    }
    // decoder
    const int FIN = new_buf(cblocks(nf0 / 2) + 1, H, W);
    for (int j = 0; j < D; ++j) {
      const int lvl = D - 1 - j, h = H >> lvl, w = W >> lvl, oc = cur_c / 2, ob = cblocks(oc);
      const int PS = new_buf(cblocks(4 * oc), h / 2, w / 2);
      N2N_TRY(conv(view(cur, 0, cblocks(cur_c)), cur_c, 0, 4 * oc, 3, true, view(PS, 0, cblocks(4 * oc)), -1.f, nullptr));
      if (run()) {
        const View sv = view(PS, 0, cblocks(4 * oc)), dv = view(F[lvl], 0, ob);
        const long long items = (long long)N * ob * (h / 2) * (w / 2);
        (void)launch_by_dtype(dt, c16_pixel_shuffle2_kernel<float>, c16_pixel_shuffle2_kernel<__nv_bfloat16>, dim3(grid_for(items, 256)), st,
                              sv, dv, oc, items);
        N2N_LAUNCH_CHECK();
      }
      const int R = new_buf(ob + 8, h, w), A = new_buf(ob, h, w), T1 = new_buf(ob, h, w), T2 = new_buf(ob, h, w);
      N2N_TRY(conv(view(F[lvl], 0, ob + cblocks(skip_c[lvl])), oc, skip_c[lvl], oc, 3, true, view(R, 0, ob), 0.2f, nullptr));
      N2N_TRY(rdb(R, oc, view(A, 0, ob)));
      int O = -1;
      View out;
      if (j == D - 1) out = view(FIN, 0, ob);
      else { O = new_buf(ob, h, w); out = view(O, 0, ob); }
      N2N_TRY(resblock(view(A, 0, ob), oc, T1, T2, out));
      cur = O; cur_c = oc;
    }
    // final: sigmoid(conv(cat[x, orig]))
    const int fb = cblocks(nf0 / 2);
    if (run()) N2N_TRY(launch_nchw_to_c16(x, in_nc, view(FIN, fb, 1), dt, st));
    const int OUT = new_buf(cblocks(p->out_nc), H, W);
    N2N_TRY(conv(view(FIN, 0, fb + 1), nf0 / 2, in_nc, p->out_nc, 3, true, view(OUT, 0, cblocks(p->out_nc)), -1.f, nullptr));
    if (run()) {
      N2N_TRY(launch_c16_to_nchw(view(OUT, 0, cblocks(p->out_nc)), dt, y, p->out_nc, st));
      N2N_TRY(n2n_act_fwd(y, y, (int64_t)N * p->out_nc * H * W, 2, 0.f, st));
    }
    return 0;
  }
};

}  // namespace

extern "C" int n2n_improved_plan_create(n2n_improved_plan** plan, int in_nc, int out_nc, int n_feature, int depth, int noise, int n,
                                        int h, int w, int dtype) {
  N2N_CHECK_ARG(plan && in_nc >= 1 && in_nc <= 15 && out_nc >= 1 && out_nc <= 16 && n_feature >= 2 && n_feature % 2 == 0 && depth >= 1 &&
                depth <= 6 && n >= 1 && h >= 1 && w >= 1, "improved_plan_create: bad arguments");
  N2N_CHECK_ARG(h % (1 << depth) == 0 && w % (1 << depth) == 0, "improved_plan_create: H and W must be multiples of 2^depth");
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, "improved_plan_create: bad dtype");
  N2N_CHECK_ARG(noise || in_nc == 1, "improved_plan_create: noise = 0 expects one input channel (arch_unet.py:496)");
  n2n_improved_plan* p = new n2n_improved_plan();
  p->in_nc = in_nc; p->out_nc = out_nc; p->nf = n_feature; p->depth = depth; p->noise = noise; p->N = n; p->H = h; p->W = w; p->dtype = dtype;
  Builder b{p, nullptr, nullptr, nullptr};
  const int r = b.network(nullptr, nullptr);
  if (r != 0) { delete p; return r; }
  size_t off = b.off;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  p->off_wp = take(b.wp_total); p->off_bias = take(b.bias_total);
  p->off_gn_partial = take(b.gn_partial_max); p->off_gn_ss = take(b.gn_ss_max);
  p->off_f32 = take((size_t)n * h * w * sizeof(float));
  p->total = off;
  p->nparams = b.pi;
  *plan = p;
  return 0;
}
extern "C" void n2n_improved_plan_destroy(n2n_improved_plan* plan) { delete plan; }
extern "C" size_t n2n_improved_workspace_bytes(const n2n_improved_plan* plan) { return plan ? plan->total : 0; }
extern "C" int n2n_improved_num_params(const n2n_improved_plan* plan) { return plan ? plan->nparams : 0; }
extern "C" int n2n_improved_launches(const n2n_improved_plan* plan) { return plan ? plan->launches : 0; }

extern "C" int n2n_improved_forward(n2n_improved_plan* p, const float* const* params, const float* x, float* y, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && x && y && ws, "improved_forward: null argument");
  const long long l0 = g_launch_count;
  cudaStream_t st = (cudaStream_t)stream;
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
  p->launches = (int)(g_launch_count - l0);
  return 0;
}
