// layers.cuh — host-side builders that turn "a conv / deconv layer on C16 views" into the
// generic engine descriptors (TapGemm / TapWgrad / PackJob / UnpackJob).  Shared by the
// single-layer C-ABI entry points (api.cu) and the network plans (unet_plan.cu).
#pragma once
#include "common.cuh"

namespace n2n {

// Channel segments of a (possibly concatenated) operand: real channel ranges and where each
// starts inside the blocked layout (every segment starts on a 16-channel block boundary).
struct ChanSegs {
  int n = 1;
  int cnt[2] = {0, 0};
  int real() const { return cnt[0] + (n > 1 ? cnt[1] : 0); }
  int blocks() const { return cblocks(cnt[0]) + (n > 1 ? cblocks(cnt[1]) : 0); }
  Segs to_segs() const {
    Segs s; s.n = n;
    s.src0[0] = 0; s.cnt[0] = cnt[0]; s.dst0[0] = 0;
    if (n > 1) { s.src0[1] = cnt[0]; s.cnt[1] = cnt[1]; s.dst0[1] = cblocks(cnt[0]) * 16; }
    return s;
  }
};
inline ChanSegs chan1(int c) { ChanSegs s; s.n = 1; s.cnt[0] = c; return s; }
inline ChanSegs chan2(int a, int b) { ChanSegs s; s.n = 2; s.cnt[0] = a; s.cnt[1] = b; return s; }

enum LayerKind { L_CONV3 = 0, L_CONV1 = 1, L_DECONV = 2 };

struct LayerGeom {
  int kind = L_CONV3;
  ChanSegs cin;          // input channels (concat aware)
  int cout = 0;          // real output channels
  int ntaps() const { return kind == L_CONV3 ? 9 : (kind == L_CONV1 ? 1 : 4); }
  int cin_blocks() const { return cin.blocks(); }
  int cout_blocks() const { return cblocks(cout); }
  // source strides of the PyTorch weight for (tap, out-channel, in-channel)
  void strides(long long& s_t, long long& s_co, long long& s_ci) const {
    const int ci = cin.real();
    if (kind == L_CONV3) { s_t = 1; s_co = (long long)ci * 9; s_ci = 9; }
    else if (kind == L_CONV1) { s_t = 0; s_co = ci; s_ci = 1; }
    else { s_t = 1; s_co = 4; s_ci = (long long)cout * 4; }       // ConvTranspose2d [ci][co][2][2]
  }
  size_t fwd_pack_bytes(int dtype) const { return packed_weight_bytes(dtype, ntaps(), cout_blocks() * 16, cin_blocks()); }
  size_t dgrad_pack_bytes(int dtype, int out_blocks) const { return packed_weight_bytes(dtype, ntaps(), out_blocks * 16, cout_blocks()); }
  size_t partial_bytes(int splits) const {
    return (size_t)splits * ntaps() * cin_blocks() * 16 * cout_blocks() * 16 * sizeof(float);
  }
  size_t bias_partial_bytes(int splits) const {
    return (size_t)splits * (kind == L_DECONV ? 4 : 1) * cout_blocks() * 16 * sizeof(float);
  }
};

inline PackJob make_fwd_pack(const LayerGeom& L, const float* w, void* dst) {
  PackJob j;
  j.src = w; j.dst = dst; j.ntaps = L.ntaps();
  j.nout_pad = L.cout_blocks() * 16; j.cin_blocks = L.cin_blocks();
  long long st, sco, sci; L.strides(st, sco, sci);
  j.s_t = st; j.s_n = sco; j.s_c = sci;
  j.nseg = chan1(L.cout).to_segs(); j.cseg = L.cin.to_segs();
  return j;
}
inline PackJob make_dgrad_pack(const LayerGeom& L, const float* w, void* dst, int out_blocks) {
  PackJob j;
  j.src = w; j.dst = dst; j.ntaps = L.ntaps();
  j.nout_pad = out_blocks * 16; j.cin_blocks = L.cout_blocks();
  long long st, sco, sci; L.strides(st, sco, sci);
  j.s_t = st; j.s_n = sci; j.s_c = sco;
  j.nseg = L.cin.to_segs(); j.cseg = chan1(L.cout).to_segs();
  return j;
}
inline UnpackJob make_unpack(const LayerGeom& L, const float* partial, const float* bias_partial, int splits,
                             float* dw, float* db) {
  UnpackJob j;
  j.partial = partial; j.bias_partial = bias_partial; j.dst_w = dw; j.dst_b = db;
  j.splits = splits; j.ntaps = L.ntaps(); j.npad = L.cout_blocks() * 16; j.cpad = L.cin_blocks() * 16;
  j.bias_rows = splits * (L.kind == L_DECONV ? 4 : 1);
  long long st, sco, sci; L.strides(st, sco, sci);
  j.s_t = st; j.s_n = sco; j.s_c = sci;
  j.nseg = chan1(L.cout).to_segs(); j.cseg = L.cin.to_segs();
  return j;
}

// y = conv(x) for conv3x3 / conv1x1 (x, y same spatial size)
inline TapGemm make_conv_fwd(const LayerGeom& L, int dtype, const View& x, const View& y, const void* wp,
                             const float* bias_pad) {
  TapGemm g;
  g.dtype = dtype; g.x[0] = x; g.ntaps = L.ntaps();
  for (int t = 0; t < g.ntaps; ++t) {
    g.tap_dy[t] = L.kind == L_CONV3 ? t / 3 - 1 : 0;
    g.tap_dx[t] = L.kind == L_CONV3 ? t % 3 - 1 : 0;
    g.tap_view[t] = 0; g.tap_slab[t] = t;
  }
  g.cin_blocks = L.cin_blocks(); g.nout = L.cout_blocks() * 16; g.w = wp; g.bias = bias_pad; g.y = y;
  return g;
}
// dx = conv_input_grad(dy): a correlation of dy with the mirrored taps and transposed weights
inline TapGemm make_conv_dgrad(const LayerGeom& L, int dtype, const View& dy, const View& dx, const void* wp_dgrad,
                               int out_blocks /* how many leading cin blocks to produce */) {
  TapGemm g;
  g.dtype = dtype; g.x[0] = dy; g.ntaps = L.ntaps();
  for (int t = 0; t < g.ntaps; ++t) {
    g.tap_dy[t] = L.kind == L_CONV3 ? -(t / 3 - 1) : 0;
    g.tap_dx[t] = L.kind == L_CONV3 ? -(t % 3 - 1) : 0;
    g.tap_view[t] = 0; g.tap_slab[t] = t;
  }
  g.cin_blocks = L.cout_blocks(); g.nout = out_blocks * 16; g.w = wp_dgrad; g.bias = nullptr; g.y = dx;
  return g;
}
// deconv forward for output parity (a,b): y_ab = x * W_ab + bias
inline TapGemm make_deconv_fwd(const LayerGeom& L, int dtype, const View& x, const View& y_full, int a, int b,
                               const void* wp, const float* bias_pad) {
  TapGemm g;
  g.dtype = dtype; g.x[0] = x; g.ntaps = 1;
  g.tap_dy[0] = 0; g.tap_dx[0] = 0; g.tap_view[0] = 0; g.tap_slab[0] = 2 * a + b;
  g.cin_blocks = L.cin_blocks(); g.nout = L.cout_blocks() * 16; g.w = wp; g.bias = bias_pad;
  g.y = parity_view(y_full, dtype, a, b);
  return g;
}
// Pair form (slab engine): output row parity a, both column parities in one GEMM with N = 2*Cout_pad.
// Weights: slab a = rows [b][co] of W[ci][co][a][b]  (make_deconv_pair_pack).
inline PackJob make_deconv_pair_pack(const LayerGeom& L, const float* w, void* dst) {
  PackJob j;
  j.src = w; j.dst = dst; j.ntaps = 2;
  const int cp = L.cout_blocks() * 16;
  j.nout_pad = 2 * cp; j.cin_blocks = L.cin_blocks();
  j.s_t = 2; j.s_n = 4; j.s_c = (long long)L.cout * 4;          // ConvTranspose2d [ci][co][a][b]
  j.nseg.n = 2;
  j.nseg.src0[0] = 0; j.nseg.cnt[0] = L.cout; j.nseg.dst0[0] = 0; j.nseg.off[0] = 0;
  j.nseg.src0[1] = 0; j.nseg.cnt[1] = L.cout; j.nseg.dst0[1] = cp; j.nseg.off[1] = 1;
  j.cseg = L.cin.to_segs();
  return j;
}
inline TapGemm make_deconv_fwd_pair(const LayerGeom& L, int dtype, const View& x, const View& y_full, int a,
                                    const void* wp_pair, const float* bias_pad) {
  TapGemm g;
  g.dtype = dtype; g.x[0] = x; g.ntaps = 1;
  g.tap_dy[0] = 0; g.tap_dx[0] = 0; g.tap_view[0] = 0; g.tap_slab[0] = a;
  g.cin_blocks = L.cin_blocks(); g.nout = 2 * L.cout_blocks() * 16; g.w = wp_pair; g.bias = bias_pad;
  g.y = parity_view(y_full, dtype, a, 0);
  g.n_split = L.cout_blocks(); g.split_stride = y_full.sX;       // column parity 1 = the next output pixel
  return g;
}
// ---- fused ConvTranspose2x2 -> conv3x3 (no-grad passes, slab engine, see pack.cu: upfuse_pack_kernel) ----------
// Packed weights of ONE output-row parity: 8 composite slabs over the deconv's input channels, then the conv's
// skip-channel slabs (nine 3x3 taps over `skip_blocks`, or the single im2col tap of the raw network input).
struct UpConvGeom {
  int ci_blocks = 0, skip_blocks = 0, co_blocks = 0;
  bool skip_im2col = false;       // level 1 on the bf16 engine: the skip is the 9-tap im2col block of the input
  int gu() const { return (ci_blocks + 2) / 3; }
  int gs() const { return (skip_blocks + 2) / 3; }
  size_t slab_bytes() const { return (size_t)3 * co_blocks * 16 * 32; }
  size_t skip_base() const { return (size_t)8 * gu() * slab_bytes(); }
  int skip_slabs() const { return skip_im2col ? 1 : 9; }
  size_t region_bytes() const { return skip_base() + (size_t)skip_slabs() * gs() * slab_bytes(); }
};
bool slab_upconv_ok(int dtype, int n, int h, int w, int ci_blocks, int skip_blocks, int co_blocks, size_t w_bytes);

// x: the ConvTranspose's input (source resolution); skip_hi: the skip channels of the concat buffer (2x resolution);
// y_full: the conv's output (2x resolution); py: output-row parity of this launch; wreg: this parity's weight region.
inline TapGemm make_upconv_fwd(const UpConvGeom& U, int dtype, const View& x, const View& skip_hi, const View& y_full,
                               int py, const void* wreg, const float* bias_full, const float* corr) {
  TapGemm g;
  g.dtype = dtype;
  const int cop = U.co_blocks * 16;
  g.nout = 2 * cop; g.mma_n = cop; g.cin_blocks = U.ci_blocks; g.w = wreg; g.bias = bias_full;
  g.border_corr = corr; g.up_py = py;
  g.y = parity_view(y_full, dtype, py, 0);
  g.n_split = U.co_blocks; g.split_stride = y_full.sX;
  g.x[0] = x;
  int t = 0;
  for (int px = 0; px < 2; ++px)
    for (int syi = 0; syi < 2; ++syi)
      for (int sxi = 0; sxi < 2; ++sxi) {
        g.tap_view[t] = 0; g.tap_dy[t] = syi + py - 1; g.tap_dx[t] = sxi + px - 1;
        g.tap_col[t] = px * cop;
        g.tap_woff[t] = (long long)(((px * 2 + syi) * 2 + sxi) * U.gu()) * (long long)U.slab_bytes();
        ++t;
      }
  if (U.skip_im2col) {
    for (int px = 0; px < 2; ++px) {
      g.x[1 + px] = parity_view(skip_hi, dtype, py, px); g.view_blocks[1 + px] = U.skip_blocks;
      g.tap_view[t] = 1 + px; g.tap_dy[t] = 0; g.tap_dx[t] = 0; g.tap_col[t] = px * cop;
      g.tap_woff[t] = (long long)U.skip_base();
      ++t;
    }
  } else {
    for (int q = 0; q < 4; ++q) { g.x[1 + q] = parity_view(skip_hi, dtype, q >> 1, q & 1); g.view_blocks[1 + q] = U.skip_blocks; }
    for (int px = 0; px < 2; ++px)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int yo = py + ky - 1, xo = px + kx - 1;          // offset on the 2x grid relative to (2i, 2j)
          g.tap_view[t] = 1 + (yo & 1) * 2 + (xo & 1);
          g.tap_dy[t] = yo >= 0 ? yo >> 1 : -1; g.tap_dx[t] = xo >= 0 ? xo >> 1 : -1;
          g.tap_col[t] = px * cop;
          g.tap_woff[t] = (long long)U.skip_base() + (long long)((ky * 3 + kx) * U.gs()) * (long long)U.slab_bytes();
          ++t;
        }
  }
  g.ntaps = t;
  return g;
}

// Training plans: input gradient of the fused up-conv w.r.t. the ConvTranspose's input — one launch over the four
// parity views of dL/dy (the conv's pre-activation output gradient, 2x resolution) with the transposed composites:
//   g_x[i, j] = sum_{py,px,sy,sx} Wc[py,px,sy,sx]^T  dL/dy[2(i - sy) + py, 2(j - sx) + px]
inline size_t upconv_wt_bytes(const UpConvGeom& U) {
  return (size_t)16 * ((U.co_blocks + 2) / 3) * 3 * U.ci_blocks * 16 * 32;
}
inline TapGemm make_upconv_dgrad(const UpConvGeom& U, int dtype, const View& gy_full, const View& gx, const void* wt) {
  TapGemm g;
  g.dtype = dtype; g.cin_blocks = U.co_blocks; g.nout = U.ci_blocks * 16; g.w = wt; g.bias = nullptr; g.y = gx;
  int t = 0;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      g.x[py * 2 + px] = parity_view(gy_full, dtype, py, px);
      for (int syi = 0; syi < 2; ++syi)
        for (int sxi = 0; sxi < 2; ++sxi) {
          g.tap_view[t] = py * 2 + px;
          g.tap_dy[t] = -(syi + py - 1); g.tap_dx[t] = -(sxi + px - 1);
          g.tap_slab[t] = ((py * 2 + px) * 2 + syi) * 2 + sxi;
          ++t;
        }
    }
  g.ntaps = t;
  return g;
}
// ... and the composite weight gradient of one output parity: dWc[py,px,sy,sx][ci][co] = sum_p dL/dy[2i+py, 2j+px][co] x[i+sy, j+sx][ci]
inline TapWgrad make_upconv_wgrad(const UpConvGeom& U, int dtype, const View& x, const View& gy_full, int py, int px,
                                  float* partial, int splits) {
  TapWgrad g;
  g.dtype = dtype; g.dy[0] = parity_view(gy_full, dtype, py, px); g.x[0] = x; g.npairs = 4;
  for (int syi = 0; syi < 2; ++syi)
    for (int sxi = 0; sxi < 2; ++sxi) {
      const int t = syi * 2 + sxi;
      g.pair_dyv[t] = 0; g.pair_xv[t] = 0; g.pair_dy[t] = syi + py - 1; g.pair_dx[t] = sxi + px - 1;
    }
  g.n_blocks = U.co_blocks; g.c_blocks = U.ci_blocks;
  g.partial = partial; g.bias_partial = nullptr; g.ndyviews = 1; g.splits = splits;
  return g;
}

bool slab_deconv_pair_ok(int dtype, int h, int w, int cin_blocks, int cout_blocks);
bool slab_weights_fit(int ntaps, int cin_blocks, int nout, bool halo);
bool slab_geometry_ok(int dtype, int h, int w);

inline TapGemm make_deconv_dgrad(const LayerGeom& L, int dtype, const View& dy_full, const View& dx,
                                 const void* wp_dgrad /* packed with out_blocks == cin_blocks */) {
  TapGemm g;
  g.dtype = dtype; g.ntaps = 4;
  for (int t = 0; t < 4; ++t) {
    g.x[t] = parity_view(dy_full, dtype, t / 2, t % 2);
    g.tap_dy[t] = 0; g.tap_dx[t] = 0; g.tap_view[t] = t; g.tap_slab[t] = t;
  }
  g.cin_blocks = L.cout_blocks(); g.nout = L.cin_blocks() * 16; g.w = wp_dgrad; g.bias = nullptr; g.y = dx;
  return g;
}
inline TapWgrad make_conv_wgrad(const LayerGeom& L, int dtype, const View& x, const View& dy, float* partial,
                                float* bias_partial, int splits) {
  TapWgrad g;
  g.dtype = dtype; g.dy[0] = dy; g.x[0] = x; g.npairs = L.ntaps();
  for (int t = 0; t < g.npairs; ++t) {
    g.pair_dyv[t] = 0; g.pair_xv[t] = 0;
    g.pair_dy[t] = L.kind == L_CONV3 ? t / 3 - 1 : 0;
    g.pair_dx[t] = L.kind == L_CONV3 ? t % 3 - 1 : 0;
  }
  g.n_blocks = L.cout_blocks(); g.c_blocks = L.cin_blocks();
  g.partial = partial; g.bias_partial = bias_partial; g.ndyviews = 1; g.splits = splits;
  return g;
}
inline TapWgrad make_deconv_wgrad(const LayerGeom& L, int dtype, const View& x, const View& dy_full, float* partial,
                                  float* bias_partial, int splits) {
  TapWgrad g;
  g.dtype = dtype; g.x[0] = x; g.npairs = 4;
  for (int t = 0; t < 4; ++t) {
    g.dy[t] = parity_view(dy_full, dtype, t / 2, t % 2);
    g.pair_dyv[t] = t; g.pair_xv[t] = 0; g.pair_dy[t] = 0; g.pair_dx[t] = 0;
  }
  g.n_blocks = L.cout_blocks(); g.c_blocks = L.cin_blocks();
  g.partial = partial; g.bias_partial = bias_partial; g.ndyviews = 4; g.splits = splits;
  return g;
}

// Pixel splits of the layer's weight-gradient GEMM for an input of n x h x w pixels: the slab
// engine wants (splits x tap groups) = one CTA per SM; other geometries keep the old default.
int wgrad_slab_splits(int npairs, int variant_blocks, int common_blocks, bool box_per_tap, long long tiles);
inline int layer_wgrad_splits(const LayerGeom& L, int dtype, int n, int h, int w) {
  if (dtype == N2N_BF16 && h >= 4 && w >= 4) {
    const bool dc = L.kind == L_DECONV;
    const int s = wgrad_slab_splits(L.ntaps(), dc ? L.cout_blocks() : L.cin_blocks(), dc ? L.cin_blocks() : L.cout_blocks(), dc,
                                    (long long)n * ((h + 15) / 16) * ((w + 7) / 8));
    if (s > 0) return s;
  }
  return wgrad_default_splits(dtype, (long long)n * h * w);
}

// bias padding (one launch for up to 32 vectors)
struct BiasPadJob { const float* src; float* dst; int n; int npad; };
int launch_bias_pad(const BiasPadJob* jobs, int njobs, cudaStream_t st);
// NCHW fp32 (all channels) <-> multi-block C16 view
int launch_nchw_to_c16_multi(const float* src, int C, const View& dst, int dtype, cudaStream_t st);
int launch_c16_to_nchw_multi(const View& src, int dtype, float* dst, int C, cudaStream_t st);

}  // namespace n2n
