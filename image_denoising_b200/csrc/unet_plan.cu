// unet_plan.cu — native executor for arch_unet.UNet (arch_unet.py:100-260, non-blindspot) and
// adapter.OutputAdapter (adapter.py:5-26): one C call runs the whole forward (or backward) as a
// fixed sequence of engine launches over C16 buffers carved out of a caller-owned workspace.
//
// Concat is never materialised: ConvTranspose epilogues write blocks [0, c_up) of the level's
// concat buffer, the encoder's pool kernels write the skip blocks behind them
// (torch.cat([upsampled, skip]) order, arch_unet.py:62), and the decoder conv reads the whole
// buffer as one operand.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "layers.cuh"

namespace n2n {

struct Buf { size_t off = 0; int Cb = 0; int lvl = 0; };

enum {
  B_CAT0, B_CAT1, B_CAT2, B_CAT3, B_CAT4,
  B_E0, B_E1, B_E2, B_E3, B_E4, B_E5, B_P5, B_E6,
  B_D5A, B_D5B, B_D4A, B_D4B, B_D3A, B_D3B, B_D2A, B_D2B, B_D1A, B_D1B, B_NA, B_NB, B_OUT,
  B_COUNT
};

struct LayerIO {
  int in_buf, in_cb0, in_cb;     // input blocks [in_cb0, in_cb0+in_cb) of in_buf
  int out_buf, out_cb0;          // output written at block offset out_cb0 of out_buf
  bool act;                      // LeakyReLU(0.2) after the layer
};

}  // namespace n2n

using namespace n2n;

enum { ARCH_UNET = 0, ARCH_RESNET = 1 };

struct n2n_unet_plan {
  int in_nc, out_nc, nf, N, H, W, dtype;
  bool bwd;
  int arch = ARCH_UNET;
  int nlayers = 25;                      // RESNET: 21 (state_dict order, arch_unet.py:279-347; layer 7 = the unused up5)
  // bf16 engine: the network input is kept as its 3x3 im2col (kb = ceil(9*in_nc/16) blocks), so
  // enc_conv0 and the raw-input part of dec_conv1a's concat are ONE K = 16*kb GEMM step each
  // instead of nine taps over a 16-channel block with in_nc real channels.
  bool im2col; int kb, skipb;
  size_t off_partial_skip = 0, off_wd20_full = 0; int splits_skip = 1;
  int nfb, c2b, inb, hb;
  LayerGeom L[25];
  LayerIO io[25];
  int dgrad_blocks[25];   // how many input blocks the layer's dgrad produces (0 = none)
  int splits[25];
  int head_splits = 0;                   // > 0: fused head backward (headbwd_umma.cu) with that many CTAs
  // 3x3 convs over a two-segment concat whose packed weights exceed shared memory (Cin = 2nf + nf):
  // forward = two launches over the K segments (the second adds the first's bf16 partial), input
  // gradient = two launches over the N segments; each half keeps its weights resident on the slab engine.
  bool ksplit[25];
  bool deconv_pair[25];   // ConvTranspose layers that run as two N = 2*Cout launches on the slab engine
  // No-grad plans on the bf16 engine: ConvTranspose layer dc and the 3x3 conv dc + 1 behind it run as ONE fused
  // layer (two launches, one per output-row parity) on composite weights — the upsampled tensor is never written
  // and the conv's nine taps over it collapse to four source-pixel taps (pack.cu: upfuse_pack_kernel).
  bool upfuse[25] = {false};
  UpConvGeom ug[25];
  size_t off_upw[25][2] = {{0}}, off_upbias[25] = {0}, off_upcorr[25] = {0};
  // training plans: transposed composites (fused input gradient), per-parity split-K partials of the composite weight
  // gradient, their dense reduction, border sums of dL/dy (chain rule back to the two layers' own gradients)
  size_t off_upwt[25] = {0}, off_dwcp[25][4] = {{0}}, off_dwc[25] = {0}, off_border[25] = {0};
  int splits_up[25] = {0};
  bool share_up[25] = {false};           // borrow the donor's forward composites of this level
  bool fused_layer(int i) const { return (L[i].kind == L_DECONV && upfuse[i]) || (i > 0 && L[i - 1].kind == L_DECONV && upfuse[i - 1]); }
  Buf act[B_COUNT], grd[B_COUNT];
  size_t off_wp[25], off_wd[25], off_bias[25], off_partial[25], off_bpartial[25];
  size_t total = 0;
  int fwd_launches = 0, bwd_launches = 0;
  // forward weights / padded biases may be borrowed from another plan of the same network that has already
  // packed them for this step (n2n_unet_share_weights): the pack is a function of the parameters only
  bool prepacked = false;                // n2n_unet_pack_weights ran for the next n2n_unet_forward (which then skips its pack step)
  const n2n_unet_plan* donor = nullptr;
  const void* donor_ws = nullptr;
  bool share[25] = {false};              // per layer: the donor packs this layer exactly as this plan would
  // backward: weight gradients run on a side stream, forked per layer from the input-gradient chain
  // (wgrad(i) and dgrad(i) both only READ grad(out_i)); the deep, launch-bound levels of the two
  // chains then overlap.  Joined before the partial reduction.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  ~n2n_unet_plan() {
    if (ev_ready) cudaEventDestroy(ev_ready);
    if (ev_done) cudaEventDestroy(ev_done);
    if (side) cudaStreamDestroy(side);
  }

  int lh(int lvl) const { return H >> lvl; }
  int lw(int lvl) const { return W >> lvl; }
  View view(const Buf* set, void* ws, int b, int cb0, int cb) const {
    const Buf& B = set[b];
    return make_view((char*)ws + B.off, dtype, N, lh(B.lvl), lw(B.lvl), B.Cb, cb0, cb);
  }
};

static void plan_layout(n2n_unet_plan* p) {
  const int nf = p->nf, in_nc = p->in_nc, out_nc = p->out_nc;
  p->nfb = cblocks(nf); p->c2b = cblocks(2 * nf); p->inb = cblocks(in_nc); p->hb = cblocks(96);
  const int nfb = p->nfb, c2b = p->c2b, inb = p->inb, hb = p->hb;
  auto setbuf = [&](int b, int Cb, int lvl) { p->act[b].Cb = Cb; p->act[b].lvl = lvl; p->grd[b].Cb = Cb; p->grd[b].lvl = lvl; };
  {
    const char* e1 = getenv("N2N_NO_SLAB"); const char* e2 = getenv("N2N_NO_IM2COL");
    p->im2col = p->dtype == N2N_BF16 && !(e1 && atoi(e1)) && !(e2 && atoi(e2));
    p->kb = cblocks(9 * in_nc);
    p->skipb = p->im2col ? p->kb : inb;
  }
  setbuf(B_CAT0, c2b + p->skipb, 0);
  setbuf(B_CAT1, c2b + nfb, 1); setbuf(B_CAT2, c2b + nfb, 2); setbuf(B_CAT3, c2b + nfb, 3);
  setbuf(B_CAT4, nfb + nfb, 4);
  setbuf(B_E0, nfb, 0); setbuf(B_E1, nfb, 0); setbuf(B_E2, nfb, 1); setbuf(B_E3, nfb, 2); setbuf(B_E4, nfb, 3);
  setbuf(B_E5, nfb, 4); setbuf(B_P5, nfb, 5); setbuf(B_E6, nfb, 5);
  setbuf(B_D5A, c2b, 4); setbuf(B_D5B, c2b, 4); setbuf(B_D4A, c2b, 3); setbuf(B_D4B, c2b, 3);
  setbuf(B_D3A, c2b, 2); setbuf(B_D3B, c2b, 2); setbuf(B_D2A, c2b, 1); setbuf(B_D2B, c2b, 1);
  setbuf(B_D1A, hb, 0); setbuf(B_D1B, hb, 0); setbuf(B_NA, hb, 0); setbuf(B_NB, hb, 0);
  setbuf(B_OUT, cblocks(out_nc), 0);

  auto conv3 = [&](int i, ChanSegs cin, int cout) { p->L[i].kind = L_CONV3; p->L[i].cin = cin; p->L[i].cout = cout; };
  auto conv1 = [&](int i, ChanSegs cin, int cout) { p->L[i].kind = L_CONV1; p->L[i].cin = cin; p->L[i].cout = cout; };
  auto deconv = [&](int i, int cin, int cout) { p->L[i].kind = L_DECONV; p->L[i].cin = chan1(cin); p->L[i].cout = cout; };
  auto setio = [&](int i, int ib, int icb0, int icb, int ob, int ocb0, bool act, int dgb) {
    p->io[i] = LayerIO{ib, icb0, icb, ob, ocb0, act};
    p->dgrad_blocks[i] = dgb;
  };
  // state_dict order (arch_unet.py:114-192)
  conv3(0, chan1(in_nc), nf);        setio(0, B_CAT0, c2b, inb, B_E0, 0, true, 0);
  conv3(1, chan1(nf), nf);           setio(1, B_E0, 0, nfb, B_E1, 0, true, nfb);
  conv3(2, chan1(nf), nf);           setio(2, B_CAT1, c2b, nfb, B_E2, 0, true, nfb);
  conv3(3, chan1(nf), nf);           setio(3, B_CAT2, c2b, nfb, B_E3, 0, true, nfb);
  conv3(4, chan1(nf), nf);           setio(4, B_CAT3, c2b, nfb, B_E4, 0, true, nfb);
  conv3(5, chan1(nf), nf);           setio(5, B_CAT4, nfb, nfb, B_E5, 0, true, nfb);
  conv3(6, chan1(nf), nf);           setio(6, B_P5, 0, nfb, B_E6, 0, true, nfb);
  deconv(7, nf, nf);                 setio(7, B_E6, 0, nfb, B_CAT4, 0, false, nfb);
  conv3(8, chan2(nf, nf), 2 * nf);   setio(8, B_CAT4, 0, 2 * nfb, B_D5A, 0, true, 2 * nfb);
  conv3(9, chan1(2 * nf), 2 * nf);   setio(9, B_D5A, 0, c2b, B_D5B, 0, true, c2b);
  const int cats[4] = {B_CAT3, B_CAT2, B_CAT1, B_CAT0};
  const int das[4] = {B_D4A, B_D3A, B_D2A, B_D1A}, dbs[4] = {B_D4B, B_D3B, B_D2B, B_D1B};
  const int prev[4] = {B_D5B, B_D4B, B_D3B, B_D2B};
  for (int k = 0; k < 4; ++k) {
    const int base = 10 + 3 * k;
    const bool last = (k == 3);
    deconv(base, 2 * nf, 2 * nf);    setio(base, prev[k], 0, c2b, cats[k], 0, false, c2b);
    if (!last) {
      conv3(base + 1, chan2(2 * nf, nf), 2 * nf);  setio(base + 1, cats[k], 0, c2b + nfb, das[k], 0, true, c2b + nfb);
      conv3(base + 2, chan1(2 * nf), 2 * nf);      setio(base + 2, das[k], 0, c2b, dbs[k], 0, true, c2b);
    } else {
      // arch_unet.py:177-181: literal 96-wide head; the skip is the raw input.
      conv3(base + 1, chan2(2 * nf, in_nc), 96);   setio(base + 1, cats[k], 0, c2b + inb, das[k], 0, true, c2b);
      conv3(base + 2, chan1(96), 96);              setio(base + 2, das[k], 0, hb, dbs[k], 0, true, hb);
    }
  }
  conv1(22, chan1(96), 96);          setio(22, B_D1B, 0, hb, B_NA, 0, true, hb);
  conv1(23, chan1(96), 96);          setio(23, B_NA, 0, hb, B_NB, 0, true, hb);
  conv1(24, chan1(96), out_nc);      setio(24, B_NB, 0, hb, B_OUT, 0, false, hb);

  for (int i = 0; i < 25; ++i) {
    const LayerGeom& G = p->L[i];
    const int lv = p->act[p->io[i].in_buf].lvl;
    const char* eks = getenv("N2N_NO_KSPLIT");
    p->ksplit[i] = !(eks && atoi(eks)) && G.kind == L_CONV3 && G.cin.n == 2 && slab_geometry_ok(p->dtype, p->lh(lv), p->lw(lv)) &&
                   !slab_weights_fit(9, G.cin_blocks(), G.cout_blocks() * 16, true) &&
                   slab_weights_fit(9, cblocks(G.cin.cnt[0]), G.cout_blocks() * 16, true) &&
                   slab_weights_fit(9, cblocks(G.cin.cnt[1]), G.cout_blocks() * 16, true) &&
                   slab_weights_fit(9, G.cout_blocks(), cblocks(G.cin.cnt[0]) * 16, true) && !(p->im2col && i == 20);
    p->deconv_pair[i] = false;
    if (p->L[i].kind == L_DECONV) {
      const int lvl = p->act[p->io[i].in_buf].lvl;
      p->deconv_pair[i] = slab_deconv_pair_ok(p->dtype, p->lh(lvl), p->lw(lvl), p->L[i].cin_blocks(), p->L[i].cout_blocks());
    }
  }
  for (int dc = 7; dc <= 19; dc += 3) {
    const int lv = p->act[p->io[dc].in_buf].lvl;                 // source (pre-upsampling) level
    UpConvGeom& U = p->ug[dc];
    U.ci_blocks = p->L[dc].cin_blocks(); U.co_blocks = p->L[dc + 1].cout_blocks();
    U.skip_im2col = dc == 19 && p->im2col;
    U.skip_blocks = dc == 19 ? p->skipb : nfb;
    { const char* e = getenv("N2N_UPFUSE_LEVELS");      // diagnostic: bit (dc - 7) / 3 enables the fusion of that level
      if (e && !((atoi(e) >> ((dc - 7) / 3)) & 1)) { p->upfuse[dc] = false; continue; } }
    p->upfuse[dc] = p->dtype == N2N_BF16 && !p->ksplit[dc + 1] &&
                    slab_upconv_ok(p->dtype, p->N, p->lh(lv), p->lw(lv), U.ci_blocks, U.skip_blocks, U.co_blocks, U.region_bytes());
    if (p->upfuse[dc] && p->bwd) {
      // the backward runs in composite form too: transposed composites resident for the input gradient, the slab
      // weight-gradient engine for the four parity launches
      const char* e = getenv("N2N_NO_UPFUSE_TRAIN");
      const long long src_tiles = (long long)p->N * ((p->lh(lv) + 15) / 16) * ((p->lw(lv) + 7) / 8);
      p->splits_up[dc] = wgrad_slab_splits(4, U.ci_blocks, U.co_blocks, false, src_tiles);
      // the composite backward trades FLOPs for launches (four parity weight-gradient launches + a skip launch instead of
      // two); an A/B of "all levels" against "levels with >= 256 / >= 1000 source tiles only" was within the box noise
      // (5.28 / 5.36 / 5.36 ms per step), so every eligible level is fused
      long long min_tiles = 0;
      { const char* m = getenv("N2N_UPFUSE_TRAIN_MIN_TILES"); if (m) min_tiles = atoll(m); }
      p->upfuse[dc] = !(e && atoi(e)) && p->splits_up[dc] > 0 && src_tiles >= min_tiles &&
                      slab_upconv_ok(p->dtype, p->N, p->lh(lv), p->lw(lv), U.co_blocks, 0, U.ci_blocks, upconv_wt_bytes(U));
    }
  }
  // ---- workspace layout ----
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t es = dtype_size(p->dtype);
  for (int b = 0; b < B_COUNT; ++b) {
    if (b == B_OUT) continue;   // forward output goes straight to the caller's NCHW tensor
    p->act[b].off = take((size_t)p->N * p->act[b].Cb * p->lh(p->act[b].lvl) * p->lw(p->act[b].lvl) * 16 * es);
  }
  for (int i = 0; i < 25; ++i) {
    size_t wbytes = p->L[i].fwd_pack_bytes(p->dtype);
    if (p->im2col && i == 20) {
      // im2col form of dec_conv1a: nine tap slabs over the c2b upsampled blocks, then ONE more slab (the im2col tap of
      // the raw input) at slab index 9 — its own group of three blocks.  Sizing this by the layer's nominal
      // (c2b + inb) blocks is one slab short whenever c2b + inb needs no more groups than c2b (n_feature 4, 16, 32, ...).
      const size_t im2col_form = packed_weight_bytes(p->dtype, 9, p->L[i].cout_blocks() * 16, c2b) +
                                 packed_weight_bytes(p->dtype, 1, p->L[i].cout_blocks() * 16, p->kb);
      if (im2col_form > wbytes) wbytes = im2col_form;
    }
    p->off_wp[i] = take(wbytes);
    p->off_bias[i] = take(p->L[i].cout_blocks() * 16 * sizeof(float));
  }
  for (int dc = 7; dc <= 19; dc += 3) {
    if (!p->upfuse[dc]) continue;
    for (int py = 0; py < 2; ++py) p->off_upw[dc][py] = take(p->ug[dc].region_bytes());
    p->off_upbias[dc] = take(p->ug[dc].co_blocks * 16 * sizeof(float));
    p->off_upcorr[dc] = take(9 * p->ug[dc].co_blocks * 16 * sizeof(float));
  }
  if (p->bwd) {
    for (int dc = 7; dc <= 19; dc += 3) {
      if (!p->upfuse[dc]) continue;
      const UpConvGeom& U = p->ug[dc];
      const size_t plane = (size_t)U.ci_blocks * 16 * U.co_blocks * 16 * sizeof(float);
      p->off_upwt[dc] = take(upconv_wt_bytes(U));
      for (int q = 0; q < 4; ++q) p->off_dwcp[dc][q] = take((size_t)p->splits_up[dc] * 4 * plane);
      p->off_dwc[dc] = take(16 * plane);
      p->off_border[dc] = take(8 * U.co_blocks * 16 * sizeof(float));
    }
    p->head_splits = head_bwd_splits(p->dtype, p->hb, p->out_nc, p->N, p->H, p->W);
    for (int b = 0; b < B_COUNT; ++b)
      p->grd[b].off = take((size_t)p->N * p->grd[b].Cb * p->lh(p->grd[b].lvl) * p->lw(p->grd[b].lvl) * 16 * es);
    for (int i = 0; i < 25; ++i) {
      const LayerIO& io = p->io[i];
      p->splits[i] = layer_wgrad_splits(p->L[i], p->dtype, p->N, p->lh(p->act[io.in_buf].lvl), p->lw(p->act[io.in_buf].lvl));
      // dgrad weights are packed for the full input width so that dL/dx can be served too
      p->off_wd[i] = take(p->L[i].dgrad_pack_bytes(p->dtype, p->L[i].cin_blocks()));
      if (p->im2col && (i == 0 || i == 20)) {
        // im2col forms: layer 0 is one pair over kb blocks; layer 20 = nine pairs over the c2b
        // upsampled blocks + one pair over the kb im2col blocks (own partial buffer)
        LayerGeom G = p->L[i];
        if (i == 0) { G.kind = L_CONV1; G.cin = chan1(16 * p->kb); }
        else G.cin = chan1(2 * nf);
        p->splits[i] = layer_wgrad_splits(G, p->dtype, p->N, p->H, p->W);
        if (i == 20) {
          LayerGeom S = p->L[i]; S.kind = L_CONV1; S.cin = chan1(16 * p->kb);
          p->splits_skip = layer_wgrad_splits(S, p->dtype, p->N, p->H, p->W);
          p->off_partial_skip = take(S.partial_bytes(p->splits_skip));
          p->off_wd20_full = take(p->L[i].dgrad_pack_bytes(p->dtype, p->L[i].cin_blocks()));   // only used when dL/dx is wanted
        }
      }
      if (i >= 8 && p->L[i - 1].kind == L_DECONV && p->upfuse[i - 1] && !(p->im2col && i == 20)) {
        // fused up-conv: this conv's own weight-gradient launch only covers its skip channels
        LayerGeom S = p->L[i]; S.cin = chan1(p->L[i].cin.cnt[1]);
        p->splits[i] = layer_wgrad_splits(S, p->dtype, p->N, p->lh(p->act[io.in_buf].lvl), p->lw(p->act[io.in_buf].lvl));
      }
      // nin_a / nin_b: their weight gradients come out of the fused head backward, one partial per CTA of it
      if ((i == 22 || i == 23) && p->head_splits > 0) p->splits[i] = p->head_splits;
      p->off_partial[i] = take(p->L[i].partial_bytes(p->splits[i]));
      // (the fused level-1 backward sums dec_conv1a's bias gradient in its im2col launch: splits_skip rows)
      p->off_bpartial[i] = take(p->L[i].bias_partial_bytes((i == 20 && p->splits_skip > p->splits[i]) ? p->splits_skip : p->splits[i]));
    }
  }
  p->total = off;
}

static void resnet_layout(n2n_unet_plan* p);
static int resnet_forward(n2n_unet_plan* p, const float* const* params, const float* x, float* y, void* ws, cudaStream_t st);
static int resnet_backward(n2n_unet_plan* p, const float* const* params, const float* dy, float* const* grads, float* dx,
                           void* ws, cudaStream_t st);

// arch_unet.RESNET (arch_unet.py:263-409): same plan type and the same forward / backward / workspace entry points as the
// UNet; params / grads are the 42 tensors in state_dict order (up5.deconv.* are registered by the reference but never
// used by forward: they are ignored, and their gradient slots are left untouched).  Needs out_nc == in_nc (global
// residual, :409); H and W are unconstrained (no pooling).
extern "C" int n2n_resnet_plan_create(n2n_unet_plan** plan, int in_nc, int out_nc, int n_feature, int n, int h, int w,
                                      int dtype, int with_backward) {
  N2N_CHECK_ARG(plan != nullptr, "resnet_plan_create: plan is NULL");
  N2N_CHECK_ARG(in_nc >= 1 && in_nc <= 16 && out_nc == in_nc, "resnet_plan_create: need 1 <= in_nc == out_nc <= 16 (x + in_, arch_unet.py:409)");
  N2N_CHECK_ARG(n_feature >= 1 && n_feature <= 256 && n >= 1 && h >= 1 && w >= 1, "resnet_plan_create: bad geometry");
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, "resnet_plan_create: bad dtype %d", dtype);
  n2n_unet_plan* p = new n2n_unet_plan();
  p->in_nc = in_nc; p->out_nc = out_nc; p->nf = n_feature; p->N = n; p->H = h; p->W = w; p->dtype = dtype;
  p->bwd = with_backward != 0;
  p->arch = ARCH_RESNET; p->nlayers = 21;
  resnet_layout(p);
  *plan = p;
  return 0;
}

extern "C" int n2n_unet_plan_create(n2n_unet_plan** plan, int in_nc, int out_nc, int n_feature, int n, int h, int w,
                                    int dtype, int with_backward) {
  N2N_CHECK_ARG(plan != nullptr, "unet_plan_create: plan is NULL");
  N2N_CHECK_ARG(in_nc >= 1 && in_nc <= 16 && out_nc >= 1 && out_nc <= 16, "unet_plan_create: in_nc/out_nc must be in [1,16]");
  N2N_CHECK_ARG(n_feature >= 1 && n_feature <= 256, "unet_plan_create: n_feature out of range");
  N2N_CHECK_ARG(n >= 1 && h >= 32 && w >= 32 && h % 32 == 0 && w % 32 == 0,
                "unet_plan_create: need n>=1 and H,W multiples of 32 (got %d,%d,%d)", n, h, w);
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, "unet_plan_create: bad dtype %d", dtype);
  n2n_unet_plan* p = new n2n_unet_plan();
  p->in_nc = in_nc; p->out_nc = out_nc; p->nf = n_feature; p->N = n; p->H = h; p->W = w; p->dtype = dtype;
  p->bwd = with_backward != 0;
  plan_layout(p);
  { const char* e = getenv("N2N_NO_SIDE");
    if (p->bwd && dtype == N2N_BF16 && !(e && atoi(e))) {
      if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        p->side = nullptr;        // no device yet / no resources: stay on one stream
      }
    }
  }
  *plan = p;
  return 0;
}
extern "C" void n2n_unet_plan_destroy(n2n_unet_plan* plan) { delete plan; }
extern "C" size_t n2n_unet_workspace_bytes(const n2n_unet_plan* plan) { return plan ? plan->total : 0; }
// Diagnostic / layer-level parity: copy one activation buffer of the last forward on `ws` out as fp32 NCHW
// [N][16 * blocks][H_l][W_l] (all channel blocks of the buffer, padding included).  buffer: 0..4 = concat buffers of
// levels 0..4, 5.. = enc_conv0..5 outputs, pool5, enc_conv6, dec_conv5a, 5b, 4a, 4b, 3a, 3b, 2a, 2b, 1a, 1b, nin_a, nin_b.
// dims (may be NULL) receives {channels, H_l, W_l}.  Returns the element count, 0 for an empty buffer, < 0 on error.
extern "C" long long n2n_unet_read_activation(const n2n_unet_plan* p, void* ws, int buffer, float* out, int* dims, void* stream) {
  if (!p || !ws || buffer < 0 || buffer >= B_COUNT || buffer == B_OUT) { set_error("unet_read_activation: bad arguments"); return N2N_ERR_ARG; }
  const Buf& B = p->act[buffer];
  const int h = p->lh(B.lvl), w = p->lw(B.lvl);
  if (dims) { dims[0] = B.Cb * 16; dims[1] = h; dims[2] = w; }
  const long long count = (long long)p->N * B.Cb * 16 * h * w;
  if (count == 0 || !out) return count;
  const int r = launch_c16_to_nchw(p->view(p->act, ws, buffer, 0, B.Cb), p->dtype, out, B.Cb * 16, (cudaStream_t)stream);
  return r < 0 ? r : count;
}
extern "C" int n2n_unet_launches(const n2n_unet_plan* plan, int backward) {
  return plan ? (backward ? plan->bwd_launches : plan->fwd_launches) : 0;
}

// Let `plan` read its forward weights and padded biases from `donor`'s workspace instead of packing its own copy:
// valid when donor's n2n_unet_forward on the SAME parameters precedes plan's on the same stream (one training step
// runs the no-grad full-resolution pass and the half-resolution pass on the same weights).  Returns 0 when
// shared (layer by layer: a layer the two plans lay out differently is still packed locally), 1 when the plans
// are not the same network (nothing changed), donor = NULL to stop.
extern "C" int n2n_unet_share_weights(n2n_unet_plan* plan, const n2n_unet_plan* donor, const void* donor_ws) {
  N2N_CHECK_ARG(plan != nullptr, "unet_share_weights: plan is NULL");
  plan->donor = nullptr; plan->donor_ws = nullptr;
  for (int i = 0; i < 25; ++i) plan->share_up[i] = false;
  if (!donor) return 0;
  if (plan->arch != ARCH_UNET || donor->arch != ARCH_UNET) return 1;
  N2N_CHECK_ARG(donor_ws != nullptr && donor != plan, "unet_share_weights: bad donor");
  const bool same = donor->in_nc == plan->in_nc && donor->out_nc == plan->out_nc && donor->nf == plan->nf &&
                    donor->dtype == plan->dtype && donor->im2col == plan->im2col && donor->kb == plan->kb;
  if (!same) return 1;
  // per layer: the deepest levels of the two plans can pick different launch forms (pair-form ConvTranspose
  // needs H, W >= 4), which changes that layer's packed layout only
  for (int i = 0; i < 25; ++i)
    plan->share[i] = !donor->fused_layer(i) && !plan->fused_layer(i) &&
                     donor->ksplit[i] == plan->ksplit[i] && donor->deconv_pair[i] == plan->deconv_pair[i] &&
                     donor->L[i].fwd_pack_bytes(donor->dtype) == plan->L[i].fwd_pack_bytes(plan->dtype);
  // forward composites of a fused level could be borrowed too, but the borrower of a training step is the plan with
  // the backward pass, which has to run the composite pack anyway (it also emits the transposed composites)
  for (int dc = 7; dc <= 19; dc += 3)
    plan->share_up[dc] = !plan->bwd && donor->upfuse[dc] && plan->upfuse[dc] &&
                         donor->ug[dc].region_bytes() == plan->ug[dc].region_bytes() && donor->ug[dc].skip_im2col == plan->ug[dc].skip_im2col;
  plan->donor = donor; plan->donor_ws = donor_ws;
  return 0;
}

// The pack step of n2n_unet_forward: PyTorch-layout fp32 parameters -> the engines' packed weight images, padded biases
// and (fused levels) the composite up-conv weights.  A function of the parameters only.
static int unet_pack_weights(n2n_unet_plan* p, const float* const* params, void* ws, cudaStream_t st) {
  const int dt = p->dtype;
  {
    std::vector<PackJob> jobs;
    BiasPadJob bj[25];
    UpFuseJob uj[5];
    int nuj = 0;
    for (int dc = 7; dc <= 19; dc += 3) {
      if (!p->upfuse[dc]) continue;
      if (p->donor && p->share_up[dc] && !p->bwd) continue;         // forward composites borrowed, nothing else to pack
      const UpConvGeom& U = p->ug[dc];
      const LayerGeom& D = p->L[dc]; const LayerGeom& A = p->L[dc + 1];
      if (p->bwd && !U.skip_im2col) {
        // input gradient of the conv's skip channels: its own transposed weights, skip segment only
        PackJob k = make_dgrad_pack(A, params[2 * (dc + 1)], (char*)ws + p->off_wd[dc + 1], U.skip_blocks);
        k.nseg.n = 1; k.nseg.src0[0] = D.cout; k.nseg.cnt[0] = A.cin.real() - D.cout; k.nseg.dst0[0] = 0;
        jobs.push_back(k);
      }
      UpFuseJob& j = uj[nuj++];
      if (p->bwd) { j.dst_t = (char*)ws + p->off_upwt[dc]; j.gco = (U.co_blocks + 2) / 3; j.ci_rows = U.ci_blocks * 16; }
      j.w3 = params[2 * (dc + 1)]; j.b3 = params[2 * (dc + 1) + 1]; j.wd = params[2 * dc]; j.bd = params[2 * dc + 1];
      j.Ci = D.cin.real(); j.Cu = D.cout; j.Cs = A.cin.real() - D.cout; j.Co = A.cout;
      j.ngroups = U.gu(); j.co_pad = U.co_blocks * 16;
      j.bias_full = (float*)((char*)ws + p->off_upbias[dc]); j.corr = (float*)((char*)ws + p->off_upcorr[dc]);
      for (int py = 0; py < 2; ++py) {
        j.dst[py] = (char*)ws + p->off_upw[dc][py];
        if (p->donor && p->share_up[dc]) continue;                  // skip slabs live in the donor's regions
        // the conv's skip-channel slabs behind the composite slabs of each parity's region
        PackJob k = make_fwd_pack(A, params[2 * (dc + 1)], (char*)ws + p->off_upw[dc][py] + U.skip_base());
        k.cin_blocks = U.skip_blocks;
        if (U.skip_im2col) { k.ntaps = 1; k.im2col_nc = p->in_nc; k.im2col_c0 = D.cout; }
        else { k.cseg.n = 1; k.cseg.src0[0] = D.cout; k.cseg.cnt[0] = j.Cs; k.cseg.dst0[0] = 0; }
        jobs.push_back(k);
      }
    }
    for (int i = 0; i < 25; ++i) {
      if (p->donor && p->share[i]) {
        // forward pack borrowed
      } else if (p->fused_layer(i)) {
        // packed above in fused form
      } else if (p->im2col && i == 0) {
        PackJob j = make_fwd_pack(p->L[0], params[0], (char*)ws + p->off_wp[0]);
        j.ntaps = 1; j.cin_blocks = p->kb; j.im2col_nc = p->in_nc; j.im2col_c0 = 0;
        jobs.push_back(j);
      } else if (p->im2col && i == 20) {
        PackJob j = make_fwd_pack(p->L[20], params[40], (char*)ws + p->off_wp[20]);
        j.cin_blocks = p->c2b; j.cseg = chan1(2 * p->nf).to_segs();          // the nine taps over the upsampled channels
        jobs.push_back(j);
        PackJob k = j;                                                         // + the im2col tap (slab 9, group 0)
        const int mg = (p->c2b + 2) / 3;
        k.dst = (char*)ws + p->off_wp[20] + (size_t)9 * mg * 3 * j.nout_pad * 32;
        k.ntaps = 1; k.cin_blocks = p->kb; k.im2col_nc = p->in_nc; k.im2col_c0 = 2 * p->nf;
        jobs.push_back(k);
      } else if (p->ksplit[i]) {
        const LayerGeom& G = p->L[i];
        const int b0 = cblocks(G.cin.cnt[0]), b1 = cblocks(G.cin.cnt[1]);
        PackJob j = make_fwd_pack(G, params[2 * i], (char*)ws + p->off_wp[i]);
        j.cin_blocks = b0; j.cseg = chan1(G.cin.cnt[0]).to_segs();
        jobs.push_back(j);
        PackJob k = make_fwd_pack(G, params[2 * i], (char*)ws + p->off_wp[i] + packed_weight_bytes(dt, 9, j.nout_pad, b0));
        k.cin_blocks = b1; k.cseg.n = 1; k.cseg.src0[0] = G.cin.cnt[0]; k.cseg.cnt[0] = G.cin.cnt[1]; k.cseg.dst0[0] = 0;
        jobs.push_back(k);
      } else if (p->deconv_pair[i]) {
        jobs.push_back(make_deconv_pair_pack(p->L[i], params[2 * i], (char*)ws + p->off_wp[i]));
      } else {
        jobs.push_back(make_fwd_pack(p->L[i], params[2 * i], (char*)ws + p->off_wp[i]));
      }
      if (p->bwd && p->fused_layer(i)) {
        // composite backward: no per-layer input-gradient pack (the skip segment's was queued above)
      } else if (p->bwd && p->ksplit[i]) {
        const LayerGeom& G = p->L[i];
        const int b0 = cblocks(G.cin.cnt[0]), b1 = cblocks(G.cin.cnt[1]);
        PackJob j = make_dgrad_pack(G, params[2 * i], (char*)ws + p->off_wd[i], b0);      // rows = first segment's channels
        j.nseg = chan1(G.cin.cnt[0]).to_segs();
        jobs.push_back(j);
        PackJob k = make_dgrad_pack(G, params[2 * i], (char*)ws + p->off_wd[i] + packed_weight_bytes(dt, 9, b0 * 16, G.cout_blocks()), b1);
        k.nseg.n = 1; k.nseg.src0[0] = G.cin.cnt[0]; k.nseg.cnt[0] = G.cin.cnt[1]; k.nseg.dst0[0] = 0;
        jobs.push_back(k);
      } else if (p->bwd)   // im2col mode: dec_conv1a's input gradient is only needed for the upsampled blocks
        jobs.push_back(make_dgrad_pack(p->L[i], params[2 * i], (char*)ws + p->off_wd[i],
                                       (p->im2col && i == 20) ? p->c2b : p->L[i].cin_blocks()));
      bj[i] = BiasPadJob{params[2 * i + 1], (float*)((char*)ws + p->off_bias[i]), p->L[i].cout, p->L[i].cout_blocks() * 16};
    }
    if (!jobs.empty()) N2N_TRY(launch_pack(jobs.data(), (int)jobs.size(), dt, st));
    N2N_TRY(launch_upfuse_pack(uj, nuj, st));
    if (!p->donor) N2N_TRY(launch_bias_pad(bj, 25, st));
  }
  return 0;
}

extern "C" int n2n_unet_pack_weights(n2n_unet_plan* p, const float* const* params, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && ws, "unet_pack_weights: null argument");
  if (p->arch != ARCH_UNET) return 0;                       // the RESNET plan packs inside its forward
  N2N_TRY(unet_pack_weights(p, params, ws, (cudaStream_t)stream));
  p->prepacked = true;
  return 0;
}

extern "C" int n2n_unet_forward(n2n_unet_plan* p, const float* const* params, const float* x, float* y,
                                void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && x && y && ws, "unet_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long launches0 = g_launch_count;
  if (p->arch == ARCH_RESNET) {
    N2N_TRY(resnet_forward(p, params, x, y, ws, st));
    p->fwd_launches = (int)(g_launch_count - launches0);
    return 0;
  }
  const int dt = p->dtype;
  // weights -> engine layout (fwd; dgrad copies too when a backward will follow) — unless n2n_unet_pack_weights already
  // did it for this call
  if (p->prepacked) p->prepacked = false;
  else N2N_TRY(unet_pack_weights(p, params, ws, st));
  // input image -> skip block of the level-0 concat buffer (pool0 = x, arch_unet.py:200)
  bool enc0_done = false;     // the fused input stage also produced enc_conv0's output
  if (p->im2col) {
    const int r = launch_input_stage(x, p->in_nc, params[0], params[1], p->nf, p->view(p->act, ws, B_CAT0, p->c2b, p->kb),
                                     p->view(p->act, ws, B_E0, 0, p->nfb), 0.2f, st);
    if (r < 0) return r;
    enc0_done = r == 0;
    if (!enc0_done) N2N_TRY(launch_nchw_to_im2col9(x, p->in_nc, p->view(p->act, ws, B_CAT0, p->c2b, p->kb), dt, st));
  } else {
    N2N_TRY(launch_nchw_to_c16(x, p->in_nc, p->view(p->act, ws, B_CAT0, p->c2b, p->inb), dt, st));
  }

  auto fw = [&](int i) -> const void* {
    return (p->donor && p->share[i]) ? (const char*)p->donor_ws + p->donor->off_wp[i] : (const char*)ws + p->off_wp[i];
  };
  auto fb = [&](int i) -> const float* {
    return (const float*)(p->donor ? (const char*)p->donor_ws + p->donor->off_bias[i] : (const char*)ws + p->off_bias[i]);
  };
  auto run_layer = [&](int i, int pool_buf = -1, int pool_cb0 = 0) -> int {
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    View xin = p->view(p->act, ws, io.in_buf, io.in_cb0, io.in_cb);
    const void* wp = fw(i);
    const float* bias = fb(i);
    if (L.kind == L_DECONV && p->upfuse[i]) return 0;            // runs inside the next layer's fused launches
    if (i > 0 && p->L[i - 1].kind == L_DECONV && p->upfuse[i - 1]) {
      const int dc = i - 1;
      const LayerIO& dio = p->io[dc];
      const UpConvGeom& U = p->ug[dc];
      View xsrc = p->view(p->act, ws, dio.in_buf, dio.in_cb0, dio.in_cb);
      View skip = p->view(p->act, ws, io.in_buf, p->L[dc].cout_blocks(), U.skip_blocks);
      View yfull = p->view(p->act, ws, io.out_buf, io.out_cb0, L.cout_blocks());
      const bool borrowed = p->donor && p->share_up[dc];
      const char* wbase = borrowed ? (const char*)p->donor_ws : (const char*)ws;
      const n2n_unet_plan* wp_plan = borrowed ? p->donor : p;
      for (int py = 0; py < 2; ++py) {
        TapGemm g = make_upconv_fwd(U, dt, xsrc, skip, yfull, py, wbase + wp_plan->off_upw[dc][py],
                                    (const float*)(wbase + wp_plan->off_upbias[dc]),
                                    (const float*)(wbase + wp_plan->off_upcorr[dc]));
        if (io.act) { g.act = 1; g.slope = 0.2f; }
        const int r = launch_tapgemm(g, st);
        if (r != 0) return r;
      }
      return 0;
    }
    if (p->im2col && (i == 0 || i == 20)) {
      TapGemm g;
      g.dtype = dt; g.nout = L.cout_blocks() * 16; g.w = wp; g.bias = bias;
      g.y = p->view(p->act, ws, io.out_buf, io.out_cb0, L.cout_blocks());
      g.act = 1; g.slope = 0.2f;
      View skip = p->view(p->act, ws, B_CAT0, p->c2b, p->kb);
      if (i == 0) {
        g.x[0] = skip; g.ntaps = 1; g.cin_blocks = p->kb;                 // tap 0 = (0, 0), slab 0
      } else {
        g.x[0] = p->view(p->act, ws, B_CAT0, 0, p->c2b); g.x[1] = skip;
        g.ntaps = 10; g.cin_blocks = p->c2b; g.view_blocks[1] = p->kb;
        for (int t = 0; t < 9; ++t) { g.tap_dy[t] = t / 3 - 1; g.tap_dx[t] = t % 3 - 1; g.tap_view[t] = 0; g.tap_slab[t] = t; }
        g.tap_view[9] = 1; g.tap_slab[9] = 9;
      }
      return launch_tapgemm(g, st);
    }
    if (L.kind == L_DECONV) {
      View yfull = p->view(p->act, ws, io.out_buf, io.out_cb0, L.cout_blocks());
      if (p->deconv_pair[i]) {
        for (int a = 0; a < 2; ++a) N2N_TRY(launch_tapgemm(make_deconv_fwd_pair(L, dt, xin, yfull, a, wp, bias), st));
        return 0;
      }
      for (int ab = 0; ab < 4; ++ab) {
        TapGemm g = make_deconv_fwd(L, dt, xin, yfull, ab / 2, ab % 2, wp, bias);
        N2N_TRY(launch_tapgemm(g, st));
      }
      return 0;
    }
    if (p->ksplit[i]) {
      const int b0 = cblocks(L.cin.cnt[0]), b1 = cblocks(L.cin.cnt[1]);
      View yv = p->view(p->act, ws, io.out_buf, io.out_cb0, L.cout_blocks());
      LayerGeom A = L; A.cin = chan1(L.cin.cnt[0]);
      LayerGeom B = L; B.cin = chan1(L.cin.cnt[1]);
      TapGemm g0 = make_conv_fwd(A, dt, p->view(p->act, ws, io.in_buf, io.in_cb0, b0), yv, wp, bias);     // partial (+bias), no act
      N2N_TRY(launch_tapgemm(g0, st));
      TapGemm g1 = make_conv_fwd(B, dt, p->view(p->act, ws, io.in_buf, io.in_cb0 + b0, b1), yv,
                                 (const char*)wp + packed_weight_bytes(dt, 9, L.cout_blocks() * 16, b0), nullptr);
      g1.has_addend = true; g1.addend = yv;
      if (io.act) { g1.act = 1; g1.slope = 0.2f; }
      return launch_tapgemm(g1, st);
    }
    TapGemm g;
    if (io.out_buf == B_OUT) {
      View dummy = p->view(p->act, ws, B_NB, 0, L.cout_blocks());   // geometry only
      g = make_conv_fwd(L, dt, xin, dummy, wp, bias);
      g.out_nchw = y; g.out_c = p->out_nc;
    } else {
      g = make_conv_fwd(L, dt, xin, p->view(p->act, ws, io.out_buf, io.out_cb0, L.cout_blocks()), wp, bias);
    }
    if (io.act) { g.act = 1; g.slope = 0.2f; }
    if (pool_buf >= 0) {
      // MaxPool2d(2) fused behind the conv (arch_unet.py:203-219); the un-pooled activation is only
      // kept when a backward pass will need it (max routing + LeakyReLU sign)
      g.has_pool = true; g.pool = p->view(p->act, ws, pool_buf, pool_cb0, p->nfb);
      g.store_y = p->bwd;
    }
    return launch_tapgemm(g, st);
  };
  if (!enc0_done) N2N_TRY(run_layer(0));
  N2N_TRY(run_layer(1, B_CAT1, p->c2b));
  N2N_TRY(run_layer(2, B_CAT2, p->c2b));
  N2N_TRY(run_layer(3, B_CAT3, p->c2b));
  N2N_TRY(run_layer(4, B_CAT4, p->nfb));
  N2N_TRY(run_layer(5, B_P5, 0));
  for (int i = 6; i < 22; ++i) N2N_TRY(run_layer(i));
  // nin_a -> nin_b -> nin_c: one fused kernel on the bf16 engine (the intermediates are only
  // written when a backward pass will read them)
  int head = kSgNotEligible;
  if (dt == N2N_BF16) {
    HeadChain h;
    h.x = p->view(p->act, ws, B_D1B, 0, p->hb);
    h.in_blocks = p->hb; h.mid_blocks = p->hb; h.mid_channels = 96; h.out_nc = p->out_nc;
    h.wa = fw(22); h.wb = fw(23);
    h.bias_a = fb(22); h.bias_b = fb(23);
    h.wc = params[2 * 24]; h.bias_c = params[2 * 24 + 1];
    h.slope = 0.2f; h.has_save = p->bwd;
    h.save_a = p->view(p->act, ws, B_NA, 0, p->hb); h.save_b = p->view(p->act, ws, B_NB, 0, p->hb);
    h.out_nchw = y;
    head = launch_head_chain(h, st);
    if (head < 0) return head;
  }
  if (head == kSgNotEligible)
    for (int i = 22; i < 25; ++i) N2N_TRY(run_layer(i));
  p->fwd_launches = (int)(g_launch_count - launches0);
  return 0;
}

extern "C" int n2n_unet_backward(n2n_unet_plan* p, const float* const* params, const float* dy,
                                 float* const* grads, float* dx, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && dy && grads && ws, "unet_backward: null argument");
  N2N_CHECK_ARG(p->bwd, "unet_backward: plan was created without with_backward");
  cudaStream_t st = (cudaStream_t)stream;
  const long long launches0 = g_launch_count;
  if (p->arch == ARCH_RESNET) {
    N2N_TRY(resnet_backward(p, params, dy, grads, dx, ws, st));
    p->bwd_launches = (int)(g_launch_count - launches0);
    return 0;
  }
  const int dt = p->dtype;
  const bool want_dx = dx != nullptr;

  N2N_TRY(launch_nchw_to_c16(dy, p->out_nc, p->view(p->grd, ws, B_OUT, 0, cblocks(p->out_nc)), dt, st));

  // grad of layer i's OUTPUT lives in grd[out_buf] blocks [out_cb0, +cout_blocks).
  cudaStream_t main_st = st;
  const bool use_side = p->side != nullptr && !profiling_active();
  auto wgrad = [&](int i) -> int {
    cudaStream_t st = main_st;
    if (use_side) {
      // fork: everything issued on the main stream so far (in particular grad(out_i)) precedes this wgrad
      N2N_CUDA(cudaEventRecord(p->ev_ready, main_st));
      N2N_CUDA(cudaStreamWaitEvent(p->side, p->ev_ready, 0));
      st = p->side;
    }
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    View xin = p->view(p->act, ws, io.in_buf, io.in_cb0, io.in_cb);
    View gy = p->view(p->grd, ws, io.out_buf, io.out_cb0, L.cout_blocks());
    float* partial = (float*)((char*)ws + p->off_partial[i]);
    float* bpartial = (float*)((char*)ws + p->off_bpartial[i]);
    if (p->im2col && (i == 0 || i == 20)) {
      View skip = p->view(p->act, ws, B_CAT0, p->c2b, p->kb);
      LayerGeom S = L; S.kind = L_CONV1; S.cin = chan1(16 * p->kb);          // one (0,0) pair over the im2col blocks
      if (i == 0) return launch_tapwgrad(make_conv_wgrad(S, dt, skip, gy, partial, bpartial, p->splits[0]), st);
      LayerGeom M = L; M.cin = chan1(2 * p->nf);                             // nine pairs over the upsampled blocks
      N2N_TRY(launch_tapwgrad(make_conv_wgrad(M, dt, p->view(p->act, ws, B_CAT0, 0, p->c2b), gy, partial, bpartial, p->splits[20]), st));
      return launch_tapwgrad(make_conv_wgrad(S, dt, skip, gy, (float*)((char*)ws + p->off_partial_skip), nullptr, p->splits_skip), st);
    }
    TapWgrad g = (L.kind == L_DECONV) ? make_deconv_wgrad(L, dt, xin, gy, partial, bpartial, p->splits[i])
                                      : make_conv_wgrad(L, dt, xin, gy, partial, bpartial, p->splits[i]);
    return launch_tapwgrad(g, st);
  };
  // dgrad of layer i into grd[in_buf] (first `blocks` input blocks); mask_by_input multiplies by
  // lrelu'(input) (valid because the input is the in-place activated output of the previous
  // layer, arch_unet.py:113); add_existing accumulates onto what the skip path already wrote.
  auto dgrad = [&](int i, int blocks, bool mask_by_input, bool add_existing) -> int {
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    View gy = p->view(p->grd, ws, io.out_buf, io.out_cb0, L.cout_blocks());
    View gx = p->view(p->grd, ws, io.in_buf, io.in_cb0, blocks);
    const void* wd = (char*)ws + p->off_wd[i];
    if (p->im2col && i == 20 && blocks != p->c2b) {
      // dL/dx requested: repack dec_conv1a's transposed weights for all input blocks (raw-input block included)
      PackJob j = make_dgrad_pack(L, params[2 * i], (char*)ws + p->off_wd20_full, L.cin_blocks());
      N2N_TRY(launch_pack(&j, 1, dt, st));
      wd = (char*)ws + p->off_wd20_full;
    }
    if (p->ksplit[i] && blocks == L.cin_blocks() && !mask_by_input && !add_existing) {
      const int b0 = cblocks(L.cin.cnt[0]), b1 = cblocks(L.cin.cnt[1]);
      TapGemm g0 = make_conv_dgrad(L, dt, gy, p->view(p->grd, ws, io.in_buf, io.in_cb0, b0), wd, b0);
      N2N_TRY(launch_tapgemm(g0, st));
      TapGemm g1 = make_conv_dgrad(L, dt, gy, p->view(p->grd, ws, io.in_buf, io.in_cb0 + b0, b1),
                                   (const char*)wd + packed_weight_bytes(dt, 9, b0 * 16, L.cout_blocks()), b1);
      return launch_tapgemm(g1, st);
    }
    TapGemm g = (L.kind == L_DECONV) ? make_deconv_dgrad(L, dt, gy, gx, wd)
                                     : make_conv_dgrad(L, dt, gy, gx, wd, L.cin_blocks());
    g.nout = blocks * 16;
    if (mask_by_input) { g.has_mask = true; g.mask = p->view(p->act, ws, io.in_buf, io.in_cb0, blocks); g.slope = 0.2f; }
    if (add_existing) { g.has_addend = true; g.addend = gx; }
    return launch_tapgemm(g, st);
  };
  auto unpool = [&](int act_buf, int gpool_buf, int gpool_cb0) -> int {
    return launch_unpool_lrelu(p->view(p->act, ws, act_buf, 0, p->nfb), p->view(p->grd, ws, gpool_buf, gpool_cb0, p->nfb),
                               p->view(p->grd, ws, act_buf, 0, p->nfb), 0.2f, dt, st);
  };

  // Fused up-conv level (deconv dc + conv dc + 1) in composite form: four parity launches of the weight-gradient engine
  // (dWc partials), the conv's own weight-gradient launch restricted to its skip channels (it also sums the bias
  // gradient), ONE input-gradient launch with the transposed composites, the skip channels' input gradient, and the
  // border sums the chain rule needs (reduced / chained after the join, see below).
  View bviews[5]; float* bouts[5]; int bpads[5]; int nborder = 0;      // border sums of dL/dy of the fused levels: one launch at the end
  auto fused_bwd = [&](int dc) -> int {
    const UpConvGeom& U = p->ug[dc];
    const int ca = dc + 1;
    const LayerIO& dio = p->io[dc];
    const LayerIO& aio = p->io[ca];
    const LayerGeom& A = p->L[ca];
    View xsrc = p->view(p->act, ws, dio.in_buf, dio.in_cb0, dio.in_cb);
    View gy = p->view(p->grd, ws, aio.out_buf, aio.out_cb0, A.cout_blocks());
    const int upb = p->L[dc].cout_blocks();
    View skip_act = p->view(p->act, ws, aio.in_buf, upb, U.skip_blocks);
    {
      cudaStream_t wst = main_st;
      if (use_side) {
        N2N_CUDA(cudaEventRecord(p->ev_ready, main_st));
        N2N_CUDA(cudaStreamWaitEvent(p->side, p->ev_ready, 0));
        wst = p->side;
      }
      for (int q = 0; q < 4; ++q)
        N2N_TRY(launch_tapwgrad(make_upconv_wgrad(U, dt, xsrc, gy, q >> 1, q & 1, (float*)((char*)ws + p->off_dwcp[dc][q]),
                                                  p->splits_up[dc]), wst));
      if (U.skip_im2col) {
        LayerGeom S = A; S.kind = L_CONV1; S.cin = chan1(16 * p->kb);
        N2N_TRY(launch_tapwgrad(make_conv_wgrad(S, dt, skip_act, gy, (float*)((char*)ws + p->off_partial_skip),
                                                (float*)((char*)ws + p->off_bpartial[ca]), p->splits_skip), wst));
      } else {
        LayerGeom S = A; S.cin = chan1(A.cin.cnt[1]);
        N2N_TRY(launch_tapwgrad(make_conv_wgrad(S, dt, skip_act, gy, (float*)((char*)ws + p->off_partial[ca]),
                                                (float*)((char*)ws + p->off_bpartial[ca]), p->splits[ca]), wst));
      }
    }
    bviews[nborder] = gy; bouts[nborder] = (float*)((char*)ws + p->off_border[dc]); bpads[nborder] = U.co_blocks * 16; ++nborder;
    // input gradients (main stream)
    TapGemm g = make_upconv_dgrad(U, dt, gy, p->view(p->grd, ws, dio.in_buf, dio.in_cb0, U.ci_blocks), (const char*)ws + p->off_upwt[dc]);
    g.has_mask = true; g.mask = xsrc; g.slope = 0.2f;          // the deconv's input is an activated conv output
    N2N_TRY(launch_tapgemm(g, st));
    if (!U.skip_im2col) {
      TapGemm gs = make_conv_dgrad(A, dt, gy, p->view(p->grd, ws, aio.in_buf, upb, U.skip_blocks), (const char*)ws + p->off_wd[ca], U.skip_blocks);
      N2N_TRY(launch_tapgemm(gs, st));
    } else if (want_dx) {
      // dL/dx requested: the raw-input channels of dec_conv1a's concat (packed on the spot; not on the training path)
      PackJob j = make_dgrad_pack(A, params[2 * ca], (char*)ws + p->off_wd20_full, p->inb);
      j.nseg.n = 1; j.nseg.src0[0] = p->L[dc].cout; j.nseg.cnt[0] = p->in_nc; j.nseg.dst0[0] = 0;
      N2N_TRY(launch_pack(&j, 1, dt, st));
      TapGemm gs = make_conv_dgrad(A, dt, gy, p->view(p->grd, ws, B_CAT0, p->c2b, p->inb), (const char*)ws + p->off_wd20_full, p->inb);
      N2N_TRY(launch_tapgemm(gs, st));
    }
    return 0;
  };

  // head + level-0 decoder
  if (p->head_splits > 0) {
    // input gradients of the three 1x1 layers and the weight gradients of nin_a / nin_b in one kernel
    HeadBwd h;
    h.blocks = p->hb; h.channels = 96; h.out_nc = p->out_nc; h.slope = 0.2f;
    h.gout = dy; h.wc = params[2 * 24];
    h.wb_dgrad = (char*)ws + p->off_wd[23]; h.wa_dgrad = (char*)ws + p->off_wd[22];
    h.act_nb = p->view(p->act, ws, B_NB, 0, p->hb); h.act_na = p->view(p->act, ws, B_NA, 0, p->hb);
    h.act_d1b = p->view(p->act, ws, B_D1B, 0, p->hb);
    h.g_d1b = p->view(p->grd, ws, B_D1B, 0, p->hb);
    h.splits = p->head_splits;
    h.partial_b = (float*)((char*)ws + p->off_partial[23]); h.bpartial_b = (float*)((char*)ws + p->off_bpartial[23]);
    h.partial_a = (float*)((char*)ws + p->off_partial[22]); h.bpartial_a = (float*)((char*)ws + p->off_bpartial[22]);
    N2N_TRY(wgrad(24));                       // reads grad(out) and nin_b's activation only
    const int hb = launch_head_bwd(h, st);
    if (hb < 0) return hb;
    N2N_CHECK_ARG(hb == 0, "unet_backward: fused head backward declined a geometry the plan was built for");
  } else {
    for (int i = 24; i >= 22; --i) {
      N2N_TRY(wgrad(i));
      N2N_TRY(dgrad(i, p->L[i].cin_blocks(), true, false));
    }
  }
  N2N_TRY(wgrad(21)); N2N_TRY(dgrad(21, p->L[21].cin_blocks(), true, false));
  // decoder levels 1..5: dec_conv{k}a then up{k} (one fused composite level when the plan fused them), dec_conv{k+1}b
  for (int dc = 19; dc >= 7; dc -= 3) {
    if (p->upfuse[dc]) {
      N2N_TRY(fused_bwd(dc));
    } else {
      N2N_TRY(wgrad(dc + 1));                                                      // dec_conv a -> concat
      if (dc == 19) N2N_TRY(dgrad(20, (want_dx || !p->im2col) ? p->c2b + p->inb : p->c2b, false, false));
      else N2N_TRY(dgrad(dc + 1, p->L[dc + 1].cin_blocks(), false, false));
      N2N_TRY(wgrad(dc)); N2N_TRY(dgrad(dc, p->L[dc].cin_blocks(), true, false));  // up_k: input is an activated conv output
    }
    if (dc > 7) { N2N_TRY(wgrad(dc - 1)); N2N_TRY(dgrad(dc - 1, p->c2b, true, false)); }   // dec_conv{k+1}b
  }
  N2N_TRY(wgrad(6)); N2N_TRY(dgrad(6, p->nfb, false, false));                     // enc_conv6: input is pool5 (no act)
  N2N_TRY(unpool(B_E5, B_P5, 0));
  // encoder: the pooled tensors also feed the skip connections, whose grads are already in the
  // concat-grad buffers -> accumulate.
  N2N_TRY(wgrad(5)); N2N_TRY(dgrad(5, p->nfb, false, true)); N2N_TRY(unpool(B_E4, B_CAT4, p->nfb));
  N2N_TRY(wgrad(4)); N2N_TRY(dgrad(4, p->nfb, false, true)); N2N_TRY(unpool(B_E3, B_CAT3, p->c2b));
  N2N_TRY(wgrad(3)); N2N_TRY(dgrad(3, p->nfb, false, true)); N2N_TRY(unpool(B_E2, B_CAT2, p->c2b));
  N2N_TRY(wgrad(2)); N2N_TRY(dgrad(2, p->nfb, false, true)); N2N_TRY(unpool(B_E1, B_CAT1, p->c2b));
  N2N_TRY(wgrad(1)); N2N_TRY(dgrad(1, p->nfb, true, false));
  N2N_TRY(wgrad(0));
  if (want_dx) {
    N2N_TRY(dgrad(0, p->inb, false, true));
    N2N_TRY(launch_c16_to_nchw(p->view(p->grd, ws, B_CAT0, p->c2b, p->inb), dt, dx, p->in_nc, st));
  }
  if (use_side) {   // join
    N2N_CUDA(cudaEventRecord(p->ev_done, p->side));
    N2N_CUDA(cudaStreamWaitEvent(st, p->ev_done, 0));
  }
  N2N_TRY(launch_border_sums(bviews, bouts, bpads, nborder, st));
  // partials -> PyTorch-layout fp32 gradients
  {
    std::vector<UnpackJob> jobs;
    UpFuseGradJob gj[5];
    int ngj = 0;
    for (int i = 0; i < 25; ++i) {
      if (p->L[i].kind == L_DECONV && p->upfuse[i]) continue;                 // gradients come out of the chain-rule kernel
      UnpackJob j = make_unpack(p->L[i], (const float*)((char*)ws + p->off_partial[i]),
                                (const float*)((char*)ws + p->off_bpartial[i]), p->splits[i], grads[2 * i], grads[2 * i + 1]);
      if (i > 0 && p->L[i - 1].kind == L_DECONV && p->upfuse[i - 1]) {
        // fused up-conv level: (1) the four parity partials of the composite weight gradient -> one dense [16][ci][co]
        // buffer, (2) this conv's skip-channel weights + its bias, (3) queue the chain rule for after the reductions
        const int dc = i - 1;
        const UpConvGeom& U = p->ug[dc];
        const int cip = U.ci_blocks * 16, cop = U.co_blocks * 16;
        float* dense = (float*)((char*)ws + p->off_dwc[dc]);
        for (int q = 0; q < 4; ++q) {
          UnpackJob d;
          d.partial = (const float*)((char*)ws + p->off_dwcp[dc][q]); d.bias_partial = nullptr;
          d.dst_w = dense + (size_t)q * 4 * cip * cop; d.dst_b = nullptr;
          d.splits = p->splits_up[dc]; d.ntaps = 4; d.npad = cop; d.cpad = cip; d.bias_rows = 0;
          d.s_t = (long long)cip * cop; d.s_n = 1; d.s_c = cop;
          d.nseg.n = 1; d.nseg.src0[0] = 0; d.nseg.cnt[0] = cop; d.nseg.dst0[0] = 0;
          d.cseg.n = 1; d.cseg.src0[0] = 0; d.cseg.cnt[0] = cip; d.cseg.dst0[0] = 0;
          jobs.push_back(d);
        }
        if (U.skip_im2col) {
          j.partial = (const float*)((char*)ws + p->off_partial_skip); j.splits = p->splits_skip; j.bias_rows = p->splits_skip;
          j.ntaps = 1; j.cpad = 16 * p->kb; j.im2col_nc = p->in_nc; j.im2col_c0 = p->L[dc].cout;
        } else {
          j.cpad = U.skip_blocks * 16;
          j.cseg.n = 1; j.cseg.src0[0] = p->L[dc].cout; j.cseg.cnt[0] = p->L[i].cin.cnt[1]; j.cseg.dst0[0] = 0;
        }
        jobs.push_back(j);
        UpFuseGradJob& g = gj[ngj++];
        g.dwc = dense; g.w3 = params[2 * i]; g.wd = params[2 * dc]; g.bd = params[2 * dc + 1];
        g.border = (const float*)((char*)ws + p->off_border[dc]); g.db3 = grads[2 * i + 1];
        g.dw3 = grads[2 * i]; g.dwd = grads[2 * dc]; g.dbd = grads[2 * dc + 1];
        g.Ci = p->L[dc].cin.real(); g.Cu = p->L[dc].cout; g.Cs = p->L[i].cin.real() - p->L[dc].cout; g.Co = p->L[i].cout;
        g.ci_pad = cip; g.co_pad = cop;
        continue;
      }
      if (p->im2col && i == 0) {
        j.ntaps = 1; j.cpad = 16 * p->kb; j.im2col_nc = p->in_nc; j.im2col_c0 = 0;
      } else if (p->im2col && i == 20) {
        j.cpad = 16 * p->c2b; j.cseg = chan1(2 * p->nf).to_segs();
        UnpackJob k = j;
        k.partial = (const float*)((char*)ws + p->off_partial_skip); k.bias_partial = nullptr; k.dst_b = nullptr;
        k.splits = p->splits_skip; k.ntaps = 1; k.cpad = 16 * p->kb; k.im2col_nc = p->in_nc; k.im2col_c0 = 2 * p->nf;
        jobs.push_back(k);
      }
      jobs.push_back(j);
    }
    N2N_TRY(launch_unpack(jobs.data(), (int)jobs.size(), st));
    N2N_TRY(launch_upfuse_grad(gj, ngj, st));
  }
  p->bwd_launches = (int)(g_launch_count - launches0);
  return 0;
}

// ------------------------------------------------------------------------------------------
// RESNET (arch_unet.py:263-409, non-blindspot): the UNet's convolutions with no pooling and no up-sampling — every
// tensor at full resolution, torch.cat([x, pool_k]) skips, global residual.  Buffers ("CAT_k" = [decoder output |
// encoder skip], both written in place by their producers, so no concat is materialised):
//   enc1 -> CAT1[c2b:], enc2 -> CAT2[c2b:], enc3 -> CAT3[c2b:], enc4 -> CAT4[nfb:], enc5 -> E5, enc6 -> CAT4[0:nfb]
//   dec5a: CAT4 -> D5A, dec5b -> CAT3[0:c2b]; dec4a: CAT3 -> D4A, dec4b -> CAT2[0:]; dec3a: CAT2 -> D3A, dec3b -> CAT1[0:];
//   dec2a: CAT1 -> D2A, dec2b -> CAT0[0:c2b]; dec1a: CAT0 = [dec2b | input] -> D1A; dec1b -> D1B; nin_a/b/c; + input.
// ------------------------------------------------------------------------------------------
static void resnet_layout(n2n_unet_plan* p) {
  const int nf = p->nf, in_nc = p->in_nc, out_nc = p->out_nc;
  p->nfb = cblocks(nf); p->c2b = cblocks(2 * nf); p->inb = cblocks(in_nc); p->hb = cblocks(96);
  p->im2col = false; p->kb = 0; p->skipb = p->inb;
  const int nfb = p->nfb, c2b = p->c2b, inb = p->inb, hb = p->hb;
  auto setbuf = [&](int b, int Cb) { p->act[b].Cb = Cb; p->act[b].lvl = 0; p->grd[b].Cb = Cb; p->grd[b].lvl = 0; };
  for (int b = 0; b < B_COUNT; ++b) setbuf(b, 0);
  setbuf(B_CAT0, c2b + inb); setbuf(B_CAT1, c2b + nfb); setbuf(B_CAT2, c2b + nfb); setbuf(B_CAT3, c2b + nfb);
  setbuf(B_CAT4, nfb + nfb); setbuf(B_E0, nfb); setbuf(B_E5, nfb);
  setbuf(B_D5A, c2b); setbuf(B_D4A, c2b); setbuf(B_D3A, c2b); setbuf(B_D2A, c2b);
  setbuf(B_D1A, hb); setbuf(B_D1B, hb); setbuf(B_NA, hb); setbuf(B_NB, hb); setbuf(B_OUT, cblocks(out_nc));
  auto conv3 = [&](int i, ChanSegs cin, int cout) { p->L[i].kind = L_CONV3; p->L[i].cin = cin; p->L[i].cout = cout; };
  auto conv1 = [&](int i, ChanSegs cin, int cout) { p->L[i].kind = L_CONV1; p->L[i].cin = cin; p->L[i].cout = cout; };
  auto setio = [&](int i, int ib, int icb0, int icb, int ob, int ocb0) { p->io[i] = LayerIO{ib, icb0, icb, ob, ocb0, true}; p->dgrad_blocks[i] = icb; };
  conv3(0, chan1(in_nc), nf);  setio(0, B_CAT0, c2b, inb, B_E0, 0);
  conv3(1, chan1(nf), nf);     setio(1, B_E0, 0, nfb, B_CAT1, c2b);
  conv3(2, chan1(nf), nf);     setio(2, B_CAT1, c2b, nfb, B_CAT2, c2b);
  conv3(3, chan1(nf), nf);     setio(3, B_CAT2, c2b, nfb, B_CAT3, c2b);
  conv3(4, chan1(nf), nf);     setio(4, B_CAT3, c2b, nfb, B_CAT4, nfb);
  conv3(5, chan1(nf), nf);     setio(5, B_CAT4, nfb, nfb, B_E5, 0);
  conv3(6, chan1(nf), nf);     setio(6, B_E5, 0, nfb, B_CAT4, 0);
  p->L[7].kind = L_DECONV; p->L[7].cin = chan1(nf); p->L[7].cout = nf; p->io[7] = LayerIO{B_E5, 0, 0, B_E5, 0, false};   // up5: unused
  conv3(8, chan2(nf, nf), 2 * nf);        setio(8, B_CAT4, 0, 2 * nfb, B_D5A, 0);
  conv3(9, chan1(2 * nf), 2 * nf);        setio(9, B_D5A, 0, c2b, B_CAT3, 0);
  conv3(10, chan2(2 * nf, nf), 2 * nf);   setio(10, B_CAT3, 0, c2b + nfb, B_D4A, 0);
  conv3(11, chan1(2 * nf), 2 * nf);       setio(11, B_D4A, 0, c2b, B_CAT2, 0);
  conv3(12, chan2(2 * nf, nf), 2 * nf);   setio(12, B_CAT2, 0, c2b + nfb, B_D3A, 0);
  conv3(13, chan1(2 * nf), 2 * nf);       setio(13, B_D3A, 0, c2b, B_CAT1, 0);
  conv3(14, chan2(2 * nf, nf), 2 * nf);   setio(14, B_CAT1, 0, c2b + nfb, B_D2A, 0);
  conv3(15, chan1(2 * nf), 2 * nf);       setio(15, B_D2A, 0, c2b, B_CAT0, 0);
  conv3(16, chan2(2 * nf, in_nc), 96);    setio(16, B_CAT0, 0, c2b + inb, B_D1A, 0);
  conv3(17, chan1(96), 96);               setio(17, B_D1A, 0, hb, B_D1B, 0);
  conv1(18, chan1(96), 96);               setio(18, B_D1B, 0, hb, B_NA, 0);
  conv1(19, chan1(96), 96);               setio(19, B_NA, 0, hb, B_NB, 0);
  conv1(20, chan1(96), out_nc);           setio(20, B_NB, 0, hb, B_OUT, 0);
  p->io[20].act = false;
  for (int i = 0; i < 25; ++i) { p->ksplit[i] = false; p->deconv_pair[i] = false; p->upfuse[i] = false; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t es = dtype_size(p->dtype);
  const size_t px = (size_t)p->N * p->H * p->W * 16 * es;
  for (int b = 0; b < B_COUNT; ++b) {
    if (b == B_OUT || p->act[b].Cb == 0) continue;
    p->act[b].off = take((size_t)p->act[b].Cb * px);
  }
  for (int i = 0; i < 21; ++i) {
    if (i == 7) continue;
    p->off_wp[i] = take(p->L[i].fwd_pack_bytes(p->dtype));
    p->off_bias[i] = take(p->L[i].cout_blocks() * 16 * sizeof(float));
  }
  if (p->bwd) {
    // the fused head backward covers the UNet's head geometry (same three layers here); tiles of 8 x 16 need H, W >= 4
    p->head_splits = (p->H >= 4 && p->W >= 4) ? head_bwd_splits(p->dtype, p->hb, p->out_nc, p->N, p->H, p->W) : 0;
    for (int b = 0; b < B_COUNT; ++b)
      if (p->grd[b].Cb) p->grd[b].off = take((size_t)p->grd[b].Cb * px);
    for (int i = 0; i < 21; ++i) {
      if (i == 7) continue;
      p->splits[i] = layer_wgrad_splits(p->L[i], p->dtype, p->N, p->H, p->W);
      p->off_wd[i] = take(p->L[i].dgrad_pack_bytes(p->dtype, p->L[i].cin_blocks()));
      if ((i == 18 || i == 19) && p->head_splits > 0) p->splits[i] = p->head_splits;
      p->off_partial[i] = take(p->L[i].partial_bytes(p->splits[i]));
      p->off_bpartial[i] = take(p->L[i].bias_partial_bytes(p->splits[i]));
    }
  }
  p->total = off;
}

static int resnet_forward(n2n_unet_plan* p, const float* const* params, const float* x, float* y, void* ws, cudaStream_t st) {
  const int dt = p->dtype;
  {
    std::vector<PackJob> jobs;
    BiasPadJob bj[21];
    int nb = 0;
    for (int i = 0; i < 21; ++i) {
      if (i == 7) continue;
      const LayerGeom& G = p->L[i];
      jobs.push_back(make_fwd_pack(G, params[2 * i], (char*)ws + p->off_wp[i]));
      if (p->bwd) {
        if (G.cin.n == 2) {
          // input gradient of a concat consumer = two launches over its two channel segments (the first is masked by the
          // decoder activation it feeds back into, the skip segment is not — see resnet_backward)
          const int b0 = cblocks(G.cin.cnt[0]), b1 = cblocks(G.cin.cnt[1]);
          PackJob j = make_dgrad_pack(G, params[2 * i], (char*)ws + p->off_wd[i], b0);
          j.nseg = chan1(G.cin.cnt[0]).to_segs();
          jobs.push_back(j);
          PackJob k = make_dgrad_pack(G, params[2 * i], (char*)ws + p->off_wd[i] + packed_weight_bytes(dt, 9, b0 * 16, G.cout_blocks()), b1);
          k.nseg.n = 1; k.nseg.src0[0] = G.cin.cnt[0]; k.nseg.cnt[0] = G.cin.cnt[1]; k.nseg.dst0[0] = 0;
          jobs.push_back(k);
        } else {
          jobs.push_back(make_dgrad_pack(G, params[2 * i], (char*)ws + p->off_wd[i], G.cin_blocks()));
        }
      }
      bj[nb++] = BiasPadJob{params[2 * i + 1], (float*)((char*)ws + p->off_bias[i]), G.cout, G.cout_blocks() * 16};
    }
    N2N_TRY(launch_pack(jobs.data(), (int)jobs.size(), dt, st));
    N2N_TRY(launch_bias_pad(bj, nb, st));
  }
  N2N_TRY(launch_nchw_to_c16(x, p->in_nc, p->view(p->act, ws, B_CAT0, p->c2b, p->inb), dt, st));
  auto run_layer = [&](int i) -> int {
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    View xin = p->view(p->act, ws, io.in_buf, io.in_cb0, io.in_cb);
    TapGemm g;
    const void* wp = (const char*)ws + p->off_wp[i];
    const float* bias = (const float*)((const char*)ws + p->off_bias[i]);
    if (io.out_buf == B_OUT) {
      View dummy = p->view(p->act, ws, B_NB, 0, L.cout_blocks());
      g = make_conv_fwd(L, dt, xin, dummy, wp, bias);
      g.out_nchw = y; g.out_c = p->out_nc;
    } else {
      g = make_conv_fwd(L, dt, xin, p->view(p->act, ws, io.out_buf, io.out_cb0, L.cout_blocks()), wp, bias);
    }
    if (io.act) { g.act = 1; g.slope = 0.2f; }
    return launch_tapgemm(g, st);
  };
  for (int i = 0; i < 18; ++i)
    if (i != 7) N2N_TRY(run_layer(i));
  int head = kSgNotEligible;
  if (dt == N2N_BF16 && p->H >= 4 && p->W >= 4) {
    HeadChain h;
    h.x = p->view(p->act, ws, B_D1B, 0, p->hb);
    h.in_blocks = p->hb; h.mid_blocks = p->hb; h.mid_channels = 96; h.out_nc = p->out_nc;
    h.wa = (const char*)ws + p->off_wp[18]; h.wb = (const char*)ws + p->off_wp[19];
    h.bias_a = (const float*)((const char*)ws + p->off_bias[18]); h.bias_b = (const float*)((const char*)ws + p->off_bias[19]);
    h.wc = params[2 * 20]; h.bias_c = params[2 * 20 + 1];
    h.slope = 0.2f; h.has_save = p->bwd;
    h.save_a = p->view(p->act, ws, B_NA, 0, p->hb); h.save_b = p->view(p->act, ws, B_NB, 0, p->hb);
    h.out_nchw = y;
    head = launch_head_chain(h, st);
    if (head < 0) return head;
  }
  if (head == kSgNotEligible)
    for (int i = 18; i < 21; ++i) N2N_TRY(run_layer(i));
  // global residual (arch_unet.py:409)
  return launch_add_inplace(y, x, (long long)p->N * p->out_nc * p->H * p->W, st);
}

static int resnet_backward(n2n_unet_plan* p, const float* const* params, const float* dy, float* const* grads, float* dx,
                           void* ws, cudaStream_t st) {
  const int dt = p->dtype;
  const bool want_dx = dx != nullptr;
  N2N_TRY(launch_nchw_to_c16(dy, p->out_nc, p->view(p->grd, ws, B_OUT, 0, cblocks(p->out_nc)), dt, st));
  auto wgrad = [&](int i) -> int {
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    return launch_tapwgrad(make_conv_wgrad(L, dt, p->view(p->act, ws, io.in_buf, io.in_cb0, io.in_cb),
                                           p->view(p->grd, ws, io.out_buf, io.out_cb0, L.cout_blocks()),
                                           (float*)((char*)ws + p->off_partial[i]), (float*)((char*)ws + p->off_bpartial[i]), p->splits[i]), st);
  };
  // grd[buf] always holds the gradient w.r.t. the producer's PRE-activation output: a dgrad multiplies by
  // lrelu'(its input) (= the in-place activated output of the producer, arch_unet.py:276) on the way out.
  auto dgrad = [&](int i, bool mask, bool add_existing) -> int {
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    View gy = p->view(p->grd, ws, io.out_buf, io.out_cb0, L.cout_blocks());
    View gx = p->view(p->grd, ws, io.in_buf, io.in_cb0, io.in_cb);
    TapGemm g = make_conv_dgrad(L, dt, gy, gx, (char*)ws + p->off_wd[i], L.cin_blocks());
    if (mask) { g.has_mask = true; g.mask = p->view(p->act, ws, io.in_buf, io.in_cb0, io.in_cb); g.slope = 0.2f; }
    if (add_existing) { g.has_addend = true; g.addend = gx; }
    return launch_tapgemm(g, st);
  };
  // concat consumer: decoder segment (masked by the decoder activation) and skip segment (NOT masked here: the skip
  // tensor also feeds the next encoder conv, whose dgrad adds its share and applies the mask once to the sum)
  auto dgrad_cat = [&](int i, bool want_skip) -> int {
    const LayerIO& io = p->io[i];
    const LayerGeom& L = p->L[i];
    const int b0 = cblocks(L.cin.cnt[0]), b1 = cblocks(L.cin.cnt[1]);
    View gy = p->view(p->grd, ws, io.out_buf, io.out_cb0, L.cout_blocks());
    TapGemm g0 = make_conv_dgrad(L, dt, gy, p->view(p->grd, ws, io.in_buf, io.in_cb0, b0), (char*)ws + p->off_wd[i], b0);
    g0.has_mask = true; g0.mask = p->view(p->act, ws, io.in_buf, io.in_cb0, b0); g0.slope = 0.2f;
    N2N_TRY(launch_tapgemm(g0, st));
    if (!want_skip) return 0;
    TapGemm g1 = make_conv_dgrad(L, dt, gy, p->view(p->grd, ws, io.in_buf, io.in_cb0 + b0, b1),
                                 (const char*)ws + p->off_wd[i] + packed_weight_bytes(dt, 9, b0 * 16, L.cout_blocks()), b1);
    return launch_tapgemm(g1, st);
  };
  if (p->head_splits > 0) {
    HeadBwd h;
    h.blocks = p->hb; h.channels = 96; h.out_nc = p->out_nc; h.slope = 0.2f;
    h.gout = dy; h.wc = params[2 * 20];
    h.wb_dgrad = (char*)ws + p->off_wd[19]; h.wa_dgrad = (char*)ws + p->off_wd[18];
    h.act_nb = p->view(p->act, ws, B_NB, 0, p->hb); h.act_na = p->view(p->act, ws, B_NA, 0, p->hb);
    h.act_d1b = p->view(p->act, ws, B_D1B, 0, p->hb);
    h.g_d1b = p->view(p->grd, ws, B_D1B, 0, p->hb);
    h.splits = p->head_splits;
    h.partial_b = (float*)((char*)ws + p->off_partial[19]); h.bpartial_b = (float*)((char*)ws + p->off_bpartial[19]);
    h.partial_a = (float*)((char*)ws + p->off_partial[18]); h.bpartial_a = (float*)((char*)ws + p->off_bpartial[18]);
    N2N_TRY(wgrad(20));
    const int hb = launch_head_bwd(h, st);
    if (hb < 0) return hb;
    N2N_CHECK_ARG(hb == 0, "resnet_backward: fused head backward declined a geometry the plan was built for");
  } else {
    for (int i = 20; i >= 18; --i) { N2N_TRY(wgrad(i)); N2N_TRY(dgrad(i, true, false)); }
  }
  N2N_TRY(wgrad(17)); N2N_TRY(dgrad(17, true, false));
  N2N_TRY(wgrad(16)); N2N_TRY(dgrad_cat(16, want_dx));
  for (int a = 14; a >= 8; a -= 2) {         // (dec_conv{k}b, dec_conv{k}a) for k = 2, 3, 4, 5
    N2N_TRY(wgrad(a + 1)); N2N_TRY(dgrad(a + 1, true, false));
    N2N_TRY(wgrad(a)); N2N_TRY(dgrad_cat(a, true));
  }
  N2N_TRY(wgrad(6)); N2N_TRY(dgrad(6, true, false));
  for (int i = 5; i >= 2; --i) { N2N_TRY(wgrad(i)); N2N_TRY(dgrad(i, true, true)); }
  N2N_TRY(wgrad(1)); N2N_TRY(dgrad(1, true, false));
  N2N_TRY(wgrad(0));
  if (want_dx) {
    N2N_TRY(dgrad(0, false, true));
    N2N_TRY(launch_c16_to_nchw(p->view(p->grd, ws, B_CAT0, p->c2b, p->inb), dt, dx, p->in_nc, st));
    N2N_TRY(launch_add_inplace(dx, dy, (long long)p->N * p->in_nc * p->H * p->W, st));      // d(x + in_)/d in_
  }
  UnpackJob jobs[21];
  int nj = 0;
  for (int i = 0; i < 21; ++i) {
    if (i == 7) continue;
    jobs[nj++] = make_unpack(p->L[i], (const float*)((char*)ws + p->off_partial[i]), (const float*)((char*)ws + p->off_bpartial[i]),
                             p->splits[i], grads[2 * i], grads[2 * i + 1]);
  }
  return launch_unpack(jobs, nj, st);
}

// ------------------------------------------------------------------------------------------
// Output adapter (adapter.py:5-26)
// ------------------------------------------------------------------------------------------
namespace n2n {      // adapter_fused.cu: CUDA-core direct convolutions for C in {1, 3}, hidden 16
struct AdapterFusedWs { size_t off_const, off_h, off_gh, off_p1, off_p2, total; };
AdapterFusedWs adapter_fused_layout(int C, int n, int h, int w, bool bwd);
bool adapter_fused_ok(int C, int hid, int w);
int adapter_fused_forward(int C, int dtype, const float* const* params, const float* noisy, const float* base_out, float* out, void* ws,
                          int n, int h, int w, bool bwd, cudaStream_t st);
int adapter_fused_backward(int C, int dtype, const float* const* params, const float* noisy, const float* base_out, const float* dout,
                           float* const* grads, void* ws, int n, int h, int w, cudaStream_t st);
}  // namespace n2n

struct n2n_adapter_plan {
  int C, hid, N, H, W, dtype; bool bwd;
  bool fused = false;      // adapter_fused.cu path (fp32 math in both precision modes)
  LayerGeom L[2];
  size_t off_cat, off_h, off_gout, off_gh, off_wp[2], off_wd1, off_bias[2], off_partial[2], off_bpartial[2];
  int splits[2];
  size_t total;
};

extern "C" int n2n_adapter_plan_create(n2n_adapter_plan** plan, int channels, int hidden, int n, int h, int w,
                                       int dtype, int with_backward) {
  N2N_CHECK_ARG(plan && channels >= 1 && channels <= 16 && hidden >= 1 && hidden <= 256 && n >= 1 && h >= 1 && w >= 1,
                "adapter_plan_create: bad arguments");
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, "adapter_plan_create: bad dtype");
  n2n_adapter_plan* p = new n2n_adapter_plan();
  p->C = channels; p->hid = hidden; p->N = n; p->H = h; p->W = w; p->dtype = dtype; p->bwd = with_backward != 0;
  { const char* e = getenv("N2N_NO_ADAPTER_FUSED"); p->fused = adapter_fused_ok(channels, hidden, w) && !(e && atoi(e)); }
  if (p->fused) {
    p->total = adapter_fused_layout(channels, n, h, w, p->bwd).total;
    *plan = p;
    return 0;
  }
  p->L[0].kind = L_CONV3; p->L[0].cin = chan2(channels, channels); p->L[0].cout = hidden;   // cat[noisy, base_out]
  p->L[1].kind = L_CONV3; p->L[1].cin = chan1(hidden); p->L[1].cout = channels;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t px = (size_t)n * h * w * 16 * dtype_size(dtype);
  p->off_cat = take(2 * px);
  p->off_h = take(cblocks(hidden) * px);
  for (int i = 0; i < 2; ++i) {
    p->off_wp[i] = take(p->L[i].fwd_pack_bytes(dtype));
    p->off_bias[i] = take(p->L[i].cout_blocks() * 16 * sizeof(float));
  }
  for (int i = 0; i < 2; ++i) p->splits[i] = layer_wgrad_splits(p->L[i], dtype, n, h, w);
  if (p->bwd) {
    p->off_gout = take(px);
    p->off_gh = take(cblocks(hidden) * px);
    p->off_wd1 = take(p->L[1].dgrad_pack_bytes(dtype, p->L[1].cin_blocks()));
    for (int i = 0; i < 2; ++i) {
      p->off_partial[i] = take(p->L[i].partial_bytes(p->splits[i]));
      p->off_bpartial[i] = take(p->L[i].bias_partial_bytes(p->splits[i]));
    }
  }
  p->total = off;
  *plan = p;
  return 0;
}
extern "C" void n2n_adapter_plan_destroy(n2n_adapter_plan* plan) { delete plan; }
extern "C" size_t n2n_adapter_workspace_bytes(const n2n_adapter_plan* plan) { return plan ? plan->total : 0; }

extern "C" int n2n_adapter_forward(n2n_adapter_plan* p, const float* const* params, const float* noisy,
                                   const float* base_out, float* out, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && noisy && base_out && out && ws, "adapter_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->fused) return adapter_fused_forward(p->C, p->dtype, params, noisy, base_out, out, ws, p->N, p->H, p->W, p->bwd, st);
  const int dt = p->dtype;
  const int hb = cblocks(p->hid);
  View cat = make_view((char*)ws + p->off_cat, dt, p->N, p->H, p->W, 2, 0, 2);
  View hv = make_view((char*)ws + p->off_h, dt, p->N, p->H, p->W, hb, 0, hb);
  PackJob pj[3];
  int nj = 0;
  pj[nj++] = make_fwd_pack(p->L[0], params[0], (char*)ws + p->off_wp[0]);
  pj[nj++] = make_fwd_pack(p->L[1], params[2], (char*)ws + p->off_wp[1]);
  if (p->bwd) pj[nj++] = make_dgrad_pack(p->L[1], params[2], (char*)ws + p->off_wd1, p->L[1].cin_blocks());
  N2N_TRY(launch_pack(pj, nj, dt, st));
  BiasPadJob bj[2] = {{params[1], (float*)((char*)ws + p->off_bias[0]), p->hid, hb * 16},
                      {params[3], (float*)((char*)ws + p->off_bias[1]), p->C, 16}};
  N2N_TRY(launch_bias_pad(bj, 2, st));
  N2N_TRY(launch_nchw_to_c16(noisy, p->C, sub_blocks(cat, dt, 0, 1), dt, st));
  N2N_TRY(launch_nchw_to_c16(base_out, p->C, sub_blocks(cat, dt, 1, 1), dt, st));
  TapGemm g0 = make_conv_fwd(p->L[0], dt, cat, hv, (char*)ws + p->off_wp[0], (const float*)((char*)ws + p->off_bias[0]));
  g0.act = 1; g0.slope = 0.f;                                   // ReLU (adapter.py:16)
  N2N_TRY(launch_tapgemm(g0, st));
  TapGemm g1 = make_conv_fwd(p->L[1], dt, hv, sub_blocks(cat, dt, 0, 1), (char*)ws + p->off_wp[1],
                             (const float*)((char*)ws + p->off_bias[1]));
  g1.has_addend = true; g1.addend = sub_blocks(cat, dt, 1, 1);  // + base_out (adapter.py:26)
  g1.out_nchw = out; g1.out_c = p->C;
  return launch_tapgemm(g1, st);
}

extern "C" int n2n_adapter_backward(n2n_adapter_plan* p, const float* const* params, const float* noisy, const float* base_out,
                                    const float* dout, float* const* grads, void* ws, void* stream) {
  N2N_CHECK_ARG(p && params && noisy && base_out && dout && grads && ws, "adapter_backward: null argument");
  N2N_CHECK_ARG(p->bwd, "adapter_backward: plan was created without with_backward");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->fused) return adapter_fused_backward(p->C, p->dtype, params, noisy, base_out, dout, grads, ws, p->N, p->H, p->W, st);
  const int dt = p->dtype;
  const int hb = cblocks(p->hid);
  View cat = make_view((char*)ws + p->off_cat, dt, p->N, p->H, p->W, 2, 0, 2);
  View hv = make_view((char*)ws + p->off_h, dt, p->N, p->H, p->W, hb, 0, hb);
  View gout = make_view((char*)ws + p->off_gout, dt, p->N, p->H, p->W, 1, 0, 1);
  View gh = make_view((char*)ws + p->off_gh, dt, p->N, p->H, p->W, hb, 0, hb);
  N2N_TRY(launch_nchw_to_c16(dout, p->C, gout, dt, st));
  float* part[2] = {(float*)((char*)ws + p->off_partial[0]), (float*)((char*)ws + p->off_partial[1])};
  float* bpart[2] = {(float*)((char*)ws + p->off_bpartial[0]), (float*)((char*)ws + p->off_bpartial[1])};
  TapWgrad w1 = make_conv_wgrad(p->L[1], dt, hv, gout, part[1], bpart[1], p->splits[1]);
  N2N_TRY(launch_tapwgrad(w1, st));
  TapGemm d1 = make_conv_dgrad(p->L[1], dt, gout, gh, (char*)ws + p->off_wd1, hb);
  d1.has_mask = true; d1.mask = hv; d1.slope = 0.f;             // ReLU'
  N2N_TRY(launch_tapgemm(d1, st));
  TapWgrad w0 = make_conv_wgrad(p->L[0], dt, cat, gh, part[0], bpart[0], p->splits[0]);
  N2N_TRY(launch_tapwgrad(w0, st));
  UnpackJob uj[2] = {make_unpack(p->L[0], part[0], bpart[0], p->splits[0], grads[0], grads[1]),
                     make_unpack(p->L[1], part[1], bpart[1], p->splits[1], grads[2], grads[3])};
  return launch_unpack(uj, 2, st);
}
