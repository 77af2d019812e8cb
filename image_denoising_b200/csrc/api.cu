// api.cu — error plumbing, engine dispatch and the single-layer C-ABI entry points
// (NCHW fp32 at the boundary; the blocked C16 layout lives in the caller's workspace).
#include <stdarg.h>

#include <vector>

#include "common.cuh"
#include "layers.cuh"

namespace n2n {

static thread_local char g_err[512] = "";
thread_local long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional per-launch CUDA-event timing of the two GEMM kernel classes (bench.py roofline) ----
struct ProfRec { cudaEvent_t a, b; int cls; double flops; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec>* g_prof = nullptr;

bool profiling_active() { return g_prof_on; }

struct ProfScope {
  ProfRec r; bool on; cudaStream_t st;
  ProfScope(int cls, double flops, cudaStream_t s) : on(g_prof_on), st(s) {
    if (!on) return;
    r.cls = cls; r.flops = flops;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, st);
    g_prof->push_back(r);
  }
};

int launch_tapgemm(const TapGemm& g, cudaStream_t st) {
  N2N_CHECK_ARG(g.ntaps >= 1 && g.ntaps <= kTapMax && g.cin_blocks >= 1 && g.nout >= 16 && g.nout % 16 == 0,
                "tapgemm: bad geometry (taps=%d cin_blocks=%d nout=%d)", g.ntaps, g.cin_blocks, g.nout);
  // executed (padded) FLOPs of this launch: 2 * pixels * nout * taps * 16*cin_blocks
  double kblocks = 0;
  for (int t = 0; t < g.ntaps; ++t) kblocks += g.view_blocks[g.tap_view[t]] ? g.view_blocks[g.tap_view[t]] : g.cin_blocks;
  const double flops = 2.0 * g.y.N * g.y.H * g.y.W * (double)(g.mma_n ? g.mma_n : g.nout) * 16.0 * kblocks;
  ProfScope ps(0, flops, st);
  if (g.dtype == N2N_BF16) {
    const int r = launch_slabgemm_umma(g, st);
    if (r != kSgNotEligible) return r;
    N2N_CHECK_ARG(g.n_split == 0 && g.view_blocks[1] == 0 && g.mma_n == 0 && g.ntaps <= 12,
                  "tapgemm: this launch form needs the slab engine (geometry not eligible)");
    N2N_TRY(launch_tapgemm_umma(g, st));
  } else {
    N2N_TRY(launch_tapgemm_simt(g, st));
  }
  if (g.has_pool) return launch_maxpool(g.y, g.pool, g.dtype, st);
  return 0;
}
int launch_head_chain(const HeadChain& h, cudaStream_t st) {
  const double px = (double)h.x.N * h.x.H * h.x.W;
  const double flops = 2.0 * px * 16.0 * h.mid_blocks * (16.0 * h.in_blocks + 16.0 * h.mid_blocks + h.out_nc);
  ProfScope ps(0, flops, st);
  return launch_head_chain_umma(h, st);
}
int launch_head_bwd(const HeadBwd& h, cudaStream_t st) {
  const double px = (double)h.g_d1b.N * h.g_d1b.H * h.g_d1b.W;
  const double flops = 2.0 * px * h.channels * (2.0 * h.channels + h.out_nc);
  ProfScope ps(0, flops, st);
  return launch_head_bwd_umma(h, st);
}
int launch_tapwgrad(const TapWgrad& g, cudaStream_t st) {
  const double flops = 2.0 * g.dy[0].N * g.dy[0].H * g.dy[0].W * 256.0 * g.n_blocks * g.c_blocks * g.npairs;
  ProfScope ps(1, flops, st);
  if (g.dtype == N2N_BF16) {
    const int r = launch_wgrad_slab_umma(g, st);
    if (r != kSgNotEligible) return r;
    return launch_tapwgrad_umma(g, st);
  }
  return launch_tapwgrad_simt(g, st);
}

// carve a workspace
struct Carver {
  char* base; size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  void* take(size_t bytes) { void* p = base ? base + off : nullptr; off += align_up(bytes, 1024); return p; }
};

static size_t c16_bytes(int dtype, int n, int c, int h, int w) {
  return (size_t)n * cblocks(c) * h * w * 16 * dtype_size(dtype);
}

}  // namespace n2n

using namespace n2n;

extern "C" const char* n2n_last_error(void) { return g_err; }
extern "C" int n2n_version(void) { return 100; }
extern "C" long long n2n_launch_count(void) { return g_launch_count; }

extern "C" int n2n_profile_active(void) { return g_prof_on ? 1 : 0; }

extern "C" int n2n_profile_begin(void) {
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  for (auto& r : *g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof->clear();
  g_prof_on = true;
  return 0;
}
// out[3*cls + {0,1,2}] = {summed milliseconds, summed executed FLOPs, launches} for cls 0 (tap GEMM:
// conv / deconv forward + input gradient) and cls 1 (weight-gradient GEMM incl. its bias reduction).
extern "C" int n2n_profile_end(double* out) {
  N2N_CHECK_ARG(out != nullptr, "profile_end: out is NULL");
  g_prof_on = false;
  for (int i = 0; i < 6; ++i) out[i] = 0.0;
  if (!g_prof) return 0;
  for (auto& r : *g_prof) {
    N2N_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    N2N_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    out[3 * r.cls] += ms; out[3 * r.cls + 1] += r.flops; out[3 * r.cls + 2] += 1.0;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof->clear();
  return 0;
}
// Per-launch form of n2n_profile_end: rows of {class, ms, executed FLOPs} in launch order.
extern "C" int n2n_profile_end_list(double* out, int max_rows) {
  N2N_CHECK_ARG(out != nullptr && max_rows >= 0, "profile_end_list: bad arguments");
  g_prof_on = false;
  int n = 0;
  if (!g_prof) return 0;
  for (auto& r : *g_prof) {
    N2N_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    N2N_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (n < max_rows) { out[3 * n] = r.cls; out[3 * n + 1] = ms; out[3 * n + 2] = r.flops; ++n; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof->clear();
  return n;
}
extern "C" int n2n_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
  cudaDeviceProp p;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// conv2d (k in {1,3})
// ------------------------------------------------------------------------------------------
static LayerGeom conv_geom(int cin, int cout, int k) {
  LayerGeom L; L.kind = k == 3 ? L_CONV3 : L_CONV1; L.cin = chan1(cin); L.cout = cout; return L;
}

struct ConvWs {
  void *xa, *ya, *wp, *wd; float *bias, *partial, *bpartial; size_t total; int splits;
};
static ConvWs conv_ws(void* ws, const LayerGeom& L, int n, int h, int w, int dtype, bool deconv) {
  Carver c(ws);
  ConvWs r;
  const int oh = deconv ? 2 * h : h, ow = deconv ? 2 * w : w;
  r.splits = layer_wgrad_splits(L, dtype, n, h, w);
  r.xa = c.take(c16_bytes(dtype, n, L.cin.real(), h, w));
  r.ya = c.take(c16_bytes(dtype, n, L.cout, oh, ow));
  r.wp = c.take(L.fwd_pack_bytes(dtype));
  r.wd = c.take(L.dgrad_pack_bytes(dtype, L.cin_blocks()));
  r.bias = (float*)c.take(L.cout_blocks() * 16 * sizeof(float));
  r.partial = (float*)c.take(L.partial_bytes(r.splits));
  r.bpartial = (float*)c.take(L.bias_partial_bytes(r.splits));
  r.total = c.off;
  return r;
}

extern "C" size_t n2n_conv2d_workspace_bytes(int n, int cin, int cout, int h, int w, int ksize, int dtype) {
  return conv_ws(nullptr, conv_geom(cin, cout, ksize), n, h, w, dtype, false).total;
}

#define CONV_ARGCHECK(name)                                                                                   \
  N2N_CHECK_ARG(n > 0 && cin > 0 && cout > 0 && h > 0 && w_ > 0 && (ksize == 1 || ksize == 3),                \
                name ": bad shape n=%d cin=%d cout=%d h=%d w=%d k=%d", n, cin, cout, h, w_, ksize);           \
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, name ": bad dtype %d", dtype);                         \
  N2N_CHECK_ARG(workspace != nullptr, name ": workspace is NULL")

extern "C" int n2n_conv2d_fwd(const float* x, const float* w, const float* b, float* y, int n, int cin, int cout,
                              int h, int w_, int ksize, float act_slope, int dtype, void* workspace, void* stream) {
  CONV_ARGCHECK("conv2d_fwd");
  N2N_CHECK_ARG(x && w && y, "conv2d_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerGeom L = conv_geom(cin, cout, ksize);
  ConvWs ws = conv_ws(workspace, L, n, h, w_, dtype, false);
  View xv = make_view(ws.xa, dtype, n, h, w_, L.cin_blocks(), 0, L.cin_blocks());
  View yv = make_view(ws.ya, dtype, n, h, w_, L.cout_blocks(), 0, L.cout_blocks());
  N2N_TRY(launch_nchw_to_c16(x, cin, xv, dtype, st));
  PackJob pj = make_fwd_pack(L, w, ws.wp);
  N2N_TRY(launch_pack(&pj, 1, dtype, st));
  BiasPadJob bj{b, ws.bias, cout, L.cout_blocks() * 16};
  N2N_TRY(launch_bias_pad(&bj, 1, st));
  TapGemm g = make_conv_fwd(L, dtype, xv, yv, ws.wp, ws.bias);
  if (act_slope >= 0.f) { g.act = 1; g.slope = act_slope; }
  N2N_TRY(launch_tapgemm(g, st));
  return launch_c16_to_nchw(yv, dtype, y, cout, st);
}

extern "C" int n2n_conv2d_dgrad(const float* dy, const float* w, float* dx, int n, int cin, int cout, int h, int w_,
                                int ksize, int dtype, void* workspace, void* stream) {
  CONV_ARGCHECK("conv2d_dgrad");
  N2N_CHECK_ARG(dy && w && dx, "conv2d_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerGeom L = conv_geom(cin, cout, ksize);
  ConvWs ws = conv_ws(workspace, L, n, h, w_, dtype, false);
  View xv = make_view(ws.xa, dtype, n, h, w_, L.cin_blocks(), 0, L.cin_blocks());
  View yv = make_view(ws.ya, dtype, n, h, w_, L.cout_blocks(), 0, L.cout_blocks());
  N2N_TRY(launch_nchw_to_c16(dy, cout, yv, dtype, st));
  PackJob pj = make_dgrad_pack(L, w, ws.wd, L.cin_blocks());
  N2N_TRY(launch_pack(&pj, 1, dtype, st));
  TapGemm g = make_conv_dgrad(L, dtype, yv, xv, ws.wd, L.cin_blocks());
  N2N_TRY(launch_tapgemm(g, st));
  return launch_c16_to_nchw(xv, dtype, dx, cin, st);
}

extern "C" int n2n_conv2d_wgrad(const float* x, const float* dy, float* dw, float* db, int n, int cin, int cout,
                                int h, int w_, int ksize, int dtype, void* workspace, void* stream) {
  CONV_ARGCHECK("conv2d_wgrad");
  N2N_CHECK_ARG(x && dy && dw, "conv2d_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerGeom L = conv_geom(cin, cout, ksize);
  ConvWs ws = conv_ws(workspace, L, n, h, w_, dtype, false);
  View xv = make_view(ws.xa, dtype, n, h, w_, L.cin_blocks(), 0, L.cin_blocks());
  View yv = make_view(ws.ya, dtype, n, h, w_, L.cout_blocks(), 0, L.cout_blocks());
  N2N_TRY(launch_nchw_to_c16(x, cin, xv, dtype, st));
  N2N_TRY(launch_nchw_to_c16(dy, cout, yv, dtype, st));
  TapWgrad g = make_conv_wgrad(L, dtype, xv, yv, ws.partial, ws.bpartial, ws.splits);
  N2N_TRY(launch_tapwgrad(g, st));
  UnpackJob uj = make_unpack(L, ws.partial, ws.bpartial, ws.splits, dw, db);
  return launch_unpack(&uj, 1, st);
}

// ------------------------------------------------------------------------------------------
// ConvTranspose2d(k=2, s=2)   x: [n,cin,h,w]  ->  y: [n,cout,2h,2w]
// ------------------------------------------------------------------------------------------
static LayerGeom deconv_geom(int cin, int cout) {
  LayerGeom L; L.kind = L_DECONV; L.cin = chan1(cin); L.cout = cout; return L;
}
extern "C" size_t n2n_deconv2x2_workspace_bytes(int n, int cin, int cout, int h, int w, int dtype) {
  return conv_ws(nullptr, deconv_geom(cin, cout), n, h, w, dtype, true).total;
}
#define DECONV_ARGCHECK(name)                                                                          \
  N2N_CHECK_ARG(n > 0 && cin > 0 && cout > 0 && h > 0 && w_ > 0, name ": bad shape");                  \
  N2N_CHECK_ARG(dtype == N2N_F32 || dtype == N2N_BF16, name ": bad dtype %d", dtype);                  \
  N2N_CHECK_ARG(workspace != nullptr, name ": workspace is NULL")

extern "C" int n2n_deconv2x2_fwd(const float* x, const float* w, const float* b, float* y, int n, int cin, int cout,
                                 int h, int w_, int dtype, void* workspace, void* stream) {
  DECONV_ARGCHECK("deconv2x2_fwd");
  N2N_CHECK_ARG(x && w && y, "deconv2x2_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerGeom L = deconv_geom(cin, cout);
  ConvWs ws = conv_ws(workspace, L, n, h, w_, dtype, true);
  View xv = make_view(ws.xa, dtype, n, h, w_, L.cin_blocks(), 0, L.cin_blocks());
  View yv = make_view(ws.ya, dtype, n, 2 * h, 2 * w_, L.cout_blocks(), 0, L.cout_blocks());
  N2N_TRY(launch_nchw_to_c16(x, cin, xv, dtype, st));
  BiasPadJob bj{b, ws.bias, cout, L.cout_blocks() * 16};
  N2N_TRY(launch_bias_pad(&bj, 1, st));
  if (slab_deconv_pair_ok(dtype, h, w_, L.cin_blocks(), L.cout_blocks())) {
    // the launch form the network plans use: one N = 2*Cout GEMM per output-row parity
    PackJob pj = make_deconv_pair_pack(L, w, ws.wp);
    N2N_TRY(launch_pack(&pj, 1, dtype, st));
    for (int a = 0; a < 2; ++a) N2N_TRY(launch_tapgemm(make_deconv_fwd_pair(L, dtype, xv, yv, a, ws.wp, ws.bias), st));
    return launch_c16_to_nchw(yv, dtype, y, cout, st);
  }
  PackJob pj = make_fwd_pack(L, w, ws.wp);
  N2N_TRY(launch_pack(&pj, 1, dtype, st));
  for (int ab = 0; ab < 4; ++ab) {
    TapGemm g = make_deconv_fwd(L, dtype, xv, yv, ab / 2, ab % 2, ws.wp, ws.bias);
    N2N_TRY(launch_tapgemm(g, st));
  }
  return launch_c16_to_nchw(yv, dtype, y, cout, st);
}

extern "C" int n2n_deconv2x2_dgrad(const float* dy, const float* w, float* dx, int n, int cin, int cout, int h,
                                   int w_, int dtype, void* workspace, void* stream) {
  DECONV_ARGCHECK("deconv2x2_dgrad");
  N2N_CHECK_ARG(dy && w && dx, "deconv2x2_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerGeom L = deconv_geom(cin, cout);
  ConvWs ws = conv_ws(workspace, L, n, h, w_, dtype, true);
  View xv = make_view(ws.xa, dtype, n, h, w_, L.cin_blocks(), 0, L.cin_blocks());
  View yv = make_view(ws.ya, dtype, n, 2 * h, 2 * w_, L.cout_blocks(), 0, L.cout_blocks());
  N2N_TRY(launch_nchw_to_c16(dy, cout, yv, dtype, st));
  PackJob pj = make_dgrad_pack(L, w, ws.wd, L.cin_blocks());
  N2N_TRY(launch_pack(&pj, 1, dtype, st));
  TapGemm g = make_deconv_dgrad(L, dtype, yv, xv, ws.wd);
  N2N_TRY(launch_tapgemm(g, st));
  return launch_c16_to_nchw(xv, dtype, dx, cin, st);
}

extern "C" int n2n_deconv2x2_wgrad(const float* x, const float* dy, float* dw, float* db, int n, int cin, int cout,
                                   int h, int w_, int dtype, void* workspace, void* stream) {
  DECONV_ARGCHECK("deconv2x2_wgrad");
  N2N_CHECK_ARG(x && dy && dw, "deconv2x2_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerGeom L = deconv_geom(cin, cout);
  ConvWs ws = conv_ws(workspace, L, n, h, w_, dtype, true);
  View xv = make_view(ws.xa, dtype, n, h, w_, L.cin_blocks(), 0, L.cin_blocks());
  View yv = make_view(ws.ya, dtype, n, 2 * h, 2 * w_, L.cout_blocks(), 0, L.cout_blocks());
  N2N_TRY(launch_nchw_to_c16(x, cin, xv, dtype, st));
  N2N_TRY(launch_nchw_to_c16(dy, cout, yv, dtype, st));
  TapWgrad g = make_deconv_wgrad(L, dtype, xv, yv, ws.partial, ws.bpartial, ws.splits);
  N2N_TRY(launch_tapwgrad(g, st));
  UnpackJob uj = make_unpack(L, ws.partial, ws.bpartial, ws.splits, dw, db);
  return launch_unpack(&uj, 1, st);
}

// ------------------------------------------------------------------------------------------
// MaxPool2d(2)
// ------------------------------------------------------------------------------------------
extern "C" size_t n2n_pool_workspace_bytes(int n, int c, int h, int w, int dtype) {
  return 2 * align_up(c16_bytes(dtype, n, c, h, w), 1024) + 2 * align_up(c16_bytes(dtype, n, c, h / 2, w / 2), 1024);
}

extern "C" int n2n_maxpool2_fwd(const float* x, float* y, int n, int c, int h, int w, int dtype, void* workspace,
                                void* stream) {
  N2N_CHECK_ARG(x && y && workspace && n > 0 && c > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0,
                "maxpool2_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace);
  void* xa = cv.take(c16_bytes(dtype, n, c, h, w));
  cv.take(c16_bytes(dtype, n, c, h, w));
  void* ya = cv.take(c16_bytes(dtype, n, c, h / 2, w / 2));
  View xv = make_view(xa, dtype, n, h, w, cblocks(c), 0, cblocks(c));
  View yv = make_view(ya, dtype, n, h / 2, w / 2, cblocks(c), 0, cblocks(c));
  N2N_TRY(launch_nchw_to_c16(x, c, xv, dtype, st));
  N2N_TRY(launch_maxpool(xv, yv, dtype, st));
  return launch_c16_to_nchw(yv, dtype, y, c, st);
}

extern "C" int n2n_maxpool2_bwd(const float* x, const float* dy, float* dx, int n, int c, int h, int w, float slope,
                                int dtype, void* workspace, void* stream) {
  N2N_CHECK_ARG(x && dy && dx && workspace && n > 0 && c > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0,
                "maxpool2_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace);
  void* xa = cv.take(c16_bytes(dtype, n, c, h, w));
  void* ga = cv.take(c16_bytes(dtype, n, c, h, w));
  void* ya = cv.take(c16_bytes(dtype, n, c, h / 2, w / 2));
  View xv = make_view(xa, dtype, n, h, w, cblocks(c), 0, cblocks(c));
  View gv = make_view(ga, dtype, n, h, w, cblocks(c), 0, cblocks(c));
  View yv = make_view(ya, dtype, n, h / 2, w / 2, cblocks(c), 0, cblocks(c));
  N2N_TRY(launch_nchw_to_c16(x, c, xv, dtype, st));
  N2N_TRY(launch_nchw_to_c16(dy, c, yv, dtype, st));
  N2N_TRY(launch_unpool_lrelu(xv, yv, gv, slope, dtype, st));
  return launch_c16_to_nchw(gv, dtype, dx, c, st);
}

// ------------------------------------------------------------------------------------------
// Device-side data path (train.py:208-228, finetune.py:94-150)
// ------------------------------------------------------------------------------------------
extern "C" int n2n_crop_patches(const float* const* images, const int32_t* dims_hw, const int32_t* sel, int batch,
                                int channels, int patch, float scale, float* out, void* stream) {
  N2N_CHECK_ARG(batch >= 0 && channels >= 1 && patch >= 1, "crop_patches: bad geometry");
  if (batch == 0) return 0;
  N2N_CHECK_ARG(images && dims_hw && sel && out, "crop_patches: null pointer");
  return launch_crop_patches(images, dims_hw, sel, batch, channels, patch, scale, out, (cudaStream_t)stream);
}
