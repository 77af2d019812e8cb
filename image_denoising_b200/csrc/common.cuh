// common.cuh — shared types for libn2n_b200.so (sm_100a only).
//
// Device activation layout ("C16"): [N][Cb][H][W][16] — channels are blocked by 16 so
// that (a) every channel count of the UNet (48/96/144/in_nc) is a whole number of
// blocks once zero-padded, (b) a (16ch x pixels) box is one contiguous run in HBM
// for TMA, and lands in shared memory as a K-major SWIZZLE_32B UMMA operand, and
// (c) concat is free: producers write into block ranges of the consumer's buffer.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/n2n_b200.h"

namespace n2n {

// ---- packed fp32x2 arithmetic (sm_100: FMUL2 / FADD2 / FFMA2 issue one instruction for two lanes' worth of values; the
// rounding of each half is that of the scalar instruction, so results are bit-identical) for issue-slot-bound epilogues ----
// LeakyReLU / ReLU with 0 <= slope <= 1 on two values: max(a, slope * a).  The two multiplies are ONE packed FMUL2
// (sm_100 f32x2 arithmetic, same rounding as FMUL): the epilogues that use this are issue-slot bound.
__device__ __forceinline__ void lrelu_pair(float& a, float& b, float slope) {
  const float2 m = __fmul2_rn(make_float2(a, b), make_float2(slope, slope));
  a = fmaxf(a, m.x);
  b = fmaxf(b, m.y);
}

// (a, b) += (x, y) as one packed FADD2 (same rounding as two FADDs).
__device__ __forceinline__ void add_pair(float& a, float& b, float x, float y) {
  const float2 r = __fadd2_rn(make_float2(a, b), make_float2(x, y));
  a = r.x;
  b = r.y;
}

// (a, b) *= lrelu'(activation): 1 where the bf16 activation (low / high half of `w`) is > 0, else slope.  One HSETP2 (two
// predicates), one FMUL2 and two selects per pair; the same values as v * (t > 0 ? 1 : slope) (v * 1 is v, FMUL2 rounds as FMUL).
__device__ __forceinline__ void lrelu_mask_pair(float& a, float& b, uint32_t w, float slope) {
  const float2 s = __fmul2_rn(make_float2(a, b), make_float2(slope, slope));
  const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w);
  const __nv_bfloat16 z = __float2bfloat16(0.f);
  const bool p0 = __hgt(__low2bfloat16(t), z), p1 = __hgt(__high2bfloat16(t), z);
  a = p0 ? a : s.x;
  b = p1 ? b : s.y;
}

void set_error(const char* fmt, ...);
bool profiling_active();       // per-launch event timing armed (n2n_profile_begin): keep every launch on one stream
extern thread_local long long g_launch_count;   // kernels launched by this host thread (for gpu_launches)

#define N2N_CHECK_ARG(cond, ...)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      n2n::set_error(__VA_ARGS__);                     \
      return N2N_ERR_ARG;                              \
    }                                                  \
  } while (0)

#define N2N_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      n2n::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                           \
      return N2N_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define N2N_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ++n2n::g_launch_count;                                                          \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      n2n::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                     __FILE__, __LINE__);                                           \
      return N2N_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define N2N_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

constexpr int kSMs = 148;  // B200

// A (possibly strided) window onto a C16 tensor.  Strides are in ELEMENTS; the 16
// channels of one block at one pixel are always contiguous.
struct View {
  void* ptr = nullptr;
  int N = 0, H = 0, W = 0, Cb = 0;
  long long sN = 0, sCb = 0, sY = 0, sX = 0;
};

inline size_t dtype_size(int dtype) { return dtype == N2N_BF16 ? 2 : 4; }
inline int cblocks(int c) { return (c + 15) / 16; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Dense C16 view over a buffer of Cb_total blocks, exposing blocks [cb0, cb0+cb).
inline View make_view(void* base, int dtype, int N, int H, int W, int Cb_total, int cb0, int cb) {
  View v;
  v.N = N; v.H = H; v.W = W; v.Cb = cb;
  v.sX = 16; v.sY = (long long)W * 16; v.sCb = (long long)H * W * 16;
  v.sN = (long long)Cb_total * v.sCb;
  v.ptr = (char*)base + (size_t)cb0 * v.sCb * dtype_size(dtype);
  return v;
}
// Parity sub-view (a,b) of a dense view: pixels (2i+a, 2j+b).
inline View parity_view(const View& v, int dtype, int a, int b) {
  View p = v;
  p.H = v.H / 2; p.W = v.W / 2;
  p.sY = v.sY * 2; p.sX = v.sX * 2;
  p.ptr = (char*)v.ptr + (size_t)(a * v.sY + b * v.sX) * dtype_size(dtype);
  return p;
}
inline View sub_blocks(const View& v, int dtype, int cb0, int cb) {
  View p = v;
  p.Cb = cb;
  p.ptr = (char*)v.ptr + (size_t)cb0 * v.sCb * dtype_size(dtype);
  return p;
}

// ---- 16-element block load/store, fp32 math -------------------------------------------
template <typename T> struct Block16;
template <> struct Block16<float> {
  static __device__ __forceinline__ void load(const float* p, float v[16]) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = q[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  }
  static __device__ __forceinline__ void store(float* p, const float v[16]) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
};
template <> struct Block16<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[16]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 t = q[i];
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
        v[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float v[16]) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      q[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- programmatic dependent launch for the small kernels (see umma.cuh for the rule) ---------
// pdl_enter(): first statement of a kernel launched with launch_pdl_v — wait for the stream
// predecessor to complete, then let our own successor begin launching.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// Producers of data that a successor loads BEFORE its own wait (packed weights, padded biases) must not
// release their dependents early: they only wait, so that the successor starts after they complete.
__device__ __forceinline__ void pdl_enter_no_release() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_v(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static int use_pdl = -1;
  if (use_pdl < 0) { const char* e = getenv("N2N_NO_PDL"); use_pdl = (e && atoi(e)) ? 0 : 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(long long items, int threads, int max_waves = 8) {
  long long blocks = (items + threads - 1) / threads;
  long long cap = (long long)kSMs * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

constexpr int kTapMax = 32, kTapViews = 6;
// ---- generic "tap GEMM" (conv3x3 / conv1x1 / deconv2x2 fwd+dgrad) ------------------------
// D[p, n] = sum_t sum_c X_{view(t)}[p + (dy_t, dx_t), c] * Wp[t][n][c]
// out = epilogue(D): v = D + bias[n] (+ addend[p,n]); act; (* (mask[p,n] > 0 ? 1 : slope)).
struct TapGemm {
  int dtype = N2N_F32;
  View x[kTapViews];
  int ntaps = 0;
  int tap_dy[kTapMax] = {0}, tap_dx[kTapMax] = {0}, tap_view[kTapMax] = {0};
  int tap_slab[kTapMax] = {0};    // which packed-weight slab each tap uses
  int cin_blocks = 0;             // K per tap = 16 * cin_blocks
  int view_blocks[kTapViews] = {0};    // per-view override of cin_blocks (0 = cin_blocks); slab engine only
  int nout = 0;                   // padded to a multiple of 16
  // Column-range form (slab engine only; mma_n > 0): every tap is an N = mma_n GEMM into accumulator columns
  // [tap_col, tap_col + mma_n) of the nout-wide tile accumulator, reading the weight slab that starts tap_woff bytes
  // into `w` (its channel-block groups follow each other, 3 * mma_n * 32 B apart).  Used by the fused
  // ConvTranspose2x2 -> conv3x3 launch (layers.cuh: make_upconv_fwd), where the two column halves are the two
  // horizontally adjacent output pixels and each has its own set of source-pixel taps.
  int mma_n = 0;
  int tap_col[kTapMax] = {0};
  long long tap_woff[kTapMax] = {0};
  // Fused up-conv only: the 3x3 conv zero-pads the UPSAMPLED image, so output pixels on the image border miss the
  // ConvTranspose bias of their out-of-image taps: border_corr[cls][mma_n] (cls = 3*ycls + xcls, 1 = first row/col,
  // 2 = last) is subtracted in the epilogue; up_py = output-row parity of this launch (y is the parity view).
  const float* border_corr = nullptr;
  int up_py = 0;
  const void* w = nullptr;        // packed weights (engine layout, see pack.cu)
  const float* bias = nullptr;    // [nout] fp32 (padded) or null
  View y;                         // output view (C16, dtype); ignored when out_nchw set
  bool has_addend = false; View addend;
  bool has_mask = false; View mask;
  int act = 0;                    // 0: none, 1: leaky (slope)
  float slope = 0.f;
  float* out_nchw = nullptr;      // optional fp32 NCHW [N][out_c][H][W] output
  int out_c = 0;
  // Column split (slab engine only): output columns [n_split*16*b, n_split*16*(b+1)) are the n_split
  // channel blocks of output pixel p + b*split_stride (elements) — ConvTranspose2x2 computes the two
  // horizontally adjacent output pixels of one input pixel as ONE N = 2*Cout GEMM.
  int n_split = 0; long long split_stride = 0;
  bool store_y = true;            // false: the un-pooled output is not needed (no-grad pass), only `pool`
  bool has_pool = false; View pool;   // also write maxpool2x2(out) here (fused in the epilogue when possible)
};

// ---- fused 1x1 head: y = Wc * lrelu(Wb * lrelu(Wa * x + ba) + bb) + bc  (arch_unet.py:257-259) -----
struct HeadChain {
  View x;                          // input activations (dec_conv1b output), in_blocks blocks
  int in_blocks = 0, mid_blocks = 0, mid_channels = 0, out_nc = 0;
  const void* wa = nullptr;        // packed bf16 1x1 weights (engine layout) of nin_a / nin_b
  const void* wb = nullptr;
  const float* bias_a = nullptr;   // padded fp32
  const float* bias_b = nullptr;
  const float* wc = nullptr;       // nin_c weight, torch layout fp32 [out_nc][mid_channels]
  const float* bias_c = nullptr;   // [out_nc]
  float slope = 0.2f;
  bool has_save = false;           // training pass: also store the two intermediate activations
  View save_a, save_b;
  float* out_nchw = nullptr;       // fp32 [N][out_nc][H][W]
};
int launch_head_chain(const HeadChain& h, cudaStream_t st);        // api.cu (profiling scope + dispatch)
int launch_head_chain_umma(const HeadChain& h, cudaStream_t st);   // head_umma.cu

// ---- fused backward of the 1x1 head (headbwd_umma.cu) -----------------------------------------------
struct HeadBwd {
  int blocks = 0, channels = 0, out_nc = 0;
  float slope = 0.2f;
  const float* gout = nullptr;       // dL/dout, fp32 NCHW [N][out_nc][H][W]
  const float* wc = nullptr;         // nin_c weight, torch layout fp32 [out_nc][channels]
  const void* wb_dgrad = nullptr;    // nin_b / nin_a weights in the input-gradient pack (rows = cin, K = cout)
  const void* wa_dgrad = nullptr;
  View act_nb, act_na, act_d1b;      // activated outputs of nin_b, nin_a, dec_conv1b (sign -> lrelu'; wgrad operands)
  View g_d1b;                        // gradient w.r.t. dec_conv1b's output (written; the other two stay on chip)
  int splits = 0;                    // = head_bwd_splits(...): CTAs of the kernel = rows of the partials below
  float* partial_b = nullptr;        // nin_b weight-gradient partial [splits][channels(c)][channels(n)]
  float* bpartial_b = nullptr;       // nin_b bias-gradient partial   [splits][channels]
  float* partial_a = nullptr;        // same for nin_a
  float* bpartial_a = nullptr;
};
int head_bwd_splits(int dtype, int blocks, int out_nc, int n, int h, int w);   // 0 = geometry not covered
int launch_head_bwd(const HeadBwd& h, cudaStream_t st);        // api.cu (profiling scope + dispatch)
int launch_head_bwd_umma(const HeadBwd& h, cudaStream_t st);   // headbwd_umma.cu

// ---- generic weight-gradient GEMM ---------------------------------------------------------
// P[s][t][c][n] = sum_{p in split s} dY_{a(t)}[p, n] * X_{b(t)}[p + (dy_t, dx_t), c]
// bias_partial[s][n] = sum_{p in split s} sum_{distinct dY views} dY[p, n]
struct TapWgrad {
  int dtype = N2N_F32;
  View dy[4]; View x[4];
  int npairs = 0;
  int pair_dyv[9] = {0}, pair_xv[9] = {0}, pair_dy[9] = {0}, pair_dx[9] = {0};
  int n_blocks = 0;               // Cout / 16 (from dY)
  int c_blocks = 0;               // Cin / 16 (from X)
  float* partial = nullptr;       // [splits][npairs][c_blocks*16][n_blocks*16]
  float* bias_partial = nullptr;  // [splits][ndyviews][n_blocks*16] or null
  int ndyviews = 1;               // how many distinct dY views feed the bias sum
  int splits = 1;
};

// Weight pack (torch layout fp32 -> engine layout) and the inverse gradient reduce.
struct Segs {
  int n = 1;
  int src0[2] = {0, 0}, cnt[2] = {0, 0}, dst0[2] = {0, 0};
  long long off[2] = {0, 0};      // extra source-element offset of the segment (pack only)
};
struct PackJob {
  // im2col mode (im2col_nc > 0): the GEMM's contraction index k of the single tap decodes as
  // (3x3 tap k / nc, input channel im2col_c0 + k % nc) of a conv3x3 weight — the layout of the
  // network input's 9-tap im2col block.
  int im2col_nc = 0, im2col_c0 = 0;
  const float* src = nullptr;     // torch-layout fp32 weights
  void* dst = nullptr;            // engine layout
  int ntaps = 1;
  int nout_pad = 16, cin_blocks = 1;
  long long s_t = 0, s_n = 0, s_c = 0;   // source element strides (tap, out-channel, in-channel)
  Segs nseg, cseg;                // channel remaps (concat skip starts on a block boundary)
};
struct UnpackJob {
  int im2col_nc = 0, im2col_c0 = 0;   // see PackJob
  const float* partial = nullptr; // [splits][ntaps][cpad][npad]
  const float* bias_partial = nullptr;  // [splits][npad]
  float* dst_w = nullptr;         // torch layout grad
  float* dst_b = nullptr;         // [nreal] or null
  int splits = 1, ntaps = 1, npad = 16, cpad = 16;
  int bias_rows = 1;              // rows of bias_partial to sum (= splits * ndyviews)
  long long s_t = 0, s_n = 0, s_c = 0;
  Segs nseg, cseg;
};

// Fused ConvTranspose2x2 -> conv3x3 weights (pack.cu: upfuse_pack_kernel), bf16 engine only.
struct UpFuseJob {
  const float* w3 = nullptr;      // dec_conv a weight  [Co][Cu + Cs][3][3]
  const float* b3 = nullptr;      // dec_conv a bias    [Co]
  const float* wd = nullptr;      // ConvTranspose2d weight [Ci][Cu][2][2]
  const float* bd = nullptr;      // ConvTranspose2d bias   [Cu]
  int Ci = 0, Cu = 0, Cs = 0, Co = 0;
  int ngroups = 1, co_pad = 16;   // channel-block groups of Ci (3 blocks each); Co padded to 16
  void* dst[2] = {nullptr, nullptr};   // per output-row parity: 8 composite slabs [(px, sy, sx)][group][3][co_pad][32 B]
  float* bias_full = nullptr;     // [co_pad]
  float* corr = nullptr;          // [9][co_pad]
  // training plans: the same 16 composites TRANSPOSED for the fused input gradient (rows = ci, K = co):
  // [(py, px, sy, sx)][co group][3][ci_rows][32 B]
  void* dst_t = nullptr;
  int gco = 1, ci_rows = 16;
};
int launch_upfuse_pack(const UpFuseJob* jobs, int njobs, cudaStream_t st);

// Backward of the fused up-conv (pack.cu: upfuse_grad_kernel): chain rule from the composite weight gradients
// dWc[(py,px,sy,sx)][ci][co] (dense fp32, already reduced over the pixel splits) to the two layers' own gradients.
struct UpFuseGradJob {
  const float* dwc = nullptr;     // [16][ci_pad][co_pad]
  const float* w3 = nullptr;      // [Co][Cu + Cs][3][3]
  const float* wd = nullptr;      // [Ci][Cu][2][2]
  const float* bd = nullptr;      // [Cu]
  const float* border = nullptr;  // [8][co_pad]: sums of dL/dy over the first / last row, first / last column, four corners
  const float* db3 = nullptr;     // [Co] = sum of dL/dy over all pixels (the conv's bias gradient, already reduced)
  float* dw3 = nullptr;           // [Co][Cu + Cs][3][3]: channels [0, Cu) are written
  float* dwd = nullptr;           // [Ci][Cu][2][2]
  float* dbd = nullptr;           // [Cu]
  int Ci = 0, Cu = 0, Cs = 0, Co = 0, ci_pad = 16, co_pad = 16;
};
int launch_upfuse_grad(const UpFuseGradJob* jobs, int njobs, cudaStream_t st);
// border[8][co_pad] of a C16 bf16 gradient tensor (see UpFuseGradJob::border); deterministic block reductions
int launch_border_sums(const View* g, float* const* border, const int* co_pad, int n, cudaStream_t st);

size_t packed_weight_bytes(int dtype, int ntaps, int nout_pad, int cin_blocks);

// engine entry points (host, async on stream)
int launch_tapgemm(const TapGemm& g, cudaStream_t st);
int launch_tapwgrad(const TapWgrad& g, cudaStream_t st);
int wgrad_default_splits(int dtype, long long pixels);
int launch_pack(const PackJob* jobs, int njobs, int dtype, cudaStream_t st);
int launch_unpack(const UnpackJob* jobs, int njobs, cudaStream_t st);

// elementwise (elementwise.cu)
int launch_nchw_to_c16(const float* src, int C, const View& dst, int dtype, cudaStream_t st);
// dst[p, (ky*3+kx)*C + c] = src[c, y+ky-1, x+kx-1] (zero outside the image): the 3x3 im2col of a
// C-channel NCHW image as ceil(9C/16) C16 blocks, so that a 3x3 conv over it is ONE K = 16 (32) GEMM step
int launch_nchw_to_im2col9(const float* src, int C, const View& dst, int dtype, cudaStream_t st);
int launch_input_stage(const float* x, int C, const float* w, const float* bias, int Cout, const View& col,
                       const View& e0, float slope, cudaStream_t st);
int launch_c16_to_nchw(const View& src, int dtype, float* dst, int C, cudaStream_t st);
int launch_maxpool(const View& src, const View& dst, int dtype, cudaStream_t st);
int launch_unpool_lrelu(const View& act, const View& gpool, const View& gact, float slope, int dtype, cudaStream_t st);
int launch_fill_zero(void* p, size_t bytes, cudaStream_t st);
int launch_add_inplace(float* y, const float* x, long long n, cudaStream_t st);
int launch_crop_patches(const float* const* imgs, const int* dims, const int* sel, int batch, int C, int ps, float scale,
                        float* out, cudaStream_t st);

// bf16 tensor-core engine (tapgemm_umma.cu / wgrad_umma.cu)
int launch_tapgemm_umma(const TapGemm& g, cudaStream_t st);
int launch_tapwgrad_umma(const TapWgrad& g, cudaStream_t st);
// 8x16-tile slab engine (slabgemm_umma.cu): 0 = launched, kSgNotEligible = use the row-slab engine
constexpr int kSgNotEligible = 1;
int launch_slabgemm_umma(const TapGemm& g, cudaStream_t st);
int launch_wgrad_slab_umma(const TapWgrad& g, cudaStream_t st);
// fp32 CUDA-core engine (tapgemm_simt.cu)
int launch_tapgemm_simt(const TapGemm& g, cudaStream_t st);
int launch_tapwgrad_simt(const TapWgrad& g, cudaStream_t st);

}  // namespace n2n
