// tapgemm_simt.cu — CUDA-core (FFMA, fp32 accumulate) engine for the generic tap GEMM and the
// weight-gradient GEMM.  This is the *precision-reference* engine: it serves dtype == N2N_F32
// (the fp32 parity mode, max-abs error against the oracle) and is deliberately simple.  The
// bf16 production path is the tcgen05 engine (tapgemm_umma.cu / wgrad_umma.cu).
#include "common.cuh"

namespace n2n {

struct TapGemmDev {
  View x[4];
  int ntaps;
  int8_t tap_dy[9], tap_dx[9], tap_view[9], tap_slab[9];
  int cin_blocks, nout;
  const float* w;
  const float* bias;
  View y;
  int has_addend; View addend;
  int has_mask; View mask;
  int act; float slope;
  float* out_nchw; int out_c;
  int tiles_x;      // 32-pixel tiles per row
};

// Block = 128 threads: lane = pixel within a 32-pixel row segment, warp = output sub-group.
// NT output channels per block (grid.y tiles nout); per stage (tap, 16-channel block) the
// A tile [32][16] and the W tile [NT][16] go through shared memory.
template <typename T, int NT>
__global__ void __launch_bounds__(128)
tapgemm_simt_kernel(const __grid_constant__ TapGemmDev g) {
  constexpr int PT = NT / 4;
  __shared__ float As[32][17];
  __shared__ float Ws[NT][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = g.y.H, W = g.y.W;
  long long tile = blockIdx.x;
  const int tx = (int)(tile % g.tiles_x); tile /= g.tiles_x;
  const int y = (int)(tile % H);
  const int img = (int)(tile / H);
  const int x0 = tx * 32;
  const int n0 = blockIdx.y * NT;
  float acc[PT];
#pragma unroll
  for (int j = 0; j < PT; ++j) acc[j] = 0.f;

  const int cpad = g.cin_blocks * 16;
  for (int t = 0; t < g.ntaps; ++t) {
    const View& xv = g.x[g.tap_view[t]];
    const int sy = y + g.tap_dy[t];
    const float* wslab = g.w + (long long)g.tap_slab[t] * g.nout * cpad;
    for (int cb = 0; cb < g.cin_blocks; ++cb) {
      {  // A tile: thread -> pixel tid/4, 4 channels (tid%4)*4
        const int px = tid >> 2, c4 = (tid & 3) * 4;
        const int sx = x0 + px + g.tap_dx[t];
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (sy >= 0 && sy < xv.H && sx >= 0 && sx < xv.W) {
          const T* p = (const T*)xv.ptr + img * xv.sN + cb * xv.sCb + sy * xv.sY + sx * xv.sX + c4;
          v0 = to_f32<T>(p[0]); v1 = to_f32<T>(p[1]); v2 = to_f32<T>(p[2]); v3 = to_f32<T>(p[3]);
        }
        As[px][c4] = v0; As[px][c4 + 1] = v1; As[px][c4 + 2] = v2; As[px][c4 + 3] = v3;
      }
      for (int i = tid; i < NT * 16; i += 128) {
        const int n = i >> 4, c = i & 15;
        Ws[n][c] = wslab[(long long)(n0 + n) * cpad + cb * 16 + c];
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float a = As[lane][c];
#pragma unroll
        for (int j = 0; j < PT; ++j) acc[j] = fmaf(a, Ws[warp * PT + j][c], acc[j]);
      }
      __syncthreads();
    }
  }
  const int x = x0 + lane;
  if (x >= W) return;
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    const int n = n0 + warp * PT + j;
    float v = acc[j];
    if (g.bias) v += g.bias[n];
    const int cb = n >> 4, e = n & 15;
    if (g.has_addend)
      v += to_f32<T>(((const T*)g.addend.ptr)[img * g.addend.sN + cb * g.addend.sCb + y * g.addend.sY + x * g.addend.sX + e]);
    if (g.act) v = v > 0.f ? v : v * g.slope;
    if (g.has_mask) {
      const float m = to_f32<T>(((const T*)g.mask.ptr)[img * g.mask.sN + cb * g.mask.sCb + y * g.mask.sY + x * g.mask.sX + e]);
      v *= (m > 0.f ? 1.f : g.slope);
    }
    if (g.out_nchw) {
      if (n < g.out_c) g.out_nchw[(((long long)img * g.out_c + n) * H + y) * W + x] = v;
    } else {
      ((T*)g.y.ptr)[img * g.y.sN + cb * g.y.sCb + y * g.y.sY + x * g.y.sX + e] = from_f32<T>(v);
    }
  }
}

template <typename T>
static int run_tapgemm(const TapGemm& g, cudaStream_t st) {
  TapGemmDev d;
  memset(&d, 0, sizeof(d));
  N2N_CHECK_ARG(g.ntaps <= 9 && g.mma_n == 0, "tapgemm_simt: launch form not supported by the fp32 engine");
  for (int i = 0; i < 4; ++i) d.x[i] = g.x[i];
  d.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) {
    d.tap_dy[t] = (int8_t)g.tap_dy[t]; d.tap_dx[t] = (int8_t)g.tap_dx[t];
    d.tap_view[t] = (int8_t)g.tap_view[t]; d.tap_slab[t] = (int8_t)g.tap_slab[t];
  }
  d.cin_blocks = g.cin_blocks; d.nout = g.nout; d.w = (const float*)g.w; d.bias = g.bias; d.y = g.y;
  d.has_addend = g.has_addend; d.addend = g.addend; d.has_mask = g.has_mask; d.mask = g.mask;
  d.act = g.act; d.slope = g.slope; d.out_nchw = g.out_nchw; d.out_c = g.out_c;
  d.tiles_x = (g.y.W + 31) / 32;
  const long long tiles = (long long)g.y.N * g.y.H * d.tiles_x;
  N2N_CHECK_ARG(tiles > 0 && tiles < (1LL << 31), "tapgemm_simt: bad tile count");
  if (g.nout % 48 == 0) {
    dim3 grid((unsigned)tiles, g.nout / 48);
    tapgemm_simt_kernel<T, 48><<<grid, 128, 0, st>>>(d);
  } else {
    dim3 grid((unsigned)tiles, g.nout / 16);
    tapgemm_simt_kernel<T, 16><<<grid, 128, 0, st>>>(d);
  }
  N2N_LAUNCH_CHECK();
  return 0;
}

int launch_tapgemm_simt(const TapGemm& g, cudaStream_t st) {
  N2N_CHECK_ARG(g.dtype == N2N_F32, "tapgemm_simt: only the fp32 engine packs weights as plain fp32");
  return run_tapgemm<float>(g, st);
}

// ------------------------------------------------------------------------------------------
// weight gradient:  P[s][t][c][n] = sum_{p in split s} dY[p, n] * X[p + off_t, c]
// Block = 128 threads computes a [16 c] x [NT n] tile for one (split, pair, c-block).
// ------------------------------------------------------------------------------------------
struct TapWgradDev {
  View dy[4]; View x[4];
  int npairs;
  int8_t pair_dyv[9], pair_xv[9], pair_dy[9], pair_dx[9];
  int n_blocks, c_blocks;
  float* partial; float* bias_partial;
  int ndyviews, splits;
  long long pixels, per_split;
};

template <typename T, int NT>
__global__ void __launch_bounds__(128)
tapwgrad_simt_kernel(const __grid_constant__ TapWgradDev g) {
  constexpr int PT = NT / 8;
  __shared__ float Ys[32][NT];
  __shared__ float Xs[32][16];
  const int tid = threadIdx.x;
  int b = blockIdx.x;
  const int ntile = b % (g.n_blocks * 16 / NT); b /= (g.n_blocks * 16 / NT);
  const int cb = b % g.c_blocks; b /= g.c_blocks;
  const int t = b % g.npairs;
  const int s = b / g.npairs;
  const View& dyv = g.dy[g.pair_dyv[t]];
  const View& xv = g.x[g.pair_xv[t]];
  const int H = dyv.H, W = dyv.W;
  const int c = tid & 15, ng = tid >> 4;      // 8 n-groups
  const int n0 = ntile * NT;
  float acc[PT];
#pragma unroll
  for (int j = 0; j < PT; ++j) acc[j] = 0.f;
  float bsum[PT];
#pragma unroll
  for (int j = 0; j < PT; ++j) bsum[j] = 0.f;
  const bool do_bias = g.bias_partial != nullptr && cb == 0 && t < g.ndyviews;

  const long long p_begin = (long long)s * g.per_split;
  long long p_end = p_begin + g.per_split;
  if (p_end > g.pixels) p_end = g.pixels;
  for (long long p0 = p_begin; p0 < p_end; p0 += 32) {
    // stage 32 pixels
    for (int i = tid; i < 32 * (NT + 16); i += 128) {
      const int pp = i / (NT + 16), k = i - pp * (NT + 16);
      const long long p = p0 + pp;
      float v = 0.f;
      if (p < p_end) {
        const int x = (int)(p % W);
        const int yy = (int)((p / W) % H);
        const int img = (int)(p / ((long long)W * H));
        if (k < NT) {
          const int n = n0 + k;
          v = to_f32<T>(((const T*)dyv.ptr)[img * dyv.sN + (n >> 4) * dyv.sCb + yy * dyv.sY + x * dyv.sX + (n & 15)]);
        } else {
          const int sx = x + g.pair_dx[t], sy = yy + g.pair_dy[t];
          if (sx >= 0 && sx < xv.W && sy >= 0 && sy < xv.H)
            v = to_f32<T>(((const T*)xv.ptr)[img * xv.sN + cb * xv.sCb + sy * xv.sY + sx * xv.sX + (k - NT)]);
        }
      }
      if (k < NT) Ys[pp][k] = v; else Xs[pp][k - NT] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < 32; ++pp) {
      const float xval = Xs[pp][c];
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        const float yv = Ys[pp][ng + 8 * j];
        acc[j] = fmaf(xval, yv, acc[j]);
        if (c == 0) bsum[j] += yv;
      }
    }
    __syncthreads();
  }
  const int npad = g.n_blocks * 16, cpad = g.c_blocks * 16;
  float* P = g.partial + (((long long)s * g.npairs + t) * cpad + (cb * 16 + c)) * npad;
#pragma unroll
  for (int j = 0; j < PT; ++j) P[n0 + ng + 8 * j] = acc[j];
  if (do_bias && c == 0) {
    float* B = g.bias_partial + ((long long)s * g.ndyviews + t) * npad;
#pragma unroll
    for (int j = 0; j < PT; ++j) B[n0 + ng + 8 * j] = bsum[j];
  }
}

int wgrad_default_splits(int dtype, long long pixels) {
  // bf16 engine: one split per 128-pixel chunk up to ~2/3 of the SMs (x tap groups fills the chip)
  long long s = dtype == N2N_BF16 ? (pixels + 127) / 128 : (pixels + 2047) / 2048;
  const long long cap = dtype == N2N_BF16 ? 96 : 32;
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return (int)s;
}

int launch_tapwgrad_simt(const TapWgrad& g, cudaStream_t st) {
  N2N_CHECK_ARG(g.dtype == N2N_F32, "tapwgrad_simt: fp32 engine only");
  TapWgradDev d;
  memset(&d, 0, sizeof(d));
  for (int i = 0; i < 4; ++i) { d.dy[i] = g.dy[i]; d.x[i] = g.x[i]; }
  d.npairs = g.npairs;
  for (int t = 0; t < g.npairs; ++t) {
    d.pair_dyv[t] = (int8_t)g.pair_dyv[t]; d.pair_xv[t] = (int8_t)g.pair_xv[t];
    d.pair_dy[t] = (int8_t)g.pair_dy[t]; d.pair_dx[t] = (int8_t)g.pair_dx[t];
  }
  d.n_blocks = g.n_blocks; d.c_blocks = g.c_blocks; d.partial = g.partial; d.bias_partial = g.bias_partial;
  d.ndyviews = g.ndyviews; d.splits = g.splits;
  d.pixels = (long long)g.dy[0].N * g.dy[0].H * g.dy[0].W;
  d.per_split = (d.pixels + g.splits - 1) / g.splits;
  const int npad = g.n_blocks * 16;
  if (npad % 48 == 0) {
    const long long blocks = (long long)g.splits * g.npairs * g.c_blocks * (npad / 48);
    tapwgrad_simt_kernel<float, 48><<<(unsigned)blocks, 128, 0, st>>>(d);
  } else {
    const long long blocks = (long long)g.splits * g.npairs * g.c_blocks * (npad / 16);
    tapwgrad_simt_kernel<float, 16><<<(unsigned)blocks, 128, 0, st>>>(d);
  }
  N2N_LAUNCH_CHECK();
  return 0;
}

}  // namespace n2n
