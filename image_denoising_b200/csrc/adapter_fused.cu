// adapter_fused.cu — the small-channel kernels of adapter.OutputAdapter (reference adapter.py:5-26, finetune.py:277-288;
// SURVEY.md §2.1 K7): out = base_out + conv3x3(relu(conv3x3(cat[noisy, base_out]; 2C -> 16)); 16 -> C) and its backward,
// for C in {1, 3}, hidden = 16, straight from / to the fp32 NCHW tensors of the reference.
//
// Why not the tensor-core engines: with 6 and 3 real channels the tap GEMMs are all padding (16-channel blocks) and the
// launches are bound by per-tile overhead and layout conversions, not by math (measured at 32 x 3 x 256 x 256: forward
// 328 us, backward 601 us for 2 x 1296 + 2 x 1728 MAC per pixel).  Here the same arithmetic runs as direct convolutions on
// the CUDA cores in fp32 (exact in both precision modes): every thread owns 4 adjacent pixels x all output channels, the
// halo'd input tile and the layer's weights sit in shared memory (one broadcast LDS.128 = four output channels of a tap);
// `adapter_prep_kernel` re-lays the four parameter tensors (and the flipped / transposed copy of conv2's weight for the
// input gradient) once per forward.  The weight gradients
// accumulate per warp in registers (one warp = one block of weights, lanes = pixels, 3 x 3 window slid down the rows),
// are reduced over the lanes by shuffles and over the persistent CTAs in a fixed order (deterministic).
#include "common.cuh"

namespace n2n {

constexpr int kAdHid = 16;
constexpr int kAdTH = 8, kAdTW = 128, kAdPX = 4;
constexpr int kAdTWP = kAdTW + 8;      // smem row: image column x0 - 4 + s at index s; s in [3, 132] is used

// Weights as the kernels read them (built per call by adapter_prep_kernel in the caller's workspace): output channel fastest,
// so that a thread fetches the weights of four output channels of one (input channel, tap) with one broadcast LDS.128.
// (A first version read them from __constant__ memory: the 3.4 KB cycled through the immediate-constant cache once per CTA
// and the LDCU misses held the kernel at 20 % of the FP32 rate.)
struct AdapterW {
  float w1[6 * 9 * kAdHid];     // [(ci*9 + k)][co], ci < 2C
  float b1[kAdHid];
  float w2[kAdHid * 9 * 4];     // [(ci*9 + k)][co padded to 4], co < C
  float b2[4];
  float w2t[3 * 9 * kAdHid];    // [(c*9 + k)][hid] = w2[c][hid][8 - k]: taps of the input gradient (correlation with the flipped kernel)
};

__global__ void adapter_prep_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                    const float* __restrict__ b2, int C, AdapterW* __restrict__ out) {
  pdl_enter_no_release();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kAdHid * 2 * C * 9) {                    // w1[co][ci][k]
    const int co = t / (2 * C * 9), r = t % (2 * C * 9);
    out->w1[r * kAdHid + co] = w1[t];
  }
  if (t < kAdHid) out->b1[t] = b1[t];
  if (t < kAdHid * 9 * 4) out->w2[t] = 0.f;          // padded columns stay zero; adapter_prep_w2_kernel scatters the real ones
  if (t < 4) out->b2[t] = t < C ? b2[t] : 0.f;
  if (t < C * kAdHid * 9) {                        // w2[co][ci][k]
    const int co = t / (kAdHid * 9), ci = (t / 9) % kAdHid, k = t % 9;
    out->w2t[(co * 9 + (8 - k)) * kAdHid + ci] = w2[t];
  }
}
__global__ void adapter_prep_w2_kernel(const float* __restrict__ w2, int C, AdapterW* __restrict__ out) {
  pdl_enter_no_release();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < C * kAdHid * 9) {
    const int co = t / (kAdHid * 9), r = t % (kAdHid * 9);
    out->w2[r * 4 + co] = w2[t];
  }
}

// MODE 0: h = relu(conv1(cat[src0, src1]) + b1)          (2C -> 16)
// MODE 1: out = conv2(src0 = h) + b2 + aux (= base_out)   (16 -> C)
// MODE 2: gh = (aux (= h) > 0) * conv2^T(src0 = dout)     (C -> 16)
template <int C, int MODE> struct AdGeom {
  static constexpr int CI = MODE == 0 ? 2 * C : (MODE == 1 ? kAdHid : C);
  static constexpr int CO = MODE == 1 ? C : kAdHid;
  static constexpr int COP = MODE == 1 ? 4 : kAdHid;          // padded output channels of the weight rows
};

template <int C, int MODE>
__global__ void __launch_bounds__(256)
adapter_conv_kernel(const float* __restrict__ src0, const float* __restrict__ src1, const float* __restrict__ aux,
                    float* __restrict__ dst, const AdapterW* __restrict__ wts, int H, int W) {
  constexpr int CI = AdGeom<C, MODE>::CI, CO = AdGeom<C, MODE>::CO, COP = AdGeom<C, MODE>::COP;
  extern __shared__ __align__(16) float ad_smem[];
  float (*tile)[kAdTH + 2][kAdTWP] = reinterpret_cast<float (*)[kAdTH + 2][kAdTWP]>(ad_smem);
  float* ws = ad_smem + CI * (kAdTH + 2) * kAdTWP;             // [CI*9][COP] weights, then COP biases
  pdl_enter();
  {
    const float* wsrc = MODE == 0 ? wts->w1 : (MODE == 1 ? wts->w2 : wts->w2t);
    for (int i = threadIdx.x; i < CI * 9 * COP; i += 256) ws[i] = wsrc[i];
    if (threadIdx.x < COP) ws[CI * 9 * COP + threadIdx.x] = MODE == 0 ? wts->b1[threadIdx.x] : (MODE == 1 ? wts->b2[threadIdx.x] : 0.f);
  }
  const int n = blockIdx.z, y0 = blockIdx.y * kAdTH, x0 = blockIdx.x * kAdTW;
  const size_t plane = (size_t)H * W;
  for (int i = threadIdx.x; i < CI * (kAdTH + 2) * (kAdTW + 2); i += 256) {
    const int s = i % (kAdTW + 2), r = (i / (kAdTW + 2)) % (kAdTH + 2), ci = i / ((kAdTW + 2) * (kAdTH + 2));
    const int y = y0 + r - 1, x = x0 + s - 1;
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      constexpr int C0 = MODE == 1 ? kAdHid : C;            // channels of src0
      const float* p = (MODE == 0 && ci >= C) ? src1 + ((size_t)n * C + (ci - C)) * plane : src0 + ((size_t)n * C0 + ci) * plane;
      v = __ldg(p + (size_t)y * W + x);
    }
    tile[ci][r][s + 3] = v;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int y = y0 + ty, x = x0 + kAdPX * tx;
  if (y >= H || x >= W) return;
  float acc[CO][kAdPX];
#pragma unroll
  for (int co = 0; co < CO; ++co) {
    const float b = ws[CI * 9 * COP + co];
#pragma unroll
    for (int p = 0; p < kAdPX; ++p) acc[co][p] = b;
  }
#pragma unroll
  for (int ci = 0; ci < CI; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float* row = &tile[ci][ty + ky][kAdPX * tx];
      float a[6];
      a[0] = row[3];
      const float4 m = *reinterpret_cast<const float4*>(row + 4);
      a[1] = m.x; a[2] = m.y; a[3] = m.z; a[4] = m.w;
      a[5] = row[8];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* wr = reinterpret_cast<const float4*>(ws + (ci * 9 + ky * 3 + kx) * COP);
#pragma unroll
        for (int c4 = 0; c4 < COP / 4; ++c4) {
          const float4 w4 = wr[c4];
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int co = 4 * c4 + j;
            if (co < CO) {
#pragma unroll
              for (int p = 0; p < kAdPX; ++p) acc[co][p] = fmaf(wv[j], a[p + kx], acc[co][p]);
            }
          }
        }
      }
    }
  // W % 4 == 0 (host check): the four pixels are inside the image together and 16-byte aligned
#pragma unroll
  for (int co = 0; co < CO; ++co) {
    const size_t o = ((size_t)n * CO + co) * plane + (size_t)y * W + x;
    float4 v = make_float4(acc[co][0], acc[co][1], acc[co][2], acc[co][3]);
    if constexpr (MODE == 0) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    } else if constexpr (MODE == 1) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(aux + o));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    } else {
      const float4 h = __ldg(reinterpret_cast<const float4*>(aux + o));
      v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f; v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
    }
    *reinterpret_cast<float4*>(dst + o) = v;
  }
}

// ---- the same three convolutions on the tensor cores (precision "bf16": TF32 operands, fp32 accumulate) ----------------------
// Warp-level mma.sync.m16n8k8.tf32: one warp = one tile row, 16 pixels x (8 input channels of one tap) per A fragment read
// straight from the planar fp32 tile (4 x LDS.32 feed 2048 MACs; the FFMA form above needs one shared-memory wavefront per
// 128 MACs and is bound by that), all weight fragments (9 taps x <= 2 k-chunks x <= 2 n-tiles) live in registers.
// tcgen05 has no business here: K = 8 and N <= 16 would leave a 128 x N x 16 UMMA >95 % empty.
constexpr int kAdPS = (kAdTH + 2) * kAdTWP + 8;     // plane stride = 8 (mod 32): A[px g][ch t] reads hit 32 distinct banks

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
// Operands are rounded to TF32 (round-to-nearest) ONCE, on their way into shared memory — every element is read ~9 times (taps).
// Feeding raw fp32 bits instead lets the tensor core truncate: biased towards zero, which on all-positive operands (images,
// ReLU outputs) and 2 M-term sums showed up as a 1-2 % error of the weight gradients.
__device__ __forceinline__ float4 round_tf32(float4 v) {
  v.x = __uint_as_float(to_tf32(v.x)); v.y = __uint_as_float(to_tf32(v.y));
  v.z = __uint_as_float(to_tf32(v.z)); v.w = __uint_as_float(to_tf32(v.w));
  return v;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// planar halo tile [CIP][TH+2][TWP] (+8 floats per plane), image column x0 - 4 + s at index s; 16-byte quads, zero outside
template <int CI, int CIP, int C0>
__device__ __forceinline__ void ad_load_tile(float* tile, const float* __restrict__ src0, const float* __restrict__ src1, int n,
                                             int y0, int x0, int H, int W, int nthreads) {
  constexpr int QUADS = kAdTWP / 4;
  const size_t plane = (size_t)H * W;
  for (int i = threadIdx.x; i < CIP * (kAdTH + 2) * QUADS; i += nthreads) {
    const int q = i % QUADS, r = (i / QUADS) % (kAdTH + 2), ci = i / (QUADS * (kAdTH + 2));
    const int y = y0 + r - 1, x = x0 - 4 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ci < CI && y >= 0 && y < H && x >= 0 && x < W) {
      const float* p = ci < C0 ? src0 + ((size_t)n * C0 + ci) * plane : src1 + ((size_t)n * (CI - C0) + (ci - C0)) * plane;
      v = __ldg(reinterpret_cast<const float4*>(p + (size_t)y * W + x));
    }
    *reinterpret_cast<float4*>(tile + ci * kAdPS + r * kAdTWP + 4 * q) = round_tf32(v);
  }
}

template <int C, int MODE>
__global__ void __launch_bounds__(256)
adapter_conv_mma_kernel(const float* __restrict__ src0, const float* __restrict__ src1, const float* __restrict__ aux,
                        float* __restrict__ dst, const AdapterW* __restrict__ wts, int H, int W) {
  constexpr int CI = AdGeom<C, MODE>::CI, CO = AdGeom<C, MODE>::CO, COP = AdGeom<C, MODE>::COP;
  constexpr int KC = (CI + 7) / 8, CIP = 8 * KC, NT = (CO + 7) / 8;
  constexpr int C0 = MODE == 0 ? C : CI;                       // channels of src0 (conv1 reads cat[src0, src1])
  extern __shared__ __align__(16) float ad_smem[];
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // weight fragments: B[k = input channel][n = output channel] of every (tap, k-chunk, n-tile)
  uint32_t bf[9][KC][NT][2];
  float bias[NT][2];
  {
    const float* wsrc = MODE == 0 ? wts->w1 : (MODE == 1 ? wts->w2 : wts->w2t);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int kc = 0; kc < KC; ++kc)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int ci = kc * 8 + t + 4 * hh, co = nt * 8 + g;
            bf[tap][kc][nt][hh] = to_tf32((ci < CI && co < CO) ? wsrc[(ci * 9 + tap) * COP + co] : 0.f);
          }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int co = nt * 8 + 2 * t + j;
        bias[nt][j] = (MODE != 2 && co < CO) ? (MODE == 0 ? wts->b1[co] : wts->b2[co]) : 0.f;
      }
  }
  const int n = blockIdx.z, y0 = blockIdx.y * kAdTH, x0 = blockIdx.x * kAdTW;
  ad_load_tile<CI, CIP, C0>(ad_smem, src0, src1, n, y0, x0, H, W, 256);
  __syncthreads();
  const int y = y0 + warp;
  if (y >= H) return;
  const size_t plane = (size_t)H * W;
#pragma unroll 1
  for (int cg = 0; cg < kAdTW / 16; ++cg) {
    const int xg = x0 + 16 * cg;
    if (xg >= W) break;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][2] = bias[nt][0]; acc[nt][1] = acc[nt][3] = bias[nt][1]; }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      // pixel (row warp, column 16 cg + m) under tap (ky, kx) sits at tile row warp + ky, index 16 cg + m + kx + 3
      const float* base = ad_smem + (warp + tap / 3) * kAdTWP + 16 * cg + tap % 3 + 3;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const float* pa = base + (kc * 8 + t) * kAdPS + g;
        uint32_t a[4];
        a[0] = __float_as_uint(pa[0]); a[1] = __float_as_uint(pa[8]);
        a[2] = __float_as_uint(pa[4 * kAdPS]); a[3] = __float_as_uint(pa[4 * kAdPS + 8]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], a, bf[tap][kc][nt][0], bf[tap][kc][nt][1]);
      }
    }
    // c0 / c1: pixel g, channels 2t, 2t+1 of the n-tile; c2 / c3: pixel g + 8
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int co = nt * 8 + 2 * t + j;
        if (co >= CO) continue;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int x = xg + g + 8 * hh;
          if (x >= W) continue;
          const size_t o = ((size_t)n * CO + co) * plane + (size_t)y * W + x;
          float v = acc[nt][2 * hh + j];
          if constexpr (MODE == 0) v = fmaxf(v, 0.f);
          else if constexpr (MODE == 1) v += __ldg(aux + o);
          else v = __ldg(aux + o) > 0.f ? v : 0.f;
          dst[o] = v;
        }
      }
  }
}

// ---- weight gradients ---------------------------------------------------------------------------------------------------
// dW[cg][cx][k] = sum_p g[cg][p] * x[cx][p + off_k], db[cg] = sum_p g[cg][p].
// One warp = one block of weights (GB g-channels x XB x-channels x 9 taps), lanes = 32 adjacent columns; a lane walks down
// the 8 rows of a tile column keeping the 3 x 3 window of x in registers (3 new values per pixel).  Persistent CTAs.
template <int CG, int CX, int GB, int XB, int CSPLIT /* x channels [0, CSPLIT) come from x0p, the rest from x1p */>
__global__ void __launch_bounds__(32 * (CG / GB) * (CX / XB))
adapter_wgrad_kernel(const float* __restrict__ gp, const float* __restrict__ x0p, const float* __restrict__ x1p,
                     float* __restrict__ partial /* [grid][CG*CX*9 + CG] */, int N, int H, int W) {
  constexpr int NWARPS = (CG / GB) * (CX / XB), NT = 32 * NWARPS;
  extern __shared__ __align__(16) float ad_smem[];
  float (*xs)[kAdTH + 2][kAdTW + 2] = reinterpret_cast<float (*)[kAdTH + 2][kAdTW + 2]>(ad_smem);
  float (*gs)[kAdTH][kAdTW] = reinterpret_cast<float (*)[kAdTH][kAdTW]>(ad_smem + CX * (kAdTH + 2) * (kAdTW + 2));
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gb = warp / (CX / XB), xb = warp % (CX / XB);
  float acc[GB][XB][9], bacc[GB];
#pragma unroll
  for (int g = 0; g < GB; ++g) {
    bacc[g] = 0.f;
#pragma unroll
    for (int c = 0; c < XB; ++c)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[g][c][k] = 0.f;
  }
  const int tiles_x = (W + kAdTW - 1) / kAdTW, tiles_y = (H + kAdTH - 1) / kAdTH;
  const long long ntiles = (long long)N * tiles_x * tiles_y;
  const size_t plane = (size_t)H * W;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int n = (int)(t / (tiles_x * tiles_y)), rem = (int)(t % (tiles_x * tiles_y));
    const int y0 = (rem / tiles_x) * kAdTH, x0 = (rem % tiles_x) * kAdTW;
    __syncthreads();
    for (int i = threadIdx.x; i < CX * (kAdTH + 2) * (kAdTW + 2); i += NT) {
      const int s = i % (kAdTW + 2), r = (i / (kAdTW + 2)) % (kAdTH + 2), c = i / ((kAdTW + 2) * (kAdTH + 2));
      const int y = y0 + r - 1, x = x0 + s - 1;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float* p = c < CSPLIT ? x0p + ((size_t)n * CSPLIT + c) * plane : x1p + ((size_t)n * (CX - CSPLIT) + (c - CSPLIT)) * plane;
        v = __ldg(p + (size_t)y * W + x);
      }
      xs[c][r][s] = v;
    }
    for (int i = threadIdx.x; i < CG * kAdTH * kAdTW; i += NT) {
      const int s = i % kAdTW, r = (i / kAdTW) % kAdTH, c = i / (kAdTW * kAdTH);
      const int y = y0 + r, x = x0 + s;
      gs[c][r][s] = (y < H && x < W) ? __ldg(gp + ((size_t)n * CG + c) * plane + (size_t)y * W + x) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < kAdTW / 32; ++j) {
      const int s = 32 * j + lane;
      float win[XB][3][3];
#pragma unroll
      for (int c = 0; c < XB; ++c)
#pragma unroll
        for (int ky = 0; ky < 2; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) win[c][ky + 1][kx] = xs[xb * XB + c][ky][s + kx];
#pragma unroll
      for (int r = 0; r < kAdTH; ++r) {
#pragma unroll
        for (int c = 0; c < XB; ++c)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            win[c][0][kx] = win[c][1][kx]; win[c][1][kx] = win[c][2][kx];
            win[c][2][kx] = xs[xb * XB + c][r + 2][s + kx];
          }
#pragma unroll
        for (int g = 0; g < GB; ++g) {
          const float gv = gs[gb * GB + g][r][s];
          bacc[g] += gv;
#pragma unroll
          for (int c = 0; c < XB; ++c)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[g][c][k] = fmaf(gv, win[c][k / 3][k % 3], acc[g][c][k]);
        }
      }
    }
  }
  float* out = partial + (size_t)blockIdx.x * (CG * CX * 9 + CG);
#pragma unroll
  for (int g = 0; g < GB; ++g) {
#pragma unroll
    for (int c = 0; c < XB; ++c)
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float v = acc[g][c][k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out[((gb * GB + g) * CX + xb * XB + c) * 9 + k] = v;
      }
    float b = bacc[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0 && xb == 0) out[CG * CX * 9 + gb * GB + g] = b;
  }
}

// ---- weight gradients on the tensor cores (precision "bf16") -------------------------------------------------------------
// K = pixels.  dW1: D[hid 16][n = (ci, tap) 54 + a ones column = db1] = G^T (16 x px) * im2col(X) (px x n): A fragments from the
// planar gradient tile, B fragments gathered from the planar halo tile through per-lane (ci, tap) offsets.
// dW2: one accumulator per tap, D_tap[ci 16][co] = shifted H^T (16 x px) * Dout (px x co).  One warp = one tile row; the
// accumulators persist over all tiles of the CTA and are reduced over warps (shared memory) and CTAs (reduce kernel).
constexpr int kAdGPS = kAdTH * kAdTW + 4;                    // interior-only planes: stride = 4 (mod 32) -> A[ch g][px t] conflict-free
constexpr int kAdHPS = (kAdTH + 2) * kAdTWP + 20;            // halo planes read as A[ch g][px t]: stride = 4 (mod 32)

template <int CH>
__device__ __forceinline__ void ad_load_interior(float* tile, const float* __restrict__ src, int n, int y0, int x0, int H, int W, int nthreads) {
  const size_t plane = (size_t)H * W;
  for (int i = threadIdx.x; i < CH * kAdTH * (kAdTW / 4); i += nthreads) {
    const int q = i % (kAdTW / 4), r = (i / (kAdTW / 4)) % kAdTH, c = i / ((kAdTW / 4) * kAdTH);
    const int y = y0 + r, x = x0 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && x < W) v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)n * CH + c) * plane + (size_t)y * W + x));
    *reinterpret_cast<float4*>(tile + c * kAdGPS + r * kAdTW + 4 * q) = round_tf32(v);
  }
}
template <int CH, int PS>
__device__ __forceinline__ void ad_load_halo(float* tile, const float* __restrict__ src0, const float* __restrict__ src1, int c0, int n,
                                             int y0, int x0, int H, int W, int nthreads) {
  constexpr int QUADS = kAdTWP / 4;
  const size_t plane = (size_t)H * W;
  for (int i = threadIdx.x; i < CH * (kAdTH + 2) * QUADS; i += nthreads) {
    const int q = i % QUADS, r = (i / QUADS) % (kAdTH + 2), c = i / (QUADS * (kAdTH + 2));
    const int y = y0 + r - 1, x = x0 - 4 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const float* p = c < c0 ? src0 + ((size_t)n * c0 + c) * plane : src1 + ((size_t)n * (CH - c0) + (c - c0)) * plane;
      v = __ldg(reinterpret_cast<const float4*>(p + (size_t)y * W + x));
    }
    *reinterpret_cast<float4*>(tile + c * PS + r * kAdTWP + 4 * q) = round_tf32(v);
  }
}

template <int C>
__global__ void __launch_bounds__(256)
adapter_wgrad1_mma_kernel(const float* __restrict__ gh, const float* __restrict__ noisy, const float* __restrict__ base_out,
                          float* __restrict__ partial /* [grid][16*2C*9 + 16] */, int N, int H, int W) {
  constexpr int CX = 2 * C, NCOL = CX * 9, NTL = (NCOL + 8) / 8;       // + the ones column
  extern __shared__ __align__(16) float ad_smem[];
  float* xs = ad_smem;                       // [CX][10][136] halo planes, stride kAdPS
  float* gs = ad_smem + CX * kAdPS;          // [16][8][128] planes, stride kAdGPS
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  int boff[NTL], bkind[NTL];                 // per lane: where column n = 8j + g of the im2col operand starts; 0 data, 1 ones, 2 zero
#pragma unroll
  for (int j = 0; j < NTL; ++j) {
    const int nn = 8 * j + g;
    bkind[j] = nn < NCOL ? 0 : (nn == NCOL ? 1 : 2);
    const int ci = nn < NCOL ? nn / 9 : 0, tap = nn < NCOL ? nn % 9 : 0;
    boff[j] = ci * kAdPS + (tap / 3) * kAdTWP + tap % 3 + 3;
  }
  float acc[NTL][4];
#pragma unroll
  for (int j = 0; j < NTL; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const int tiles_x = (W + kAdTW - 1) / kAdTW, tiles_y = (H + kAdTH - 1) / kAdTH;
  const long long ntiles = (long long)N * tiles_x * tiles_y;
  for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int n = (int)(tl / (tiles_x * tiles_y)), rem = (int)(tl % (tiles_x * tiles_y));
    const int y0 = (rem / tiles_x) * kAdTH, x0 = (rem % tiles_x) * kAdTW;
    __syncthreads();
    ad_load_halo<CX, kAdPS>(xs, noisy, base_out, C, n, y0, x0, H, W, 256);
    ad_load_interior<kAdHid>(gs, gh, n, y0, x0, H, W, 256);
    __syncthreads();
    const float* grow = gs + warp * kAdTW;                    // this warp's row of the gradient tile
    const float* xrow = xs + warp * kAdTWP;                   // tile row warp + ky via boff
#pragma unroll 2
    for (int ks = 0; ks < kAdTW / 8; ++ks) {
      const int px = 8 * ks;
      uint32_t a[4];
      a[0] = __float_as_uint(grow[g * kAdGPS + px + t]);       a[1] = __float_as_uint(grow[(g + 8) * kAdGPS + px + t]);
      a[2] = __float_as_uint(grow[g * kAdGPS + px + t + 4]);   a[3] = __float_as_uint(grow[(g + 8) * kAdGPS + px + t + 4]);
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        float b0 = xrow[boff[j] + px + t], b1 = xrow[boff[j] + px + t + 4];
        if (8 * j + 7 >= NCOL) {                               // only the last n-tile(s) hold the ones / padding columns
          b0 = bkind[j] == 0 ? b0 : (bkind[j] == 1 ? 1.f : 0.f);
          b1 = bkind[j] == 0 ? b1 : (bkind[j] == 1 ? 1.f : 0.f);
        }
        mma_tf32(acc[j], a, __float_as_uint(b0), __float_as_uint(b1));
      }
    }
  }
  // reduce over the 8 warps: red[warp][hid 16][NTL*8]
  __syncthreads();
  float* red = ad_smem;
#pragma unroll
  for (int j = 0; j < NTL; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) red[(warp * 16 + g + 8 * (e >> 1)) * (NTL * 8) + 8 * j + 2 * t + (e & 1)] = acc[j][e];
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * (16 * NCOL + 16);
  for (int i = threadIdx.x; i < 16 * (NCOL + 1); i += 256) {
    const int hid = i / (NCOL + 1), nn = i % (NCOL + 1);
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[(w * 16 + hid) * (NTL * 8) + nn];
    if (nn < NCOL) out[hid * NCOL + nn] = sum; else out[16 * NCOL + hid] = sum;
  }
}

template <int C>
__global__ void __launch_bounds__(256)
adapter_wgrad2_mma_kernel(const float* __restrict__ dout, const float* __restrict__ hbuf, float* __restrict__ partial /* [grid][C*16*9 + C] */,
                          int N, int H, int W) {
  extern __shared__ __align__(16) float ad_smem[];
  float* hs = ad_smem;                       // [16][10][136] halo planes of h, stride kAdHPS
  float* ds = ad_smem + kAdHid * kAdHPS;     // [C][8][128] planes of dL/dout, stride kAdGPS
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float acc[9][4], bsum = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  const int tiles_x = (W + kAdTW - 1) / kAdTW, tiles_y = (H + kAdTH - 1) / kAdTH;
  const long long ntiles = (long long)N * tiles_x * tiles_y;
  for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int n = (int)(tl / (tiles_x * tiles_y)), rem = (int)(tl % (tiles_x * tiles_y));
    const int y0 = (rem / tiles_x) * kAdTH, x0 = (rem % tiles_x) * kAdTW;
    __syncthreads();
    ad_load_halo<kAdHid, kAdHPS>(hs, hbuf, hbuf, kAdHid, n, y0, x0, H, W, 256);
    ad_load_interior<C>(ds, dout, n, y0, x0, H, W, 256);
    __syncthreads();
    const float* drow = ds + warp * kAdTW + (g < C ? g : 0) * kAdGPS;
#pragma unroll 2
    for (int ks = 0; ks < kAdTW / 8; ++ks) {
      const int px = 8 * ks;
      float b0 = 0.f, b1 = 0.f;                               // B[k = px t][n = co g]
      if (g < C) { b0 = drow[px + t]; b1 = drow[px + t + 4]; }
      bsum += b0 + b1;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float* pa = hs + (warp + tap / 3) * kAdTWP + px + tap % 3 + 3 + t;     // A[m = ci g][k = px t], shifted by the tap
        uint32_t a[4];
        a[0] = __float_as_uint(pa[g * kAdHPS]);     a[1] = __float_as_uint(pa[(g + 8) * kAdHPS]);
        a[2] = __float_as_uint(pa[g * kAdHPS + 4]); a[3] = __float_as_uint(pa[(g + 8) * kAdHPS + 4]);
        mma_tf32(acc[tap], a, __float_as_uint(b0), __float_as_uint(b1));
      }
    }
  }
  bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
  bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);               // lanes with the same g: sum over this warp's pixels of dout[co = g]
  __syncthreads();
  float* red = ad_smem;                                        // [warp][tap][ci 16][co 8], then [warp][8] bias sums
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int e = 0; e < 4; ++e) red[((warp * 9 + tap) * 16 + g + 8 * (e >> 1)) * 8 + 2 * t + (e & 1)] = acc[tap][e];
  if (t == 0) red[8 * 9 * 16 * 8 + warp * 8 + g] = bsum;
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * (C * 16 * 9 + C);
  for (int i = threadIdx.x; i < C * 16 * 9 + C; i += 256) {
    float sum = 0.f;
    if (i < C * 16 * 9) {
      const int co = i / (16 * 9), ci = (i / 9) % 16, tap = i % 9;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[((w * 9 + tap) * 16 + ci) * 8 + co];
    } else {
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[8 * 9 * 16 * 8 + w * 8 + (i - C * 16 * 9)];
    }
    out[i] = sum;
  }
}

// grads = sum over the persistent CTAs' partials, fixed order; one thread per gradient element
// block = 32 outputs x 8 slices of the CTA range (coalesced 128-byte rows), slices folded in a fixed order
__global__ void __launch_bounds__(256)
adapter_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int nw, int nb, float* __restrict__ dw, float* __restrict__ db) {
  pdl_enter();
  __shared__ double red[8][32];
  const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  double s = 0.0;
  if (i < nw + nb)
    for (int p = slice; p < nparts; p += 8) s += (double)partial[(size_t)p * (nw + nb) + i];
  red[slice][col] = s;
  __syncthreads();
  if (slice == 0 && i < nw + nb) {
    for (int k = 1; k < 8; ++k) s += red[k][col];
    if (i < nw) dw[i] = (float)s; else db[i - nw] = (float)s;
  }
}

}  // namespace n2n

using namespace n2n;

namespace n2n {
constexpr int kAdWgradCtas = kSMs * 2;
struct AdapterFusedWs {       // byte offsets into the caller's workspace
  size_t off_const, off_h, off_gh, off_p1, off_p2, total;
};
AdapterFusedWs adapter_fused_layout(int C, int n, int h, int w, bool bwd) {
  AdapterFusedWs L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  L.off_const = take(sizeof(AdapterW));
  L.off_h = take((size_t)n * kAdHid * h * w * sizeof(float));
  L.off_gh = L.off_p1 = L.off_p2 = 0;
  if (bwd) {
    L.off_gh = take((size_t)n * kAdHid * h * w * sizeof(float));
    L.off_p1 = take((size_t)kAdWgradCtas * (kAdHid * 2 * C * 9 + kAdHid) * sizeof(float));
    L.off_p2 = take((size_t)kAdWgradCtas * (C * kAdHid * 9 + C) * sizeof(float));
  }
  L.total = off;
  return L;
}
bool adapter_fused_ok(int C, int hid, int w) { return hid == kAdHid && (C == 1 || C == 3) && w % 4 == 0; }

template <int C, int MODE>
static int ad_launch_conv(const float* s0, const float* s1, const float* aux, float* dst, const AdapterW* wts, int n, int h, int w,
                          cudaStream_t st) {
  constexpr int CI = AdGeom<C, MODE>::CI, COP = AdGeom<C, MODE>::COP;
  const size_t smem = ((size_t)CI * (kAdTH + 2) * kAdTWP + (size_t)CI * 9 * COP + COP) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    N2N_CUDA(cudaFuncSetAttribute(adapter_conv_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((w + kAdTW - 1) / kAdTW, (h + kAdTH - 1) / kAdTH, n);
  (void)launch_pdl_v(adapter_conv_kernel<C, MODE>, grid, dim3(256), smem, st, s0, s1, aux, dst, wts, h, w);
  N2N_LAUNCH_CHECK();
  return 0;
}
template <int C, int MODE>
static int ad_launch_conv_mma(const float* s0, const float* s1, const float* aux, float* dst, const AdapterW* wts, int n, int h, int w,
                              cudaStream_t st) {
  constexpr int CIP = 8 * ((AdGeom<C, MODE>::CI + 7) / 8);
  const size_t smem = (size_t)CIP * kAdPS * sizeof(float);
  static bool attr = false;
  if (!attr) {
    N2N_CUDA(cudaFuncSetAttribute(adapter_conv_mma_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((w + kAdTW - 1) / kAdTW, (h + kAdTH - 1) / kAdTH, n);
  (void)launch_pdl_v(adapter_conv_mma_kernel<C, MODE>, grid, dim3(256), smem, st, s0, s1, aux, dst, wts, h, w);
  N2N_LAUNCH_CHECK();
  return 0;
}
template <int CG, int CX, int GB, int XB, int CSPLIT>
static int ad_launch_wgrad(const float* g, const float* x0, const float* x1, float* partial, int n, int h, int w, cudaStream_t st) {
  constexpr int NT = 32 * (CG / GB) * (CX / XB);
  const size_t smem = ((size_t)CX * (kAdTH + 2) * (kAdTW + 2) + (size_t)CG * kAdTH * kAdTW) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    N2N_CUDA(cudaFuncSetAttribute(adapter_wgrad_kernel<CG, CX, GB, XB, CSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  (void)launch_pdl_v(adapter_wgrad_kernel<CG, CX, GB, XB, CSPLIT>, dim3(kAdWgradCtas), dim3(NT), smem, st, g, x0, x1, partial, n, h, w);
  N2N_LAUNCH_CHECK();
  return 0;
}

template <int C>
static int ad_launch_wgrad_mma(const float* gh, const float* noisy, const float* base_out, const float* dout, const float* hbuf,
                               float* p1, float* p2, int n, int h, int w, cudaStream_t st) {
  const size_t smem1 = ((size_t)2 * C * kAdPS + (size_t)kAdHid * kAdGPS) * sizeof(float);
  size_t smem2 = ((size_t)kAdHid * kAdHPS + (size_t)C * kAdGPS) * sizeof(float);
  const size_t red2 = ((size_t)8 * 9 * 16 * 8 + 64) * sizeof(float);
  if (smem2 < red2) smem2 = red2;
  static bool attr = false;
  if (!attr) {
    N2N_CUDA(cudaFuncSetAttribute(adapter_wgrad1_mma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    N2N_CUDA(cudaFuncSetAttribute(adapter_wgrad2_mma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    attr = true;
  }
  (void)launch_pdl_v(adapter_wgrad2_mma_kernel<C>, dim3(kAdWgradCtas), dim3(256), smem2, st, dout, hbuf, p2, n, h, w);
  N2N_LAUNCH_CHECK();
  (void)launch_pdl_v(adapter_wgrad1_mma_kernel<C>, dim3(kAdWgradCtas), dim3(256), smem1, st, gh, noisy, base_out, p1, n, h, w);
  N2N_LAUNCH_CHECK();
  return 0;
}

static int ad_prepare_weights(int C, const float* const* params, AdapterW* stage, cudaStream_t st) {
  (void)launch_pdl_v(adapter_prep_kernel, dim3((kAdHid * 6 * 9 + 255) / 256), dim3(256), 0, st, params[0], params[1], params[2], params[3], C, stage);
  N2N_LAUNCH_CHECK();
  (void)launch_pdl_v(adapter_prep_w2_kernel, dim3((3 * kAdHid * 9 + 255) / 256), dim3(256), 0, st, params[2], C, stage);
  N2N_LAUNCH_CHECK();
  return 0;
}

template <int C>
static int ad_forward(const float* const* params, const float* noisy, const float* base_out, float* out, void* ws,
                      const AdapterFusedWs& L, int n, int h, int w, bool tensor, cudaStream_t st) {
  float* hbuf = (float*)((char*)ws + L.off_h);
  AdapterW* wts = (AdapterW*)((char*)ws + L.off_const);
  N2N_TRY(ad_prepare_weights(C, params, wts, st));
  if (tensor) {
    N2N_TRY((ad_launch_conv_mma<C, 0>(noisy, base_out, nullptr, hbuf, wts, n, h, w, st)));
    return ad_launch_conv_mma<C, 1>(hbuf, nullptr, base_out, out, wts, n, h, w, st);
  }
  N2N_TRY((ad_launch_conv<C, 0>(noisy, base_out, nullptr, hbuf, wts, n, h, w, st)));
  return ad_launch_conv<C, 1>(hbuf, nullptr, base_out, out, wts, n, h, w, st);
}
template <int C>
static int ad_backward(const float* const* params, const float* noisy, const float* base_out, const float* dout,
                       float* const* grads, void* ws, const AdapterFusedWs& L, int n, int h, int w, bool tensor, cudaStream_t st) {
  float* hbuf = (float*)((char*)ws + L.off_h);
  float* gh = (float*)((char*)ws + L.off_gh);
  float* p1 = (float*)((char*)ws + L.off_p1);
  float* p2 = (float*)((char*)ws + L.off_p2);
  const AdapterW* wts = (const AdapterW*)((char*)ws + L.off_const);    // left in the workspace by the forward of this plan
  if (tensor) N2N_TRY((ad_launch_conv_mma<C, 2>(dout, nullptr, hbuf, gh, wts, n, h, w, st)));
  else N2N_TRY((ad_launch_conv<C, 2>(dout, nullptr, hbuf, gh, wts, n, h, w, st)));
  if (tensor) {
    N2N_TRY((ad_launch_wgrad_mma<C>(gh, noisy, base_out, dout, hbuf, p1, p2, n, h, w, st)));
  } else {
    N2N_TRY((ad_launch_wgrad<C, kAdHid, C, 2, kAdHid>(dout, hbuf, nullptr, p2, n, h, w, st)));
    N2N_TRY((ad_launch_wgrad<kAdHid, 2 * C, 8, 1, C>(gh, noisy, base_out, p1, n, h, w, st)));
  }
  const int nw1 = kAdHid * 2 * C * 9, nw2 = C * kAdHid * 9;
  (void)launch_pdl_v(adapter_wgrad_reduce_kernel, dim3((nw1 + kAdHid + 31) / 32), dim3(256), 0, st, (const float*)p1, kAdWgradCtas, nw1,
                     kAdHid, grads[0], grads[1]);
  N2N_LAUNCH_CHECK();
  (void)launch_pdl_v(adapter_wgrad_reduce_kernel, dim3((nw2 + C + 31) / 32), dim3(256), 0, st, (const float*)p2, kAdWgradCtas, nw2, C,
                     grads[2], grads[3]);
  N2N_LAUNCH_CHECK();
  return 0;
}

int adapter_fused_forward(int C, int dtype, const float* const* params, const float* noisy, const float* base_out, float* out, void* ws,
                          int n, int h, int w, bool bwd, cudaStream_t st) {
  const AdapterFusedWs L = adapter_fused_layout(C, n, h, w, bwd);
  const bool tensor = dtype == N2N_BF16;
  return C == 1 ? ad_forward<1>(params, noisy, base_out, out, ws, L, n, h, w, tensor, st)
                : ad_forward<3>(params, noisy, base_out, out, ws, L, n, h, w, tensor, st);
}
int adapter_fused_backward(int C, int dtype, const float* const* params, const float* noisy, const float* base_out, const float* dout,
                           float* const* grads, void* ws, int n, int h, int w, cudaStream_t st) {
  const AdapterFusedWs L = adapter_fused_layout(C, n, h, w, true);
  const bool tensor = dtype == N2N_BF16;
  return C == 1 ? ad_backward<1>(params, noisy, base_out, dout, grads, ws, L, n, h, w, tensor, st)
                : ad_backward<3>(params, noisy, base_out, dout, grads, ws, L, n, h, w, tensor, st);
}
}  // namespace n2n
