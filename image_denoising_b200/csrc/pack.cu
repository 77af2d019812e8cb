// pack.cu — weight repack (PyTorch fp32 layouts -> engine layouts) and the inverse reduction of
// weight-gradient partials back into PyTorch-layout fp32 gradients.
//
// Engine weight layouts (one "slab" per tap t; n = GEMM output channel, c = contraction channel):
//   fp32 engine : Wp[t][n][c]                     n < nout_pad, c < 16*cin_blocks, zero padded
//   bf16 engine : the exact shared-memory image the tcgen05 kernel wants, so that one
//                 cp.async.bulk per pipeline stage brings a ready-to-use B operand:
//                 Wp[t][g][j][n][16] with g = channel-block group (3 blocks = 48 channels per
//                 stage), j = block in group, rows of 32 bytes in the K-major SWIZZLE_32B
//                 pattern (16-byte chunk index XOR bit 2 of the row index).
#include "common.cuh"

namespace n2n {

constexpr int kGroupBlocks = 3;   // channel blocks per pipeline stage of the bf16 engine

size_t packed_weight_bytes(int dtype, int ntaps, int nout_pad, int cin_blocks) {
  if (dtype == N2N_BF16) {
    const int ngroups = (cin_blocks + kGroupBlocks - 1) / kGroupBlocks;
    return (size_t)ntaps * ngroups * kGroupBlocks * nout_pad * 32;
  }
  return (size_t)ntaps * nout_pad * cin_blocks * 16 * sizeof(float);
}

__device__ __forceinline__ int seg_lookup(const Segs& s, int dst, long long* extra = nullptr) {
  for (int i = 0; i < s.n; ++i)
    if (dst >= s.dst0[i] && dst < s.dst0[i] + s.cnt[i]) {
      if (extra) *extra = s.off[i];
      return s.src0[i] + (dst - s.dst0[i]);
    }
  return -1;
}

struct PackBatch { PackJob j[8]; int n; };
struct UnpackBatch { UnpackJob j[16]; int n; };

template <bool BF16>
__global__ void pack_kernel(const __grid_constant__ PackBatch b) {
  pdl_enter_no_release();   // its output is prefetched by the next GEMM's prologue
  const PackJob& J = b.j[blockIdx.y];
  const int cpad = J.cin_blocks * 16;
  const long long total = (long long)J.ntaps * J.nout_pad * cpad;
  const int ngroups = (J.cin_blocks + kGroupBlocks - 1) / kGroupBlocks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpad);
    const int n = (int)((i / cpad) % J.nout_pad);
    const int t = (int)(i / ((long long)cpad * J.nout_pad));
    long long noff = 0;
    const int sn = seg_lookup(J.nseg, n, &noff);
    float v = 0.f;
    if (J.im2col_nc > 0) {
      const int tap = c / J.im2col_nc, ch = c - tap * J.im2col_nc;
      if (sn >= 0 && tap < 9) v = J.src[tap * J.s_t + sn * J.s_n + (long long)(J.im2col_c0 + ch) * J.s_c];
    } else {
      const int sc = seg_lookup(J.cseg, c);
      if (sn >= 0 && sc >= 0) v = J.src[t * J.s_t + sn * J.s_n + sc * J.s_c + noff];
    }
    if (BF16) {
      const int cb = c >> 4, e = c & 15;
      const int g = cb / kGroupBlocks, jj = cb - g * kGroupBlocks;
      const size_t byte = ((size_t)((t * ngroups + g) * kGroupBlocks + jj) * J.nout_pad + n) * 32 +
                          ((((e >> 3) ^ ((n >> 2) & 1))) << 4) + (e & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>((char*)J.dst + byte) = __float2bfloat16_rn(v);
    } else {
      ((float*)J.dst)[i] = v;
    }
  }
  if (BF16) {
    // zero the channel blocks that pad the last group (the MMA never reads them, the bulk copy does)
    const int padb = ngroups * kGroupBlocks - J.cin_blocks;
    if (padb > 0) {
      const long long ptotal = (long long)J.ntaps * padb * J.nout_pad * 16;
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ptotal;
           i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i & 15);
        const int n = (int)((i >> 4) % J.nout_pad);
        const int pb = (int)((i / (16LL * J.nout_pad)) % padb);
        const int t = (int)(i / (16LL * J.nout_pad * padb));
        const int jj = (J.cin_blocks % kGroupBlocks) + pb;
        const size_t byte = ((size_t)((t * ngroups + (ngroups - 1)) * kGroupBlocks + jj) * J.nout_pad + n) * 32 + e * 2;
        *reinterpret_cast<__nv_bfloat16*>((char*)J.dst + byte) = __float2bfloat16_rn(0.f);
      }
    }
  }
}

// Four consecutive lanes share one group of four output channels: each sums a fixed quarter of the splits, the
// quarters are combined by shuffles in a fixed order (deterministic).  One thread per output made the kernel a
// chain of splits / 8 dependent load batches (19 DRAM round trips at 148 splits) whatever the data volume.
constexpr int kUnpackLanes = 4;

__global__ void unpack_kernel(const __grid_constant__ UnpackBatch b) {
  pdl_enter();
  const UnpackJob& J = b.j[blockIdx.y];
  const int plane = J.cpad * J.npad;
  const int total = J.ntaps * plane;
  const int npad4 = J.npad >> 2;
  const int sub = threadIdx.x & (kUnpackLanes - 1);
  const unsigned gmask = ((1u << kUnpackLanes) - 1u) << (threadIdx.x & (32 - kUnpackLanes) & 31);   // this group's lanes within the warp
  const int per = (J.splits + kUnpackLanes - 1) / kUnpackLanes;
  const int s0 = sub * per < J.splits ? sub * per : J.splits;
  const int s1 = s0 + per < J.splits ? s0 + per : J.splits;
  const int groups = (gridDim.x * blockDim.x) / kUnpackLanes;
  for (int i4 = (blockIdx.x * blockDim.x + threadIdx.x) / kUnpackLanes; i4 < (total >> 2); i4 += groups) {
    const int n0 = (i4 % npad4) << 2;
    const int ct = i4 / npad4;
    const int c = ct % J.cpad;
    const int t = ct / J.cpad;
    long long cdst;
    if (J.im2col_nc > 0) {
      const int tap = c / J.im2col_nc, ch = c - tap * J.im2col_nc;
      if (tap >= 9) continue;
      cdst = tap * J.s_t + (long long)(J.im2col_c0 + ch) * J.s_c;
    } else {
      const int sc = seg_lookup(J.cseg, c);
      if (sc < 0) continue;
      cdst = t * J.s_t + sc * J.s_c;
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = reinterpret_cast<const float4*>(J.partial) + i4;
    const long long stride4 = total >> 2;
    int sp = s0;
    for (; sp + 8 <= s1; sp += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (long long)(sp + u) * stride4);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; sp < s1; ++sp) {
      const float4 v = __ldg(src + (long long)sp * stride4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
#pragma unroll
    for (int o = kUnpackLanes / 2; o > 0; o >>= 1) {
      s.x += __shfl_down_sync(gmask, s.x, o, kUnpackLanes); s.y += __shfl_down_sync(gmask, s.y, o, kUnpackLanes);
      s.z += __shfl_down_sync(gmask, s.z, o, kUnpackLanes); s.w += __shfl_down_sync(gmask, s.w, o, kUnpackLanes);
    }
    if (sub == 0) {
      const float r[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int sn = seg_lookup(J.nseg, n0 + q);
        if (sn >= 0) J.dst_w[cdst + sn * J.s_n] = r[q];
      }
    }
  }
  if (J.dst_b && J.bias_partial) {
    // one warp per channel: lane l sums rows l, l + 32, ... (up to splits x taps = 592 rows for the deconvs; one
    // thread per channel made that a 74-deep chain of dependent load batches, the longest thing in the launch),
    // then a fixed-order shuffle tree
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < J.npad; n += nwarps) {
      float s = 0.f;
      int r = lane;
      for (; r + 96 < J.bias_rows; r += 128) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(J.bias_partial + (long long)(r + 32 * u) * J.npad + n);
#pragma unroll
        for (int u = 0; u < 4; ++u) s += v[u];
      }
      for (; r < J.bias_rows; r += 32) s += __ldg(J.bias_partial + (long long)r * J.npad + n);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) {
        const int sn = seg_lookup(J.nseg, n);
        if (sn >= 0) J.dst_b[sn] = s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused ConvTranspose2x2(s2) -> conv3x3 (arch_unet.py:57-62 followed by dec_conv{k}a, :230-242), no-grad passes.
// conv3x3(cat[deconv(x), skip]) is linear in x: for the output pixel (2i+py, 2j+px) the nine taps of the 3x3 conv
// land on a 2x2 neighbourhood of SOURCE pixels (i+sy, j+sx), sy in {py-1, py}, sx in {px-1, px}, so
//   y[:, 2i+py, 2j+px] = sum_{sy,sx} Wc[py][px][sy][sx] x[:, i+sy, j+sx] + (skip part) + bias,
//   Wc[py][px][sy][sx][co][ci] = sum_{ky: floor((py+ky-1)/2) = sy} sum_{kx: ...} sum_c
//                                 W3[co][c][ky][kx] * Wd[ci][c][(py+ky-1)&1][(px+kx-1)&1]
// — 4 source taps of K = Ci instead of 9 taps of K = Cu on the upsampled image, and the upsampled tensor is never
// written.  This kernel builds the 16 composite matrices per layer (bf16, engine layout, one region per py), the
// full bias b3 + sum_taps W3[tap] bd, and the border table: the 3x3 conv zero-pads the UPSAMPLED image, so a
// border pixel's out-of-image taps contribute no ConvTranspose bias: corr[3*ycls + xcls][co] (1 = first row / column
// -> ky / kx = 0 outside, 2 = last -> ky / kx = 2 outside) is what the epilogue subtracts there.
struct UpFuseBatch { UpFuseJob j[5]; int n; };

// One block = one (16 co x 16 ci) tile of ALL 16 composite matrices of a layer: the nine taps of W3[co][c][.] and the four
// taps of Wd[ci][c][.] of a 16-channel chunk are staged in shared memory with contiguous (coalesced) global reads, and
// every thread accumulates its (co, ci) entry of the 16 composites in registers: 36 FMAs per reduction channel.
constexpr int kUfTile = 16;
__global__ void __launch_bounds__(kUfTile * kUfTile)
upfuse_pack_kernel(const __grid_constant__ UpFuseBatch b) {
  pdl_enter_no_release();   // its output is prefetched by the next GEMM's prologue
  const UpFuseJob& J = b.j[blockIdx.y];
  __shared__ float sA[kUfTile][9][kUfTile];       // [c][tap][co]
  __shared__ float sB[kUfTile][4][kUfTile];       // [c][a*2+b][ci]
  const int ci_pad = J.ngroups * kGroupBlocks * 16;        // whole groups (the bulk copy reads the padding too)
  const int co_span = (J.dst_t && J.gco * kGroupBlocks * 16 > J.co_pad) ? J.gco * kGroupBlocks * 16 : J.co_pad;
  const int tiles_ci = ci_pad / kUfTile, tiles_co = co_span / kUfTile;
  const int cin3 = J.Cu + J.Cs;                            // dec_conv a's input channels
  if ((int)blockIdx.x < tiles_ci * tiles_co) {
    const int co0 = ((int)blockIdx.x / tiles_ci) * kUfTile, ci0 = ((int)blockIdx.x % tiles_ci) * kUfTile;
    const int tco = threadIdx.x / kUfTile, tci = threadIdx.x % kUfTile;
    float acc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) acc[u] = 0.f;
    for (int c0 = 0; c0 < J.Cu; c0 += kUfTile) {
      __syncthreads();
      // W3[co0 + r][c0 + c][t]: for a fixed co the 16 x 9 values are contiguous in global memory
      for (int i = threadIdx.x; i < kUfTile * kUfTile * 9; i += kUfTile * kUfTile) {
        const int r = i / (kUfTile * 9), rem = i - r * (kUfTile * 9);
        const int c = rem / 9, t = rem - c * 9;
        const int co = co0 + r, cc = c0 + c;
        sA[c][t][r] = (co < J.Co && cc < J.Cu) ? __ldg(J.w3 + ((long long)co * cin3 + cc) * 9 + t) : 0.f;
      }
      // Wd[ci0 + r][c0 + c][ab]: 16 x 4 contiguous values per ci
      for (int i = threadIdx.x; i < kUfTile * kUfTile * 4; i += kUfTile * kUfTile) {
        const int r = i / (kUfTile * 4), rem = i - r * (kUfTile * 4);
        const int c = rem / 4, ab = rem - c * 4;
        const int ci = ci0 + r, cc = c0 + c;
        sB[c][ab][r] = (ci < J.Ci && cc < J.Cu) ? __ldg(J.wd + ((long long)ci * J.Cu + cc) * 4 + ab) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int c = 0; c < kUfTile; ++c) {
        float a[9], w[4];
#pragma unroll
        for (int t = 0; t < 9; ++t) a[t] = sA[c][t][tco];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = sB[c][q][tci];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int yo = py + ky - 1, xo = px + kx - 1;
                const int syi = (yo >= 0 ? yo >> 1 : -1) - (py - 1), sxi = (xo >= 0 ? xo >> 1 : -1) - (px - 1);
                const int u = ((py * 2 + px) * 2 + syi) * 2 + sxi;
                acc[u] = fmaf(a[ky * 3 + kx], w[(yo & 1) * 2 + (xo & 1)], acc[u]);
              }
      }
    }
    const int co = co0 + tco, ci = ci0 + tci;
    const int cb = ci >> 4, e = ci & 15;
    const int g = cb / kGroupBlocks, jj = cb - g * kGroupBlocks;
    if (co < J.co_pad) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int py = u >> 3, u8 = u & 7;                 // region per output-row parity; slab (px, syi, sxi) inside it
        const size_t byte = ((size_t)((u8 * J.ngroups + g) * kGroupBlocks + jj) * J.co_pad + co) * 32 +
                            ((((e >> 3) ^ ((co >> 2) & 1))) << 4) + (e & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>((char*)J.dst[py] + byte) = __float2bfloat16_rn(acc[u]);
      }
    }
    if (J.dst_t && ci < J.ci_rows) {
      // transposed pack for the fused input gradient: row = ci, contraction index = co
      const int cbo = co >> 4, eo = co & 15;
      const int go = cbo / kGroupBlocks, jo = cbo - go * kGroupBlocks;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const size_t byte = ((size_t)((u * J.gco + go) * kGroupBlocks + jo) * J.ci_rows + ci) * 32 +
                            ((((eo >> 3) ^ ((ci >> 2) & 1))) << 4) + (eo & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>((char*)J.dst_t + byte) = __float2bfloat16_rn(acc[u]);
      }
    }
  }
  if (blockIdx.x == gridDim.x - 1) {
    for (int co = threadIdx.x; co < J.co_pad; co += blockDim.x) {
      float tap[9];
      float full = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) tap[t] = 0.f;
      if (co < J.Co)
#pragma unroll 4
        for (int c = 0; c < J.Cu; ++c) {                  // the nine taps of (co, c) are contiguous
          const float bdc = J.bd[c];
          const float* w = J.w3 + ((long long)co * cin3 + c) * 9;
#pragma unroll
          for (int t = 0; t < 9; ++t) tap[t] = fmaf(__ldg(w + t), bdc, tap[t]);
        }
#pragma unroll
      for (int t = 0; t < 9; ++t) full += tap[t];
      J.bias_full[co] = co < J.Co ? J.b3[co] + full : 0.f;
      for (int cls = 0; cls < 9; ++cls) {
        const int yc = cls / 3, xc = cls % 3;
        float s = 0.f;
        for (int t = 0; t < 9; ++t) {
          const int ky = t / 3, kx = t % 3;
          const bool out = (yc == 1 && ky == 0) || (yc == 2 && ky == 2) || (xc == 1 && kx == 0) || (xc == 2 && kx == 2);
          if (out) s += tap[t];
        }
        J.corr[cls * J.co_pad + co] = s;
      }
    }
  }
}

int launch_upfuse_pack(const UpFuseJob* jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return 0;
  N2N_CHECK_ARG(njobs <= 5, "upfuse_pack: too many jobs");
  UpFuseBatch b;
  b.n = njobs;
  int maxtiles = 1;
  for (int i = 0; i < njobs; ++i) {
    b.j[i] = jobs[i];
    const int co_span = (jobs[i].dst_t && jobs[i].gco * kGroupBlocks * 16 > jobs[i].co_pad) ? jobs[i].gco * kGroupBlocks * 16 : jobs[i].co_pad;
    const int tiles = (co_span / kUfTile) * (jobs[i].ngroups * kGroupBlocks * 16 / kUfTile);
    if (tiles > maxtiles) maxtiles = tiles;
  }
  dim3 grid(maxtiles + 1, njobs);            // + one block per layer for the bias / border tables
  (void)launch_pdl_v(upfuse_pack_kernel, grid, dim3(kUfTile * kUfTile), 0, st, b);
  N2N_LAUNCH_CHECK();
  return 0;
}

// ---- backward of the fused up-conv: chain rule (see UpFuseGradJob) -----------------------------------------------
// Wc[u][co][ci] = sum over the taps (ky,kx) landing on source offset (sy,sx) of W3[co][c][ky][kx] Wd[ci][c][a][b], so
//   dW3[co][c][ky][kx] = sum_{py,px} sum_ci dWc[u(py,px,ky,kx)][ci][co] Wd[ci][c][a][b]  +  S_{ky,kx}[co] bd[c]
//   dWd[ci][c][a][b]   = sum_{(py,ky): a, (px,kx): b} sum_co dWc[u][ci][co] W3[co][c][ky][kx]
//   dbd[c]             = sum_{ky,kx} sum_co W3[co][c][ky][kx] S_{ky,kx}[co]
// S_t[co] = sum of dL/dy[co] over the output pixels whose tap t lies inside the upsampled image (the 3x3 conv zero-pads
// it): total minus the excluded border row / column plus the doubly excluded corner.
struct UpFuseGradBatch { UpFuseGradJob j[5]; int n; };

__device__ __forceinline__ float upfuse_S(const UpFuseGradJob& J, int ky, int kx, int co) {
  const float* B = J.border;
  float s = J.db3[co];
  if (ky == 0) s -= B[0 * J.co_pad + co];
  if (ky == 2) s -= B[1 * J.co_pad + co];
  if (kx == 0) s -= B[2 * J.co_pad + co];
  if (kx == 2) s -= B[3 * J.co_pad + co];
  if (ky != 1 && kx != 1) s += B[(4 + (ky >> 1) * 2 + (kx >> 1)) * J.co_pad + co];
  return s;
}

// Tiled like the pack kernel (a one-thread-per-output version spent 470 us per step on strided global reads):
//   blocks [0, T3):      (16 co x 16 c) tiles of dW3 — all nine taps per thread, reduction over ci in chunks of 16
//   blocks [T3, T3+Td):  (16 ci x 16 c) tiles of dWd — all four (a, b) per thread, reduction over co in chunks of 16
//   last block:          dbd
__global__ void __launch_bounds__(kUfTile * kUfTile)
upfuse_grad_kernel(const __grid_constant__ UpFuseGradBatch b) {
  pdl_enter();
  const UpFuseGradJob& J = b.j[blockIdx.y];
  __shared__ float sD[kUfTile][16][kUfTile + 1];     // [reduction index][composite u][co or ci]
  __shared__ float sW[kUfTile][9][kUfTile];          // [reduction index][tap or ab][c]
  const int cin3 = J.Cu + J.Cs;
  const long long plane = (long long)J.ci_pad * J.co_pad;
  const int tco = (J.Co + kUfTile - 1) / kUfTile, tc = (J.Cu + kUfTile - 1) / kUfTile, tci = (J.Ci + kUfTile - 1) / kUfTile;
  const int T3 = tco * tc, Td = tci * tc;
  const int tr = threadIdx.x / kUfTile, tcx = threadIdx.x % kUfTile;      // tile row (co or ci), tile column (c)
  if ((int)blockIdx.x < T3) {
    const int co0 = ((int)blockIdx.x / tc) * kUfTile, c0 = ((int)blockIdx.x % tc) * kUfTile;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int i0 = 0; i0 < J.Ci; i0 += kUfTile) {
      __syncthreads();
      for (int i = threadIdx.x; i < kUfTile * 16 * kUfTile; i += kUfTile * kUfTile) {
        const int co = i % kUfTile, u = (i / kUfTile) % 16, ci = i / (kUfTile * 16);
        sD[ci][u][co] = (i0 + ci < J.Ci && co0 + co < J.Co) ? __ldg(J.dwc + u * plane + (long long)(i0 + ci) * J.co_pad + co0 + co) : 0.f;
      }
      for (int i = threadIdx.x; i < kUfTile * kUfTile * 4; i += kUfTile * kUfTile) {
        const int ab = i % 4, c = (i / 4) % kUfTile, ci = i / (4 * kUfTile);
        sW[ci][ab][c] = (i0 + ci < J.Ci && c0 + c < J.Cu) ? __ldg(J.wd + ((long long)(i0 + ci) * J.Cu + c0 + c) * 4 + ab) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int ci = 0; ci < kUfTile; ++ci) {
        float d[16], w[4];
#pragma unroll
        for (int u = 0; u < 16; ++u) d[u] = sD[ci][u][tr];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = sW[ci][q][tcx];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int yo = py + ky - 1, xo = px + kx - 1;
                const int syi = (yo >= 0 ? yo >> 1 : -1) - (py - 1), sxi = (xo >= 0 ? xo >> 1 : -1) - (px - 1);
                acc[ky * 3 + kx] = fmaf(d[((py * 2 + px) * 2 + syi) * 2 + sxi], w[(yo & 1) * 2 + (xo & 1)], acc[ky * 3 + kx]);
              }
      }
    }
    const int co = co0 + tr, c = c0 + tcx;
    if (co < J.Co && c < J.Cu) {
      const float bdc = J.bd[c];
#pragma unroll
      for (int t = 0; t < 9; ++t) J.dw3[((long long)co * cin3 + c) * 9 + t] = acc[t] + upfuse_S(J, t / 3, t % 3, co) * bdc;
    }
  } else if ((int)blockIdx.x < T3 + Td) {
    const int bi = (int)blockIdx.x - T3;
    const int ci0 = (bi / tc) * kUfTile, c0 = (bi % tc) * kUfTile;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int o0 = 0; o0 < J.Co; o0 += kUfTile) {
      __syncthreads();
      for (int i = threadIdx.x; i < kUfTile * 16 * kUfTile; i += kUfTile * kUfTile) {
        const int co = i % kUfTile, ci = (i / kUfTile) % kUfTile, u = i / (kUfTile * kUfTile);      // co fastest: coalesced
        sD[co][u][ci] = (o0 + co < J.Co && ci0 + ci < J.Ci) ? __ldg(J.dwc + u * plane + (long long)(ci0 + ci) * J.co_pad + o0 + co) : 0.f;
      }
      for (int i = threadIdx.x; i < kUfTile * kUfTile * 9; i += kUfTile * kUfTile) {
        const int t = i % 9, c = (i / 9) % kUfTile, co = i / (9 * kUfTile);
        sW[co][t][c] = (o0 + co < J.Co && c0 + c < J.Cu) ? __ldg(J.w3 + ((long long)(o0 + co) * cin3 + c0 + c) * 9 + t) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int co = 0; co < kUfTile; ++co) {
        float d[16], w[9];
#pragma unroll
        for (int u = 0; u < 16; ++u) d[u] = sD[co][u][tr];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = sW[co][t][tcx];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int yo = py + ky - 1, xo = px + kx - 1;
                const int syi = (yo >= 0 ? yo >> 1 : -1) - (py - 1), sxi = (xo >= 0 ? xo >> 1 : -1) - (px - 1);
                acc[(yo & 1) * 2 + (xo & 1)] = fmaf(d[((py * 2 + px) * 2 + syi) * 2 + sxi], w[ky * 3 + kx], acc[(yo & 1) * 2 + (xo & 1)]);
              }
      }
    }
    const int ci = ci0 + tr, c = c0 + tcx;
    if (ci < J.Ci && c < J.Cu) {
#pragma unroll
      for (int ab = 0; ab < 4; ++ab) J.dwd[((long long)ci * J.Cu + c) * 4 + ab] = acc[ab];
    }
  } else if ((int)blockIdx.x == T3 + Td) {
    for (int c = threadIdx.x; c < J.Cu; c += blockDim.x) {
      float acc = 0.f;
      for (int co = 0; co < J.Co; ++co)
        for (int t = 0; t < 9; ++t) acc = fmaf(J.w3[((long long)co * cin3 + c) * 9 + t], upfuse_S(J, t / 3, t % 3, co), acc);
      J.dbd[c] = acc;
    }
  }
}

int launch_upfuse_grad(const UpFuseGradJob* jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return 0;
  N2N_CHECK_ARG(njobs <= 5, "upfuse_grad: too many jobs");
  UpFuseGradBatch b;
  b.n = njobs;
  int maxblocks = 1;
  for (int i = 0; i < njobs; ++i) {
    b.j[i] = jobs[i];
    const int tco = (jobs[i].Co + kUfTile - 1) / kUfTile, tc = (jobs[i].Cu + kUfTile - 1) / kUfTile, tci = (jobs[i].Ci + kUfTile - 1) / kUfTile;
    const int blocks = tco * tc + tci * tc + 1;
    if (blocks > maxblocks) maxblocks = blocks;
  }
  (void)launch_pdl_v(upfuse_grad_kernel, dim3(maxblocks, njobs), dim3(kUfTile * kUfTile), 0, st, b);
  N2N_LAUNCH_CHECK();
  return 0;
}

// border[k][co_pad], k = 0: first row, 1: last row, 2: first column, 3: last column, 4..7: corners (top-left, top-right,
// bottom-left, bottom-right) of a C16 bf16 tensor, summed over the batch.  One block per (k, channel block): fixed-order sums.
struct BorderBatch { View g[5]; float* out[5]; int co_pad[5]; int n; };

__global__ void __launch_bounds__(256) border_sums_kernel(const __grid_constant__ BorderBatch b) {
  pdl_enter();
  __shared__ float red[8][16];
  const View& g = b.g[blockIdx.z];
  const int k = blockIdx.x, cb = blockIdx.y;
  if (cb >= g.Cb) return;
  const int H = g.H, W = g.W;
  const long long count = k < 2 ? (long long)g.N * W : (k < 4 ? (long long)g.N * H : g.N);
  float acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = 0.f;
  const __nv_bfloat16* base = (const __nv_bfloat16*)g.ptr + (long long)cb * g.sCb;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) {
    int n, y, x;
    if (k < 2) { n = (int)(i / W); x = (int)(i - (long long)n * W); y = k == 0 ? 0 : H - 1; }
    else if (k < 4) { n = (int)(i / H); y = (int)(i - (long long)n * H); x = k == 2 ? 0 : W - 1; }
    else { n = (int)i; y = (k - 4) >> 1 ? H - 1 : 0; x = (k - 4) & 1 ? W - 1 : 0; }
    float v[16];
    Block16<__nv_bfloat16>::load(base + (long long)n * g.sN + (long long)y * g.sY + (long long)x * g.sX, v);
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] += v[q];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    float v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    b.out[blockIdx.z][(long long)k * b.co_pad[blockIdx.z] + cb * 16 + threadIdx.x] = s;
  }
}

// one launch for up to five tensors (the fused levels of a backward pass)
int launch_border_sums(const View* g, float* const* border, const int* co_pad, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  N2N_CHECK_ARG(n <= 5, "border_sums: too many tensors");
  BorderBatch b;
  b.n = n;
  int maxcb = 1;
  for (int i = 0; i < n; ++i) {
    N2N_CHECK_ARG(g[i].Cb * 16 <= co_pad[i], "border_sums: bad channel count");
    b.g[i] = g[i]; b.out[i] = border[i]; b.co_pad[i] = co_pad[i];
    if (g[i].Cb > maxcb) maxcb = g[i].Cb;
  }
  (void)launch_pdl_v(border_sums_kernel, dim3(8, maxcb, n), dim3(256), 0, st, b);
  N2N_LAUNCH_CHECK();
  return 0;
}

int launch_pack(const PackJob* jobs, int njobs, int dtype, cudaStream_t st) {
  for (int base = 0; base < njobs; base += 8) {
    PackBatch b;
    b.n = njobs - base < 8 ? njobs - base : 8;
    long long maxtotal = 1;
    for (int i = 0; i < b.n; ++i) {
      b.j[i] = jobs[base + i];
      long long tot = (long long)b.j[i].ntaps * b.j[i].nout_pad * b.j[i].cin_blocks * 16;
      if (tot > maxtotal) maxtotal = tot;
    }
    dim3 grid(grid_for(maxtotal, 256, 2), b.n);
    if (dtype == N2N_BF16) (void)launch_pdl_v(pack_kernel<true>, grid, dim3(256), 0, st, b);
    else (void)launch_pdl_v(pack_kernel<false>, grid, dim3(256), 0, st, b);
    N2N_LAUNCH_CHECK();
  }
  return 0;
}

int launch_unpack(const UnpackJob* jobs, int njobs, cudaStream_t st) {
  for (int base = 0; base < njobs; base += 16) {
    UnpackBatch b;
    b.n = njobs - base < 16 ? njobs - base : 16;
    long long maxtotal = 1;
    for (int i = 0; i < b.n; ++i) {
      b.j[i] = jobs[base + i];
      long long tot = (long long)b.j[i].ntaps * b.j[i].npad * b.j[i].cpad / 4 * kUnpackLanes;
      if (tot > maxtotal) maxtotal = tot;
    }
    dim3 grid(grid_for(maxtotal, 256, 8), b.n);
    (void)launch_pdl_v(unpack_kernel, grid, dim3(256), 0, st, b);
    N2N_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace n2n
