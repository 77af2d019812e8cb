// wgrad_umma.cu — bf16 tensor-core engine for the weight-gradient GEMM
//   P[t][c][n] = sum_p dY_a(t)[p, n] * X_b(t)[p + (dy_t, dx_t), c]
// i.e. a GEMM whose K dimension is the PIXEL axis.  Both operands come straight from the C16
// activation tensors through the same 5-D TMA boxes as the forward engine ([block][pixel][16 ch]
// tiles, SWIZZLE_32B); because the contraction runs over the tile's rows, the tiles are consumed
// as MN-major UMMA operands (channels contiguous, pixels strided) — no transposes anywhere.
//
// One CTA owns a contiguous range of 128-pixel chunks (split-K over pixels) and a group of taps;
// for each chunk it loads the "common" operand once (conv: dY; deconv: X) and one "variant" tile
// per tap (conv: X shifted by the tap; deconv: the dY parity view), and accumulates one
// [128 x 16*n_blocks] fp32 tile per tap in TMEM (up to 512 columns).  The epilogue stores the
// per-split partial sums; pack.cu's unpack kernel reduces them in a fixed order (deterministic).
#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kCStages = 2;
constexpr int kThreadsW = 192;

struct UmmaWgradParams {
  CUtensorMap tmap_common;
  CUtensorMap tmap_var[4];
  int npairs, pairs_per_cta;
  int8_t pair_view[9], pair_dy[9], pair_dx[9];
  int m_blocks, n_blocks;       // common / variant operand widths in 16-channel blocks
  int swap;                     // 0: rows = dY channel (n), cols = X channel (c); 1: rows = c, cols = n
  float* partial;
  int npad, cpad;
  int bw, bh, rows, tiles_x, tiles_y;
  long long chunks, chunks_per_split;
  int vstages;
  uint32_t c_slot_bytes, v_slot_bytes, tmem_cols, idesc;
};

__global__ void __launch_bounds__(kThreadsW)
tapwgrad_umma_kernel(const __grid_constant__ UmmaWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * kCStages + 2 * 4 + 1];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t v_ring0 = smem0 + kCStages * p.c_slot_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto cfull = [&](int s) { return bar0 + 8u * s; };
  auto cempty = [&](int s) { return bar0 + 8u * (kCStages + s); };
  auto vfull = [&](int s) { return bar0 + 8u * (2 * kCStages + s); };
  auto vempty = [&](int s) { return bar0 + 8u * (2 * kCStages + 4 + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * kCStages + 8);

  const int split = blockIdx.x, tgroup = blockIdx.y;
  const int pair0 = tgroup * p.pairs_per_cta;
  int npair = p.npairs - pair0;
  if (npair > p.pairs_per_cta) npair = p.pairs_per_cta;
  const long long chunk_begin = (long long)split * p.chunks_per_split;
  long long chunk_end = chunk_begin + p.chunks_per_split;
  if (chunk_end > p.chunks) chunk_end = p.chunks;
  const int ncols_pair = p.n_blocks * 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCStages; ++s) { mbar_init(cfull(s), 1); mbar_init(cempty(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(vfull(s), 1); mbar_init(vempty(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const bool has_work = chunk_end > chunk_begin;
  pdl_wait();
  if (threadIdx.x == 32) pdl_release();

  if (warp == 0) {
    if (lane == 0 && has_work) {
      prefetch_tensormap(&p.tmap_common);
      for (int v = 0; v < 4; ++v) prefetch_tensormap(&p.tmap_var[v]);
      int cs = 0, vs = 0; uint32_t cph = 0, vph = 0;
      for (long long ch = chunk_begin; ch < chunk_end; ++ch) {
        long long r = ch;
        const int tx = (int)(r % p.tiles_x); r /= p.tiles_x;
        const int ty = (int)(r % p.tiles_y);
        const int img = (int)(r / p.tiles_y);
        const int x0 = tx * p.bw, y0 = ty * p.bh;
        mbar_wait(cempty(cs), cph ^ 1u);
        mbar_arrive_expect_tx(cfull(cs), (uint32_t)(p.m_blocks * p.rows * 32));
        tma_load_5d(smem0 + cs * p.c_slot_bytes, &p.tmap_common, cfull(cs), 0, x0, y0, 0, img);
        if (++cs == kCStages) { cs = 0; cph ^= 1u; }
        for (int pi = 0; pi < npair; ++pi) {
          const int t = pair0 + pi;
          mbar_wait(vempty(vs), vph ^ 1u);
          mbar_arrive_expect_tx(vfull(vs), (uint32_t)(p.n_blocks * p.rows * 32));
          tma_load_5d(v_ring0 + vs * p.v_slot_bytes, &p.tmap_var[p.pair_view[t]], vfull(vs), 0, x0 + p.pair_dx[t],
                      y0 + p.pair_dy[t], 0, img);
          if (++vs == p.vstages) { vs = 0; vph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && has_work) {
      int cs = 0, vs = 0; uint32_t cph = 0, vph = 0;
      const uint32_t atom = (uint32_t)p.rows * 32u;       // bytes between 16-channel blocks of a tile
      const int ksteps = p.rows / 16;
      for (long long ch = chunk_begin; ch < chunk_end; ++ch) {
        mbar_wait(cfull(cs), cph);
        fence_after_sync();
        const uint32_t a_base = smem0 + cs * p.c_slot_bytes;
        for (int pi = 0; pi < npair; ++pi) {
          mbar_wait(vfull(vs), vph);
          fence_after_sync();
          const uint32_t b_base = v_ring0 + vs * p.v_slot_bytes;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t ad = make_smem_desc(a_base + kk * 512u, atom, 256, kSwizzle32);
            const uint64_t bd = make_smem_desc(b_base + kk * 512u, atom, 256, kSwizzle32);
            mma_bf16(tmem_base + pi * ncols_pair, ad, bd, p.idesc, (ch != chunk_begin) || (kk != 0));
          }
          mma_commit(vempty(vs));
          if (++vs == p.vstages) { vs = 0; vph ^= 1u; }
        }
        mma_commit(cempty(cs));
        if (++cs == kCStages) { cs = 0; cph ^= 1u; }
      }
      mma_commit(tmem_full_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    if (has_work) {
      mbar_wait(tmem_full_bar, 0);
      fence_after_sync();
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool row_ok = m < p.m_blocks * 16;
    for (int pi = 0; pi < npair; ++pi) {
      const int t = pair0 + pi;
      float* P = p.partial + ((long long)split * p.npairs + t) * p.cpad * p.npad;
      for (int cb = 0; cb < p.n_blocks; ++cb) {
        float v[16];
        if (has_work) {
          tmem_ld16(lane_addr + pi * ncols_pair + cb * 16, v);
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0.f;
        }
        if (row_ok) {
          if (p.swap == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) P[(long long)(cb * 16 + q) * p.npad + m] = v[q];
          } else {
            float4* dst = reinterpret_cast<float4*>(P + (long long)m * p.npad + cb * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// Bias gradient: bias_partial[row = split * ndyviews + view][n] = sum over the split's pixel range
// of dY_view[p, n].  One block per (row, 16-channel block): consecutive threads read consecutive
// pixels (32 B each -> 1 KB per warp, fully coalesced), four independent 256-bit loads in flight
// per thread, 32-bit index arithmetic with no division on dense views; HBM-bound single pass over dY.
__device__ __forceinline__ void bg_accum(const __nv_bfloat16* p, float s[16]) {
  uint32_t w[8];
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[2 * j] += __uint_as_float(w[j] << 16);
    s[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(256)
bias_grad_kernel(View dy0, View dy1, View dy2, View dy3, int ndyviews, long long pixels, long long per_split,
                 float* __restrict__ bias_partial, int npad) {
  pdl_enter();
  __shared__ float red[8][16];
  const int cb = blockIdx.y;
  const int split = blockIdx.x / ndyviews, vi = blockIdx.x - split * ndyviews;
  const View& dv = vi == 0 ? dy0 : vi == 1 ? dy1 : vi == 2 ? dy2 : dy3;
  const long long p0 = (long long)split * per_split;
  long long p1 = p0 + per_split;
  if (p1 > pixels) p1 = pixels;
  float s[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) s[q] = 0.f;
  const int hw = dv.W * dv.H;
  const bool dense = dv.sY == (long long)dv.W * dv.sX;
  // walk the images the split touches; inside an image the pixel index is a 32-bit offset
  for (long long pbase = p0; pbase < p1;) {
    const int img = (int)(pbase / hw);
    const int r0 = (int)(pbase - (long long)img * hw);
    long long pend = (long long)(img + 1) * hw;
    if (pend > p1) pend = p1;
    const int r1 = r0 + (int)(pend - pbase);
    const __nv_bfloat16* base = (const __nv_bfloat16*)dv.ptr + (long long)img * dv.sN + (long long)cb * dv.sCb;
    if (dense) {
      const int sx = (int)dv.sX;
      int r = r0 + threadIdx.x;
      for (; r + 768 < r1; r += 1024) {
        bg_accum(base + (long long)r * sx, s);
        bg_accum(base + (long long)(r + 256) * sx, s);
        bg_accum(base + (long long)(r + 512) * sx, s);
        bg_accum(base + (long long)(r + 768) * sx, s);
      }
      for (; r < r1; r += 256) bg_accum(base + (long long)r * sx, s);
    } else {
      for (int r = r0 + threadIdx.x; r < r1; r += 256) {
        const int y = r / dv.W, x = r - y * dv.W;
        bg_accum(base + (long long)y * dv.sY + (long long)x * dv.sX, s);
      }
    }
    pbase = pend;
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 16; ++q) red[warp][q] = s[q];
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    bias_partial[(long long)blockIdx.x * npad + cb * 16 + threadIdx.x] = t;
  }
}

void choose_tile(int H, int W, int& bw, int& bh);

// HBM-bound bias-gradient pass shared by both weight-gradient engines.
int launch_bias_grad(const TapWgrad& g, cudaStream_t st) {
  if (!g.bias_partial) return 0;
  const int splits = g.splits;
  const long long pixels = (long long)g.dy[0].N * g.dy[0].H * g.dy[0].W;
  const long long per_split = (pixels + splits - 1) / splits;
  dim3 bgrid(splits * g.ndyviews, g.n_blocks);
  (void)launch_pdl_v(bias_grad_kernel, bgrid, dim3(256), 0, st, g.dy[0], g.dy[1], g.dy[2], g.dy[3], g.ndyviews, pixels, per_split,
                     g.bias_partial, g.n_blocks * 16);
  N2N_LAUNCH_CHECK();
  return 0;
}

int launch_tapwgrad_umma(const TapWgrad& g, cudaStream_t st) {
  static bool attr_set = false;
  // conv: common = dY (view 0), variants = X shifted; deconv (ndyviews == 4): common = X, variants = dY parity views
  const bool swap = g.ndyviews > 1;
  const View& common = swap ? g.x[0] : g.dy[0];
  const int m_blocks = swap ? g.c_blocks : g.n_blocks;
  const int n_blocks = swap ? g.n_blocks : g.c_blocks;
  N2N_CHECK_ARG(m_blocks >= 1 && m_blocks <= 8, "tapwgrad_umma: common operand has %d blocks (max 8)", m_blocks);
  N2N_CHECK_ARG(n_blocks >= 1 && n_blocks <= 16, "tapwgrad_umma: variant operand has %d blocks (max 16)", n_blocks);
  UmmaWgradParams p;
  memset(&p, 0, sizeof(p));
  const int H = common.H, W = common.W;
  int bw, bh;
  choose_tile(H, W, bw, bh);
  if (bh > H) bh = H;
  p.bw = bw; p.bh = bh; p.rows = bw * bh;
  N2N_CHECK_ARG(p.rows % 16 == 0, "tapwgrad_umma: tile of %d pixels is not a multiple of 16", p.rows);
  p.tiles_x = (W + bw - 1) / bw; p.tiles_y = (H + bh - 1) / bh;
  p.chunks = (long long)common.N * p.tiles_x * p.tiles_y;
  int splits = g.splits;
  p.chunks_per_split = (p.chunks + splits - 1) / splits;
  p.npairs = g.npairs;
  p.pairs_per_cta = 512 / (n_blocks * 16);
  if (p.pairs_per_cta > g.npairs) p.pairs_per_cta = g.npairs;
  const int tgroups = (g.npairs + p.pairs_per_cta - 1) / p.pairs_per_cta;
  p.m_blocks = m_blocks; p.n_blocks = n_blocks; p.swap = swap ? 1 : 0;
  p.partial = g.partial; p.npad = g.n_blocks * 16; p.cpad = g.c_blocks * 16;
  N2N_TRY(encode_c16_tensor_map(&p.tmap_common, common, bw, bh, m_blocks));
  int nvar = 0;
  for (int t = 0; t < g.npairs; ++t) {
    const int vi = swap ? g.pair_dyv[t] : g.pair_xv[t];
    p.pair_view[t] = (int8_t)vi;
    p.pair_dy[t] = (int8_t)(swap ? 0 : g.pair_dy[t]);
    p.pair_dx[t] = (int8_t)(swap ? 0 : g.pair_dx[t]);
    if (vi + 1 > nvar) nvar = vi + 1;
  }
  for (int v = 0; v < 4; ++v) {
    const View& vv = swap ? g.dy[v < nvar ? v : 0] : g.x[v < nvar ? v : 0];
    N2N_CHECK_ARG(vv.H == H && vv.W == W, "tapwgrad_umma: variant view %d geometry mismatch", v);
    N2N_TRY(encode_c16_tensor_map(&p.tmap_var[v], vv, bw, bh, n_blocks));
  }
  p.c_slot_bytes = (uint32_t)align_up((size_t)m_blocks * p.rows * 32, 1024);
  p.v_slot_bytes = (uint32_t)align_up((size_t)n_blocks * p.rows * 32, 1024);
  const size_t budget = 200 * 1024;
  int vstages = (int)((budget - 1024 - kCStages * p.c_slot_bytes) / p.v_slot_bytes);
  if (vstages > 4) vstages = 4;
  N2N_CHECK_ARG(vstages >= 2, "tapwgrad_umma: operands too wide for shared memory");
  p.vstages = vstages;
  p.tmem_cols = tmem_cols_for(p.pairs_per_cta * n_blocks * 16);
  p.idesc = make_idesc_bf16(128, n_blocks * 16, true, true);
  size_t smem = 1024 + (size_t)kCStages * p.c_slot_bytes + (size_t)vstages * p.v_slot_bytes;
  // the M=128 MMA always walks 8 channel blocks of the common tile; keep that window inside the allocation
  const size_t window = 1024 + (size_t)(kCStages - 1) * p.c_slot_bytes + 8u * p.rows * 32 + 1024;
  if (smem < window) smem = window;
  if (!attr_set) {
    N2N_CUDA(cudaFuncSetAttribute(tapwgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  dim3 grid(splits, tgroups);
  N2N_CUDA(launch_pdl(tapwgrad_umma_kernel, grid, dim3(kThreadsW), smem, st, p));
  N2N_LAUNCH_CHECK();
  return launch_bias_grad(g, st);
}

}  // namespace n2n
