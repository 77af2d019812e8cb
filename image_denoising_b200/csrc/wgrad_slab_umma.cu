// wgrad_slab_umma.cu — second-generation bf16 tensor-core engine for the weight-gradient GEMM
//   P[t][c][n] = sum_p dY_a(t)[p, n] * X_b(t)[p + (dy_t, dx_t), c]          (K = the pixel axis)
// in the 8 x 16-pixel tile geometry of slabgemm_umma.cu.
//
// Per tile a CTA stages ONE "common" box (conv: the dY tile, 8 x 16 pixels; deconv: the X tile) and
// the "variant" boxes of its tap group (conv: the halo'd 10 x 18 X window — all nine taps are
// operand views of it, tap (dy, dx) starting (10*dy + dx) * 32 B further in; deconv: one dY parity
// tile per tap).  Both operands are consumed MN-major straight from the C16 boxes: the 16 channels
// of a block are the contiguous M/N run (leading-byte-offset = one block of the box), the eight
// pixels of an image row are the eight K rows of a SWIZZLE_32B group and the stride-byte-offset
// (256 B for a tile, 320 B for a halo'd window) walks the image rows; one tcgen05.mma covers
// K = 16 pixels = two image rows, so a tile is 8 K-steps per tap.
//
// Grid = (pixel splits) x (tap groups): a CTA owns every splits-th tile and as many taps as
// fit in TMEM (512 / (16 * variant blocks) accumulators of [128 x N] fp32); it accumulates over its
// whole tile range in TMEM and stores its fp32 partial once at the end (pack.cu's unpack kernel
// reduces the splits in a fixed order -> deterministic).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kWsThreads = 192;
constexpr int kWsMaxRing = 8;
constexpr int kWsTileW = 8, kWsTileH = 16;
constexpr size_t kWsSmemMax = 232448;
constexpr size_t kWsStaticSlack = 4096;

struct WsParams {
  int npairs, taps_per_cta, tgroups;
  int m_blocks, n_blocks, swap;
  int npad, cpad;
  int tiles_x, tiles_y;
  long long tiles; int nsplits;
  int ring, halo, box_per_tap;           // box_per_tap: 1 = every tap has its own variant box (deconv), 0 = one shared box
  uint32_t slot_bytes, a_bytes, var_box_bytes, tmem_cols, idesc;
  uint32_t a_lbo, a_hi, a_kstep16;       // LBO field (already << 16) of the A descriptor low word; high word; K-step (bytes/16)
  uint32_t b_lbo, b_hi, b_kstep16;
  float* partial;
  float* bias_partial;                   // fused bias gradient: conv [splits][npad] (tap group 0 sums the dY tile); deconv
                                         // [splits * npairs][npad] (every tap group sums its dY parity boxes into its first row)
  int dbg_flags;                         // N2N_DBG_FLAGS: 1 = skip the loads, 2 = skip the MMAs, 4 = skip the bias sums
  uint16_t tap_off16[12];                // start of each tap's view inside its variant box (bytes/16)
  int8_t tap_view[12];                   // tensor map of the tap's variant view
  CUtensorMap tmap_common;
  CUtensorMap tmap_var[4];
};

__device__ __forceinline__ void ws_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ void ws_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate));
}

// All MMAs of one tile for NT taps: 8 K-steps (two image rows each) x NT accumulators.
template <int NT>
__device__ __forceinline__ void ws_issue(uint32_t d_tmem, uint32_t ncols, uint32_t a_lo0, uint32_t b_lo0, uint32_t a_kstep16,
                                         uint32_t b_kstep16, uint32_t a_hi, uint32_t b_hi, uint32_t idesc, uint32_t acc,
                                         const uint32_t (&boff)[10]) {
#pragma unroll 1
  for (int kk = 0; kk < 8; ++kk) {
    const uint32_t a_lo = a_lo0 + kk * a_kstep16;
    const uint32_t b_lok = b_lo0 + kk * b_kstep16;
    const uint32_t a1 = (acc | (uint32_t)kk) ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < NT; ++i) ws_mma(d_tmem + (uint32_t)i * ncols, a_lo, a_hi, b_lok + boff[i], b_hi, idesc, a1);
  }
}

__global__ void __launch_bounds__(kWsThreads, 1)
wgrad_slab_umma_kernel(const __grid_constant__ WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * kWsMaxRing + 1];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_red[4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kWsMaxRing + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * kWsMaxRing);

  const int split = blockIdx.x / p.tgroups, tg = blockIdx.x - split * p.tgroups;
  const int tap0 = tg * p.taps_per_cta;
  int ntap = p.npairs - tap0;
  if (ntap > p.taps_per_cta) ntap = p.taps_per_cta;
  // split s owns tiles s, s + splits, s + 2*splits, ...: at any moment the CTAs read neighbouring tiles (adjacent
  // 256-byte row segments of the same DRAM pages), not 148 far-apart streams
  const int tile_begin = split, tile_end = (int)p.tiles, tile_step = p.nsplits;
  const bool has_work = tile_end > tile_begin;
  const int ncols = p.n_blocks * 16;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int nbox = p.box_per_tap ? ntap : 1;
  // bias gradient = per-channel pixel sum of dY: the four otherwise idle epilogue warps add it up from
  // the dY tile the pipeline already staged in shared memory (tap group 0 only), so dY is not read twice
  const bool do_bias = p.bias_partial != nullptr && (p.swap || tg == 0) && !(p.dbg_flags & 32);
  const int bias_blocks = p.swap ? p.n_blocks : p.m_blocks;      // channel blocks of dY

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWsMaxRing; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), do_bias ? 5 : 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                           // everything below reads the predecessor's tensors or writes the partials
  if (threadIdx.x == 32) pdl_release();

  if (warp == 0) {
    // ---- TMA producer ----
    if (elect_one_sync()) {
      prefetch_tensormap(&p.tmap_common);
      for (int v = 0; v < 4; ++v) prefetch_tensormap(&p.tmap_var[v]);
    }
    __syncwarp();
    const uint32_t tx = p.a_bytes + (uint32_t)nbox * p.var_box_bytes;
    const int org = p.halo ? -1 : 0;
    int slot = 0; uint32_t phase = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      int img, r, ty, txi;
      if (p.dbg_flags & 8) { img = 0; r = 0; ty = 0; txi = 0; }
      else {
        img = tile / tiles_per_img;
        r = tile - img * tiles_per_img;
        ty = r / p.tiles_x; txi = r - ty * p.tiles_x;
      }
      const int x0 = txi * kWsTileW, y0 = ty * kWsTileH;
      ws_wait(empty_bar(slot), phase ^ 1u);
      if (p.dbg_flags & 1) {            // diagnostic: no loads, the pipeline runs on stale shared memory
        if (elect_one_sync()) mbar_arrive(full_bar(slot));
      } else if (elect_one_sync()) {
        const uint32_t dst = smem0 + slot * p.slot_bytes;
        mbar_arrive_expect_tx(full_bar(slot), tx);
        tma_load_5d(dst, &p.tmap_common, full_bar(slot), 0, x0, y0, 0, img);
        for (int i = 0; i < nbox; ++i)
          tma_load_5d(dst + p.a_bytes + i * p.var_box_bytes, &p.tmap_var[p.tap_view[tap0 + i]], full_bar(slot), 0,
                      x0 + org, y0 + org, 0, img);
      }
      __syncwarp();
      if (++slot == p.ring) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    if (has_work) {
      // per-tap B start offsets (16-byte units, relative to the slot's variant area) in registers
      uint32_t boff[10];
#pragma unroll
      for (int i = 0; i < 10; ++i)
        boff[i] = i < ntap ? (uint32_t)p.tap_off16[tap0 + i] + (p.box_per_tap ? (uint32_t)i * (p.var_box_bytes >> 4) : 0u) : 0u;
      const uint32_t a_hi = p.a_hi, b_hi = p.b_hi, idesc = p.idesc;
      int slot = 0; uint32_t phase = 0;
      uint32_t acc = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const uint32_t s0 = smem0 + slot * p.slot_bytes;
        const uint32_t a_lo0 = ((s0 & 0x3FFFFu) >> 4) | p.a_lbo;
        const uint32_t b_lo0 = (((s0 + p.a_bytes) & 0x3FFFFu) >> 4) | p.b_lbo;
        ws_wait(full_bar(slot), phase);
        fence_after_sync();
        if (elect_one_sync()) {
          if (!(p.dbg_flags & 2)) {
            const uint32_t d0 = tmem_base, nc = (uint32_t)ncols;
            // compile-time tap count: the issuing lane runs exactly 8 * ntap MMAs and nothing else
            switch (ntap) {
              case 1: ws_issue<1>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 2: ws_issue<2>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 3: ws_issue<3>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 4: ws_issue<4>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 5: ws_issue<5>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 6: ws_issue<6>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 7: ws_issue<7>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 8: ws_issue<8>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              case 9: ws_issue<9>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
              default: ws_issue<10>(d0, nc, a_lo0, b_lo0, p.a_kstep16, p.b_kstep16, a_hi, b_hi, idesc, acc, boff); break;
            }
          }
          if (p.dbg_flags & 16) mbar_arrive(empty_bar(slot)); else mma_commit(empty_bar(slot));
        }
        __syncwarp();
        acc = 1;
        if (++slot == p.ring) { slot = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) mma_commit(tfull_bar);
      __syncwarp();
    }
  } else {
    // ---- epilogue (once): TMEM -> fp32 partial ----
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    if (do_bias) {
      // thread m owns pixel m of every tile; the pixel's 32-byte row of block cb is two 16-byte chunks
      // whose order is swapped when address bit 7 (= bit 2 of the pixel index) is set
      float acc[8][16];
#pragma unroll
      for (int cb = 0; cb < 8; ++cb)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[cb][q] = 0.f;
      const uint32_t sw = ((uint32_t)m >> 2) & 1u;
      int slot = 0; uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        if (quarter == 0) ws_wait(full_bar(slot), phase);          // one polling warp, three parked on a named barrier
        asm volatile("bar.sync 2, 128;" ::: "memory");
        // conv: dY is the common tile at the head of the slot; deconv: dY are this group's parity boxes behind it
        const int nb_boxes = (p.dbg_flags & 4) ? 0 : (p.swap ? nbox : 1);
        for (int bx = 0; bx < nb_boxes; ++bx) {
          const uint8_t* row = smem_raw + (smem0 - smem_u32(smem_raw)) + (size_t)slot * p.slot_bytes +
                               (p.swap ? (size_t)p.a_bytes + (size_t)bx * p.var_box_bytes : (size_t)0) + (size_t)m * 32;
#pragma unroll
          for (int cb = 0; cb < 8; ++cb) {
            if (cb < bias_blocks) {
              const uint4* c4 = reinterpret_cast<const uint4*>(row + (size_t)cb * (kWsTileW * kWsTileH * 32));
              const uint4 lo = c4[sw], hi = c4[sw ^ 1u];
              const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                acc[cb][2 * j] += __uint_as_float(w[j] << 16);
                acc[cb][2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(slot));
        if (++slot == p.ring) { slot = 0; phase ^= 1u; }
      }
      // fixed-order reduction over the 128 pixels: lanes (shuffles), then the four warps
#pragma unroll
      for (int cb = 0; cb < 8; ++cb) {
        if (cb < bias_blocks) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float v = acc[cb][q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[quarter][cb * 16 + q] = v;
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (m < bias_blocks * 16) {
        const float tot = s_red[0][m] + s_red[1][m] + s_red[2][m] + s_red[3][m];
        if (p.swap) {
          // rows [split * npairs + tap0, + ntap): the group's sum goes to its first row, the others are zero
          for (int i = 0; i < ntap; ++i)
            p.bias_partial[((long long)split * p.npairs + tap0 + i) * p.npad + m] = i == 0 ? tot : 0.f;
        } else {
          p.bias_partial[(long long)split * p.npad + m] = tot;
        }
      }
    }
    if (has_work) {
      ws_wait(tfull_bar, 0);
      fence_after_sync();
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool row_ok = m < p.m_blocks * 16;
    for (int i = 0; i < ntap; ++i) {
      const int t = tap0 + i;
      float* P = p.partial + ((long long)split * p.npairs + t) * p.cpad * p.npad;
      for (int cb = 0; cb < p.n_blocks; ++cb) {
        float v[16];
        if (has_work) {
          tmem_ld16(lane_addr + i * ncols + cb * 16, v);
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

int launch_bias_grad(const TapWgrad& g, cudaStream_t st);
static int g_ring_override = 0;

static int ws_taps_per_cta(int npairs, int variant_blocks, int common_blocks, bool box_per_tap, bool halo) {
  const int bw = halo ? kWsTileW + 2 : kWsTileW, bh = halo ? kWsTileH + 2 : kWsTileH;
  const size_t a_bytes = (size_t)common_blocks * kWsTileW * kWsTileH * 32;
  const size_t var_box = (size_t)variant_blocks * bw * bh * 32;
  int tpc = 512 / (variant_blocks * 16);
  if (tpc > npairs) tpc = npairs;
  if (tpc > 10) tpc = 10;
  const size_t budget = kWsSmemMax - kWsStaticSlack - 1024;
  auto slot_for = [&](int taps) { return align_up(a_bytes + (size_t)(box_per_tap ? taps : 1) * var_box, 1024); };
  while (tpc > 1 && 2 * slot_for(tpc) > budget) --tpc;
  if (tpc < 1 || 2 * slot_for(tpc) > budget) return 0;
  return tpc;
}

// How many pixel splits fill the chip exactly once given the tap grouping (0 = geometry not eligible).
int wgrad_slab_splits(int npairs, int variant_blocks, int common_blocks, bool box_per_tap, long long tiles) {
  if (variant_blocks < 1 || variant_blocks > 16 || common_blocks < 1 || common_blocks > 8 || tiles < 1) return 0;
  const int tpc = ws_taps_per_cta(npairs, variant_blocks, common_blocks, box_per_tap, !box_per_tap && npairs > 1);
  if (tpc < 1) return 0;
  const int tg = (npairs + tpc - 1) / tpc;
  long long s = kSMs / tg;
  if (s > tiles) s = tiles;
  if (s < 1) s = 1;
  return (int)s;
}

// Returns 0 when launched, kSgNotEligible when this geometry belongs to the first-generation engine.
int launch_wgrad_slab_umma(const TapWgrad& g, cudaStream_t st) {
  static bool attr_set = false;
  { const char* e = getenv("N2N_NO_SLAB"); if (e && atoi(e)) return kSgNotEligible; }
  const bool swap = g.ndyviews > 1;       // deconv: common = X, variants = dY parity views
  const View& common = swap ? g.x[0] : g.dy[0];
  const int m_blocks = swap ? g.c_blocks : g.n_blocks;
  const int n_blocks = swap ? g.n_blocks : g.c_blocks;
  if (g.dtype != N2N_BF16 || m_blocks < 1 || m_blocks > 8 || n_blocks < 1 || n_blocks > 16) return kSgNotEligible;
  if (g.npairs < 1 || g.npairs > 9) return kSgNotEligible;
  const int H = common.H, W = common.W;
  if (H < 4 || W < 4) return kSgNotEligible;    // edge tiles: out-of-image pixels are zero-filled by TMA and add nothing
  bool halo = false;
  int nvar = 0;
  for (int t = 0; t < g.npairs; ++t) {
    if (g.pair_dy[t] < -1 || g.pair_dy[t] > 1 || g.pair_dx[t] < -1 || g.pair_dx[t] > 1) return kSgNotEligible;
    if (g.pair_dy[t] || g.pair_dx[t]) halo = true;
    const int vi = swap ? g.pair_dyv[t] : g.pair_xv[t];
    if (vi + 1 > nvar) nvar = vi + 1;
    if ((swap ? g.pair_xv[t] : g.pair_dyv[t]) != 0) return kSgNotEligible;
  }
  if (nvar > 4) return kSgNotEligible;
  // conv: every tap is a view of ONE variant tensor; deconv: one variant tensor per tap, no offsets
  const bool box_per_tap = nvar > 1;
  if (box_per_tap && (halo || nvar != g.npairs)) return kSgNotEligible;

  WsParams p;
  memset(&p, 0, sizeof(p));
  p.npairs = g.npairs; p.m_blocks = m_blocks; p.n_blocks = n_blocks; p.swap = swap ? 1 : 0;
  p.npad = g.n_blocks * 16; p.cpad = g.c_blocks * 16;
  p.partial = g.partial;
  p.bias_partial = (g.bias_partial && (!swap || (box_per_tap && g.ndyviews == g.npairs && n_blocks <= 8))) ? g.bias_partial : nullptr;
  { static const char* const df = getenv("N2N_DBG_FLAGS"); p.dbg_flags = df ? atoi(df) : 0; }
  { const char* e = getenv("N2N_WS_RING"); if (e && atoi(e) >= 2) g_ring_override = atoi(e); }
  p.halo = halo ? 1 : 0; p.box_per_tap = box_per_tap ? 1 : 0;
  const int bw = halo ? kWsTileW + 2 : kWsTileW, bh = halo ? kWsTileH + 2 : kWsTileH;
  p.a_bytes = (uint32_t)(m_blocks * kWsTileW * kWsTileH * 32);
  p.var_box_bytes = (uint32_t)(n_blocks * bw * bh * 32);
  // taps per CTA: bounded by TMEM columns, and (deconv) by the shared memory one slot may take
  const int tpc = ws_taps_per_cta(g.npairs, n_blocks, m_blocks, box_per_tap, halo);
  if (tpc < 1) return kSgNotEligible;
  const size_t budget = kWsSmemMax - kWsStaticSlack - 1024;
  auto slot_for = [&](int taps) { return align_up((size_t)p.a_bytes + (size_t)(box_per_tap ? taps : 1) * p.var_box_bytes, 1024); };
  p.taps_per_cta = tpc;
  p.tgroups = (g.npairs + tpc - 1) / tpc;
  p.slot_bytes = (uint32_t)slot_for(tpc);
  int ring = (int)(budget / p.slot_bytes);
  if (ring > kWsMaxRing) ring = kWsMaxRing;
  if (g_ring_override && g_ring_override < ring) ring = g_ring_override;
  p.ring = ring;
  p.tiles_x = (W + kWsTileW - 1) / kWsTileW; p.tiles_y = (H + kWsTileH - 1) / kWsTileH;
  p.tiles = (long long)common.N * p.tiles_x * p.tiles_y;
  const int splits = g.splits;
  p.nsplits = splits;
  if (p.tiles >= (1LL << 31)) return kSgNotEligible;
  p.tmem_cols = tmem_cols_for(tpc * n_blocks * 16);
  p.idesc = make_idesc_bf16(128, n_blocks * 16, true, true);
  // MN-major descriptors: LBO = bytes between 16-channel blocks of a box, SBO = bytes between image rows
  const uint32_t a_blk = (uint32_t)(kWsTileW * kWsTileH * 32), b_blk = (uint32_t)(bw * bh * 32);
  p.a_lbo = ((a_blk >> 4) & 0x3FFFu) << 16;
  p.b_lbo = ((b_blk >> 4) & 0x3FFFu) << 16;
  p.a_hi = (((uint32_t)(kWsTileW * 32) >> 4) & 0x3FFFu) | (1u << 14) | (kSwizzle32 << 29);
  p.b_hi = (((uint32_t)(bw * 32) >> 4) & 0x3FFFu) | (1u << 14) | (kSwizzle32 << 29);
  p.a_kstep16 = (uint32_t)(2 * kWsTileW * 32) >> 4;
  p.b_kstep16 = (uint32_t)(2 * bw * 32) >> 4;
  for (int t = 0; t < g.npairs; ++t) {
    const int oy = halo ? g.pair_dy[t] + 1 : 0, ox = halo ? g.pair_dx[t] + 1 : 0;
    p.tap_off16[t] = (uint16_t)(((oy * bw + ox) * 32) >> 4);
    p.tap_view[t] = (int8_t)(swap ? g.pair_dyv[t] : g.pair_xv[t]);
  }
  N2N_TRY(encode_c16_tensor_map(&p.tmap_common, common, kWsTileW, kWsTileH, m_blocks));
  for (int v = 0; v < 4; ++v) {
    const View& vv = swap ? g.dy[v < nvar ? v : 0] : g.x[v < nvar ? v : 0];
    if (vv.H != H || vv.W != W) return kSgNotEligible;
    N2N_TRY(encode_c16_tensor_map(&p.tmap_var[v], vv, bw, bh, n_blocks));
  }
  // the M = 128 MMA always walks 8 channel blocks of the common tile: keep that window inside the allocation
  size_t smem = 1024 + (size_t)ring * p.slot_bytes;
  const size_t window = 1024 + (size_t)(ring - 1) * p.slot_bytes + 8u * a_blk + 1024;
  if (smem < window) smem = window;
  if (smem > kWsSmemMax - kWsStaticSlack) return kSgNotEligible;
  if (!attr_set) {
    N2N_CUDA(cudaFuncSetAttribute(wgrad_slab_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kWsSmemMax - kWsStaticSlack)));
    attr_set = true;
  }
  N2N_CUDA(launch_pdl(wgrad_slab_umma_kernel, dim3(splits * p.tgroups), dim3(kWsThreads), smem, st, p));
  N2N_LAUNCH_CHECK();
  if (g.bias_partial && !p.bias_partial) return launch_bias_grad(g, st);   // deconv: dY is the variant operand
  return 0;
}

}  // namespace n2n
