// headbwd_umma.cu — the whole backward of the 1x1 head (nin_c -> nin_b -> nin_a, arch_unet.py:186-190 /
// :257-259) below nin_c's own weight gradient, as ONE persistent tcgen05 kernel, the mirror image of
// head_umma.cu:
//
//   g_nb  = (Wc^T  g_out) * lrelu'(NB)        rank-out_nc outer product per pixel, CUDA cores (stage S0)
//   g_na  = (Wb^T  g_nb ) * lrelu'(NA)        GEMM-1 on the tensor core, masked in stage S1
//   g_d1b = (Wa^T  g_na ) * lrelu'(D1B)       GEMM-2, masked in stage S2
//   dWb   = g_nb^T  NA ,  dbb = sum g_nb      weight-gradient GEMMs over the pixel axis, accumulated in
//   dWa   = g_na^T  D1B,  dba = sum g_na      TMEM across all tiles of the CTA (bias = a column of ones)
//
// Unfused these are three HBM-bound input-gradient launches plus two weight-gradient launches that
// each read back a gradient tensor the previous one just wrote; fused, g_nb and g_na never leave the
// SM: they are written once into shared memory in the swizzled 32-byte-row tile layout, where the
// SAME bytes serve as the K-major A operand of the next input-gradient GEMM and as the MN-major A
// operand (K = pixels) of the weight-gradient GEMM.  The saved activations NA / D1B arrive by TMA in
// that layout too: they are the B operand of the weight-gradient GEMM and, read back by the stage
// threads, the sign that gives lrelu'(.) (LeakyReLU is in-place in the reference, arch_unet.py:113,
// slope > 0).  Stage S1 overwrites the NA tile in place with g_na once GEMM-1 / dWb have consumed it.
//
// Per 8 x 16-pixel tile: warp 0 is the TMA producer (activation tiles, prefetched into L2 three tiles ahead),
// warp 1 issues dWb + GEMM-1 and warp 14 dWa + GEMM-2 (two issuers: one in-order issuer would queue the GEMM
// that frees a tile buffer behind the GEMM that waits for a TMA load); warps 2-5 run S0, 6-9 S1, 10-13 S2
// (thread = pixel = TMEM lane).  The g_nb and D1B tiles are double-buffered, the NA tile (which lives from its
// TMA load through S1's in-place rewrite to GEMM-2) triple-buffered, so the three stages work on consecutive
// tiles; the two input-gradient accumulators are single TMEM buffers that a stage drains into registers and
// hands back before it starts its arithmetic.  Each CTA stores its fp32
// weight-gradient partial once at the end ([grid][c][n]; pack.cu's unpack kernel reduces the CTAs in
// a fixed order -> deterministic).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kHbThreads = 480;
constexpr int kHbNaBufs = 3;             // NA tiles live longest (TMA -> GEMM-1 -> S1 in place -> GEMM-2)
constexpr int kHbMaxOut = 4;
constexpr int kHbBlocks = 6;              // the head is literally 96 channels wide (arch_unet.py:177-190)
constexpr int kHbCh = kHbBlocks * 16;
constexpr uint32_t kHbBlk = 4096;         // one 16-channel block of a 128-pixel tile (32-byte rows)
constexpr uint32_t kHbHBuf = kHbBlocks * kHbBlk;          // g_nb tile
constexpr uint32_t kHbXBuf = (kHbBlocks + 1) * kHbBlk;    // activation tile + the block of ones behind it
constexpr uint32_t kHbWBytes = kHbBlocks * kHbCh * 32;    // one packed 96 x 96 input-gradient weight
constexpr int kHbWN = kHbCh + 16;         // weight-gradient accumulator width: 96 input channels + bias block
// TMEM columns: D1 | D2 | dWb | dWa
constexpr uint32_t kHbColD1 = 0, kHbColD2 = kHbCh, kHbColWb = 2 * kHbCh, kHbColWa = 2 * kHbCh + kHbWN;
constexpr uint32_t kHbTmemCols = 512;
static_assert(kHbColWa + kHbWN <= kHbTmemCols, "TMEM budget");
constexpr size_t kHbSmem = 1024 + 2 * (size_t)kHbWBytes + 2 * (size_t)kHbHBuf + (2 + kHbNaBufs) * (size_t)kHbXBuf;

struct HbParams {
  int out_nc;
  int tiles_x, tiles_y, ntiles, H, W;
  float slope;
  const uint8_t *wb, *wa;
  const float *wc, *gout;
  View act_nb, g_d1b;
  float *partial_b, *bpartial_b, *partial_a, *bpartial_a;
  CUtensorMap tmap_na, tmap_d1b;
};

__device__ __forceinline__ void hb_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void hb_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void hb_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void hb_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void hb_ld_global_32B(const void* ptr, uint32_t w[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(ptr)
               : "memory");
}
__device__ __forceinline__ void hb_st_global_32B(void* ptr, const uint32_t w[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// v[16] (fp32 gradient of one channel block of one pixel) * lrelu'(activation words) -> packed bf16.
// lrelu' = 1 where the bf16 activation is > 0 (sign clear and non-zero, tested on the raw bits), else slope:
// one compare and one predicated multiply per element.
__device__ __forceinline__ void hb_mask_pack(const float v[16], const uint32_t mk[8], float slope, uint32_t w[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float g0 = v[2 * j], g1 = v[2 * j + 1];
    if ((int)(mk[j] << 16) <= 0) g0 *= slope;
    if ((int)mk[j] <= 0xFFFF) g1 *= slope;
    __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
}

template <int OUT_NC>
__global__ void __launch_bounds__(kHbThreads, 1)
head_bwd_umma_kernel(const __grid_constant__ HbParams p) {
  extern __shared__ uint8_t smem_raw[];
  // h0_full[2] h0_empty[2] db_full[2] db_empty[2] | na_full[3] na_empty[3] h1_full[3] | d1_full d1_empty d2_full
  // d2_empty w_full acc_full
  __shared__ uint64_t bars[23];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_wc[OUT_NC * kHbCh];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t wb0 = smem0, wa0 = wb0 + kHbWBytes;
  const uint32_t h00 = wa0 + kHbWBytes;                  // g_nb tiles (written by S0)
  const uint32_t na0 = h00 + 2 * kHbHBuf;                // NA tiles (TMA), overwritten in place with g_na by S1
  const uint32_t db0 = na0 + kHbNaBufs * kHbXBuf;        // D1B tiles (TMA)
  const uint32_t bar0 = smem_u32(bars);
  enum { H0F = 0, H0E, DBF, DBE };
  enum { NAF = 0, NAE, H1F };
  auto bar = [&](int kind, int b) { return bar0 + 8u * (2 * kind + b); };
  auto bar3 = [&](int kind, int b) { return bar0 + 8u * (8 + 3 * kind + b); };
  const uint32_t d1_full = bar0 + 8u * 17, d1_empty = bar0 + 8u * 18, d2_full = bar0 + 8u * 19, d2_empty = bar0 + 8u * 20;
  const uint32_t w_full = bar0 + 8u * 21;
  const uint32_t acc_full = bar0 + 8u * 22;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(H0F, b), 4); mbar_init(bar(H0E, b), 1);
      mbar_init(bar(DBF, b), 1); mbar_init(bar(DBE, b), 4);
    }
    for (int b = 0; b < kHbNaBufs; ++b) {
      mbar_init(bar3(NAF, b), 1); mbar_init(bar3(NAE, b), 1); mbar_init(bar3(H1F, b), 4);
    }
    mbar_init(d1_full, 1); mbar_init(d1_empty, 4);
    mbar_init(d2_full, 1); mbar_init(d2_empty, 4);
    mbar_init(w_full, 1); mbar_init(acc_full, 2);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < OUT_NC * kHbCh; i += kHbThreads) s_wc[i] = p.wc[i];
  // the block of ones behind every activation tile: column 96 of the weight-gradient GEMM = bias gradient
  for (int i = threadIdx.x; i < (2 + kHbNaBufs) * (int)(kHbBlk / 4); i += kHbThreads) {
    const int buf = i / (int)(kHbBlk / 4), wd = i - buf * (int)(kHbBlk / 4);
    const uint32_t base = na0 + buf * kHbXBuf + kHbBlocks * kHbBlk;      // the D1B buffers follow the NA buffers
    reinterpret_cast<uint32_t*>(smem_gen + (base - smem0))[wd] = 0x3F803F80u;
  }
  fence_proxy_async();
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), kHbTmemCols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int niter = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ---- TMA producer: weights once, then the NA / D1B activation tiles ----
    if (elect_one_sync()) {
      prefetch_tensormap(&p.tmap_na);
      prefetch_tensormap(&p.tmap_d1b);
      mbar_arrive_expect_tx(w_full, 2 * kHbWBytes);
      bulk_load(wb0, p.wb, kHbWBytes, w_full);
      bulk_load(wa0, p.wa, kHbWBytes, w_full);
    }
    __syncwarp();
    pdl_wait();
    // A tile's buffer is held from the TMA issue until the GEMMs and the stage that read it are done, so
    // with two buffers the DRAM latency would sit inside the per-tile cycle: pull the boxes into L2 a few
    // tiles ahead and let the shared-memory load pay only the L2 latency.
    constexpr int kAhead = 3;
    auto prefetch = [&](int lt) {
      if (lt < niter && elect_one_sync()) {
        const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
        const int img = tile / tiles_per_img;
        const int r = tile - img * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        tma_prefetch_l2_5d(&p.tmap_na, 0, tx * 8, ty * 16, 0, img);
        tma_prefetch_l2_5d(&p.tmap_d1b, 0, tx * 8, ty * 16, 0, img);
      }
      __syncwarp();
    };
    for (int lt = 0; lt < kAhead; ++lt) prefetch(lt);
    for (int lt = 0; lt < niter; ++lt) {
      const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
      const int img = tile / tiles_per_img;
      const int r = tile - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int b = lt & 1;
      const uint32_t par = ((uint32_t)lt >> 1) & 1u;
      const int b3 = lt % kHbNaBufs;
      const uint32_t par3 = (uint32_t)(lt / kHbNaBufs) & 1u;
      prefetch(lt + kAhead);
      hb_wait(bar3(NAE, b3), par3 ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar3(NAF, b3), kHbHBuf);
        tma_load_5d(na0 + b3 * kHbXBuf, &p.tmap_na, bar3(NAF, b3), 0, tx * 8, ty * 16, 0, img);
      }
      __syncwarp();
      hb_wait(bar(DBE, b), par ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar(DBF, b), kHbHBuf);
        tma_load_5d(db0 + b * kHbXBuf, &p.tmap_d1b, bar(DBF, b), 0, tx * 8, ty * 16, 0, img);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 14) {
    // ---- MMA issuers.  Warp 1: dWb + GEMM-1 per tile; warp 14: dWa + GEMM-2.  Two issuers, because one
    // in-order issuer would hold GEMM-2 of tile i-1 (which frees an NA buffer for the next TMA load) behind
    // GEMM-1 of tile i (which waits for a TMA load): the loads would run strictly one at a time.
    const bool second = warp == 14;
    pdl_wait();
    if (!second) pdl_release();
    hb_wait(w_full, 0);
    const uint32_t hi = (256u >> 4) | (1u << 14) | (kSwizzle32 << 29);   // SBO = one 8-pixel image row, both majors
    const uint32_t idesc_g = make_idesc_bf16(128, kHbCh, false, false);
    const uint32_t idesc_w = make_idesc_bf16(128, kHbWN, true, true);
    const uint32_t lbo_k = 1u << 16, lbo_mn = (kHbBlk >> 4) << 16;
    auto lo = [](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
    // weight gradient: K = pixels, two image rows (512 B) per MMA; input gradient: K = channels, one block per MMA
    auto issue = [&](uint32_t g_buf, uint32_t x_buf, uint32_t w_buf, uint32_t col_w, uint32_t col_d, uint32_t first) {
      const uint32_t ga = lo(g_buf), xa = lo(x_buf), wa = lo(w_buf);
#pragma unroll 1
      for (int kk = 0; kk < 8; ++kk)
        hb_mma(tmem_base + col_w, (ga + kk * 32u) | lbo_mn, (xa + kk * 32u) | lbo_mn, hi, idesc_w, (first | (uint32_t)kk) ? 1u : 0u);
#pragma unroll 1
      for (int cb = 0; cb < kHbBlocks; ++cb)
        hb_mma(tmem_base + col_d, (ga + cb * (kHbBlk >> 4)) | lbo_k, (wa + cb * (uint32_t)(kHbCh * 2)) | lbo_k, hi, idesc_g,
               cb ? 1u : 0u);
    };
    for (int lt = 0; lt < niter; ++lt) {
      const int b = lt & 1;
      const uint32_t par = ((uint32_t)lt >> 1) & 1u;
      const int b3 = lt % kHbNaBufs;
      const uint32_t par3 = (uint32_t)(lt / kHbNaBufs) & 1u;
      if (!second) {
        hb_wait(d1_empty, ((uint32_t)lt & 1u) ^ 1u);
        hb_wait(bar3(NAF, b3), par3);
        hb_wait(bar(H0F, b), par);
        fence_after_sync();
        if (elect_one_sync()) {
          issue(h00 + b * kHbHBuf, na0 + b3 * kHbXBuf, wb0, kHbColWb, kHbColD1, (uint32_t)lt);
          mma_commit(bar(H0E, b));
          mma_commit(d1_full);
        }
      } else {
        hb_wait(d2_empty, ((uint32_t)lt & 1u) ^ 1u);
        hb_wait(bar(DBF, b), par);
        hb_wait(bar3(H1F, b3), par3);
        fence_after_sync();
        if (elect_one_sync()) {
          issue(na0 + b3 * kHbXBuf, db0 + b * kHbXBuf, wa0, kHbColWa, kHbColD2, (uint32_t)lt);
          mma_commit(bar3(NAE, b3));
          mma_commit(d2_full);
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(acc_full);
    __syncwarp();
  } else if (warp < 14) {
    const int stage = (warp - 2) >> 2;                 // 0: S0, 1: S1, 2: S2
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int py = m >> 3, px = m & 7;
    const uint32_t sw = ((uint32_t)m >> 2) & 1u;       // 16-byte chunk swap of this pixel's rows (address bit 7)
    pdl_wait();
    if (stage == 0) {
      // Software pipeline over tiles: the global loads of tile i+1 (activation words that give lrelu', the
      // output gradient) are issued while tile i is processed, each into the registers its block just freed.
      const long long hw = (long long)p.H * p.W;
      uint32_t mk[kHbBlocks][8];
      float go_next[OUT_NC];
      long long apix = 0;
      bool valid = false;
      auto locate = [&](int lt, long long& gidx) {
        const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
        const int img = tile / tiles_per_img;
        const int r = tile - img * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int y = ty * 16 + py, x = tx * 8 + px;
        valid = lt < niter && y < p.H && x < p.W;
        apix = (long long)img * p.act_nb.sN + (long long)y * p.act_nb.sY + (long long)x * p.act_nb.sX;
        gidx = (long long)img * OUT_NC * hw + (long long)y * p.W + x;
      };
      auto load_block = [&](int cb) {
        if (valid) hb_ld_global_32B((const __nv_bfloat16*)p.act_nb.ptr + apix + cb * p.act_nb.sCb, mk[cb]);
        else {
#pragma unroll
          for (int q = 0; q < 8; ++q) mk[cb][q] = 0u;
        }
      };
      auto load_gout = [&](long long gidx) {
#pragma unroll
        for (int oc = 0; oc < OUT_NC; ++oc) go_next[oc] = valid ? p.gout[gidx + oc * hw] : 0.f;
      };
      {
        long long gidx;
        locate(0, gidx);
#pragma unroll
        for (int cb = 0; cb < kHbBlocks; ++cb) load_block(cb);
        load_gout(gidx);
      }
      for (int lt = 0; lt < niter; ++lt) {
        const int b = lt & 1;
        const uint32_t par = ((uint32_t)lt >> 1) & 1u;
        float go[OUT_NC];
#pragma unroll
        for (int oc = 0; oc < OUT_NC; ++oc) go[oc] = go_next[oc];
        long long gidx;
        locate(lt + 1, gidx);                       // from here on `valid` / `apix` describe the NEXT tile
        load_gout(gidx);
        if (quarter == 0) hb_wait(bar(H0E, b), par ^ 1u);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint8_t* row = smem_gen + (h00 - smem0) + (uint32_t)b * kHbHBuf + (uint32_t)m * 32u;
#pragma unroll
        for (int cb = 0; cb < kHbBlocks; ++cb) {
          float v[16];
#pragma unroll
          for (int oc = 0; oc < OUT_NC; ++oc) {
            const float4* wr = reinterpret_cast<const float4*>(&s_wc[oc * kHbCh + cb * 16]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 wv = wr[q];
              if (oc == 0) {
                v[4 * q] = wv.x * go[0]; v[4 * q + 1] = wv.y * go[0]; v[4 * q + 2] = wv.z * go[0]; v[4 * q + 3] = wv.w * go[0];
              } else {
                v[4 * q] += wv.x * go[oc]; v[4 * q + 1] += wv.y * go[oc]; v[4 * q + 2] += wv.z * go[oc]; v[4 * q + 3] += wv.w * go[oc];
              }
            }
          }
          uint32_t w[8];
          hb_mask_pack(v, mk[cb], p.slope, w);
          load_block(cb);
          uint4* dst = reinterpret_cast<uint4*>(row + (uint32_t)cb * kHbBlk);
          dst[sw] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[sw ^ 1u] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(H0F, b));
      }
    } else {
      const bool s1 = stage == 1;
      const uint32_t dfull = s1 ? d1_full : d2_full, dempty = s1 ? d1_empty : d2_empty;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t dcol = s1 ? kHbColD1 : kHbColD2;
      for (int lt = 0; lt < niter; ++lt) {
        const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
        const int img = tile / tiles_per_img;
        const int r = tile - img * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int y = ty * 16 + py, x = tx * 8 + px;
        const bool valid = y < p.H && x < p.W;
        const int b = lt & 1;
        const uint32_t par = ((uint32_t)lt >> 1) & 1u;
        const int b3 = lt % kHbNaBufs;
        const uint32_t par3 = (uint32_t)(lt / kHbNaBufs) & 1u;
        if (quarter == 0) {
          hb_wait(s1 ? bar3(NAF, b3) : bar(DBF, b), s1 ? par3 : par);   // the TMA'd activation tile is visible to this group
          hb_wait(dfull, (uint32_t)lt & 1u);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + stage) : "memory");
        fence_after_sync();
        // drain the accumulator into registers and hand it back before any arithmetic
        uint32_t rr[kHbBlocks][16];
#pragma unroll
        for (int cb = 0; cb < kHbBlocks; ++cb) hb_ld16_nowait(lane_addr + dcol + cb * 16, rr[cb]);
        hb_ld_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(dempty);
        uint8_t* row = smem_gen + ((s1 ? na0 + (uint32_t)b3 * kHbXBuf : db0 + (uint32_t)b * kHbXBuf) - smem0) + (uint32_t)m * 32u;
        const long long gpix = (long long)img * p.g_d1b.sN + (long long)y * p.g_d1b.sY + (long long)x * p.g_d1b.sX;
#pragma unroll
        for (int cb = 0; cb < kHbBlocks; ++cb) {
          uint4* rowb = reinterpret_cast<uint4*>(row + (uint32_t)cb * kHbBlk);
          const uint4 m0 = rowb[sw], m1 = rowb[sw ^ 1u];
          const uint32_t mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rr[cb][q]);
          uint32_t w[8];
          hb_mask_pack(v, mk, p.slope, w);
          if (s1) {
            rowb[sw] = make_uint4(w[0], w[1], w[2], w[3]);
            rowb[sw ^ 1u] = make_uint4(w[4], w[5], w[6], w[7]);
          } else if (valid) {
            hb_st_global_32B((__nv_bfloat16*)p.g_d1b.ptr + gpix + cb * p.g_d1b.sCb, w);
          }
        }
        if (s1) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(s1 ? bar3(H1F, b3) : bar(DBE, b));
      }
      // ---- this CTA's weight-gradient partial: TMEM lane = output channel n, column = input channel c ----
      if (quarter == 0) hb_wait(acc_full, 0);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + stage) : "memory");
      fence_after_sync();
      float* P = (s1 ? p.partial_b : p.partial_a) + (size_t)blockIdx.x * kHbCh * kHbCh;
      float* Bp = (s1 ? p.bpartial_b : p.bpartial_a) + (size_t)blockIdx.x * kHbCh;
      const uint32_t wcol = s1 ? kHbColWb : kHbColWa;
      for (int cb = 0; cb <= kHbBlocks; ++cb) {
        float v[16];
        tmem_ld16(lane_addr + wcol + cb * 16, v);
        if (m < kHbCh) {
          if (cb < kHbBlocks) {
#pragma unroll
            for (int q = 0; q < 16; ++q) P[(cb * 16 + q) * kHbCh + m] = v[q];
          } else {
            Bp[m] = v[0];
          }
        }
      }
      fence_before_sync();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, kHbTmemCols); }
}

// Geometry the fused kernel covers, and how many CTAs (= weight-gradient partial rows) it runs: the
// plan sizes the nin_a / nin_b partial buffers and the unpack reduction from this.
int head_bwd_splits(int dtype, int blocks, int out_nc, int n, int h, int w) {
  { const char* e = getenv("N2N_NO_HEAD_FUSION"); if (e && atoi(e)) return 0; }
  if (dtype != N2N_BF16 || blocks != kHbBlocks || out_nc < 1 || out_nc > kHbMaxOut || h < 4 || w < 4) return 0;
  const long long tiles = (long long)n * ((w + 7) / 8) * ((h + 15) / 16);
  if (tiles < 1 || tiles >= (1LL << 31)) return 0;
  return tiles < kSMs ? (int)tiles : kSMs;
}

// Returns 0 when launched, kSgNotEligible when the geometry is not covered (caller runs the three
// input-gradient and two weight-gradient launches one by one).
int launch_head_bwd_umma(const HeadBwd& h, cudaStream_t st) {
  static int attr_set = 0;
  const int grid = head_bwd_splits(N2N_BF16, h.blocks, h.out_nc, h.g_d1b.N, h.g_d1b.H, h.g_d1b.W);
  if (grid < 1 || h.channels != kHbCh || h.splits != grid) return kSgNotEligible;
  N2N_CHECK_ARG(h.partial_a && h.partial_b && h.bpartial_a && h.bpartial_b, "head_bwd: null partial buffer");
  HbParams p;
  memset(&p, 0, sizeof(p));
  p.out_nc = h.out_nc; p.H = h.g_d1b.H; p.W = h.g_d1b.W;
  p.tiles_x = (p.W + 7) / 8; p.tiles_y = (p.H + 15) / 16;
  p.ntiles = h.g_d1b.N * p.tiles_x * p.tiles_y;
  p.slope = h.slope;
  p.wb = (const uint8_t*)h.wb_dgrad; p.wa = (const uint8_t*)h.wa_dgrad; p.wc = h.wc; p.gout = h.gout;
  p.act_nb = h.act_nb; p.g_d1b = h.g_d1b;
  p.partial_b = h.partial_b; p.bpartial_b = h.bpartial_b; p.partial_a = h.partial_a; p.bpartial_a = h.bpartial_a;
  N2N_TRY(encode_c16_tensor_map(&p.tmap_na, h.act_na, 8, 16, kHbBlocks));
  N2N_TRY(encode_c16_tensor_map(&p.tmap_d1b, h.act_d1b, 8, 16, kHbBlocks));
  void (*kernel)(HbParams) = h.out_nc == 1 ? head_bwd_umma_kernel<1> : h.out_nc == 2 ? head_bwd_umma_kernel<2>
                             : h.out_nc == 3 ? head_bwd_umma_kernel<3> : head_bwd_umma_kernel<4>;
  if (!(attr_set & (1 << h.out_nc))) {
    N2N_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHbSmem));
    attr_set |= 1 << h.out_nc;
  }
  N2N_CUDA(launch_pdl(kernel, dim3(grid), dim3(kHbThreads), kHbSmem, st, p));
  N2N_LAUNCH_CHECK();
  return 0;
}

}  // namespace n2n
