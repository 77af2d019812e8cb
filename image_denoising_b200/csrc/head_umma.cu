// head_umma.cu — the UNet's 1x1 head (nin_a -> LeakyReLU -> nin_b -> LeakyReLU -> nin_c,
// arch_unet.py:186-190 / :257-259) as ONE persistent tcgen05 kernel.  Unfused, the chain is three
// HBM round trips over a full-resolution 96-channel tensor (the largest in the network) for 13
// MFLOP/pixel; fused, the tile of dec_conv1b's output is read once and only the out_nc-channel
// fp32 result leaves the SM (plus, in a training pass, the two bf16 activations the backward needs).
//
// Per 8 x 16-pixel tile (M = 128):
//   warp 0     TMA: the X tile [in_blocks][128 px][16 ch] (K-major SWIZZLE_32B operand) -> ring slot
//   warp 1     MMA-1: D1 = X * Wa^T + bias_a   (TMEM, double buffered; the bias is one more K block)
//   warp 18    MMA-2: D2 = H1 * Wb^T + bias_b  (H1 = shared-memory operand written by stage E1; triple buffered —
//              E2 is the longest stage; its own issuer so that neither GEMM queues behind the other one's wait)
//   warps 2-9  E1: D1 -> LeakyReLU -> bf16 -> H1 tile in shared memory, in exactly the swizzled K-major
//              layout MMA-2 wants (16-byte chunk XOR address bit 7), then fence.proxy.async so the tensor
//              core sees it; D1 goes back to MMA-1 as soon as its last block is in registers
//   warps 10-17 E2: D2 -> LeakyReLU -> dot with nin_c's rows (fp32, registers) -> +bias_c -> fp32 NCHW
//              store (training pass: the bf16-rounded activation is what is saved and what nin_c sees)
// E1 and E2 work on different tiles at the same time; both weight matrices stay resident in
// shared memory; nin_c's weights are broadcast reads of a small shared-memory table.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kHdThreads = 608;            // TMA warp, MMA-1 warp, 8 warps for stage E1, 8 for stage E2, MMA-2 warp
constexpr int kHdD2Bufs = 3;               // E2 is the longest stage: a third accumulator keeps MMA-2 ahead of it
constexpr int kHdRing = 5;
constexpr int kHdMaxOut = 4;

struct HdParams {
  int in_blocks, mid_blocks, out_nc;
  int tiles_x, tiles_y, ntiles;
  int has_save;
  float slope;
  uint32_t wa_bytes, wb_bytes, x_bytes, h_bytes, tmem_cols, idesc;
  const uint8_t *wa, *wb;
  const float *bias_a, *bias_b, *wc, *bias_c;
  float* out_nchw;
  View x, save_a, save_b;
  CUtensorMap tmap_x;
};

__device__ __forceinline__ void hd_wait(uint32_t bar, uint32_t parity) {
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

__device__ __forceinline__ void hd_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ void hd_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive accumulator columns (two 16-channel blocks) per instruction; completion via hd_ld_wait
__device__ __forceinline__ void hd_ld32_issue(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void hd_ld16_issue(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void hd_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void hd_st_global_32B(void* ptr, const uint32_t w[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

template <bool SAVE>
__global__ void __launch_bounds__(kHdThreads, 1)
head_chain_umma_kernel(const __grid_constant__ HdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // barriers: x_full[ring] x_empty[ring] d1_full[2] d1_empty[2] h_full[2] h_empty[2] d2_full[3] d2_empty[3] w_full
  __shared__ uint64_t bars[2 * kHdRing + 8 + 2 * kHdD2Bufs + 1];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_wc[kHdMaxOut * 128], s_bc[kHdMaxOut];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t wa0 = smem0, wb0 = wa0 + p.wa_bytes;
  const uint32_t x0s = wb0 + p.wb_bytes;                     // X ring
  const uint32_t h0s = x0s + kHdRing * p.x_bytes;            // H1 double buffer
  // Bias through the tensor core: one extra K block whose activation operand is constant (channels 0, 1 = 1)
  // and whose weight rows carry the bias split into bf16 hi + lo parts — 16 FADD + 4 LDS less per
  // epilogue block, for one more (cheap) MMA per tile and GEMM.
  const uint32_t ones0 = h0s + 2 * p.h_bytes;                // [128 rows][32 B]
  const uint32_t bia0 = ones0 + 4096u, bib0 = bia0 + (uint32_t)p.mid_blocks * 16u * 32u;
  const uint32_t bar0 = smem_u32(bars);
  auto x_full = [&](int s) { return bar0 + 8u * s; };
  auto x_empty = [&](int s) { return bar0 + 8u * (kHdRing + s); };
  auto d1_full = [&](int b) { return bar0 + 8u * (2 * kHdRing + b); };
  auto d1_empty = [&](int b) { return bar0 + 8u * (2 * kHdRing + 2 + b); };
  auto h_full = [&](int b) { return bar0 + 8u * (2 * kHdRing + 4 + b); };
  auto h_empty = [&](int b) { return bar0 + 8u * (2 * kHdRing + 6 + b); };
  auto d2_full = [&](int b) { return bar0 + 8u * (2 * kHdRing + 8 + b); };
  auto d2_empty = [&](int b) { return bar0 + 8u * (2 * kHdRing + 8 + kHdD2Bufs + b); };
  const uint32_t w_full = bar0 + 8u * (2 * kHdRing + 8 + 2 * kHdD2Bufs);
  const int nmid = p.mid_blocks * 16;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHdRing; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d1_full(b), 1); mbar_init(d1_empty(b), 4);
      mbar_init(h_full(b), 4);  mbar_init(h_empty(b), 1);
    }
    for (int b = 0; b < kHdD2Bufs; ++b) { mbar_init(d2_full(b), 1); mbar_init(d2_empty(b), 4); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kHdMaxOut * 128; i += kHdThreads) {
    const int oc = i >> 7, c = i & 127;
    s_wc[i] = (oc < p.out_nc && c < nmid) ? p.wc[oc * nmid + c] : 0.f;
  }
  if (threadIdx.x < kHdMaxOut) s_bc[threadIdx.x] = threadIdx.x < p.out_nc ? p.bias_c[threadIdx.x] : 0.f;
  {
    uint8_t* base = smem_raw + (smem0 - smem_u32(smem_raw));
    for (int r = threadIdx.x; r < 128 + 2 * nmid; r += kHdThreads) {
      // rows 0..127: the ones block; then nmid rows of bias_a, then nmid rows of bias_b
      const uint32_t addr = r < 128 ? ones0 + r * 32u : (r < 128 + nmid ? bia0 + (r - 128) * 32u : bib0 + (r - 128 - nmid) * 32u);
      float v = 1.0f;
      if (r >= 128) v = r < 128 + nmid ? p.bias_a[r - 128] : p.bias_b[r - 128 - nmid];
      const __nv_bfloat16 hi_part = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo_part = r < 128 ? hi_part : __float2bfloat16_rn(v - __bfloat162float(hi_part));
      const uint32_t sw = (addr >> 7) & 1u;                  // logical chunk 0 sits in physical chunk sw
      uint4* row = reinterpret_cast<uint4*>(base + (addr - smem0));
      __nv_bfloat162 h2 = __halves2bfloat162(hi_part, lo_part);
      row[sw] = make_uint4(*reinterpret_cast<uint32_t*>(&h2), 0u, 0u, 0u);
      row[sw ^ 1u] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t hi = (256u >> 4) | (1u << 14) | (kSwizzle32 << 29);

  if (warp == 0) {
    // ---- TMA producer ----
    if (elect_one_sync()) {
      prefetch_tensormap(&p.tmap_x);
      mbar_arrive_expect_tx(w_full, p.wa_bytes + p.wb_bytes);
      bulk_load(wa0, p.wa, p.wa_bytes, w_full);
      bulk_load(wb0, p.wb, p.wb_bytes, w_full);
    }
    __syncwarp();
    pdl_wait();
    int slot = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img;
      const int r = tile - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      hd_wait(x_empty(slot), phase ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(x_full(slot), p.x_bytes);
        tma_load_5d(x0s + slot * p.x_bytes, &p.tmap_x, x_full(slot), 0, tx * 8, ty * 16, 0, img);
      }
      __syncwarp();
      if (++slot == kHdRing) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == 1 || warp == 18) {
    // ---- MMA issuers: warp 1 runs MMA-1 (needs an X tile and a free D1), warp 18 MMA-2 (needs E1's H1 tile and a
    // free D2).  Two warps so that neither GEMM queues behind the other one's wait.
    pdl_wait();
    if (warp == 1) pdl_release();
    hd_wait(w_full, 0);
    const uint32_t idesc = p.idesc;
    const uint32_t ba16 = (uint32_t)nmid * 2u;              // bytes/16 of one K block of Wa / Wb (nmid rows x 32 B)
    const uint32_t ones_lo = ((ones0 & 0x3FFFFu) >> 4) | (1u << 16);
    const int niter = p.ntiles > (int)blockIdx.x ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == 1) {
      int slot = 0; uint32_t xph = 0;
      for (int lt = 0; lt < niter; ++lt) {
        const int b = lt & 1;
        hd_wait(d1_empty(b), (((uint32_t)lt >> 1) & 1u) ^ 1u);
        hd_wait(x_full(slot), xph);
        fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t a_lo = (((x0s + slot * p.x_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t b_lo = ((wa0 & 0x3FFFFu) >> 4) | (1u << 16);
          for (int cb = 0; cb < p.in_blocks; ++cb)
            hd_mma(tmem_base + (uint32_t)(b * nmid), a_lo + cb * 256u, b_lo + cb * ba16, hi, idesc, cb ? 1u : 0u);
          hd_mma(tmem_base + (uint32_t)(b * nmid), ones_lo, ((bia0 & 0x3FFFFu) >> 4) | (1u << 16), hi, idesc, 1u);
          mma_commit(x_empty(slot));
          mma_commit(d1_full(b));
        }
        __syncwarp();
        if (++slot == kHdRing) { slot = 0; xph ^= 1u; }
      }
    } else {
      for (int lt = 0; lt < niter; ++lt) {
        const int b = lt & 1, b2 = lt % kHdD2Bufs;
        hd_wait(d2_empty(b2), ((uint32_t)(lt / kHdD2Bufs) & 1u) ^ 1u);
        hd_wait(h_full(b), ((uint32_t)lt >> 1) & 1u);
        fence_after_sync();
        if (elect_one_sync()) {
          const uint32_t a_lo = (((h0s + b * p.h_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t b_lo = ((wb0 & 0x3FFFFu) >> 4) | (1u << 16);
          for (int cb = 0; cb < p.mid_blocks; ++cb)
            hd_mma(tmem_base + (uint32_t)((2 + b2) * nmid), a_lo + cb * 256u, b_lo + cb * ba16, hi, idesc, cb ? 1u : 0u);
          hd_mma(tmem_base + (uint32_t)((2 + b2) * nmid), ones_lo, ((bib0 & 0x3FFFFu) >> 4) | (1u << 16), hi, idesc, 1u);
          mma_commit(h_empty(b));
          mma_commit(d2_full(b2));
        }
        __syncwarp();
      }
    }
  } else {
    // warps 2-9: stage E1, warps 10-17: stage E2; each stage has two groups of four warps that take
    // alternate tiles (group g <-> buffer g of D1 / H1 / D2), so both buffers are worked on at once
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int py = m >> 3, px = m & 7;
    const bool is_e1 = warp < 10;
    const int group = ((warp - 2) >> 2) & 1;
    pdl_wait();
    const int cb_lo = 0, cb_hi = p.mid_blocks;
    int lt = group;
    for (int tile = blockIdx.x + group * (int)gridDim.x; tile < p.ntiles; tile += 2 * (int)gridDim.x, lt += 2) {
      const int img = tile / tiles_per_img;
      const int r = tile - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y = ty * 16 + py, x = tx * 8 + px;
      const bool valid = y < p.x.H && x < p.x.W;          // edge tiles
      const int b = lt & 1;
      const uint32_t par = ((uint32_t)lt >> 1) & 1u;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
      if (is_e1) {
        // ---- E1: D1 -> H1 (shared-memory operand of MMA-2) ----
        // one warp of the group polls the mbarriers, the other three park on a hardware named barrier
        // (every mbarrier event wakes every polling warp of the CTA: fewer pollers, fewer wasted issue slots)
        if (quarter == 0) { hd_wait(d1_full(b), par); hd_wait(h_empty(b), par ^ 1u); }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
        fence_after_sync();
        const long long spix = SAVE ? (long long)img * p.save_a.sN + (long long)y * p.save_a.sY + (long long)x * p.save_a.sX : 0;
        uint8_t* const h_row = smem_gen + (h0s - smem0) + (size_t)b * p.h_bytes + (size_t)m * 32u;
        const uint32_t h_sw = ((h0s + (uint32_t)m * 32u) >> 7) & 1u;
        // two accumulator-read buffers in ping-pong: the read of block cb+1 is in flight while block cb is processed
        uint32_t ra[16], rb[16];
        auto e1_block = [&](int cb, const uint32_t* rv) {
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a0 = __uint_as_float(rv[4 * q]), a1 = __uint_as_float(rv[4 * q + 1]);      // bias already in the accumulator
              float a2 = __uint_as_float(rv[4 * q + 2]), a3 = __uint_as_float(rv[4 * q + 3]);
              // LeakyReLU with 0 <= slope <= 1 is max(a, slope * a): 1.5 instructions per element (packed multiply)
              lrelu_pair(a0, a1, p.slope); lrelu_pair(a2, a3, p.slope);
              __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h23 = __floats2bfloat162_rn(a2, a3);
              w[2 * q] = *reinterpret_cast<uint32_t*>(&h01);
              w[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h23);
            }
            // row m of K block cb: 32 bytes at [cb][m]; the two 16-byte chunks swap when address bit 7 is set
            // (a per-thread constant: buffers and blocks are multiples of 256 B apart)
            uint4* dst = reinterpret_cast<uint4*>(h_row + (size_t)cb * 4096u);
            dst[h_sw] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[h_sw ^ 1u] = make_uint4(w[4], w[5], w[6], w[7]);
            if (SAVE && valid) hd_st_global_32B((__nv_bfloat16*)p.save_a.ptr + spix + cb * p.save_a.sCb, w);
        };
        // D1 goes back to the MMA-1 warp as soon as its last block is in registers, before that block is processed
        auto release_d1 = [&]() {
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(d1_empty(b));
        };
        if (cb_lo < cb_hi) hd_ld16_issue(lane_base + (uint32_t)(b * nmid + cb_lo * 16), ra);
#pragma unroll 1
        for (int cb = cb_lo; cb < cb_hi; cb += 2) {
          hd_ld_wait();
          if (cb + 1 < cb_hi) hd_ld16_issue(lane_base + (uint32_t)(b * nmid + (cb + 1) * 16), rb);
          else release_d1();
          e1_block(cb, ra);
          if (cb + 1 < cb_hi) {
            hd_ld_wait();
            if (cb + 2 < cb_hi) hd_ld16_issue(lane_base + (uint32_t)(b * nmid + (cb + 2) * 16), ra);
            else release_d1();
            e1_block(cb + 1, rb);
          }
        }
        fence_proxy_async();                 // generic-proxy writes of H1 -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full(b));
      } else {
        // ---- E2: D2 -> nin_c -> fp32 NCHW ----
        const int b2 = lt % kHdD2Bufs;
        if (quarter == 0) hd_wait(d2_full(b2), (uint32_t)(lt / kHdD2Bufs) & 1u);
        asm volatile("bar.sync %0, 128;" ::"r"(3 + group) : "memory");
        fence_after_sync();
        const long long spix = SAVE ? (long long)img * p.save_b.sN + (long long)y * p.save_b.sY + (long long)x * p.save_b.sX : 0;
        float o[kHdMaxOut];
#pragma unroll
        for (int oc = 0; oc < kHdMaxOut; ++oc) o[oc] = s_bc[oc];
        uint32_t ra[16], rb[16];
        auto e2_block = [&](int cb, const uint32_t* rv) {
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a0 = __uint_as_float(rv[4 * q]), a1 = __uint_as_float(rv[4 * q + 1]);
              float a2 = __uint_as_float(rv[4 * q + 2]), a3 = __uint_as_float(rv[4 * q + 3]);
              // LeakyReLU with 0 <= slope <= 1 is max(a, slope * a): 1.5 instructions per element (packed multiply)
              lrelu_pair(a0, a1, p.slope); lrelu_pair(a2, a3, p.slope);
              v[4 * q] = a0; v[4 * q + 1] = a1; v[4 * q + 2] = a2; v[4 * q + 3] = a3;
            }
            if (SAVE) {
              // training pass: nin_c sees the bf16-rounded activation that is saved for the backward
              uint32_t w[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
                w[q] = *reinterpret_cast<uint32_t*>(&h);
                v[2 * q] = __uint_as_float(w[q] << 16); v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
              }
              if (valid) hd_st_global_32B((__nv_bfloat16*)p.save_b.ptr + spix + cb * p.save_b.sCb, w);
            }
#pragma unroll
            for (int oc = 0; oc < kHdMaxOut; ++oc) {
              if (oc < p.out_nc) {
                const float4* wr = reinterpret_cast<const float4*>(&s_wc[oc * 128 + cb * 16]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 wv = wr[q];
                  o[oc] += v[4 * q] * wv.x + v[4 * q + 1] * wv.y + v[4 * q + 2] * wv.z + v[4 * q + 3] * wv.w;
                }
              }
            }
        };
        if (cb_lo < cb_hi) hd_ld16_issue(lane_base + (uint32_t)((2 + b2) * nmid + cb_lo * 16), ra);
#pragma unroll 1
        for (int cb = cb_lo; cb < cb_hi; cb += 2) {
          hd_ld_wait();
          if (cb + 1 < cb_hi) hd_ld16_issue(lane_base + (uint32_t)((2 + b2) * nmid + (cb + 1) * 16), rb);
          e2_block(cb, ra);
          if (cb + 1 < cb_hi) {
            hd_ld_wait();
            if (cb + 2 < cb_hi) hd_ld16_issue(lane_base + (uint32_t)((2 + b2) * nmid + (cb + 2) * 16), ra);
            e2_block(cb + 1, rb);
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty(b2));
        const long long hw = (long long)p.x.H * p.x.W;
#pragma unroll
        for (int oc = 0; oc < kHdMaxOut; ++oc)
          if (oc < p.out_nc && valid) p.out_nchw[((long long)img * p.out_nc + oc) * hw + (long long)y * p.x.W + x] = o[oc];
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// Returns 0 when launched, kSgNotEligible when the geometry is not covered (caller runs the three
// layers one by one).
int launch_head_chain_umma(const HeadChain& h, cudaStream_t st) {
  static int attr_set = 0;
  { const char* e = getenv("N2N_NO_HEAD_FUSION"); if (e && atoi(e)) return kSgNotEligible; }
  if (h.in_blocks < 1 || h.in_blocks > 8 || h.mid_blocks < 1 || h.mid_blocks > 8 || h.out_nc < 1 || h.out_nc > kHdMaxOut)
    return kSgNotEligible;
  if (h.x.H < 4 || h.x.W < 4 || h.mid_channels != h.mid_blocks * 16) return kSgNotEligible;
  HdParams p;
  memset(&p, 0, sizeof(p));
  p.in_blocks = h.in_blocks; p.mid_blocks = h.mid_blocks; p.out_nc = h.out_nc;
  p.tiles_x = (h.x.W + 7) / 8; p.tiles_y = (h.x.H + 15) / 16;
  const long long tiles = (long long)h.x.N * p.tiles_x * p.tiles_y;
  N2N_CHECK_ARG(tiles > 0 && tiles < (1LL << 31), "head_chain: bad tile count");
  p.ntiles = (int)tiles;
  p.has_save = h.has_save ? 1 : 0; p.slope = h.slope;
  const int nmid = h.mid_blocks * 16;
  // packed 1x1 weights: [group][3 blocks][nmid rows][32 B]; block cb sits at cb * nmid * 32
  p.wa_bytes = (uint32_t)(((h.in_blocks + 2) / 3) * 3 * nmid * 32);
  p.wb_bytes = (uint32_t)(((h.mid_blocks + 2) / 3) * 3 * nmid * 32);
  p.x_bytes = (uint32_t)(h.in_blocks * 4096);
  p.h_bytes = (uint32_t)(h.mid_blocks * 4096);
  p.tmem_cols = tmem_cols_for((2 + kHdD2Bufs) * nmid);
  if ((2 + kHdD2Bufs) * nmid > 512) return kSgNotEligible;
  p.idesc = make_idesc_bf16(128, nmid, false, false);
  p.wa = (const uint8_t*)h.wa; p.wb = (const uint8_t*)h.wb;
  p.bias_a = h.bias_a; p.bias_b = h.bias_b; p.wc = h.wc; p.bias_c = h.bias_c;
  p.out_nchw = h.out_nchw; p.x = h.x; p.save_a = h.save_a; p.save_b = h.save_b;
  N2N_TRY(encode_c16_tensor_map(&p.tmap_x, h.x, 8, 16, h.in_blocks));
  const size_t smem = 1024 + (size_t)p.wa_bytes + p.wb_bytes + (size_t)kHdRing * p.x_bytes + 2 * (size_t)p.h_bytes +
                      4096 + 2 * (size_t)nmid * 32;
  if (smem > 220 * 1024) return kSgNotEligible;
  void (*kernel)(HdParams) = h.has_save ? head_chain_umma_kernel<true> : head_chain_umma_kernel<false>;
  if (!(attr_set & (h.has_save ? 2 : 1))) {
    N2N_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set |= h.has_save ? 2 : 1;
  }
  int nsm = 0, dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = kSMs;
  const int grid = tiles < nsm ? (int)tiles : nsm;
  N2N_CUDA(launch_pdl(kernel, dim3(grid), dim3(kHdThreads), smem, st, p));
  N2N_LAUNCH_CHECK();
  return 0;
}

}  // namespace n2n
