// headbwd_umma.cu — input gradients of the 1x1 head (nin_c -> nin_b -> nin_a, arch_unet.py:186-190 /
// :257-259) as ONE persistent tcgen05 kernel, the mirror image of head_umma.cu:
//
//   g_nb  = (Wc^T  g_out) * lrelu'(NB)        rank-out_nc outer product per pixel, CUDA cores (stage S0)
//   g_na  = (Wb^T  g_nb ) * lrelu'(NA)        GEMM-1 on the tensor core, masked in stage S1
//   g_d1b = (Wa^T  g_na ) * lrelu'(D1B)       GEMM-2, masked in stage S2
//
// Unfused these are three HBM-bound launches that each read a gradient tensor back that the previous
// one just wrote; fused, every gradient is written once (the weight-gradient kernels need all three)
// and never re-read, and the 96-channel intermediates feed the next GEMM straight from shared memory
// in the swizzled K-major operand layout.  LeakyReLU is in-place in the reference (arch_unet.py:113),
// so lrelu'(.) is taken from the sign of the saved activated outputs (slope > 0).
//
// Per 8 x 16-pixel tile: warp 1 issues the MMAs; warps 2-5 run S0, 6-9 S1, 10-13 S2 (thread = pixel =
// TMEM lane); H0 / H1 / D1 / D2 are double-buffered so the three stages work on consecutive tiles.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kHbThreads = 448;
constexpr int kHbMaxOut = 4;
constexpr int kHbBlocks = 6;              // the head is literally 96 channels wide (arch_unet.py:177-190)

struct HbParams {
  int blocks, out_nc;
  int tiles_x, tiles_y, ntiles, H, W;
  float slope;
  uint32_t wb_bytes, wa_bytes, h_bytes, tmem_cols, idesc;
  const uint8_t *wb, *wa;
  const float *wc, *gout;
  View act_nb, act_na, act_d1b;
  View g_nb, g_na, g_d1b;
};

__device__ __forceinline__ void hb_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
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
__device__ __forceinline__ void hb_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
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

// v[16] (fp32 gradient of one channel block of one pixel) * lrelu'(activation words) -> packed bf16
__device__ __forceinline__ void hb_mask_pack(const float v[16], const uint32_t mk[8], float slope, uint32_t w[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a0 = __uint_as_float(mk[j] << 16), a1 = __uint_as_float(mk[j] & 0xffff0000u);
    const float g0 = v[2 * j] * (a0 > 0.f ? 1.f : slope), g1 = v[2 * j + 1] * (a1 > 0.f ? 1.f : slope);
    __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
}

__global__ void __launch_bounds__(kHbThreads, 1)
head_bwd_umma_kernel(const __grid_constant__ HbParams p) {
  extern __shared__ uint8_t smem_raw[];
  // h0_full[2] h0_empty[2] d1_full[2] d1_empty[2] h1_full[2] h1_empty[2] d2_full[2] d2_empty[2] w_full
  __shared__ uint64_t bars[17];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_wc[kHbMaxOut * 128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  const uint32_t wb0 = smem0, wa0 = wb0 + p.wb_bytes;
  const uint32_t h00 = wa0 + p.wa_bytes, h10 = h00 + 2 * p.h_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int kind, int b) { return bar0 + 8u * (2 * kind + b); };
  enum { H0F = 0, H0E, D1F, D1E, H1F, H1E, D2F, D2E };
  const uint32_t w_full = bar0 + 8u * 16;
  const int nch = p.blocks * 16;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(H0F, b), 4); mbar_init(bar(H0E, b), 1);
      mbar_init(bar(D1F, b), 1); mbar_init(bar(D1E, b), 4);
      mbar_init(bar(H1F, b), 4); mbar_init(bar(H1E, b), 1);
      mbar_init(bar(D2F, b), 1); mbar_init(bar(D2E, b), 4);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < kHbMaxOut * 128; i += kHbThreads) {
    const int oc = i >> 7, c = i & 127;
    s_wc[i] = (oc < p.out_nc && c < nch) ? p.wc[oc * nch + c] : 0.f;
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t hi = (256u >> 4) | (1u << 14) | (kSwizzle32 << 29);
  const int niter = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_full, p.wb_bytes + p.wa_bytes);
      bulk_load(wb0, p.wb, p.wb_bytes, w_full);
      bulk_load(wa0, p.wa, p.wa_bytes, w_full);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer: GEMM-1 of tile i, then GEMM-2 of tile i-1 ----
    pdl_wait();
    pdl_release();
    hb_wait(w_full, 0);
    const uint32_t idesc = p.idesc;
    const uint32_t bsub16 = (uint32_t)nch * 2u;
    auto gemm = [&](int lt, uint32_t h_base, uint32_t w_base, int d_slot, int kind_hfull, int kind_dempty, int kind_hempty,
                    int kind_dfull) {
      const int b = lt & 1;
      const uint32_t par = ((uint32_t)lt >> 1) & 1u;
      hb_wait(bar(kind_dempty, b), par ^ 1u);
      hb_wait(bar(kind_hfull, b), par);
      fence_after_sync();
      if (elect_one_sync()) {
        const uint32_t a_lo = (((h_base + b * p.h_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo = ((w_base & 0x3FFFFu) >> 4) | (1u << 16);
        for (int cb = 0; cb < p.blocks; ++cb)
          hb_mma(tmem_base + (uint32_t)((d_slot + b) * nch), a_lo + cb * 256u, b_lo + cb * bsub16, hi, idesc, cb ? 1u : 0u);
        mma_commit(bar(kind_hempty, b));
        mma_commit(bar(kind_dfull, b));
      }
      __syncwarp();
    };
    for (int lt = 0; lt <= niter; ++lt) {
      if (lt < niter) gemm(lt, h00, wb0, 0, H0F, D1E, H0E, D1F);
      if (lt >= 1) gemm(lt - 1, h10, wa0, 2, H1F, D2E, H1E, D2F);
    }
  } else if (warp < 14) {
    const int stage = (warp - 2) >> 2;                 // 0: S0, 1: S1, 2: S2
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int py = m >> 3, px = m & 7;
    const View& act = stage == 0 ? p.act_nb : (stage == 1 ? p.act_na : p.act_d1b);
    const View& gdst = stage == 0 ? p.g_nb : (stage == 1 ? p.g_na : p.g_d1b);
    pdl_wait();
    for (int lt = 0; lt < niter; ++lt) {
      const int tile = (int)blockIdx.x + lt * (int)gridDim.x;
      const int img = tile / tiles_per_img;
      const int r = tile - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y = ty * 16 + py, x = tx * 8 + px;
      const bool valid = y < p.H && x < p.W;
      const int b = lt & 1;
      const uint32_t par = ((uint32_t)lt >> 1) & 1u;
      const long long apix = (long long)img * act.sN + (long long)y * act.sY + (long long)x * act.sX;
      const long long gpix = (long long)img * gdst.sN + (long long)y * gdst.sY + (long long)x * gdst.sX;
      // the activation words that give lrelu' for this pixel: issue all loads before waiting on the pipeline
      uint32_t mk[kHbBlocks][8];
#pragma unroll
      for (int cb = 0; cb < kHbBlocks; ++cb) {
        if (valid) hb_ld_global_32B((const __nv_bfloat16*)act.ptr + apix + cb * act.sCb, mk[cb]);
        else {
#pragma unroll
          for (int q = 0; q < 8; ++q) mk[cb][q] = 0u;
        }
      }
      if (stage == 0) {
        float go[kHbMaxOut];
        const long long hw = (long long)p.H * p.W;
#pragma unroll
        for (int oc = 0; oc < kHbMaxOut; ++oc)
          go[oc] = (oc < p.out_nc && valid) ? p.gout[((long long)img * p.out_nc + oc) * hw + (long long)y * p.W + x] : 0.f;
        if (quarter == 0) hb_wait(bar(H0E, b), par ^ 1u);
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int cb = 0; cb < kHbBlocks; ++cb) {
          {
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = 0.f;
#pragma unroll
            for (int oc = 0; oc < kHbMaxOut; ++oc) {
              if (oc < p.out_nc) {
                const float4* wr = reinterpret_cast<const float4*>(&s_wc[oc * 128 + cb * 16]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 wv = wr[q];
                  v[4 * q] += wv.x * go[oc]; v[4 * q + 1] += wv.y * go[oc]; v[4 * q + 2] += wv.z * go[oc]; v[4 * q + 3] += wv.w * go[oc];
                }
              }
            }
            uint32_t w[8];
            hb_mask_pack(v, mk[cb], p.slope, w);
            const uint32_t off = (uint32_t)b * p.h_bytes + (uint32_t)cb * 4096u + (uint32_t)m * 32u;
            const uint32_t sw = ((h00 + off) >> 7) & 1u;
            uint4* dst = reinterpret_cast<uint4*>(smem_gen + (h00 - smem0) + off);
            dst[sw] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[sw ^ 1u] = make_uint4(w[4], w[5], w[6], w[7]);
            if (valid) hb_st_global_32B((__nv_bfloat16*)gdst.ptr + gpix + cb * gdst.sCb, w);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(H0F, b));
      } else {
        const int dfull = stage == 1 ? D1F : D2F, dempty = stage == 1 ? D1E : D2E;
        if (quarter == 0) {
          hb_wait(bar(dfull, b), par);
          if (stage == 1) hb_wait(bar(H1E, b), par ^ 1u);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + stage) : "memory");
        fence_after_sync();
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(((stage == 1 ? 0 : 2) + b) * nch);
#pragma unroll
        for (int cb = 0; cb < kHbBlocks; ++cb) {
          {
            uint32_t rr[16];
            hb_ld16(lane_addr + cb * 16, rr);
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rr[q]);
            uint32_t w[8];
            hb_mask_pack(v, mk[cb], p.slope, w);
            if (stage == 1) {
              const uint32_t off = (uint32_t)b * p.h_bytes + (uint32_t)cb * 4096u + (uint32_t)m * 32u;
              const uint32_t sw = ((h10 + off) >> 7) & 1u;
              uint4* dst = reinterpret_cast<uint4*>(smem_gen + (h10 - smem0) + off);
              dst[sw] = make_uint4(w[0], w[1], w[2], w[3]);
              dst[sw ^ 1u] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            if (valid) hb_st_global_32B((__nv_bfloat16*)gdst.ptr + gpix + cb * gdst.sCb, w);
          }
        }
        fence_before_sync();
        if (stage == 1) fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(dempty, b));
          if (stage == 1) mbar_arrive(bar(H1F, b));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// Returns 0 when launched, kSgNotEligible when the geometry is not covered (caller runs the three
// input-gradient launches one by one).
int launch_head_bwd_umma(const HeadBwd& h, cudaStream_t st) {
  static bool attr_set = false;
  { const char* e = getenv("N2N_NO_HEAD_FUSION"); if (e && atoi(e)) return kSgNotEligible; }
  if (h.blocks != kHbBlocks || h.channels != h.blocks * 16 || h.out_nc < 1 || h.out_nc > kHbMaxOut) return kSgNotEligible;
  if (h.g_d1b.H < 4 || h.g_d1b.W < 4) return kSgNotEligible;
  HbParams p;
  memset(&p, 0, sizeof(p));
  p.blocks = h.blocks; p.out_nc = h.out_nc; p.H = h.g_d1b.H; p.W = h.g_d1b.W;
  p.tiles_x = (p.W + 7) / 8; p.tiles_y = (p.H + 15) / 16;
  const long long tiles = (long long)h.g_d1b.N * p.tiles_x * p.tiles_y;
  N2N_CHECK_ARG(tiles > 0 && tiles < (1LL << 31), "head_bwd: bad tile count");
  p.ntiles = (int)tiles;
  p.slope = h.slope;
  const int nch = h.blocks * 16;
  if (4 * nch > 512) return kSgNotEligible;
  p.wb_bytes = p.wa_bytes = (uint32_t)(((h.blocks + 2) / 3) * 3 * nch * 32);
  p.h_bytes = (uint32_t)(h.blocks * 4096);
  p.tmem_cols = tmem_cols_for(4 * nch);
  p.idesc = make_idesc_bf16(128, nch, false, false);
  p.wb = (const uint8_t*)h.wb_dgrad; p.wa = (const uint8_t*)h.wa_dgrad; p.wc = h.wc; p.gout = h.gout;
  p.act_nb = h.act_nb; p.act_na = h.act_na; p.act_d1b = h.act_d1b;
  p.g_nb = h.g_nb; p.g_na = h.g_na; p.g_d1b = h.g_d1b;
  const size_t smem = 1024 + (size_t)p.wb_bytes + p.wa_bytes + 4 * (size_t)p.h_bytes;
  if (smem > 200 * 1024) return kSgNotEligible;
  if (!attr_set) {
    N2N_CUDA(cudaFuncSetAttribute(head_bwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int nsm = 0, dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = kSMs;
  const int grid = tiles < nsm ? (int)tiles : nsm;
  N2N_CUDA(launch_pdl(head_bwd_umma_kernel, dim3(grid), dim3(kHbThreads), smem, st, p));
  N2N_LAUNCH_CHECK();
  return 0;
}

}  // namespace n2n
