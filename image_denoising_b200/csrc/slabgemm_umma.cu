// slabgemm_umma.cu — second-generation bf16 tensor-core engine for the tap GEMM (conv3x3 / conv1x1 /
// ConvTranspose2x2, forward and input gradient).  Same contract as tapgemm_umma.cu (TapGemm in,
// C16 tensor out), different geometry:
//
//   * output tile = 8 x 16 pixels (M = 128, pixel m = 8*y + x).  One TMA box brings the tile's whole
//     halo'd input window — 10 x 18 pixels x <= 3 channel blocks — into shared memory ONCE, and all
//     nine taps of a 3x3 conv are tcgen05 operand views of that one box: tap (dy, dx) starts
//     (10*dy + dx) * 32 B into it, the eight pixels of an image row are the eight rows of a
//     SWIZZLE_32B group, and the descriptor's stride-byte-offset (320 B) walks the image rows.
//     Every input pixel crosses L2 -> SMEM 1.4 times per conv (9 in a tap-by-tap implicit GEMM,
//     3.2 in the row-slab engine).
//   * the packed weights of the layer stay resident in shared memory for the whole persistent CTA;
//   * the per-tile MMA schedule (A / B descriptor offsets of every tap) is tabulated in shared
//     memory once per CTA and pulled into registers per stage, so the issuing lane does two integer
//     adds per tcgen05.mma;
//   * default form: clusters of two CTAs and tcgen05 cta_group::2 (see the kernel's comment) — each
//     SM keeps half of the weight rows, which also makes the Cin = 144 layers fit;
//   * accumulators are triple-buffered in TMEM when 3 * N <= 512 columns (else double); one group of
//     four epilogue warps per accumulator takes every nbuf-th tile; the bias rides in the GEMM as one
//     more K block (a constant ones block x bias rows split into bf16 hi + lo) when shared memory
//     allows, else it is added here; the groups fuse LeakyReLU / ReLU, the input-gradient's activation mask and skip-gradient
//     addend (prefetched), the 2x2 max-pool (a 2x2 cell is lanes {l, l^1, l^8} of one warp -> two
//     shuffles), the 256-bit bf16 C16 store (two adjacent output pixels for the pair-form
//     ConvTranspose) and the fp32 NCHW store of the network head; one warp per group polls the
//     mbarrier, the others park on a named barrier;
//   * images need not be multiples of the tile: out-of-image rows / columns are zero-filled by TMA
//     on the way in and masked on the way out.
//
// Only images smaller than 4 x 4 (and N2N_NO_SLAB=1) return kSgNotEligible and run on the
// first-generation row-slab engine (tapgemm_umma.cu).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kSgEpiGroups = 3;           // groups of four epilogue warps, each takes every third tile
constexpr int kSgThreads = 64 + 128 * kSgEpiGroups + 32;   // TMA warp, MMA warp, epilogue warps, second MMA issuer (dual mode)
constexpr int kSgIssuer2 = 2 + 4 * kSgEpiGroups;           // warp index of the second issuer
constexpr int kSgMaxStages = 12;          // pipeline stages (TMA boxes) per tile
constexpr int kSgMaxRing = 8;
constexpr int kSgGroup = 3;               // channel blocks per stage = one packed-weight group
constexpr int kTileW = 8, kTileH = 16;
constexpr int kHaloW = kTileW + 2, kHaloH = kTileH + 2;
constexpr size_t kSgSmemMax = 232448;     // 227 KB per CTA (static + dynamic)
// CTA-pair (cta_group::2) form, N2N_PAIR bit 0: N <= 64 layers, bit 1: wider layers.  Measured at
// 32x256x256: dec_conv1b 274 -> 268 us (MMA phase alone 249 -> 203 us), enc_conv1 144 -> 139 us, and the
// Cin = 144 convs — whose 249 KB of weights only fit when split over the pair — 177 -> 93 us (1.40 PFLOP/s).
constexpr int kSgPairDefault = 3;
static int sg_pair_mode() {
  static const char* const e = getenv("N2N_PAIR");
  return e ? atoi(e) : kSgPairDefault;
}
static long long sg_pair_min_tiles() {
  static long long v = -1;
  if (v < 0) { const char* e = getenv("N2N_PAIR_MIN_TILES"); v = e ? atoll(e) : 2; }
  return v;
}
static bool sg_use_pair(int nout, long long tiles) {
  const int mode = sg_pair_mode();
  return tiles >= sg_pair_min_tiles() && nout % 16 == 0 && ((nout <= 64 && (mode & 1)) || (nout > 64 && (mode & 2)));
}
constexpr size_t kSgStaticSlack = 6144;   // static shared memory of the kernel, rounded up

struct SgStage {
  int16_t view, cb0, nb, ox, oy, ntaps;
  uint16_t cb_bytes16;       // bytes/16 between channel blocks inside the staged box
  uint16_t sbo16;            // bytes/16 between image rows of the box (A descriptor stride-byte-offset)
  uint16_t a_off16[9];       // start of each tap's operand view inside the box (bytes/16)
  uint16_t first;            // bit t: tap t is the first MMA into its accumulator columns (stage 0 only)
  uint8_t col16[9];          // accumulator column (/16) each tap's MMA starts at (column-range form, else 0)
  uint8_t pad8[3];
  uint32_t b_off16[9];       // start of each tap's weight slab inside the resident weights (bytes/16)
  uint32_t tx_bytes;         // bytes the TMA box delivers
  uint32_t nt0;              // taps [0, nt0) start at accumulator column 0 (issuer 0), [nt0, ntaps) at column mma_n (issuer 1)
};

struct SgParams {
  // hot (epilogue / loop) fields first: they stay in the first constant-cache lines
  int nst, nout, ring, dbg_flags, nbuf;
  int dual;                  // column-range form with two ranges: warp kSgIssuer2 issues the taps of the upper range (disjoint
                             // accumulator columns, so the two issuers need no ordering between them)
  int esplit;                // 3: every tile's accumulator columns are drained by ALL THREE epilogue groups (a third each), used
                             // when only two accumulators fit in TMEM (N = 144 / 192): with one group per tile the MMAs of tile
                             // i+2 wait for the whole 12-block epilogue of tile i (ncu: tensor pipe 35 % active, third group idle)
  int mma_n, up_py;          // MMA width (= nout unless the launch is in column-range form); fused up-conv: output-row parity
  const float* corr;         // fused up-conv: border bias correction [9][mma_n] (see TapGemm::border_corr)
  int tiles_x, tiles_y, ntiles;
  int act; float slope;
  int store_y, has_addend, has_mask, has_pool, out_c, n_split;
  long long split_stride;
  uint32_t slot_bytes, tmem_cols, idesc, w_bytes, w_region;
  uint32_t bias_off;         // > 0: bias rides in the GEMM (ones block at smem0 + bias_off, bias rows 4 KB behind it)
  const uint8_t* w;
  const float* bias;
  float* out_nchw;
  View y, addend, mask, pool;
  CUtensorMap tmap[kTapViews];
  SgStage st[kSgMaxStages];
};

// Bounded wait without clock reads: a protocol bug becomes a trap, never a hang.
__device__ __forceinline__ void sg_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)     // suspend-time hint (ns): sleep in hardware instead of spinning
        : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ void sg_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      );
}

// ---- CTA-pair (cta_group::2) primitives: two CTAs of a cluster compute one 256-row tile pair; each
// holds its own 128 activation rows and HALF of the weight rows, the leader (rank 0) issues the MMAs.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> rank 0's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sg_mma2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate));
}
// arrive on the same barrier of BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void mma_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// TMA load whose completion bytes are credited to the LEADER's barrier (bar already peer-masked)
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst_smem, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Relaxed form for the "accumulator drained" signal: it hands over no memory (the TMEM reads are ordered
// by tcgen05.wait::ld + tcgen05.fence::before_thread_sync), and a release at cluster scope would make
// the epilogue wait for all of its global stores to drain first (measured: 16 % of the kernel in ERRBAR).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t result_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void sg_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// One pixel's 16-channel block (32 B) moves as a single 256-bit access (full 32 B sectors).
__device__ __forceinline__ void st_global_32B(void* ptr, const uint32_t w[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_32B(const void* ptr, uint32_t w[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(ptr)
               : "memory");
}
__device__ __forceinline__ void unpack_bf16x16(const uint32_t w[8], float v[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[2 * j] = __uint_as_float(w[j] << 16);
    v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

struct SgPix {
  int img, y, x;
  bool valid;          // pixel inside the image (edge tiles of images that are not multiples of 8 x 16)
  int cls0, cls1;      // fused up-conv: border class of the two output pixels of the column halves (0 = interior)
  long long ypix, apix, mpix, ppix;
};

// Stores of one packed 16-channel block: the activation (when the launch keeps it) and the 2x2 max-pooled activation.
__device__ __forceinline__ void sg_store_block(const SgParams& p, const SgPix& c, int cb, int lane, uint32_t w[8],
                                               bool keep_y = true) {
  if (p.store_y && c.valid && keep_y) {
    long long o = c.ypix;
    int cbr = cb;
    if (p.n_split && cb >= p.n_split) { cbr = cb - p.n_split; o += p.split_stride; }
    st_global_32B((__nv_bfloat16*)p.y.ptr + o + cbr * p.y.sCb, w);
  }
  if (p.has_pool) {
    // 2x2 max over lanes {l, l^1 (x neighbour), l^8 (y neighbour)}; max commutes with bf16 rounding
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&w[j]);
      uint32_t o = __shfl_xor_sync(0xffffffffu, w[j], 1);
      a = __hmax2(a, *reinterpret_cast<__nv_bfloat162*>(&o));
      uint32_t aw = *reinterpret_cast<uint32_t*>(&a);
      o = __shfl_xor_sync(0xffffffffu, aw, 8);
      a = __hmax2(a, *reinterpret_cast<__nv_bfloat162*>(&o));
      w[j] = *reinterpret_cast<uint32_t*>(&a);
    }
    if ((lane & 9) == 0 && c.valid) st_global_32B((__nv_bfloat16*)p.pool.ptr + c.ppix + cb * p.pool.sCb, w);
  }
}

// One 16-channel block of one pixel: accumulator -> bias -> (+addend) -> act -> (*mask) -> stores.
// `aux` (when non-null) holds the block's mask words (or addend words when there is no mask), loaded
// from global memory before the accumulator wait so that their DRAM latency is off the critical path.
// PLAIN (compile time): the launch has none of bias-in-epilogue / border correction / addend / mask / NCHW output / debug
// flags — every forward convolution whose bias rides in the GEMM.  The generic form tests each of them per block with a
// uniform branch (constant load + compare + branch, ~20 instructions per block on issue-slot-bound layers).
template <bool PLAIN>
__device__ __forceinline__ void sg_epilogue_block(const SgParams& p, const SgPix& c, int cb, const uint32_t r[16],
                                                  int lane, const float* s_bias, const uint32_t* aux) {
  float v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
  if (PLAIN) {
    if (p.act) {
#pragma unroll
      for (int q = 0; q < 16; q += 2) lrelu_pair(v[q], v[q + 1], p.slope);
    }
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    sg_store_block(p, c, cb, lane, w);
    return;
  }
  if (!(p.dbg_flags & 8) && !p.bias_off) {
    const float4* b4 = reinterpret_cast<const float4*>(s_bias + cb * 16);   // shared-memory broadcast
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = b4[q];
      add_pair(v[4 * q], v[4 * q + 1], b.x, b.y); add_pair(v[4 * q + 2], v[4 * q + 3], b.z, b.w);
    }
  }
  if (p.corr) {
    const bool hi = p.n_split && cb >= p.n_split;
    const int cls = hi ? c.cls1 : c.cls0;
    if (cls) {                                             // image-border pixel: drop the out-of-image taps' ConvTranspose bias
      const float4* c4 = reinterpret_cast<const float4*>(p.corr + cls * p.mma_n + (hi ? cb - p.n_split : cb) * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(c4 + q);
        v[4 * q] -= b.x; v[4 * q + 1] -= b.y; v[4 * q + 2] -= b.z; v[4 * q + 3] -= b.w;
      }
    }
  }
  if (p.has_addend) {
    uint32_t aw[8]; float t[16];
    if (aux && !p.has_mask) {
#pragma unroll
      for (int q = 0; q < 8; ++q) aw[q] = aux[q];
    } else if (c.valid) {
      ld_global_32B((const __nv_bfloat16*)p.addend.ptr + c.apix + cb * p.addend.sCb, aw);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) aw[q] = 0u;
    }
    unpack_bf16x16(aw, t);
#pragma unroll
    for (int q = 0; q < 16; q += 2) add_pair(v[q], v[q + 1], t[q], t[q + 1]);
  }
  if (p.act && !(p.dbg_flags & 16)) {
#pragma unroll
    for (int q = 0; q < 16; q += 2) lrelu_pair(v[q], v[q + 1], p.slope);     // LeakyReLU / ReLU, 0 <= slope <= 1
  }
  if (p.has_mask) {
    uint32_t aw[8];
    if (aux) {
#pragma unroll
      for (int q = 0; q < 8; ++q) aw[q] = aux[q];
    } else if (c.valid) {
      ld_global_32B((const __nv_bfloat16*)p.mask.ptr + c.mpix + cb * p.mask.sCb, aw);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) aw[q] = 0u;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) lrelu_mask_pair(v[2 * j], v[2 * j + 1], aw[j], p.slope);
  }
  if (p.out_nchw) {
    const long long hw = (long long)p.y.H * p.y.W;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int n = cb * 16 + q;
      if (n < p.out_c && c.valid) p.out_nchw[((long long)c.img * p.out_c + n) * hw + (long long)c.y * p.y.W + c.x] = v[q];
    }
    return;
  }
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  sg_store_block(p, c, cb, lane, w, !(p.dbg_flags & 32));       // flag 32 (diagnostic): no activation store
}

// CG = 1: one CTA per tile.  CG = 2: clusters of two CTAs, tcgen05 cta_group::2 — rank r of a pair owns
// tile 2*pair + r (its activations, its accumulator, its epilogue) and rows [r*N/2, (r+1)*N/2) of every
// weight sub-tile; per MMA each SM then reads 4 KB of A + 16*N B of B instead of 4 KB + 32*N B.
template <int CG>
__global__ void __launch_bounds__(kSgThreads, 1)
slabgemm_umma_kernel(const __grid_constant__ SgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * kSgMaxRing + 8];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) uint2 s_tap[kSgMaxStages * 10];  // per (stage, tap): (A offset, B offset) in 16-byte units
  __shared__ int4 s_ld[kSgMaxStages];                 // per stage: view, cb0, ox | oy << 16, tx_bytes
  __shared__ uint4 s_mm[kSgMaxStages];
  __shared__ __align__(16) float s_bias[256];                // per stage: ntaps, nb, cb_bytes16, A descriptor high word

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slots0 = smem0 + p.w_region;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kSgMaxRing + s); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kSgMaxRing);
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kSgMaxRing + 1 + b); };      // up to three accumulators
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kSgMaxRing + 4 + b); };
  const uint32_t wready_bar = bar0 + 8u * (2 * kSgMaxRing + 7);   // CG = 2: both CTAs' weight halves have landed
  const uint32_t b_sub16 = (uint32_t)p.mma_n * 2u / CG;  // weight rows per CTA x 32 B, in 16-byte units
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSgMaxRing; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), p.dual ? 2 : 1); }
    mbar_init(wfull_bar, 1);
    mbar_init(wready_bar, CG);
    for (int b = 0; b < 3; ++b) { mbar_init(tfull_bar(b), p.dual ? 2 : 1); mbar_init(tempty_bar(b), 4 * CG * (p.esplit == 3 ? 3 : 1)); }
    fence_barrier_init();
  }
  // per-CTA schedule tables
  for (int i = threadIdx.x; i < p.nst * 10; i += kSgThreads) {
    const int s = i / 10, t = i - s * 10;
    const SgStage& S = p.st[s];
    s_tap[i] = t < S.ntaps ? make_uint2((uint32_t)S.a_off16[t] | ((uint32_t)S.col16[t] << 20),
                                        (S.b_off16[t] / CG) | (((uint32_t)(S.first >> t) & 1u) << 31))
                           : make_uint2(0u, 0u);
  }
  for (int i = threadIdx.x; i < p.nout; i += kSgThreads)
    s_bias[i] = p.bias ? p.bias[(p.n_split && i >= p.n_split * 16) ? i - p.n_split * 16 : i] : 0.f;
  if (p.bias_off) {
    // Bias through the tensor core: one extra K block per tile whose A operand is constant (elements 0, 1 of every
    // row = 1) and whose B rows carry the bias split into bf16 hi + lo parts (relative error 2^-17) — the epilogue
    // saves 4 LDS + 16 FADD per channel block, which is what bounds the narrow layers.
    uint8_t* base = smem_raw + (smem0 - smem_u32(smem_raw));
    const int nrows = p.mma_n / CG;                        // this CTA's half of the B rows in pair form
    for (int r = threadIdx.x; r < 128 + nrows; r += kSgThreads) {
      const uint32_t off = r < 128 ? p.bias_off + (uint32_t)r * 32u : p.bias_off + 4096u + (uint32_t)(r - 128) * 32u;
      float v = 1.0f;
      if (r >= 128) {
        const int n = (int)rank * nrows + (r - 128);
        v = p.bias[(p.n_split && n >= p.n_split * 16) ? n - p.n_split * 16 : n];
      }
      const __nv_bfloat16 hi_part = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo_part = r < 128 ? hi_part : __float2bfloat16_rn(v - __bfloat162float(hi_part));
      const uint32_t sw = ((smem0 + off) >> 7) & 1u;       // logical 16-byte chunk 0 sits in physical chunk sw
      uint4* row = reinterpret_cast<uint4*>(base + off);
      __nv_bfloat162 h2 = __halves2bfloat162(hi_part, lo_part);
      row[sw] = make_uint4(*reinterpret_cast<uint32_t*>(&h2), 0u, 0u, 0u);
      row[sw ^ 1u] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  if (threadIdx.x < p.nst) {
    const SgStage& S = p.st[threadIdx.x];
    s_ld[threadIdx.x] = make_int4(S.view, S.cb0, (int)(uint16_t)S.ox | ((int)(uint16_t)S.oy << 16), (int)S.tx_bytes);
    s_mm[threadIdx.x] = make_uint4((uint32_t)S.ntaps | (S.nt0 << 16), (uint32_t)S.nb, (uint32_t)S.cb_bytes16,
                                   (uint32_t)S.sbo16 | (1u << 14) | (kSwizzle32 << 29));
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc2(smem_u32(&tmem_base_smem), p.tmem_cols);
    } else {
      tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols);
      tmem_relinquish();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // the peer's barriers exist before anything signals them
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int first_tile = (int)(blockIdx.x / CG) * CG + (int)rank, tile_step = (int)gridDim.x;   // == blockIdx.x
  const int npair_iters = (p.ntiles + tile_step - 1 - (int)(blockIdx.x / CG) * CG) / tile_step;   // same trip count for both ranks

  if (warp == 0) {
    // ---- TMA producer (whole warp walks the loop, one elected lane issues) ----
    if (elect_one_sync()) {
      for (int v = 0; v < kTapViews; ++v) prefetch_tensormap(&p.tmap[v]);
      if (CG == 2) {
        // this CTA's half of the rows of every weight sub-tile ([nout][32 B] each)
        const uint32_t sub = (uint32_t)p.mma_n * 32u, half = sub / 2u;
        mbar_arrive_expect_tx(wfull_bar, p.w_bytes / 2u);
        for (uint32_t k = 0; k * sub < p.w_bytes; ++k)
          bulk_load(smem0 + k * half, p.w + (size_t)k * sub + rank * half, half, wfull_bar);
      } else {
        mbar_arrive_expect_tx(wfull_bar, p.w_bytes);
        for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
          const uint32_t n = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
          bulk_load(smem0 + off, p.w + off, n, wfull_bar);
        }
      }
    }
    __syncwarp();
    pdl_wait();                         // activations below are the predecessor's output
    int slot = 0; uint32_t phase = 0;
    for (int it = 0; it < npair_iters; ++it) {
      int tile = first_tile + it * tile_step;
      if (tile >= p.ntiles) tile = p.ntiles - 1;           // odd tile count: the peer reloads the last tile, its epilogue is masked
      const int img = tile / tiles_per_img;
      const int r = tile - img * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int x0 = tx * kTileW, y0 = ty * kTileH;
      for (int s = 0; s < p.nst; ++s) {
        const int4 ld = s_ld[s];
        sg_wait(empty_bar(slot), phase ^ 1u);
        if (elect_one_sync()) {
          if (p.dbg_flags & 1) {
            mbar_arrive(full_bar(slot));
          } else if (CG == 2) {
            // both CTAs' boxes are credited to the leader's barrier, which expects twice the bytes
            if (rank == 0) mbar_arrive_expect_tx(full_bar(slot), 2u * (uint32_t)ld.w);
            tma_load_5d_pair(slots0 + slot * p.slot_bytes, &p.tmap[ld.x], full_bar(slot) & kPeerMask, 0,
                             x0 + (int)(int16_t)(ld.z & 0xffff), y0 + (int)(int16_t)(ld.z >> 16), ld.y, img);
          } else {
            mbar_arrive_expect_tx(full_bar(slot), (uint32_t)ld.w);
            tma_load_5d(slots0 + slot * p.slot_bytes, &p.tmap[ld.x], full_bar(slot), 0, x0 + (int)(int16_t)(ld.z & 0xffff),
                        y0 + (int)(int16_t)(ld.z >> 16), ld.y, img);
          }
        }
        __syncwarp();
        if (++slot == p.ring) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 || (warp == kSgIssuer2 && p.dual)) {
    // ---- MMA issuer(s) ----
    const uint32_t q = warp == 1 ? 0u : 1u;
    int slot = 0; uint32_t phase = 0;
    pdl_wait();
    if (q == 0) pdl_release();          // our own dependents may begin their prologue
    sg_wait(wfull_bar, 0);
    if (CG == 2) {
      // tell the leader that this CTA's weight half is in place; only the leader issues MMAs
      if (q == 0) {
        if (elect_one_sync()) mbar_arrive_cluster(wready_bar & kPeerMask);
        __syncwarp();
      }
      if (rank == 0) sg_wait(wready_bar, 0);
    }
    const uint32_t b_hi = (256u >> 4) | (1u << 14) | (kSwizzle32 << 29);
    const uint32_t w_lo = ((smem0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t idesc = p.idesc;
    const bool skip = (p.dbg_flags & 2) != 0;
    for (int lt = 0; lt < ((CG == 2 && rank != 0) ? 0 : npair_iters); ++lt) {
      const int buf = lt % p.nbuf;
      sg_wait(tempty_bar(buf), ((uint32_t)(lt / p.nbuf) & 1u) ^ 1u);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.nout);
      uint32_t acc = 0;
      for (int s = 0; s < p.nst; ++s) {
        const uint4 mm = s_mm[s];                            // ntaps | nt0 << 16, nb, cb_bytes16, a_hi
        // dual mode: issuer 0 takes the taps of the lower column range [0, nt0), issuer 1 the rest
        const int nt_all = (int)(mm.x & 0xffffu), nt0 = (int)(mm.x >> 16);
        const int t_begin = q ? nt0 : 0, t_end = (p.dual && !q) ? nt0 : nt_all;
        const uint32_t a_lo = (((slots0 + slot * p.slot_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
        // the stage's tap table goes to registers up front (5 LDS.128), so the issue loop below is
        // two integer adds per tcgen05.mma
        uint2 tp[10];
        {
          const uint4* t4 = reinterpret_cast<const uint4*>(&s_tap[s * 10]);
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const uint4 t = t4[q];
            tp[2 * q] = make_uint2(t.x, t.y); tp[2 * q + 1] = make_uint2(t.z, t.w);
          }
        }
        sg_wait(full_bar(slot), phase);
        fence_after_sync();
        if (elect_one_sync()) {
          if (!skip) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t >= t_begin && t < t_end) {
                // x: A offset (bits 0..19) | accumulator column (bits 20..); y: B offset | "first MMA into these columns"
                const uint32_t al = a_lo + (tp[t].x & 0xFFFFFu), bl = w_lo + (tp[t].y & 0x7FFFFFFFu);
                const uint32_t dt = d_tmem + (tp[t].x >> 20) * 16u;
                const uint32_t a1 = acc | ((tp[t].y >> 31) ^ 1u);
                if (CG == 2) {
                  sg_mma2(dt, al, mm.w, bl, b_hi, idesc, a1);
                  if (mm.y > 1) sg_mma2(dt, al + mm.z, mm.w, bl + b_sub16, b_hi, idesc, 1);
                  if (mm.y > 2) sg_mma2(dt, al + 2 * mm.z, mm.w, bl + 2 * b_sub16, b_hi, idesc, 1);
                } else {
                  sg_mma(dt, al, mm.w, bl, b_hi, idesc, a1);
                  if (mm.y > 1) sg_mma(dt, al + mm.z, mm.w, bl + b_sub16, b_hi, idesc, 1);
                  if (mm.y > 2) sg_mma(dt, al + 2 * mm.z, mm.w, bl + 2 * b_sub16, b_hi, idesc, 1);
                }
              }
            }
          }
          if (CG == 2) mma_commit2(empty_bar(slot)); else mma_commit(empty_bar(slot));   // frees the slot in both CTAs
        }
        __syncwarp();
        acc = 1;
        if (++slot == p.ring) { slot = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) {
        if (p.bias_off && !skip) {
          const uint32_t one_lo = (((smem0 + p.bias_off) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t bia_lo = (((smem0 + p.bias_off + 4096u) & 0x3FFFFu) >> 4) | (1u << 16);
          // column-range form: the same bias rows serve every column range (dual mode: each issuer its own range)
          for (int col = p.dual ? (int)q * p.mma_n : 0; col < (p.dual ? ((int)q + 1) * p.mma_n : p.nout); col += p.mma_n) {
            if (CG == 2) sg_mma2(d_tmem + col, one_lo, b_hi, bia_lo, b_hi, idesc, 1); else sg_mma(d_tmem + col, one_lo, b_hi, bia_lo, b_hi, idesc, 1);
          }
        }
        if (CG == 2) mma_commit2(tfull_bar(buf)); else mma_commit(tfull_bar(buf));
      }
      __syncwarp();
    }
  } else if (warp >= 2 && warp < kSgIssuer2) {
    // ---- epilogue: warp w owns TMEM lanes [32*(w%4), +32) = image rows 4*(w%4) .. +3 of the tile;
    //      groups of four warps take tiles round-robin (group g <-> accumulator buffer g), so the
    //      per-tile bookkeeping is paid once per group and all accumulators drain concurrently ----
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int py = m >> 3, px = m & 7;
    const int group = (warp - 2) >> 2;
    pdl_wait();                         // mask / addend reads and all stores touch the predecessor's data
    const bool split = p.esplit == 3;
    const int nblk_all = p.nout >> 4;
    // split: this group's third of the columns of EVERY tile; else all columns of every nbuf-th tile
    const int cb_lo = split ? group * (nblk_all / 3) : 0;
    const int nblk = split ? cb_lo + nblk_all / 3 : nblk_all;
    const bool skip = (p.dbg_flags & 4) != 0;
    const bool plain = p.bias_off && !p.corr && !p.has_addend && !p.has_mask && !p.out_nchw && p.dbg_flags == 0;
    // group g <-> accumulator g: a group only ever waits on consecutive phases of its own barriers (with more
    // groups than accumulators a group could run two phases ahead, which a parity wait cannot tell apart)
    // tile coordinates advance incrementally (two integer divisions per tile were ~9 % of this warp's instructions
    // on the narrow layers): one division up front, then carries
    int t_img, t_ty, t_tx;
    {
      const int tile0 = first_tile + ((!split && group < p.nbuf) ? group : 0) * tile_step;
      t_img = tile0 / tiles_per_img;
      const int r0 = tile0 - t_img * tiles_per_img;
      t_ty = r0 / p.tiles_x; t_tx = r0 - t_ty * p.tiles_x;
    }
    const int adv = (split ? 1 : p.nbuf) * tile_step;
    const int adv_img = adv / tiles_per_img, adv_r = adv - adv_img * tiles_per_img;
    const int adv_ty = adv_r / p.tiles_x, adv_tx = adv_r - adv_ty * p.tiles_x;
    for (int lt = split ? 0 : (group < p.nbuf ? group : npair_iters); lt < npair_iters; lt += split ? 1 : p.nbuf) {
      int tile = first_tile + lt * tile_step;
      const bool tile_ok = tile < p.ntiles;                 // odd tile count: the peer's last accumulator is a duplicate
      SgPix c;
      int ty = t_ty, tx = t_tx;
      c.img = t_img;
      if (!tile_ok) {                                       // (rare) recompute for the clamped tile
        tile = p.ntiles - 1;
        c.img = tile / tiles_per_img;
        const int r = tile - c.img * tiles_per_img;
        ty = r / p.tiles_x; tx = r - ty * p.tiles_x;
      }
      t_tx += adv_tx; if (t_tx >= p.tiles_x) { t_tx -= p.tiles_x; ++t_ty; }
      t_ty += adv_ty; if (t_ty >= p.tiles_y) { t_ty -= p.tiles_y; ++t_img; }
      t_img += adv_img;
      c.y = ty * kTileH + py; c.x = tx * kTileW + px;
      c.valid = tile_ok && c.y < p.y.H && c.x < p.y.W;
      c.ypix = (long long)c.img * p.y.sN + (long long)c.y * p.y.sY + (long long)c.x * p.y.sX;
      c.cls0 = c.cls1 = 0;
      if (p.corr) {
        // output pixels (2y + up_py, 2x) and (2y + up_py, 2x + 1) of the upsampled image (2H x 2W)
        const int Y = 2 * c.y + p.up_py;
        const int ycls = Y == 0 ? 3 : (Y == 2 * p.y.H - 1 ? 6 : 0);
        c.cls0 = ycls + (c.x == 0 ? 1 : 0);
        c.cls1 = ycls + (c.x == p.y.W - 1 ? 2 : 0);
      }
      c.apix = c.mpix = c.ppix = 0;                        // only the operands this launch has (uniform branches)
      if (p.has_addend) c.apix = (long long)c.img * p.addend.sN + (long long)c.y * p.addend.sY + (long long)c.x * p.addend.sX;
      if (p.has_mask) c.mpix = (long long)c.img * p.mask.sN + (long long)c.y * p.mask.sY + (long long)c.x * p.mask.sX;
      if (p.has_pool) c.ppix = (long long)c.img * p.pool.sN + (long long)(c.y >> 1) * p.pool.sY + (long long)(c.x >> 1) * p.pool.sX;
      const int buf = lt % p.nbuf;
      // The epilogue's global operand (activation mask, else skip-gradient addend): warm L2 with the
      // whole tile's worth now, while this tile's MMAs still run, and keep the register loads one
      // block pair ahead of their use, so their latency stays off the critical path.
      const bool pre = (p.has_mask || p.has_addend) && !skip && c.valid;
      const __nv_bfloat16* ab = p.has_mask ? (const __nv_bfloat16*)p.mask.ptr + c.mpix : (const __nv_bfloat16*)p.addend.ptr + c.apix;
      const long long as = p.has_mask ? p.mask.sCb : p.addend.sCb;
      uint32_t ax0[8], ax1[8];
      if (pre) {
        for (int cb = cb_lo + 2; cb < nblk; ++cb)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ab + cb * as));
        if (cb_lo < nblk) ld_global_32B(ab + cb_lo * as, ax0);
        if (cb_lo + 1 < nblk) ld_global_32B(ab + (cb_lo + 1) * as, ax1);
      }
      // one warp of the group polls the mbarrier, the other three park on a hardware named barrier
      if (quarter == 0) sg_wait(tfull_bar(buf), (uint32_t)(lt / p.nbuf) & 1u);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      fence_after_sync();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.nout);
#pragma unroll 1
      for (int cb = cb_lo; cb < nblk; cb += 2) {
        uint32_t r0[16], r1[16], nx0[8], nx1[8];
        sg_ld16(lane_addr + cb * 16, r0);
        const bool two = cb + 1 < nblk;
        if (two) sg_ld16(lane_addr + (cb + 1) * 16, r1);
        if (pre && cb + 2 < nblk) {
          ld_global_32B(ab + (cb + 2) * as, nx0);
          if (cb + 3 < nblk) ld_global_32B(ab + (cb + 3) * as, nx1);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (skip) continue;
        if (plain) {
          sg_epilogue_block<true>(p, c, cb, r0, lane, s_bias, nullptr);
          if (two) sg_epilogue_block<true>(p, c, cb + 1, r1, lane, s_bias, nullptr);
        } else {
          sg_epilogue_block<false>(p, c, cb, r0, lane, s_bias, pre ? ax0 : nullptr);
          if (two) sg_epilogue_block<false>(p, c, cb + 1, r1, lane, s_bias, pre ? ax1 : nullptr);
        }
        if (pre) {
#pragma unroll
          for (int q = 0; q < 8; ++q) { ax0[q] = nx0[q]; ax1[q] = nx1[q]; }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) { if (CG == 2) mbar_arrive_cluster_relaxed(tempty_bar(buf) & kPeerMask); else mbar_arrive(tempty_bar(buf)); }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // neither CTA may retire its barriers / TMEM while the other still signals them
  if (warp == 1) {
    fence_after_sync();
    if (CG == 2) tmem_dealloc2(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int sg_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kSMs;
  }
  return n;
}

// Do the packed weights of a conv (taps x cin_blocks x nout) fit resident beside two pipeline slots?
bool slab_weights_fit(int ntaps, int cin_blocks, int nout, bool halo) {
  const int ngroups = (cin_blocks + kSgGroup - 1) / kSgGroup;
  const size_t w_bytes = (size_t)ntaps * ngroups * kSgGroup * nout * 32 / (sg_use_pair(nout, 1 << 20) ? 2 : 1);
  const int gb = cin_blocks < kSgGroup ? cin_blocks : kSgGroup;
  const size_t slot = align_up((size_t)gb * (halo ? kHaloW * kHaloH : kTileW * kTileH) * 32, 1024);
  return align_up(w_bytes, 1024) + 2 * slot <= kSgSmemMax - kSgStaticSlack - 1024;
}
bool slab_geometry_ok(int dtype, int h, int w) {
  { const char* e = getenv("N2N_NO_SLAB"); if (e && atoi(e)) return false; }
  return dtype == N2N_BF16 && h >= 4 && w >= 4 && h % 2 == 0 && w % 2 == 0;
}

// Can ConvTranspose2x2 run as two N = 2*Cout launches on this engine?  (geometry + shared-memory fit)
bool slab_deconv_pair_ok(int dtype, int h, int w, int cin_blocks, int cout_blocks) {
  { const char* e = getenv("N2N_NO_SLAB"); if (e && atoi(e)) return false; }
  { const char* e = getenv("N2N_NO_DECONV_PAIR"); if (e && atoi(e)) return false; }
  if (dtype != N2N_BF16 || h < 4 || w < 4) return false;
  const int nout = 2 * cout_blocks * 16;
  if (nout > 256) return false;
  const int ngroups = (cin_blocks + kSgGroup - 1) / kSgGroup;
  const size_t w_bytes = (size_t)2 * ngroups * kSgGroup * nout * 32;
  const int gb = cin_blocks < kSgGroup ? cin_blocks : kSgGroup;
  const size_t slot = align_up((size_t)gb * kTileW * kTileH * 32, 1024);
  return align_up(w_bytes, 1024) + 2 * slot <= kSgSmemMax - kSgStaticSlack - 1024;
}

// Can ConvTranspose2x2 -> conv3x3 run as the fused column-range launch (layers.cuh: make_upconv_fwd)?  h, w = SOURCE
// (pre-upsampling) image size; w_bytes = packed weights of one output-row parity.  Needs the CTA-pair form (each SM
// keeps half of the weight rows).
bool slab_upconv_ok(int dtype, int n, int h, int w, int ci_blocks, int skip_blocks, int co_blocks, size_t w_bytes) {
  { const char* e = getenv("N2N_NO_SLAB"); if (e && atoi(e)) return false; }
  { const char* e = getenv("N2N_NO_UPFUSE"); if (e && atoi(e)) return false; }
  if (dtype != N2N_BF16 || h < 4 || w < 4) return false;
  const int mma_n = co_blocks * 16;
  if (2 * mma_n > 256) return false;
  const long long tiles = (long long)n * ((w + kTileW - 1) / kTileW) * ((h + kTileH - 1) / kTileH);
  if (!sg_use_pair(mma_n, tiles)) return false;
  const int gb = ci_blocks < kSgGroup ? ci_blocks : kSgGroup, gs = skip_blocks < kSgGroup ? skip_blocks : kSgGroup;
  const size_t slot = align_up((size_t)(gb > gs ? gb : gs) * kHaloW * kHaloH * 32, 1024);
  return align_up(w_bytes / 2, 1024) + 2 * slot <= kSgSmemMax - kSgStaticSlack - 1024;
}

// Returns 0 when launched, kSgNotEligible when this geometry belongs to the row-slab engine, < 0 on error.
int launch_slabgemm_umma(const TapGemm& g, cudaStream_t st) {
  static bool attr_set = false;
  if (g.dtype != N2N_BF16 || g.nout < 16 || g.nout > 256 || g.nout % 16) return kSgNotEligible;
  const int H = g.y.H, W = g.y.W;
  if (H < 4 || W < 4 || (g.has_pool && ((H | W) & 1))) return kSgNotEligible;     // edge tiles are masked in the epilogue
  { const char* e = getenv("N2N_NO_SLAB"); if (e && atoi(e)) return kSgNotEligible; }
  const int ngroups = (g.cin_blocks + kSgGroup - 1) / kSgGroup;
  const int mma_n = g.mma_n ? g.mma_n : g.nout;
  if (mma_n % 16 || g.nout % mma_n) return kSgNotEligible;
  const uint32_t b_sub = (uint32_t)mma_n * 32u;
  const size_t slab_bytes = (size_t)kSgGroup * b_sub;
  int max_slab = 0, nviews = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.tap_slab[t] > max_slab) max_slab = g.tap_slab[t];
    if (g.tap_view[t] + 1 > nviews) nviews = g.tap_view[t] + 1;
    if (g.tap_dy[t] < -1 || g.tap_dy[t] > 1 || g.tap_dx[t] < -1 || g.tap_dx[t] > 1) return kSgNotEligible;
  }
  if (nviews > kTapViews) return kSgNotEligible;
  size_t w_bytes = 0;   // = end of the last weight byte any tap reads (set below)

  SgParams p;
  memset(&p, 0, sizeof(p));
  int nst = 0;
  size_t slot_bytes = 0;
  for (int v = 0; v < nviews; ++v) {
    bool halo = false; int nt = 0;
    for (int t = 0; t < g.ntaps; ++t)
      if (g.tap_view[t] == v) { ++nt; if (g.tap_dy[t] || g.tap_dx[t]) halo = true; }
    if (nt == 0) continue;
    const View& xv = g.x[v];
    const int vblocks = g.view_blocks[v] ? g.view_blocks[v] : g.cin_blocks;   // channel blocks this view contributes
    const int vgroups = (vblocks + kSgGroup - 1) / kSgGroup;
    if (xv.H != H || xv.W != W || xv.Cb < vblocks || vgroups > ngroups) return kSgNotEligible;
    const int gb = vblocks < kSgGroup ? vblocks : kSgGroup;
    const int bw = halo ? kHaloW : kTileW, bh = halo ? kHaloH : kTileH;
    N2N_TRY(encode_c16_tensor_map(&p.tmap[v], xv, bw, bh, gb));
    const uint32_t cb_bytes = (uint32_t)(bw * bh * 32);
    for (int grp = 0; grp < vgroups; ++grp) {
      if (nst >= kSgMaxStages) return kSgNotEligible;
      SgStage& S = p.st[nst++];
      S.view = (int16_t)v; S.cb0 = (int16_t)(grp * kSgGroup);
      const int nb = vblocks - grp * kSgGroup;
      S.nb = (int16_t)(nb < kSgGroup ? nb : kSgGroup);
      S.ox = S.oy = (int16_t)(halo ? -1 : 0);
      S.cb_bytes16 = (uint16_t)(cb_bytes >> 4);
      S.sbo16 = (uint16_t)((bw * 32) >> 4);
      S.tx_bytes = (uint32_t)gb * cb_bytes;
      int k = 0;
      for (int t = 0; t < g.ntaps; ++t) {
        if (g.tap_view[t] != v) continue;
        if (k >= 9) return kSgNotEligible;
        const int oy = halo ? g.tap_dy[t] + 1 : 0, ox = halo ? g.tap_dx[t] + 1 : 0;
        S.a_off16[k] = (uint16_t)(((oy * bw + ox) * 32) >> 4);
        const size_t boff = g.mma_n ? (size_t)g.tap_woff[t] + (size_t)grp * slab_bytes
                                    : ((size_t)g.tap_slab[t] * ngroups + grp) * slab_bytes;
        S.b_off16[k] = (uint32_t)(boff >> 4);
        const int col = g.mma_n ? g.tap_col[t] : 0;
        if (col % 16 || col + mma_n > g.nout) return kSgNotEligible;
        S.col16[k] = (uint8_t)(col >> 4);
        if (boff + (size_t)S.nb * b_sub > w_bytes) w_bytes = boff + (size_t)S.nb * b_sub;
        ++k;
      }
      S.ntaps = (int16_t)k;
      if ((size_t)gb * cb_bytes > slot_bytes) slot_bytes = (size_t)gb * cb_bytes;
    }
  }
  for (int v = 0; v < kTapViews; ++v)      // unused descriptor slots must still be valid for prefetch.tensormap
    if (v >= nviews) p.tmap[v] = p.tmap[0];
  // "first MMA into these accumulator columns" flags: every column range must be opened by a tap of stage 0
  // (the only stage whose MMAs may overwrite instead of accumulate)
  {
    uint32_t opened = 0;       // bit c: columns [16c, ...) of a range starting at 16c already written
    for (int k = 0; k < p.st[0].ntaps; ++k) {
      const uint32_t bit = 1u << p.st[0].col16[k];
      if (!(opened & bit)) { p.st[0].first |= (uint16_t)(1u << k); opened |= bit; }
    }
    for (int col = 0; col < g.nout; col += mma_n)
      if (!(opened & (1u << (col >> 4)))) return kSgNotEligible;
  }
  slot_bytes = align_up(slot_bytes, 1024);
  // CTA pairs (cta_group::2): opt-in per shape via N2N_PAIR (bit 0: N <= 64 layers, bit 1: wider layers).
  // Each CTA of a pair keeps half of the weight rows, which leaves room for a deeper activation ring.
  const long long tiles0 = (long long)g.y.N * ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  const int cg = sg_use_pair(mma_n, tiles0) ? 2 : 1;
  size_t w_region = align_up(w_bytes / cg, 1024);
  const size_t budget = kSgSmemMax - kSgStaticSlack - 1024;
  if (w_region + 2 * slot_bytes > budget) return kSgNotEligible;
  // bias through the GEMM when the ones block + bias rows still leave a ring of at least four slots
  uint32_t bias_off = 0;
  {
    const size_t extra = 4096 + align_up((size_t)mma_n / cg * 32, 1024);
    static const char* const e = getenv("N2N_NO_BIAS_MMA");
    if (g.bias && !(e && atoi(e)) && w_region + extra + 4 * slot_bytes <= budget) { bias_off = (uint32_t)w_region; w_region += extra; }
  }
  int ring = (int)((budget - w_region) / slot_bytes);
  if (ring > kSgMaxRing) ring = kSgMaxRing;
  { static const char* const e = getenv("N2N_SG_RING"); if (e && atoi(e) >= 2 && atoi(e) < ring) ring = atoi(e); }

  p.nst = nst; p.nout = g.nout; p.mma_n = mma_n; p.corr = g.border_corr; p.up_py = g.up_py;
  p.w = (const uint8_t*)g.w; p.w_bytes = (uint32_t)w_bytes; p.w_region = (uint32_t)w_region; p.bias_off = bias_off;
  p.bias = g.bias; p.y = g.y; p.store_y = g.store_y ? 1 : 0;
  p.has_addend = g.has_addend; p.addend = g.addend; p.has_mask = g.has_mask; p.mask = g.mask;
  p.has_pool = g.has_pool; p.pool = g.pool;
  p.n_split = g.n_split; p.split_stride = g.split_stride;
  if (g.n_split && (g.has_pool || g.has_mask || g.has_addend || g.out_nchw || 2 * g.n_split * 16 != g.nout)) return kSgNotEligible;
  p.act = g.act; p.slope = g.slope; p.out_nchw = g.out_nchw; p.out_c = g.out_c;
  p.tiles_x = (W + kTileW - 1) / kTileW; p.tiles_y = (H + kTileH - 1) / kTileH;
  const long long tiles = (long long)g.y.N * p.tiles_x * p.tiles_y;
  N2N_CHECK_ARG(tiles > 0 && tiles < (1LL << 31), "slabgemm: bad tile count");
  p.ntiles = (int)tiles;
  p.ring = ring; p.slot_bytes = (uint32_t)slot_bytes;
  // accumulators: three when they fit in TMEM, so the MMAs of tile i+2 need not wait for the epilogue group
  // that is still draining tile i (two groups on alternate tiles); else two
  p.nbuf = 3 * g.nout <= 512 ? 3 : 2;
  { static const char* const e = getenv("N2N_SG_NBUF"); if (e && atoi(e) == 2) p.nbuf = 2; }
  {
    bool dual_ok = g.mma_n != 0 && g.nout == 2 * mma_n;
    for (int si = 0; si < nst; ++si) {
      SgStage& S = p.st[si];
      uint32_t nt0 = 0;
      for (int k = 0; k < S.ntaps; ++k)
        if (S.col16[k] == 0) { if (nt0 != (uint32_t)k) dual_ok = false; ++nt0; }      // lower range first
      S.nt0 = nt0;
    }
    static const char* const e = getenv("N2N_NO_DUAL_ISSUE");
    p.dual = (dual_ok && !(e && atoi(e))) ? 1 : 0;
  }
  { static const char* const e = getenv("N2N_NO_ESPLIT");
    p.esplit = (p.nbuf == 2 && (g.nout >> 4) % 3 == 0 && !(e && atoi(e))) ? 3 : 1; }
  p.tmem_cols = tmem_cols_for(p.nbuf * g.nout);
  p.idesc = make_idesc_bf16(128 * cg, mma_n, false, false);
  { static const char* const df = getenv("N2N_DBG_FLAGS"); p.dbg_flags = df ? atoi(df) : 0; }
  const size_t smem = 1024 + w_region + (size_t)ring * slot_bytes;
  if (!attr_set) {
    N2N_CUDA(cudaFuncSetAttribute(slabgemm_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kSgSmemMax - kSgStaticSlack)));
    N2N_CUDA(cudaFuncSetAttribute(slabgemm_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kSgSmemMax - kSgStaticSlack)));
    attr_set = true;
  }
  if (cg == 2) {
    int grid = (int)((tiles + 1) / 2) * 2;
    const int cap = sg_num_sms() & ~1;
    if (grid > cap) grid = cap;
    N2N_CUDA(launch_pdl_cluster2(slabgemm_umma_kernel<2>, dim3(grid), dim3(kSgThreads), smem, st, p));
  } else {
    const int grid = tiles < sg_num_sms() ? (int)tiles : sg_num_sms();
    N2N_CUDA(launch_pdl(slabgemm_umma_kernel<1>, dim3(grid), dim3(kSgThreads), smem, st, p));
  }
  N2N_LAUNCH_CHECK();
  return 0;
}

}  // namespace n2n
