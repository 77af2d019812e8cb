// tapgemm_umma.cu — the bf16 tensor-core engine for the generic tap GEMM (conv3x3 / conv1x1 /
// ConvTranspose2x2, forward and input-gradient): an implicit GEMM on tcgen05 with the accumulator
// in TMEM, the activation operand fed by TMA straight from the C16 tensor (zero padding = TMA
// out-of-bounds fill, no im2col buffer), and the packed weights fed by cp.async.bulk.
//
//   D[128 pixels, nout] = sum over (tap, 48-channel group)  A_tap[128 px, 48 ch] * W_tap[nout, 48 ch]^T
//
// Persistent kernel: one CTA per SM walks the 128-pixel output tiles (bw x bh pixels of one image)
// round-robin.  6 warps:
//   warp 0 : TMA producer (one lane): per stage one 5-D tensor load (A) [+ one bulk copy (B)];
//            when the whole packed weight tensor fits next to >= 4 A stages it is loaded ONCE per
//            CTA and stays resident in shared memory (L2 traffic = activations only)
//   warp 1 : MMA issuer (one lane): <= 3 tcgen05.mma (K = 16) per stage, commit -> stage free;
//            accumulators are double-buffered in TMEM so tile i+1's MMAs overlap tile i's epilogue
//   warps 2-5: epilogue: tcgen05.ld -> bias / addend / LeakyReLU / mask -> bf16 C16 store
//            (32 B per pixel per block, contiguous across the warp) or fp32 NCHW
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace n2n {

using namespace umma;

constexpr int kMaxStages = 8;
constexpr int kGroupBlocks = 3;
constexpr int kThreads = 192;
constexpr int kSlabRows = 136;                    // 128 pixels + left/right halo, padded to a multiple of 8 rows
constexpr int kMaxEntries = 160;      // 9 taps x 16 channel-block groups (the 768-channel layers of ImprovedUNet) fit
constexpr size_t kSmemBudget = 232448 - 4096;     // 227 KB per CTA minus static/alignment slack

// One pipeline stage of the main loop: one staged activation tile (<= 3 channel blocks of one view
// at row offset dy) and the `ndx` horizontally shifted taps that consume it.  In slab mode
// (128-pixel single-row tiles) the staged tile is 136 pixels wide and the three dx taps are views
// of it that start 0 / 1 / 2 rows (32 B) into the slab, so every input pixel crosses L2->SMEM
// three times per 3x3 conv instead of nine; otherwise ndx == 1 and the tile is exactly the tap.
struct Entry {
  int8_t view, dy, dx0, ndx;
  int8_t cb0, nb, pf, pad1;      // pf: issue an L2 prefetch for this entry's box of a future tile
  uint32_t b_off[3];          // byte offsets (in the packed weight tensor) of the B sub-tiles, per dx tap
};

struct UmmaGemmParams {
  CUtensorMap tmap[4];
  Entry e[kMaxEntries];
  int nentries;
  int nout;
  const uint8_t* w;
  const float* bias;
  View y;
  int has_addend; View addend;
  int has_mask; View mask;
  int act; float slope;
  float* out_nchw; int out_c;
  int bw, bh, tiles_x, tiles_y, rows, ntiles;
  int resident_b, nstages, prefetch_dist;
  uint32_t a_sub;             // bytes between channel-block sub-tiles of a staged A tile
  uint32_t a_tx[4];           // bytes one TMA box of each view delivers
  uint32_t a_stage_bytes;     // offset of the streamed-B area inside a stage
  uint32_t b_total_bytes, b_region_bytes, stage_bytes, tmem_cols, idesc;
  long long* dbg;             // optional stall counters (CTA 0): see n2n_debug_stall_buffer
  int dbg_flags;              // diagnostic (N2N_DBG_FLAGS): 1 skip TMA loads, 2 skip MMA issue, 4 skip epilogue stores
};

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct EpiCtx {
  const UmmaGemmParams* p;
  int img, y, x;
  bool valid;
  long long ypix, apix, mpix;
};

__device__ __forceinline__ void epilogue_block(const EpiCtx& c, int cb, const uint32_t r[16]) {
  const UmmaGemmParams& p = *c.p;
  float v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
  if (!c.valid) return;
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + cb * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = __ldg(b4 + q);
      v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
    }
  }
  if (p.has_addend) {
    float a[16];
    Block16<__nv_bfloat16>::load((const __nv_bfloat16*)p.addend.ptr + c.apix + cb * p.addend.sCb, a);
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] += a[q];
  }
  if (p.act) {
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * p.slope;
  }
  if (p.has_mask) {
    float mk[16];
    Block16<__nv_bfloat16>::load((const __nv_bfloat16*)p.mask.ptr + c.mpix + cb * p.mask.sCb, mk);
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] *= (mk[q] > 0.f ? 1.f : p.slope);
  }
  if (p.out_nchw) {
    const long long hw = (long long)p.y.H * p.y.W;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int n = cb * 16 + q;
      if (n < p.out_c) p.out_nchw[((long long)c.img * p.out_c + n) * hw + (long long)c.y * p.y.W + c.x] = v[q];
    }
  } else {
    Block16<__nv_bfloat16>::store((__nv_bfloat16*)p.y.ptr + c.ypix + cb * p.y.sCb, v);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
tapgemm_umma_kernel(const __grid_constant__ UmmaGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * kMaxStages + 5];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stages0 = smem0 + p.b_region_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  const uint32_t bfull_bar = bar0 + 8u * (2 * kMaxStages);
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * kMaxStages + 1 + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * kMaxStages + 3 + b); };
  const uint32_t b_sub = (uint32_t)p.nout * 32u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(bfull_bar, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_smem), p.tmem_cols);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ---- TMA producer: the whole warp runs the loop (uniform control flow), one elected lane issues ----
    if (elect_one_sync()) {
      for (int v = 0; v < 4; ++v) prefetch_tensormap(&p.tmap[v]);
      if (p.resident_b) {
        mbar_arrive_expect_tx(bfull_bar, p.b_total_bytes);
        for (uint32_t off = 0; off < p.b_total_bytes; off += 16384u) {
          const uint32_t n = p.b_total_bytes - off < 16384u ? p.b_total_bytes - off : 16384u;
          bulk_load(smem0 + off, p.w + off, n, bfull_bar);
        }
      }
    }
    __syncwarp();
    pdl_wait();
    int stage = 0; uint32_t phase = 0;
    long long st_prod = 0;
    const long long k0 = clock64();
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      int r = tile;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int img = r / p.tiles_y;
      const int x0 = tx * p.bw, y0 = ty * p.bh;
      // L2 prefetch of the tile this CTA will process `prefetch_dist` rounds from now
      const int ptile = tile + p.prefetch_dist * (int)gridDim.x;
      int pimg = 0, px0 = 0, py0 = 0;
      const bool do_pf = p.prefetch_dist > 0 && ptile < p.ntiles;
      if (do_pf) {
        int q = ptile;
        px0 = (q % p.tiles_x) * p.bw; q /= p.tiles_x;
        py0 = (q % p.tiles_y) * p.bh;
        pimg = q / p.tiles_y;
      }
      for (int ei = 0; ei < p.nentries; ++ei) {
        const int view = p.e[ei].view, dy = p.e[ei].dy, dx0 = p.e[ei].dx0, ndx = p.e[ei].ndx, cb0 = p.e[ei].cb0,
                  nb = p.e[ei].nb, pf = p.e[ei].pf;
        if (p.dbg) { const long long w0 = clock64(); mbar_wait(empty_bar(stage), phase ^ 1u); st_prod += clock64() - w0; }
        else mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t a_dst = stages0 + stage * p.stage_bytes;
        uint32_t tx_bytes = p.a_tx[view];
        if (!p.resident_b) tx_bytes += (uint32_t)(ndx * nb) * b_sub;
        if (p.dbg_flags & 1) {
          if (elect_one_sync()) mbar_arrive(full_bar(stage));
        } else if (elect_one_sync()) {
          if (do_pf && pf) tma_prefetch_l2_5d(&p.tmap[view], 0, px0 + dx0, py0 + dy, cb0, pimg);
          mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
          tma_load_5d(a_dst, &p.tmap[view], full_bar(stage), 0, x0 + dx0, y0 + dy, cb0, img);
          if (!p.resident_b) {
            for (int i = 0; i < ndx; ++i)
              bulk_load(a_dst + p.a_stage_bytes + i * kGroupBlocks * b_sub, p.w + p.e[ei].b_off[i], (uint32_t)nb * b_sub,
                        full_bar(stage));
          }
        }
        __syncwarp();
        if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
      }
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[0] = st_prod; p.dbg[1] = clock64() - k0; }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp converged, one elected lane issues tcgen05.mma / commit ----
    int stage = 0; uint32_t phase = 0;
    pdl_wait();
    pdl_release();
    if (p.resident_b) mbar_wait(bfull_bar, 0);
    const uint32_t dhi = desc_hi(256, kSwizzle32);
    const uint32_t a_sub16 = p.a_sub >> 4, b_sub16 = b_sub >> 4;
    const uint32_t idesc = p.idesc;
    long long st_full = 0, st_tempty = 0;
    const long long k0 = clock64();
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      if (p.dbg) { const long long w0 = clock64(); mbar_wait(tempty_bar(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u); st_tempty += clock64() - w0; }
      else mbar_wait(tempty_bar(buf), (((uint32_t)lt >> 1) & 1u) ^ 1u);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.nout);
      uint32_t acc = 0;
      for (int ei = 0; ei < p.nentries; ++ei) {
        const int ndx = p.e[ei].ndx, nb = p.e[ei].nb;
        const uint32_t bo0 = p.e[ei].b_off[0], bo1 = p.e[ei].b_off[1], bo2 = p.e[ei].b_off[2];
        const uint32_t a_base = stages0 + stage * p.stage_bytes;
        const uint32_t a_lo = desc_lo(a_base, 16);
        const uint32_t bs = a_base + p.a_stage_bytes;       // streamed-B area of this stage
        const uint32_t b_lo0 = desc_lo(p.resident_b ? smem0 + bo0 : bs, 16);
        const uint32_t b_lo1 = desc_lo(p.resident_b ? smem0 + bo1 : bs + kGroupBlocks * b_sub, 16);
        const uint32_t b_lo2 = desc_lo(p.resident_b ? smem0 + bo2 : bs + 2 * kGroupBlocks * b_sub, 16);
        if (p.dbg) { const long long w0 = clock64(); mbar_wait(full_bar(stage), phase); st_full += clock64() - w0; }
        else mbar_wait(full_bar(stage), phase);
        fence_after_sync();
        if (elect_one_sync()) {
          // dx tap i of a slab = the same staged tile started i rows (32 B = 2 descriptor units) further in
          uint32_t a1 = acc;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            if (i < ndx) {
              const uint32_t bl = i == 0 ? b_lo0 : (i == 1 ? b_lo1 : b_lo2);
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                if (j < nb && !(p.dbg_flags & 2)) {
                  mma_bf16_lo(d_tmem, a_lo + j * a_sub16 + 2u * i, bl + j * b_sub16, dhi, idesc, a1);
                  a1 = 1;
                }
              }
            }
          }
          mma_commit(empty_bar(stage));                 // smem stage reusable once these MMAs finish
        }
        __syncwarp();
        acc = 1;
        if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) mma_commit(tfull_bar(buf)); // accumulator ready for the epilogue
      __syncwarp();
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[2] = st_full; p.dbg[3] = st_tempty; p.dbg[4] = clock64() - k0; p.dbg[5] = lt; }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), +32) ----
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    pdl_wait();
    const int py = m / p.bw, px = m - py * p.bw;
    const int nblk = p.nout / 16;
    long long st_tfull = 0;
    const long long k0 = clock64();
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++lt) {
      int r = tile;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      EpiCtx c;
      c.p = &p;
      c.img = r / p.tiles_y;
      c.y = ty * p.bh + py; c.x = tx * p.bw + px;
      c.valid = (m < p.rows) && (c.y < p.y.H) && (c.x < p.y.W);
      c.ypix = (long long)c.img * p.y.sN + (long long)c.y * p.y.sY + (long long)c.x * p.y.sX;
      c.apix = (long long)c.img * p.addend.sN + (long long)c.y * p.addend.sY + (long long)c.x * p.addend.sX;
      c.mpix = (long long)c.img * p.mask.sN + (long long)c.y * p.mask.sY + (long long)c.x * p.mask.sX;
      const int buf = lt & 1;
      { const long long w0 = clock64(); mbar_wait(tfull_bar(buf), ((uint32_t)lt >> 1) & 1u); st_tfull += clock64() - w0; }
      fence_after_sync();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.nout);
      for (int cb = 0; cb < nblk; cb += 2) {
        uint32_t r0[16], r1[16];
        tmem_ld16_issue(lane_addr + cb * 16, r0);
        const bool two = cb + 1 < nblk;
        if (two) tmem_ld16_issue(lane_addr + (cb + 1) * 16, r1);
        tmem_ld_wait();
        if (p.dbg_flags & 4) continue;
        epilogue_block(c, cb, r0);
        if (two) epilogue_block(c, cb + 1, r1);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));      // this warp's quarter of the accumulator is drained
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[6] = st_tfull; p.dbg[7] = clock64() - k0; }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

int encode_c16_tensor_map(CUtensorMap* out, const View& v, int bw, int bh, int cbox) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return N2N_ERR_CUDA; }
  cuuint64_t dims[5] = {16, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.Cb, (cuuint64_t)v.N};
  cuuint64_t strides[4] = {(cuuint64_t)v.sX * 2, (cuuint64_t)v.sY * 2, (cuuint64_t)v.sCb * 2, (cuuint64_t)v.sN * 2};
  cuuint32_t box[5] = {16, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)cbox, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (v.N == 1) strides[3] = strides[2] * (cuuint64_t)(v.Cb > 0 ? v.Cb : 1);   // any valid value
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p W=%d H=%d Cb=%d N=%d strides=%lld,%lld,%lld,%lld box=%d,%d,%d",
              (int)r, v.ptr, v.W, v.H, v.Cb, v.N, (long long)strides[0], (long long)strides[1], (long long)strides[2],
              (long long)strides[3], bw, bh, cbox);
    return N2N_ERR_CUDA;
  }
  return 0;
}

// Choose the 128-pixel tile shape (bw x bh, bw a power of two) that wastes the fewest pixels.
void choose_tile(int H, int W, int& bw, int& bh) {
  long long best = -1;
  for (int cand = 128; cand >= 8; cand >>= 1) {
    const int cb = 128 / cand;
    const long long cost = (long long)((W + cand - 1) / cand) * cand * ((H + cb - 1) / cb) * cb;
    if (best < 0 || cost < best) { best = cost; bw = cand; bh = cb; }
  }
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kSMs;
  }
  return n;
}

static long long* g_dbg_buf = nullptr;
void set_debug_stall_buffer(long long* p) { g_dbg_buf = p; }

int launch_tapgemm_umma(const TapGemm& g, cudaStream_t st) {
  N2N_CHECK_ARG(g.nout >= 16 && g.nout <= 256 && g.nout % 16 == 0, "tapgemm_umma: nout=%d unsupported", g.nout);
  static bool attr_set = false;
  UmmaGemmParams p;
  memset(&p, 0, sizeof(p));
  const int H = g.y.H, W = g.y.W;
  int bw, bh;
  choose_tile(H, W, bw, bh);
  if (bh > H) bh = H;
  p.bw = bw; p.bh = bh; p.rows = bw * bh;
  p.tiles_x = (W + bw - 1) / bw; p.tiles_y = (H + bh - 1) / bh;
  p.nout = g.nout;
  const int ngroups = (g.cin_blocks + kGroupBlocks - 1) / kGroupBlocks;
  const uint32_t b_sub = (uint32_t)g.nout * 32u;
  const size_t slab_bytes = (size_t)kGroupBlocks * b_sub;               // one (tap, group) slab of the packed tensor
  auto slab_off = [&](int slab, int grp) { return (uint32_t)(((size_t)slab * ngroups + grp) * slab_bytes); };

  // slab mode: plain 3x3 tap pattern on one view, single-row 128-pixel tiles
  bool slab = (g.ntaps == 9 && bw == 128 && bh == 1);
  for (int t = 0; t < g.ntaps && slab; ++t)
    slab = g.tap_view[t] == 0 && g.tap_dx[t] == (t % 3 - 1) * (g.tap_dx[1] - g.tap_dx[0] == 1 ? 1 : -1) &&
           g.tap_dy[t] == g.tap_dy[3 * (t / 3)];
  int ne = 0;
  if (slab) {
    // taps come dy-major; within a dy row dx is ascending (forward) or descending (input gradient)
    const bool asc = g.tap_dx[0] == -1;
    for (int r = 0; r < 3; ++r)
      for (int grp = 0; grp < ngroups; ++grp) {
        Entry& e = p.e[ne++];
        e.view = 0; e.dy = (int8_t)g.tap_dy[3 * r]; e.dx0 = -1; e.ndx = 3;
        e.cb0 = (int8_t)(grp * kGroupBlocks);
        const int nb = g.cin_blocks - grp * kGroupBlocks;
        e.nb = (int8_t)(nb < kGroupBlocks ? nb : kGroupBlocks);
        e.pf = 1;
        for (int i = 0; i < 3; ++i) {               // i-th view of the slab = tap with dx = i - 1
          const int t = 3 * r + (asc ? i : 2 - i);
          e.b_off[i] = slab_off(g.tap_slab[t], grp);
        }
      }
  } else {
    N2N_CHECK_ARG(g.ntaps * ngroups <= kMaxEntries, "tapgemm_umma: too many pipeline entries");
    for (int t = 0; t < g.ntaps; ++t)
      for (int grp = 0; grp < ngroups; ++grp) {
        Entry& e = p.e[ne++];
        e.view = (int8_t)g.tap_view[t]; e.dy = (int8_t)g.tap_dy[t]; e.dx0 = (int8_t)g.tap_dx[t]; e.ndx = 1;
        e.cb0 = (int8_t)(grp * kGroupBlocks);
        const int nb = g.cin_blocks - grp * kGroupBlocks;
        e.nb = (int8_t)(nb < kGroupBlocks ? nb : kGroupBlocks);
        e.b_off[0] = slab_off(g.tap_slab[t], grp);
        e.pf = (g.tap_dx[t] == 0) ? 1 : 0;          // the dx = +-1 boxes touch the same lines as dx = 0
      }
  }
  p.nentries = ne;
  {
    static const char* const pd = getenv("N2N_PREFETCH_DIST");
    p.prefetch_dist = pd ? atoi(pd) : 2;
  }
  const int boxw = slab ? kSlabRows : bw;
  const int gb = g.cin_blocks < kGroupBlocks ? g.cin_blocks : kGroupBlocks;
  p.a_sub = (uint32_t)(boxw * bh * 32);
  int nviews = 0;
  for (int t = 0; t < g.ntaps; ++t)
    if (g.tap_view[t] + 1 > nviews) nviews = g.tap_view[t] + 1;
  for (int v = 0; v < 4; ++v) {
    const View& xv = g.x[v < nviews ? v : 0];
    N2N_CHECK_ARG(xv.H == H && xv.W == W && xv.Cb >= g.cin_blocks, "tapgemm_umma: view %d geometry mismatch", v);
    N2N_TRY(encode_c16_tensor_map(&p.tmap[v], xv, boxw, bh, gb));
    p.a_tx[v] = (uint32_t)(gb * boxw * bh * 32);
  }
  p.w = (const uint8_t*)g.w; p.bias = g.bias; p.y = g.y;
  p.has_addend = g.has_addend; p.addend = g.addend; p.has_mask = g.has_mask; p.mask = g.mask;
  p.act = g.act; p.slope = g.slope; p.out_nchw = g.out_nchw; p.out_c = g.out_c;
  const long long tiles = (long long)g.y.N * p.tiles_x * p.tiles_y;
  N2N_CHECK_ARG(tiles > 0 && tiles < (1LL << 31), "tapgemm_umma: bad tile count");
  p.ntiles = (int)tiles;

  // Weights stay resident in shared memory when the packed tensor (all slabs this launch may
  // touch) fits beside >= 4 activation stages and there are enough tiles per CTA to amortise it.
  int max_slab = 0;
  for (int t = 0; t < g.ntaps; ++t) if (g.tap_slab[t] > max_slab) max_slab = g.tap_slab[t];
  const size_t b_total = (size_t)(max_slab + 1) * ngroups * slab_bytes;
  p.a_stage_bytes = (uint32_t)align_up((size_t)kGroupBlocks * boxw * bh * 32, 1024);
  if (b_total + 4 * (size_t)p.a_stage_bytes <= kSmemBudget && tiles >= 2 * num_sms()) {
    p.resident_b = 1;
    p.b_total_bytes = (uint32_t)b_total;
    p.b_region_bytes = (uint32_t)align_up(b_total, 1024);
    p.stage_bytes = p.a_stage_bytes;
  } else {
    p.resident_b = 0;
    p.b_region_bytes = 0;
    p.stage_bytes = (uint32_t)(p.a_stage_bytes + (slab ? 3 : 1) * align_up(slab_bytes, 1024));
  }
  int nst = (int)((kSmemBudget - p.b_region_bytes) / p.stage_bytes);
  if (nst > kMaxStages) nst = kMaxStages;
  N2N_CHECK_ARG(nst >= 2, "tapgemm_umma: not enough shared memory for a pipeline (nout=%d)", g.nout);
  p.nstages = nst;
  p.dbg = g_dbg_buf;
  { static const char* const df = getenv("N2N_DBG_FLAGS"); p.dbg_flags = df ? atoi(df) : 0; }
  { static const char* const ns = getenv("N2N_STAGES"); if (ns && atoi(ns) >= 2 && atoi(ns) < nst) p.nstages = nst = atoi(ns); }
  p.tmem_cols = tmem_cols_for(2 * g.nout);
  p.idesc = make_idesc_bf16(128, g.nout, false, false);
  const size_t smem = 1024 + (size_t)p.b_region_bytes + (size_t)nst * p.stage_bytes;
  if (!attr_set) {
    N2N_CUDA(cudaFuncSetAttribute(tapgemm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
    attr_set = true;
  }
  const int grid = tiles < num_sms() ? (int)tiles : num_sms();
  N2N_CUDA(launch_pdl(tapgemm_umma_kernel, dim3(grid), dim3(kThreads), smem, st, p));
  N2N_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Bring-up probe: a single-CTA GEMM D[128, N] = A * B^T (bf16 -> fp32) whose operands are written
// to shared memory by ordinary stores in the exact layouts the engines assume.  It validates the
// descriptor encodings independently of TMA.  variant = base | shift << 8 | use_base_offset << 16
//   base 0: K-major  SWIZZLE_32B  ([k/16][row][32 B])            — forward / dgrad engine
//   base 1: MN-major SWIZZLE_32B  ([mn/16][k][32 B], LBO = atom) — weight-gradient engine
//   base 2: as 1 with LBO/SBO swapped (diagnostic)
//   base 3: as 0, but A holds 128+8 rows and the MMA starts `shift` rows (32 B each) into it:
//           D[m] = A[m + shift] * B^T   (the dx-shifted view of one staged activation slab)
//   base 4: as 1, but B holds K+8 pixel rows and the MMA starts `shift` rows into it:
//           D[m][n] = sum_k A[k][m] * B[k + shift][n]   (tap-shifted X slab in the wgrad engine)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
probe_umma_kernel(int variant, const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                  float* __restrict__ D, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* s = smem_raw + (smem0 - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = variant & 0xff, shift = (variant >> 8) & 0xff, use_bo = (variant >> 16) & 1;
  const uint32_t ncols = 256;
  const int a_rows = base == 3 ? 136 : 128;          // rows of A (K-major) ...
  const int b_krows = base == 4 ? K + 8 : K;          // ... pixel rows of B (MN-major)
  const uint32_t a_bytes = base == 0 || base == 3 ? (uint32_t)a_rows * K * 2u : 128u * K * 2u;
  uint8_t* sa = s; uint8_t* sb = s + ((a_bytes + 1023u) & ~1023u);
  const uint32_t sb_addr = smem0 + ((a_bytes + 1023u) & ~1023u);
  if (base == 0 || base == 3) {
    for (int i = threadIdx.x; i < a_rows * K; i += 128) {
      const int r = i / K, k = i % K, kb = k >> 4, e = k & 15;
      const uint32_t off = (uint32_t)kb * a_rows * 32 + r * 32 + ((((e >> 3) ^ ((r >> 2) & 1))) << 4) + (e & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(sa + off) = A[i];
    }
    for (int i = threadIdx.x; i < N * K; i += 128) {
      const int r = i / K, k = i % K, kb = k >> 4, e = k & 15;
      const uint32_t off = (uint32_t)kb * N * 32 + r * 32 + ((((e >> 3) ^ ((r >> 2) & 1))) << 4) + (e & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(sb + off) = B[i];
    }
  } else {
    // A given as [128][K] (row m, col k) -> stored [m/16][k][m%16]
    for (int i = threadIdx.x; i < 128 * K; i += 128) {
      const int r = i / K, k = i % K, mb = r >> 4, e = r & 15;
      const uint32_t off = (uint32_t)mb * K * 32 + k * 32 + ((((e >> 3) ^ ((k >> 2) & 1))) << 4) + (e & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(sa + off) = A[i];
    }
    // B given as [N][b_krows] -> stored [n/16][k][n%16]
    for (int i = threadIdx.x; i < N * b_krows; i += 128) {
      const int r = i / b_krows, k = i % b_krows, nb = r >> 4, e = r & 15;
      const uint32_t off = (uint32_t)nb * b_krows * 32 + k * 32 + ((((e >> 3) ^ ((k >> 2) & 1))) << 4) + (e & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(sb + off) = B[i];
    }
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_smem), ncols); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    const bool mn = !(base == 0 || base == 3);
    const uint32_t idesc = make_idesc_bf16(128, N, mn, mn);
    for (int kk = 0; kk < K / 16; ++kk) {
      uint64_t ad, bd;
      if (base == 0 || base == 3) {
        const uint32_t a_start = smem0 + kk * a_rows * 32 + (base == 3 ? shift * 32 : 0);
        ad = make_smem_desc(a_start, 16, 256, kSwizzle32, use_bo ? (a_start >> 7) & 7 : 0);
        bd = make_smem_desc(sb_addr + kk * N * 32, 16, 256, kSwizzle32);
      } else if (base == 1 || base == 4) {
        const uint32_t b_start = sb_addr + kk * 16 * 32 + (base == 4 ? shift * 32 : 0);
        ad = make_smem_desc(smem0 + kk * 16 * 32, (uint32_t)K * 32, 256, kSwizzle32);
        bd = make_smem_desc(b_start, (uint32_t)b_krows * 32, 256, kSwizzle32, use_bo ? (b_start >> 7) & 7 : 0);
      } else {
        ad = make_smem_desc(smem0 + kk * 16 * 32, 256, (uint32_t)K * 32, kSwizzle32);
        bd = make_smem_desc(sb_addr + kk * 16 * 32, 256, (uint32_t)K * 32, kSwizzle32);
      }
      mma_bf16(tmem_base, ad, bd, idesc, kk != 0);
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after_sync();
  const int m = warp * 32 + lane;
  for (int cb = 0; cb < N / 16; ++cb) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cb * 16, v);
    for (int q = 0; q < 16; ++q) D[(long long)m * N + cb * 16 + q] = v[q];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
}

// Throughput probe: every SM issues `iters` back-to-back tcgen05.mma (M=128, N, K=16, bf16) on
// shared-memory operands in the given K-major layout (contents irrelevant) and reports cycles.
//   layout 0: SWIZZLE_32B rows of 32 B (one 16-channel block per row)   [what the conv engine uses]
//   layout 1: SWIZZLE_128B rows of 128 B, the MMA consuming 32 B slices of them (k advances inside the row)
//   layout 2: SWIZZLE_64B rows of 64 B
__global__ void __launch_bounds__(128)
probe_mma_rate_kernel(int layout, int N, int iters, int naccum, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 100 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const int distinct = layout >> 3;      // layout + 8: every MMA reads a different A and B tile (8 of each, cycled)
  layout &= 7;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t rowb = layout == 0 ? 32u : (layout == 1 ? 128u : 64u);
    const uint32_t sw = layout == 0 ? kSwizzle32 : (layout == 1 ? kSwizzle128 : kSwizzle64);
    const uint32_t ksteps = rowb / 32u;                       // K=16 slices per staged row
    const uint32_t a0 = smem0, b0 = smem0 + 128u * rowb;
    const uint32_t hi = desc_hi(8u * rowb, sw);
    const uint32_t a_lo = desc_lo(a0, 16), b_lo = desc_lo(b0, 16);
    const uint32_t d1 = tmem_base + (naccum > 1 ? (uint32_t)N : 0u);
    (void)ksteps;
    const long long t0 = clock64();
    mma_bf16_lo(tmem_base, a_lo, b_lo, hi, idesc, 0);
    mma_bf16_lo(d1, a_lo, b_lo, hi, idesc, 0);
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t ao = distinct ? (uint32_t)u * (4096u >> 4) : ((u & (ksteps - 1)) << 1);
        const uint32_t bo = distinct ? (32768u >> 4) + (uint32_t)u * ((uint32_t)N * 32u >> 4) - ((128u * rowb) >> 4) : ((u & (ksteps - 1)) << 1);
        mma_bf16_lo((u & 1) ? d1 : tmem_base, a_lo + ao, b_lo + bo, hi, idesc, 1);
      }
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) { fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace n2n

using namespace n2n;

extern "C" int n2n_probe_umma(int variant, const void* a_bf16, const void* b_bf16, float* d, int m, int n, int k,
                              void* stream) {
  N2N_CHECK_ARG(m == 128 && n >= 16 && n <= 256 && n % 16 == 0 && k >= 16 && k % 16 == 0 && k <= 256,
                "probe_umma: need m=128, n%%16==0 (<=256), k%%16==0 (<=256)");
  const int base = variant & 0xff;
  N2N_CHECK_ARG(base >= 0 && base <= 4 && ((variant >> 8) & 0xff) <= 8 && a_bf16 && b_bf16 && d, "probe_umma: bad arguments");
  const size_t smem = 4096 + (size_t)136 * k * 2 + (size_t)n * (k + 8) * 2;
  N2N_CUDA(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(variant, (const __nv_bfloat16*)a_bf16,
                                                            (const __nv_bfloat16*)b_bf16, d, n, k);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_probe_mma_rate(int layout, int n, int iters, int naccum, long long* cycles_dev, int nblocks, void* stream) {
  N2N_CHECK_ARG((layout & 7) >= 0 && (layout & 7) <= 2 && layout < 16 && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && naccum >= 1 &&
                    naccum * n <= 512 && cycles_dev && nblocks > 0, "probe_mma_rate: bad arguments");
  N2N_CUDA(cudaFuncSetAttribute(probe_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
  probe_mma_rate_kernel<<<nblocks, 128, 102 * 1024, (cudaStream_t)stream>>>(layout, n, iters, naccum, cycles_dev);
  N2N_LAUNCH_CHECK();
  return 0;
}

// Diagnostic: 8 int64 counters written by CTA 0 of every tap-GEMM launch (cycles): [0] producer waiting
// for a free stage, [1] producer total, [2] MMA thread waiting for data, [3] MMA thread waiting for a
// free accumulator, [4] MMA thread total, [5] tiles, [6] epilogue waiting for an accumulator, [7] epilogue total.
extern "C" int n2n_debug_stall_buffer(long long* dev_counters) {
  n2n::set_debug_stall_buffer(dev_counters);
  return 0;
}
