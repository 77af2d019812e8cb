// subsample.cu — the 2x2 neighbour sub-sampler (train.py:141-190) as coalesced, vectorised
// HBM-bound gather kernels.  Pure copies: bit-exact in every dtype.
//
// Work decomposition: one thread owns 4 horizontally adjacent 2x2 cells (when w % 8 == 0
// and everything is 16-byte aligned): it reads 2 rows x 8 elements with vector loads, reads
// the selector bytes of its 4 cells with one load, and writes 4 contiguous output elements
// per sub-image with one vector store.  A scalar path covers every other shape.
#include "common.cuh"

namespace n2n {

// k in [0,4): position inside the 2x2 cell, k = 2*ky + kx (F.unfold order, train.py:134-138).
__device__ __forceinline__ int sel_from_mask4(uint32_t m) {
  // m = 4 bool bytes of one cell (little endian), exactly one of them non-zero.
  return (m & 0x000000ffu) ? 0 : (m & 0x0000ff00u) ? 1 : (m & 0x00ff0000u) ? 2 : 3;
}

__constant__ int8_t c_pair_table[8][2] = {{0, 1}, {0, 2}, {1, 3}, {2, 3}, {1, 0}, {2, 0}, {3, 1}, {3, 2}};

// train.py:151-172 — masks (and the packed selector) straight from rd_idx: every byte of
// both masks is written here, so no separate zero-fill pass is needed.
__global__ void mask_pair_kernel(const int64_t* __restrict__ rd_idx, long long cells, uint32_t* __restrict__ mask1,
                                 uint32_t* __restrict__ mask2, uint8_t* __restrict__ packed) {
  pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cells;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(rd_idx[i] & 7);
    const int k1 = c_pair_table[r][0], k2 = c_pair_table[r][1];
    if (mask1) mask1[i] = 1u << (8 * k1);
    if (mask2) mask2[i] = 1u << (8 * k2);
    if (packed) packed[i] = (uint8_t)(k1 | (k2 << 2));
  }
}

template <typename E> struct Vec4;   // 4 elements of E as one vector store
template <> struct Vec4<uint8_t> { using type = uchar4; };
template <> struct Vec4<uint16_t> { using type = ushort4; };
template <> struct Vec4<uint32_t> { using type = uint4; };
template <> struct Vec4<uint64_t> { using type = ulonglong4; };

// MODE 0: one mask -> one output.  MODE 1: two masks -> two outputs.  MODE 2: packed selector.
template <typename E, int MODE>
__global__ void subsample_vec_kernel(const E* __restrict__ img, const uint8_t* __restrict__ m1,
                                     const uint8_t* __restrict__ m2, E* __restrict__ o1, E* __restrict__ o2,
                                     int n, int c, int h, int w, long long groups) {
  pdl_enter();
  const int hh = h / 2, ww = w / 2, gw = ww / 4;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x) {
    long long r = g;
    const int gx = (int)(r % gw); r /= gw;
    const int i = (int)(r % hh); r /= hh;
    const int ch = (int)(r % c);
    const int b = (int)(r / c);
    const long long cell = ((long long)b * hh + i) * ww + gx * 4;
    int k1[4], k2[4];
    if (MODE == 2) {
      const uint32_t s = *reinterpret_cast<const uint32_t*>(m1 + cell);
#pragma unroll
      for (int q = 0; q < 4; ++q) { k1[q] = (s >> (8 * q)) & 3; k2[q] = (s >> (8 * q + 2)) & 3; }
    } else {
      const uint4 a = *reinterpret_cast<const uint4*>(m1 + cell * 4);
      k1[0] = sel_from_mask4(a.x); k1[1] = sel_from_mask4(a.y); k1[2] = sel_from_mask4(a.z); k1[3] = sel_from_mask4(a.w);
      if (MODE == 1) {
        const uint4 d = *reinterpret_cast<const uint4*>(m2 + cell * 4);
        k2[0] = sel_from_mask4(d.x); k2[1] = sel_from_mask4(d.y); k2[2] = sel_from_mask4(d.z); k2[3] = sel_from_mask4(d.w);
      }
    }
    const E* row0 = img + (((long long)b * c + ch) * h + 2 * i) * w + gx * 8;
    E top[8], bot[8];
    using V = typename Vec4<E>::type;
    *reinterpret_cast<V*>(&top[0]) = *reinterpret_cast<const V*>(row0);
    *reinterpret_cast<V*>(&top[4]) = *reinterpret_cast<const V*>(row0 + 4);
    *reinterpret_cast<V*>(&bot[0]) = *reinterpret_cast<const V*>(row0 + w);
    *reinterpret_cast<V*>(&bot[4]) = *reinterpret_cast<const V*>(row0 + w + 4);
    const long long obase = (((long long)b * c + ch) * hh + i) * ww + gx * 4;
    E r1[4], r2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const E t0 = top[2 * q], t1 = top[2 * q + 1], b0 = bot[2 * q], b1 = bot[2 * q + 1];
      r1[q] = (k1[q] & 2) ? ((k1[q] & 1) ? b1 : b0) : ((k1[q] & 1) ? t1 : t0);
      if (MODE != 0) r2[q] = (k2[q] & 2) ? ((k2[q] & 1) ? b1 : b0) : ((k2[q] & 1) ? t1 : t0);
    }
    *reinterpret_cast<V*>(o1 + obase) = *reinterpret_cast<V*>(&r1[0]);
    if (MODE != 0) *reinterpret_cast<V*>(o2 + obase) = *reinterpret_cast<V*>(&r2[0]);
  }
}

template <typename E, int MODE>
__global__ void subsample_scalar_kernel(const E* __restrict__ img, const uint8_t* __restrict__ m1,
                                        const uint8_t* __restrict__ m2, E* __restrict__ o1, E* __restrict__ o2,
                                        int n, int c, int h, int w, long long items) {
  pdl_enter();
  const int hh = h / 2, ww = w / 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < items;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int j = (int)(r % ww); r /= ww;
    const int i = (int)(r % hh); r /= hh;
    const int ch = (int)(r % c);
    const int b = (int)(r / c);
    const long long cell = ((long long)b * hh + i) * ww + j;
    int k1, k2 = 0;
    if (MODE == 2) {
      const int s = m1[cell]; k1 = s & 3; k2 = (s >> 2) & 3;
    } else {
      const uint8_t* q = m1 + cell * 4;
      k1 = q[0] ? 0 : q[1] ? 1 : q[2] ? 2 : 3;
      if (MODE == 1) { const uint8_t* p = m2 + cell * 4; k2 = p[0] ? 0 : p[1] ? 1 : p[2] ? 2 : 3; }
    }
    const E* base = img + (((long long)b * c + ch) * h + 2 * i) * w + 2 * j;
    o1[t] = base[(k1 >> 1) * w + (k1 & 1)];
    if (MODE != 0) o2[t] = base[(k2 >> 1) * w + (k2 & 1)];
  }
}

// train.py:134-138: F.unfold(x, bs, stride=bs).view(n, c*bs*bs, h/bs, w/bs):
// y[n, c*bs*bs + ky*bs + kx, i, j] = x[n, c, i*bs + ky, j*bs + kx].  One thread per output element; a warp's
// 32 consecutive j read a 32*bs-element span of one input row (every byte of it is used by the bs kx-planes
// scheduled back to back, so it stays in L1/L2) and write 32 consecutive elements.
template <typename E>
__global__ void space_to_depth_kernel(const E* __restrict__ x, E* __restrict__ y, int c, int h, int w, int bs,
                                      long long items) {
  pdl_enter();
  const int hh = h / bs, ww = w / bs, b2 = bs * bs;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < items;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int j = (int)(r % ww); r /= ww;
    const int i = (int)(r % hh); r /= hh;
    const int k = (int)(r % b2); r /= b2;
    const int ch = (int)(r % c);
    const long long b = r / c;
    y[t] = x[((b * c + ch) * h + (long long)i * bs + k / bs) * w + (long long)j * bs + k % bs];
  }
}

template <typename E>
static int run_space_to_depth(const void* x, void* y, int n, int c, int h, int w, int bs, cudaStream_t st) {
  const long long items = (long long)n * c * bs * bs * (h / bs) * (w / bs);
  if (items == 0) return 0;
  (void)launch_pdl_v(space_to_depth_kernel<E>, dim3(grid_for(items, 256)), dim3(256), 0, st, (const E*)x, (E*)y, c, h, w, bs, items);
  N2N_LAUNCH_CHECK();
  return 0;
}

template <typename E>
static int run_subsample(const void* img, const uint8_t* m1, const uint8_t* m2, const uint8_t* packed, void* o1,
                         void* o2, int n, int c, int h, int w, cudaStream_t st) {
  const int mode = packed ? 2 : (m2 ? 1 : 0);
  const uint8_t* a = packed ? packed : m1;
  const int hh = h / 2, ww = w / 2;
  const long long items = (long long)n * c * hh * ww;
  if (items == 0) return 0;
  auto aligned = [](const void* p, size_t al) { return p == nullptr || ((uintptr_t)p % al) == 0; };
  const size_t va = sizeof(E) * 4 > 16 ? 32 : sizeof(E) * 4;   // ulonglong4 needs 32-byte alignment
  const bool vec = (w % 8 == 0) && aligned(img, va) && aligned(o1, va) && aligned(o2, va) &&
                   aligned(a, mode == 2 ? 4 : 16) && aligned(m2, 16);
  if (vec) {
    const long long groups = items / 4;
    const int grid = grid_for(groups, 256);
#define N2N_SS_LAUNCH(M) \
    (void)launch_pdl_v(subsample_vec_kernel<E, M>, dim3(grid), dim3(256), 0, st, (const E*)img, a, m2, (E*)o1, (E*)o2, n, c, h, w, groups)
    if (mode == 0) N2N_SS_LAUNCH(0); else if (mode == 1) N2N_SS_LAUNCH(1); else N2N_SS_LAUNCH(2);
#undef N2N_SS_LAUNCH
  } else {
    const int grid = grid_for(items, 256);
#define N2N_SS_LAUNCH(M) \
    (void)launch_pdl_v(subsample_scalar_kernel<E, M>, dim3(grid), dim3(256), 0, st, (const E*)img, a, m2, (E*)o1, (E*)o2, n, c, h, w, items)
    if (mode == 0) N2N_SS_LAUNCH(0); else if (mode == 1) N2N_SS_LAUNCH(1); else N2N_SS_LAUNCH(2);
#undef N2N_SS_LAUNCH
  }
  N2N_LAUNCH_CHECK();
  return 0;
}

static int dispatch_subsample(const void* img, const uint8_t* m1, const uint8_t* m2, const uint8_t* packed,
                              void* o1, void* o2, int n, int c, int h, int w, int es, cudaStream_t st) {
  switch (es) {
    case 1: return run_subsample<uint8_t>(img, m1, m2, packed, o1, o2, n, c, h, w, st);
    case 2: return run_subsample<uint16_t>(img, m1, m2, packed, o1, o2, n, c, h, w, st);
    case 4: return run_subsample<uint32_t>(img, m1, m2, packed, o1, o2, n, c, h, w, st);
    case 8: return run_subsample<uint64_t>(img, m1, m2, packed, o1, o2, n, c, h, w, st);
  }
  set_error("subsample: unsupported element size %d", es);
  return N2N_ERR_ARG;
}

}  // namespace n2n

using namespace n2n;

extern "C" int n2n_mask_pair_from_rdidx(const int64_t* rd_idx, int64_t cells, uint8_t* mask1, uint8_t* mask2,
                                        uint8_t* packed_sel, void* stream) {
  N2N_CHECK_ARG(rd_idx != nullptr && cells >= 0, "mask_pair_from_rdidx: bad arguments");
  N2N_CHECK_ARG(((uintptr_t)mask1 % 4) == 0 && ((uintptr_t)mask2 % 4) == 0, "mask_pair_from_rdidx: masks must be 4-byte aligned");
  if (cells == 0) return 0;
  (void)launch_pdl_v(mask_pair_kernel, dim3(grid_for(cells, 256)), dim3(256), 0, (cudaStream_t)stream, rd_idx, cells,
                     (uint32_t*)mask1, (uint32_t*)mask2, packed_sel);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_subsample(const void* img, const uint8_t* mask, void* out, int n, int c, int h, int w,
                             int elem_size, void* stream) {
  N2N_CHECK_ARG(n >= 0 && c >= 0 && h >= 0 && w >= 0, "subsample: negative dimension");
  if ((long long)n * c * (h / 2) * (w / 2) == 0) return 0;
  N2N_CHECK_ARG(img && mask && out, "subsample: null pointer");
  return dispatch_subsample(img, mask, nullptr, nullptr, out, nullptr, n, c, h, w, elem_size, (cudaStream_t)stream);
}

extern "C" int n2n_subsample_pair(const void* img, const uint8_t* mask1, const uint8_t* mask2,
                                  const uint8_t* packed_sel, void* out1, void* out2, int n, int c, int h, int w,
                                  int elem_size, void* stream) {
  N2N_CHECK_ARG(n >= 0 && c >= 0 && h >= 0 && w >= 0, "subsample_pair: negative dimension");
  if ((long long)n * c * (h / 2) * (w / 2) == 0) return 0;
  N2N_CHECK_ARG(img && out1 && out2, "subsample_pair: null pointer");
  N2N_CHECK_ARG(packed_sel || (mask1 && mask2), "subsample_pair: need both masks or the packed selector");
  return dispatch_subsample(img, mask1, mask2, packed_sel, out1, out2, n, c, h, w, elem_size, (cudaStream_t)stream);
}

extern "C" int n2n_space_to_depth(const void* x, void* y, int n, int c, int h, int w, int block_size, int elem_size,
                                  void* stream) {
  N2N_CHECK_ARG(n >= 0 && c >= 0 && h >= 0 && w >= 0 && block_size >= 1, "space_to_depth: bad dimensions");
  N2N_CHECK_ARG(h % block_size == 0 && w % block_size == 0, "space_to_depth: H and W must be multiples of block_size (%d)", block_size);
  if ((long long)n * c * h * w == 0) return 0;
  N2N_CHECK_ARG(x && y, "space_to_depth: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (elem_size) {
    case 1: return run_space_to_depth<uint8_t>(x, y, n, c, h, w, block_size, st);
    case 2: return run_space_to_depth<uint16_t>(x, y, n, c, h, w, block_size, st);
    case 4: return run_space_to_depth<uint32_t>(x, y, n, c, h, w, block_size, st);
    case 8: return run_space_to_depth<uint64_t>(x, y, n, c, h, w, block_size, st);
  }
  set_error("space_to_depth: unsupported element size %d", elem_size);
  return N2N_ERR_ARG;
}
