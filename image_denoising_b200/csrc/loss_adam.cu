// loss_adam.cu — fused N2N loss (training_script.md:146-153), fused L1 + gradient loss
// (finetune.py:153-162, :283-285) and multi-tensor Adam (train.py:332).  All HBM-bound:
// one pass over the operands, vector loads, deterministic two-stage reductions (per-block
// partials in double, summed in a fixed order by the last block to finish).
#include "common.cuh"

namespace n2n {

constexpr int kRedThreads = 256;
constexpr int kMaxRedBlocks = kSMs * 4;

struct RedWs {              // layout of the loss workspace
  unsigned int counter;     // must be zero on entry; reset by the finishing block
  unsigned int pad[3];
  double partial[kMaxRedBlocks][4];
};

template <int NQ>
__device__ __forceinline__ void block_reduce(double (&v)[NQ], double* smem /* [NQ][8] */) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int q = 0; q < NQ; ++q) smem[q * 8 + warp] = v[q];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < NQ; ++q) {
      double s = 0;
      for (int w = 0; w < kRedThreads / 32; ++w) s += smem[q * 8 + w];
      v[q] = s;
    }
  }
}

// Returns true (in thread 0 of exactly one block) once every block has published its partials.
__device__ __forceinline__ bool last_block_done(RedWs* ws) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&ws->counter, 1u);
    last = (t == gridDim.x - 1);
    if (last) __threadfence();
  }
  __syncthreads();
  return last;
}

// Second stage, run by every thread of the last block: thread t adds partials t, t + 256, ... and the block tree
// finishes (fixed order -> deterministic).  One thread walking up to 592 partials was a chain of dependent L2
// loads several times longer than the streaming pass itself.
template <int NQ>
__device__ __forceinline__ void final_reduce(const RedWs* ws, double (&v)[NQ], double* smem) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) v[q] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += kRedThreads) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] += __ldcg(&ws->partial[b][q]);
  }
  block_reduce<NQ>(v, smem);
}

// ---- N2N loss --------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
n2n_loss_kernel(const float* __restrict__ out, const float* __restrict__ sub2, const float* __restrict__ den1,
                const float* __restrict__ den2, float lam, float gscale, long long count, float* __restrict__ loss3,
                float* __restrict__ grad, RedWs* ws, const float* __restrict__ dev_scalars) {
  pdl_enter();
  __shared__ double red[2 * 8];
  if (dev_scalars) lam = dev_scalars[0];      // graph-replayed steps read Lambda from device memory
  // per-thread sums stay fp32 over short runs (<= 64 elements between flushes; fp64 throughput on this
  // part is a small fraction of fp32), then accumulate in double
  double acc[2] = {0.0, 0.0};
  float a0 = 0.f, a1 = 0.f;
  int run = 0;
  const float k = gscale * 2.0f / (float)count;
  const long long nvec = count / 4;
  const bool vec_ok = ((((uintptr_t)out | (uintptr_t)sub2 | (uintptr_t)den1 | (uintptr_t)den2 | (uintptr_t)grad) & 15) == 0);
  long long start_tail = 0;
  if (vec_ok) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
      const float4 o = reinterpret_cast<const float4*>(out)[i];
      const float4 s = reinterpret_cast<const float4*>(sub2)[i];
      const float4 a = reinterpret_cast<const float4*>(den1)[i];
      const float4 b = reinterpret_cast<const float4*>(den2)[i];
      const float d[4] = {o.x - s.x, o.y - s.y, o.z - s.z, o.w - s.w};
      const float e[4] = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
      float g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float r = d[q] - e[q];
        a0 += d[q] * d[q];
        a1 += r * r;
        g[q] = k * (d[q] + lam * r);
      }
      if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g[0], g[1], g[2], g[3]);
      if (++run == 16) { acc[0] += (double)a0; acc[1] += (double)a1; a0 = a1 = 0.f; run = 0; }
    }
    start_tail = nvec * 4;
  }
  for (long long i = start_tail + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float d = out[i] - sub2[i];
    const float r = d - (den1[i] - den2[i]);
    a0 += d * d;
    a1 += r * r;
    if (grad) grad[i] = k * (d + lam * r);
  }
  acc[0] += (double)a0; acc[1] += (double)a1;
  block_reduce<2>(acc, red);
  if (threadIdx.x == 0) { ws->partial[blockIdx.x][0] = acc[0]; ws->partial[blockIdx.x][1] = acc[1]; }
  if (!last_block_done(ws)) return;
  final_reduce<2>(ws, acc, red);
  if (threadIdx.x == 0) {
    const double s0 = acc[0], s1 = acc[1];
    const float l1 = (float)(s0 / (double)count);
    const float l2 = lam * (float)(s1 / (double)count);
    loss3[0] = l1 + l2; loss3[1] = l1; loss3[2] = l2;
    ws->counter = 0;
  }
}

// ---- L1 + gradient-consistency loss ------------------------------------------------------
// e = pred - target.  loss = mean|e| + lg * ( mean_x |e[x+1]-e[x]| + mean_y |e[y+1]-e[y]| )
// d/de[y,x] = sgn(e)/M + lg/Mx * (sgn(dx[x-1]) - sgn(dx[x])) + lg/My * (sgn(dy[y-1]) - sgn(dy[y]))
__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <typename IDX>
__global__ void __launch_bounds__(kRedThreads)
l1grad_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int planes, int h, int w,
                   float lg, float gscale, float* __restrict__ loss3, float* __restrict__ grad, RedWs* ws) {
  pdl_enter();
  __shared__ double red[3 * 8];
  double acc[3] = {0.0, 0.0, 0.0};
  const long long hw = (long long)h * w, count = (long long)planes * hw;
  const long long cx = (long long)planes * h * (w - 1), cy = (long long)planes * (h - 1) * w;
  const float k1 = gscale / (float)count;
  const float kx = cx > 0 ? gscale * lg / (float)cx : 0.f;
  const float ky = cy > 0 ? gscale * lg / (float)cy : 0.f;
  // per-thread sums stay fp32 over short runs, then accumulate in double (fp64 throughput is a small fraction of fp32);
  // IDX = 32-bit index arithmetic whenever the tensor allows it (two 64-bit divisions per element dominated the kernel)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  int run = 0;
  const IDX n = (IDX)count, ihw = (IDX)hw, step = (IDX)gridDim.x * (IDX)blockDim.x;
  for (IDX i = (IDX)blockIdx.x * (IDX)blockDim.x + (IDX)threadIdx.x; i < n; i += step) {
    const IDX r = i % ihw;
    const int y = (int)(r / (IDX)w), x = (int)(r - (IDX)y * (IDX)w);
    const float p0 = pred[i], t0 = tgt[i];
    const float e = p0 - t0;
    a0 += fabsf(e);
    float g = k1 * sgnf(e);
    if (x + 1 < w) {
      const float dd = (pred[i + 1] - p0) - (tgt[i + 1] - t0);
      a1 += fabsf(dd);
      g -= kx * sgnf(dd);
    }
    if (x > 0) {
      const float dd = (p0 - pred[i - 1]) - (t0 - tgt[i - 1]);
      g += kx * sgnf(dd);
    }
    if (y + 1 < h) {
      const float dd = (pred[i + w] - p0) - (tgt[i + w] - t0);
      a2 += fabsf(dd);
      g -= ky * sgnf(dd);
    }
    if (y > 0) {
      const float dd = (p0 - pred[i - w]) - (t0 - tgt[i - w]);
      g += ky * sgnf(dd);
    }
    if (grad) grad[i] = g;
    if (++run == 16) { acc[0] += (double)a0; acc[1] += (double)a1; acc[2] += (double)a2; a0 = a1 = a2 = 0.f; run = 0; }
  }
  acc[0] += (double)a0; acc[1] += (double)a1; acc[2] += (double)a2;
  block_reduce<3>(acc, red);
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.x][0] = acc[0]; ws->partial[blockIdx.x][1] = acc[1]; ws->partial[blockIdx.x][2] = acc[2];
  }
  if (!last_block_done(ws)) return;
  final_reduce<3>(ws, acc, red);
  if (threadIdx.x == 0) {
    const double s0 = acc[0], s1 = acc[1], s2 = acc[2];
    const float l1 = (float)(s0 / (double)count);
    const float gx = cx > 0 ? (float)(s1 / (double)cx) : 0.f;
    const float gy = cy > 0 ? (float)(s2 / (double)cy) : 0.f;
    const float lgr = gx + gy;
    loss3[0] = l1 + lg * lgr; loss3[1] = l1; loss3[2] = lgr;
    ws->counter = 0;
  }
}

// ---- Structure loss (util.py:41-70) ---------------------------------------------------------
// loss = alpha * mean|p - t| + beta * (mean_y|p2[y+1]-p2[y]| + mean_x|p2[x+1]-p2[x]|) / 2 + gamma * mean|p2 - t|
// with p = network(noisy), p2 = network(clean), t = clean (train.py:361-363).  One pass: both gradients and the four
// sums.  d/dp = alpha sgn(p-t)/M;  d/dp2 = gamma sgn(p2-t)/M + beta/2 ((sgn(dy[y-1]) - sgn(dy[y]))/My + (sgn(dx[x-1]) - sgn(dx[x]))/Mx).
__global__ void __launch_bounds__(kRedThreads)
structure_loss_kernel(const float* __restrict__ pred, const float* __restrict__ pred2, const float* __restrict__ tgt,
                      int planes, int h, int w, float alpha, float beta, float gamma, float gscale,
                      float* __restrict__ loss4, float* __restrict__ grad1, float* __restrict__ grad2, RedWs* ws) {
  pdl_enter();
  __shared__ double red[4 * 8];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};        // |p-t|, |dy p2|, |dx p2|, |p2-t|
  const long long hw = (long long)h * w, count = (long long)planes * hw;
  const long long cy = (long long)planes * (h - 1) * w, cx = (long long)planes * h * (w - 1);
  const float k1 = gscale * alpha / (float)count, k3 = gscale * gamma / (float)count;
  const float ky = cy > 0 ? gscale * beta * 0.5f / (float)cy : 0.f;
  const float kx = cx > 0 ? gscale * beta * 0.5f / (float)cx : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % hw;
    const int y = (int)(r / w), x = (int)(r - (long long)y * w);
    const float t = tgt[i], q = pred2[i];
    const float e1 = pred[i] - t, e2 = q - t;
    acc[0] += (double)fabsf(e1);
    acc[3] += (double)fabsf(e2);
    float g2 = k3 * sgnf(e2);
    if (y + 1 < h) { const float d = pred2[i + w] - q; acc[1] += (double)fabsf(d); g2 -= ky * sgnf(d); }
    if (y > 0) g2 += ky * sgnf(q - pred2[i - w]);
    if (x + 1 < w) { const float d = pred2[i + 1] - q; acc[2] += (double)fabsf(d); g2 -= kx * sgnf(d); }
    if (x > 0) g2 += kx * sgnf(q - pred2[i - 1]);
    if (grad1) grad1[i] = k1 * sgnf(e1);
    if (grad2) grad2[i] = g2;
  }
  block_reduce<4>(acc, red);
  if (threadIdx.x == 0)
    for (int q = 0; q < 4; ++q) ws->partial[blockIdx.x][q] = acc[q];
  if (!last_block_done(ws)) return;
  final_reduce<4>(ws, acc, red);
  if (threadIdx.x == 0) {
    const float pixel = (float)(acc[0] / (double)count);
    const float tv1 = cy > 0 ? (float)(acc[1] / (double)cy) : 0.f;
    const float tv2 = cx > 0 ? (float)(acc[2] / (double)cx) : 0.f;
    const float tv = (tv1 + tv2) / 2.f;
    const float cst = (float)(acc[3] / (double)count);
    loss4[0] = alpha * pixel + beta * tv + gamma * cst; loss4[1] = pixel; loss4[2] = tv; loss4[3] = cst;
    ws->counter = 0;
  }
}

// ---- IQSL: intensity-quantised structural loss (finetune_iqsl.py:291-383) ----------------------------------------
// Three soft classes (dark / mid / bright) around the centres c = {t1/2, (t1+t2)/2, (t2+1)/2}: p = softmax(-|yhat - c| / tau);
// target one-hot from the thresholds t1, t2 (an optional margin around them is "don't care"); loss = multi-class Dice over the
// whole batch + ce_factor * soft cross-entropy.  Dice couples every pixel through nine global sums, so the gradient needs
// them first: pass 1 reduces {I_k, P_k, T_k, CE, V} (deterministic two-stage) and writes the loss, pass 2 writes dL/dyhat.
struct IqslWs {
  unsigned int counter;
  unsigned int pad[3];
  double sums[12];                 // I[3], P[3], T[3], ce_sum, valid_count, (unused)
  double partial[kMaxRedBlocks][11];
};
struct IqslParams { float t1, t2, inv_tau, margin, ce_factor, eps, gscale; };

__device__ __forceinline__ void iqsl_pixel(const IqslParams& q, float yh, float y, float (&p)[3], float (&t)[3], float& valid) {
  valid = 1.f;
  if (q.margin > 0.f)
    valid = ((y <= q.t1 - q.margin) || (y >= q.t1 + q.margin && y <= q.t2 - q.margin) || (y >= q.t2 + q.margin)) ? 1.f : 0.f;
  t[0] = (y <= q.t1) ? valid : 0.f;
  t[1] = (y > q.t1 && y < q.t2) ? valid : 0.f;
  t[2] = (y >= q.t2) ? valid : 0.f;
  const float c0 = q.t1 / 2.0f, c1 = (q.t1 + q.t2) / 2.0f, c2 = (q.t2 + 1.0f) / 2.0f;
  const float l0 = -fabsf(yh - c0) * q.inv_tau, l1 = -fabsf(yh - c1) * q.inv_tau, l2 = -fabsf(yh - c2) * q.inv_tau;
  const float mx = fmaxf(l0, fmaxf(l1, l2));
  const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
  const float inv = 1.0f / (e0 + e1 + e2);
  p[0] = e0 * inv * valid; p[1] = e1 * inv * valid; p[2] = e2 * inv * valid;
}

__global__ void __launch_bounds__(kRedThreads)
iqsl_sums_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long count, IqslParams q,
                 float* __restrict__ loss3, IqslWs* ws) {
  pdl_enter();
  __shared__ double red[11 * 8];
  double acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  float a[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) a[k] = 0.f;
  int run = 0;
  auto one = [&](float yh, float y) {
    float p[3], t[3], valid;
    iqsl_pixel(q, yh, y, p, t, valid);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[k] += p[k] * t[k]; a[3 + k] += p[k]; a[6 + k] += t[k];
      a[9] -= t[k] * logf(p[k] + q.eps);
    }
    a[10] += valid;
  };
  const long long stride = (long long)gridDim.x * blockDim.x, tid0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(tgt)) & 15) == 0;
  const long long n4 = vec ? count / 4 : 0;
  for (long long i = tid0; i < n4; i += stride) {            // 16-byte loads: the scalar form was latency-bound at 14 pixels per thread
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(pred) + i), t4 = __ldg(reinterpret_cast<const float4*>(tgt) + i);
    one(p4.x, t4.x); one(p4.y, t4.y); one(p4.z, t4.z); one(p4.w, t4.w);
    if (++run == 4) {
#pragma unroll
      for (int k = 0; k < 11; ++k) { acc[k] += (double)a[k]; a[k] = 0.f; }
      run = 0;
    }
  }
  for (long long i = 4 * n4 + tid0; i < count; i += stride) one(pred[i], tgt[i]);
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] += (double)a[k];
  block_reduce<11>(acc, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 11; ++k) ws->partial[blockIdx.x][k] = acc[k];
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&ws->counter, 1u) == gridDim.x - 1;
    if (last) __threadfence();
  }
  __syncthreads();
  if (!last) return;
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += kRedThreads)
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] += __ldcg(&ws->partial[b][k]);
  __syncthreads();
  block_reduce<11>(acc, red);
  if (threadIdx.x == 0) {
    double dice_mean = 0.0;
    for (int k = 0; k < 3; ++k) dice_mean += (2.0 * acc[k] + q.eps) / (acc[3 + k] + acc[6 + k] + q.eps);
    dice_mean /= 3.0;
    const double ce = acc[9] / (acc[10] * 3.0 + q.eps);
    const float ld = (float)(1.0 - dice_mean), lc = (float)ce;
    loss3[0] = ld + q.ce_factor * lc; loss3[1] = ld; loss3[2] = lc;
    for (int k = 0; k < 11; ++k) ws->sums[k] = acc[k];
    ws->counter = 0;
  }
}

__global__ void __launch_bounds__(kRedThreads)
iqsl_grad_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, long long count, IqslParams q,
                 float* __restrict__ grad, const IqslWs* ws) {
  pdl_enter();
  // d total / d prob_k = -(1/3) (2 t_k D_k - N_k) / D_k^2  -  ce_factor t_k / (prob_k + eps) / (3 V + eps)
  float N[3], D[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { N[k] = (float)(2.0 * ws->sums[k] + q.eps); D[k] = (float)(ws->sums[3 + k] + ws->sums[6 + k] + q.eps); }
  const float ce_w = q.ce_factor / (float)(ws->sums[10] * 3.0 + q.eps);
  const float c[3] = {q.t1 / 2.0f, (q.t1 + q.t2) / 2.0f, (q.t2 + 1.0f) / 2.0f};
  auto one = [&](float yh, float y) -> float {
    float p[3], t[3], valid;
    iqsl_pixel(q, yh, y, p, t, valid);
    float G[3], gp = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      G[k] = valid * (-(2.0f * t[k] * D[k] - N[k]) / (3.0f * D[k] * D[k]) - ce_w * t[k] / (p[k] + q.eps));
      gp += G[k] * p[k];
    }
    // softmax: d total / d l_j = p_j (G_j - sum_k G_k p_k) (p already carries `valid`; for valid = 0 everything is 0);
    // l_j = -|yhat - c_j| / tau
    float g = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) g -= p[j] * (G[j] - gp) * sgnf(yh - c[j]) * q.inv_tau;
    return q.gscale * g;
  };
  const long long stride = (long long)gridDim.x * blockDim.x, tid0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(tgt) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  const long long n4 = vec ? count / 4 : 0;
  for (long long i = tid0; i < n4; i += stride) {
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(pred) + i), t4 = __ldg(reinterpret_cast<const float4*>(tgt) + i);
    reinterpret_cast<float4*>(grad)[i] = make_float4(one(p4.x, t4.x), one(p4.y, t4.y), one(p4.z, t4.z), one(p4.w, t4.w));
  }
  for (long long i = 4 * n4 + tid0; i < count; i += stride) grad[i] = one(pred[i], tgt[i]);
}

// ---- multi-tensor Adam ---------------------------------------------------------------------
// torch.optim.Adam (defaults, no amsgrad / weight decay), same operation order as
// torch/optim/adam.py: m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
// denom = sqrt(v)/sqrt(bc2) + eps; p.addcdiv_(m, denom, -lr/bc1).
__global__ void __launch_bounds__(256)
adam_multi_kernel(const long long* __restrict__ table, const int* __restrict__ blocks, float b1, float b2,
                  float eps, float step_size, float inv_sqrt_bc2_recip /* sqrt(bc2) */, float gscale,
                  const float* __restrict__ dev_scalars) {
  pdl_enter();
  if (dev_scalars) { step_size = dev_scalars[1]; inv_sqrt_bc2_recip = dev_scalars[2]; }
  const int t = blocks[2 * blockIdx.x], chunk = blocks[2 * blockIdx.x + 1];
  const long long* row = table + 5 * (long long)t;
  float* p = reinterpret_cast<float*>(row[0]);
  const float* g = reinterpret_cast<const float*>(row[1]);
  float* m = reinterpret_cast<float*>(row[2]);
  float* v = reinterpret_cast<float*>(row[3]);
  const long long n = row[4];
  const long long base = (long long)chunk * N2N_ADAM_CHUNK;
  const float w1 = 1.0f - b1, w2 = 1.0f - b2;
  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gscale;
    mi = mi + w1 * (gi - mi);
    vi = vi * b2 + w2 * gi * gi;
    const float denom = sqrtf(vi) / inv_sqrt_bc2_recip + eps;
    pi = pi - step_size * (mi / denom);
  };
  // whole, 16-byte aligned chunk: two float4 per thread and array, all eight loads in flight before the first use
  // (the scalar loop was a chain of eight dependent 4-byte round trips per thread: 14 us for 35 MB)
  if (base + N2N_ADAM_CHUNK <= n && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0)) {
    static_assert(N2N_ADAM_CHUNK == 2048, "two float4 per thread at 256 threads");
    const long long i0 = base / 4 + threadIdx.x, i1 = i0 + 256;
    float4 P0 = reinterpret_cast<float4*>(p)[i0], P1 = reinterpret_cast<float4*>(p)[i1];
    const float4 G0 = reinterpret_cast<const float4*>(g)[i0], G1 = reinterpret_cast<const float4*>(g)[i1];
    float4 M0 = reinterpret_cast<float4*>(m)[i0], M1 = reinterpret_cast<float4*>(m)[i1];
    float4 V0 = reinterpret_cast<float4*>(v)[i0], V1 = reinterpret_cast<float4*>(v)[i1];
    update(P0.x, G0.x, M0.x, V0.x); update(P0.y, G0.y, M0.y, V0.y); update(P0.z, G0.z, M0.z, V0.z); update(P0.w, G0.w, M0.w, V0.w);
    update(P1.x, G1.x, M1.x, V1.x); update(P1.y, G1.y, M1.y, V1.y); update(P1.z, G1.z, M1.z, V1.z); update(P1.w, G1.w, M1.w, V1.w);
    reinterpret_cast<float4*>(p)[i0] = P0; reinterpret_cast<float4*>(p)[i1] = P1;
    reinterpret_cast<float4*>(m)[i0] = M0; reinterpret_cast<float4*>(m)[i1] = M1;
    reinterpret_cast<float4*>(v)[i0] = V0; reinterpret_cast<float4*>(v)[i1] = V1;
    return;
  }
  for (long long i = base + threadIdx.x; i < base + N2N_ADAM_CHUNK && i < n; i += blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(pi, g[i], mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace n2n

using namespace n2n;

extern "C" size_t n2n_loss_workspace_bytes(int64_t) { return sizeof(RedWs); }

extern "C" int n2n_loss_n2n_fwdbwd(const float* out, const float* sub2, const float* den1, const float* den2,
                                   float lam, float grad_scale, int64_t count, float* loss3, float* grad,
                                   void* workspace, void* stream) {
  N2N_CHECK_ARG(out && sub2 && den1 && den2 && loss3 && workspace && count > 0, "loss_n2n: bad arguments");
  int grid = grid_for(count / 4 + 1, kRedThreads, 4);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  (void)launch_pdl_v(n2n_loss_kernel, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, out, sub2, den1, den2, lam, grad_scale, count,
                                                                  loss3, grad, (RedWs*)workspace, nullptr);
  N2N_LAUNCH_CHECK();
  return 0;
}

// Graph-friendly form: Lambda comes from dev_scalars[0] (see n2n_set_step_scalars).
extern "C" int n2n_loss_n2n_fwdbwd_dev(const float* out, const float* sub2, const float* den1, const float* den2,
                                       const float* dev_scalars, float grad_scale, int64_t count, float* loss3,
                                       float* grad, void* workspace, void* stream) {
  N2N_CHECK_ARG(out && sub2 && den1 && den2 && loss3 && workspace && dev_scalars && count > 0, "loss_n2n_dev: bad arguments");
  int grid = grid_for(count / 4 + 1, kRedThreads, 4);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  (void)launch_pdl_v(n2n_loss_kernel, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, out, sub2, den1, den2, 0.f, grad_scale, count,
                                                                  loss3, grad, (RedWs*)workspace, dev_scalars);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_loss_l1grad_fwdbwd(const float* pred, const float* target, int n, int c, int h, int w,
                                      float lambda_grad, float grad_scale, float* loss3, float* grad,
                                      void* workspace, void* stream) {
  N2N_CHECK_ARG(pred && target && loss3 && workspace && n > 0 && c > 0 && h > 0 && w > 0, "loss_l1grad: bad arguments");
  const long long count = (long long)n * c * h * w;
  int grid = grid_for(count, kRedThreads, 4);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  if (count < (1LL << 31) - (long long)grid * kRedThreads)
    (void)launch_pdl_v(l1grad_loss_kernel<unsigned int>, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, pred, target, n * c, h, w,
                       lambda_grad, grad_scale, loss3, grad, (RedWs*)workspace);
  else
    (void)launch_pdl_v(l1grad_loss_kernel<long long>, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, pred, target, n * c, h, w,
                       lambda_grad, grad_scale, loss3, grad, (RedWs*)workspace);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" int n2n_loss_structure_fwdbwd(const float* pred, const float* pred2, const float* target, int n, int c, int h,
                                         int w, float alpha, float beta, float gamma, float grad_scale, float* loss4,
                                         float* grad_pred, float* grad_pred2, void* workspace, void* stream) {
  N2N_CHECK_ARG(pred && pred2 && target && loss4 && workspace && n > 0 && c > 0 && h > 0 && w > 0, "loss_structure: bad arguments");
  const long long count = (long long)n * c * h * w;
  int grid = grid_for(count, kRedThreads, 4);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  (void)launch_pdl_v(structure_loss_kernel, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, pred, pred2, target, n * c, h, w,
                     alpha, beta, gamma, grad_scale, loss4, grad_pred, grad_pred2, (RedWs*)workspace);
  N2N_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t n2n_loss_iqsl_workspace_bytes(void) { return sizeof(IqslWs); }

extern "C" int n2n_loss_iqsl_fwdbwd(const float* pred, const float* target, int64_t count, float t1, float t2, float tau, float margin,
                                    float ce_factor, float eps, float grad_scale, float* loss3, float* grad, void* workspace,
                                    void* stream) {
  N2N_CHECK_ARG(pred && target && loss3 && workspace && count > 0, "loss_iqsl: bad arguments");
  N2N_CHECK_ARG(t1 < t2, "loss_iqsl: need t1 < t2");
  IqslParams q;
  q.t1 = t1; q.t2 = t2; q.inv_tau = 1.0f / (tau > 1e-6f ? tau : 1e-6f); q.margin = margin; q.ce_factor = ce_factor; q.eps = eps;
  q.gscale = grad_scale;
  int grid = grid_for(count, kRedThreads, 4);
  if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
  (void)launch_pdl_v(iqsl_sums_kernel, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, pred, target, (long long)count, q, loss3,
                     (IqslWs*)workspace);
  N2N_LAUNCH_CHECK();
  if (grad) {
    (void)launch_pdl_v(iqsl_grad_kernel, dim3(grid), dim3(kRedThreads), 0, (cudaStream_t)stream, pred, target, (long long)count, q, grad,
                       (const IqslWs*)workspace);
    N2N_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int n2n_adam_multi(const int64_t* table, int ntensors, const int32_t* blocks, int nblocks, float lr,
                              float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  N2N_CHECK_ARG(table && blocks && ntensors > 0 && nblocks > 0 && step >= 1, "adam_multi: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float sqrt_bc2 = (float)sqrt(bc2);
  (void)launch_pdl_v(adam_multi_kernel, dim3(nblocks), dim3(256), 0, (cudaStream_t)stream, (const long long*)table, blocks, beta1, beta2, eps,
                                                              step_size, sqrt_bc2, grad_scale, nullptr);
  N2N_LAUNCH_CHECK();
  return 0;
}

// Graph-friendly form: lr / bias-correction terms come from dev_scalars[1..2].
extern "C" int n2n_adam_multi_dev(const int64_t* table, int ntensors, const int32_t* blocks, int nblocks,
                                  const float* dev_scalars, float beta1, float beta2, float eps, float grad_scale,
                                  void* stream) {
  N2N_CHECK_ARG(table && blocks && dev_scalars && ntensors > 0 && nblocks > 0, "adam_multi_dev: bad arguments");
  (void)launch_pdl_v(adam_multi_kernel, dim3(nblocks), dim3(256), 0, (cudaStream_t)stream, (const long long*)table, blocks, beta1, beta2, eps,
                                                              0.f, 1.f, grad_scale, dev_scalars);
  N2N_LAUNCH_CHECK();
  return 0;
}

__global__ void set_step_scalars_kernel(float* dst, float lam, float step_size, float sqrt_bc2) {
  dst[0] = lam; dst[1] = step_size; dst[2] = sqrt_bc2; dst[3] = 0.f;
}

// dev_scalars[4] = {Lambda, lr / (1 - beta1^step), sqrt(1 - beta2^step), 0}: the per-step values a
// captured CUDA graph of the training step cannot carry as kernel arguments.
extern "C" int n2n_set_step_scalars(float* dev_scalars, float lam, float lr, float beta1, float beta2, int step,
                                    void* stream) {
  N2N_CHECK_ARG(dev_scalars && step >= 1, "set_step_scalars: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  set_step_scalars_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_scalars, lam, (float)((double)lr / bc1), (float)sqrt(bc2));
  N2N_LAUNCH_CHECK();
  return 0;
}
