// loss_aug.cu -- DiffAugment (diff_aug.py:23-102) and the scalar losses (loss.py:8-132) as fused
// forward / backward kernels.  Everything here is tiny or purely bandwidth bound: one block per
// image (augment) or per event (losses), warp-shuffle reductions, fixed summation order.
#include "common.cuh"
using namespace iea;

namespace {
// ------------------------------------------------------------------ DiffAugment
struct Aug {
  const float* br; const float* ct; const int64_t* tx; const int64_t* ty; const int64_t* ox; const int64_t* oy;
  int cut_h, cut_w;
};

__device__ __forceinline__ bool cut_hit(const Aug& a, int64_t n, int i, int j, int h, int w) {
  if (!a.ox) return false;
  int r0 = (int)a.ox[n] - a.cut_h / 2, c0 = (int)a.oy[n] - a.cut_w / 2;
  int ra = min(max(r0, 0), h - 1), rb = min(max(r0 + a.cut_h - 1, 0), h - 1);
  int ca = min(max(c0, 0), w - 1), cb = min(max(c0 + a.cut_w - 1, 0), w - 1);
  return i >= ra && i <= rb && j >= ca && j <= cb;
}

__global__ void __launch_bounds__(256) img_mean_kernel(const float* x, int64_t hw, float* mean) {
  __shared__ float red[33];
  const int64_t n = blockIdx.x;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) s += x[n * hw + i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) mean[n] = s / (float)hw;
}

__global__ void diffaug_fwd_kernel(const float* x, Aug a, int64_t n, int h, int w, float* y, const float* mean) {
  const int64_t total = n * h * (int64_t)w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int j = idx % w; int64_t t = idx / w; int i = t % h; int64_t nn = t / h;
    int si = i + (a.tx ? (int)a.tx[nn] : 0), sj = j + (a.ty ? (int)a.ty[nn] : 0);
    float v = 0.f;
    if (si >= 0 && si < h && sj >= 0 && sj < w && !cut_hit(a, nn, i, j, h, w)) {
      float b = a.br ? a.br[nn] - 0.5f : 0.f;
      v = x[(nn * h + si) * (int64_t)w + sj] + b;
      if (a.ct) {
        float m = mean[nn] + b, c = a.ct[nn] + 0.5f;
        v = (v - m) * c + m;
      }
    }
    y[idx] = v;
  }
}

// gradient w.r.t. the contrast stage's output, at source pixel (i, j)
__device__ __forceinline__ float aug_ga(const float* dy, const Aug& a, int64_t nn, int i, int j, int h, int w) {
  int di = i - (a.tx ? (int)a.tx[nn] : 0), dj = j - (a.ty ? (int)a.ty[nn] : 0);
  if (di < 0 || di >= h || dj < 0 || dj >= w || cut_hit(a, nn, di, dj, h, w)) return 0.f;
  return dy[(nn * h + di) * (int64_t)w + dj];
}

__global__ void __launch_bounds__(256) diffaug_bwd_mean_kernel(const float* dy, Aug a, int h, int w, float* mean) {
  __shared__ float red[33];
  const int64_t nn = blockIdx.x;
  float s = 0.f;
  for (int64_t p = threadIdx.x; p < (int64_t)h * w; p += blockDim.x) s += aug_ga(dy, a, nn, p / w, p % w, h, w);
  s = block_sum(s, red);
  if (threadIdx.x == 0) mean[nn] = s / (float)((int64_t)h * w);
}

__global__ void diffaug_bwd_kernel(const float* dy, Aug a, int64_t n, int h, int w, float* dx, const float* mean) {
  const int64_t total = n * h * (int64_t)w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int j = idx % w; int64_t t = idx / w; int i = t % h; int64_t nn = t / h;
    float g = aug_ga(dy, a, nn, i, j, h, w);
    if (a.ct) {
      float c = a.ct[nn] + 0.5f;
      g = c * g + (1.f - c) * mean[nn];
    }
    dx[idx] = g;
  }
}

Aug to_aug(const iea_aug_draws* d) {
  Aug a;
  a.br = d->brightness; a.ct = d->contrast; a.tx = d->tx; a.ty = d->ty; a.ox = d->ox; a.oy = d->oy;
  a.cut_h = d->cut_h; a.cut_w = d->cut_w;
  return a;
}
inline int ew_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : (int)b;
}

// ------------------------------------------------------------------ losses
__global__ void __launch_bounds__(256) hinge_dis_kernel(const float* fake, const float* real, int64_t n, float* out) {
  __shared__ float red[33];
  float a = 0.f, b = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    a += fmaxf(1.f - real[i], 0.f);
    b += fmaxf(1.f + fake[i], 0.f);
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) { out[0] = a / n; out[1] = b / n; }
}
__global__ void hinge_dis_bwd_kernel(const float* fake, const float* real, const float* dout, int64_t n, float* dfake,
                                     float* dreal) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dreal[i] = (1.f - real[i] > 0.f) ? -dout[0] / n : 0.f;
  dfake[i] = (1.f + fake[i] > 0.f) ? dout[1] / n : 0.f;
}
__global__ void __launch_bounds__(256) mean_kernel(const float* x, int64_t n, float scale, float* out) {
  __shared__ float red[33];
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) out[0] = scale * a / n;
}
__global__ void mean_bwd_kernel(const float* dout, int64_t n, float scale, float* dx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dout[0] * scale / n;
}
// loss.py:41-44 l2_loss = MSELoss(a, b): single block, fixed-order reduction (the inputs are (40E,) logits)
__global__ void __launch_bounds__(256) l2_kernel(const float* a, const float* b, int64_t n, float* out) {
  __shared__ float red[33];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float d = a[i] - b[i]; s = fmaf(d, d, s); }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s / n;
}
__global__ void l2_bwd_kernel(const float* a, const float* b, const float* dout, int64_t n, float* da, float* db) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = 2.f * (a[i] - b[i]) * dout[0] / n;
  if (da) da[i] = g;
  if (db) db[i] = -g;
}
__global__ void sum_events_kernel(const float* ev, int events, int stride, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int e = 0; e < events; ++e) t += ev[(int64_t)e * stride];
    out[0] = t / events;
  }
}

// Gram matrices of an event, spread over the chip: grid (events, ceil(dim / GSL)) CTAs each take a GSL-column
// slice of the (seq x dim) operands into shared memory and write the partial seq x seq products
//   part[e][s][0 .. seq*seq)        = A_e[:, slice] . B_e[:, slice]^T
//   part[e][s][seq*seq + 2i, +2i+1] = (a_i . c_i, c_i . c_i) over the slice        (when Cm != NULL)
// the per-event loss kernels then add the slices in a fixed order (deterministic).  One CTA per event computing
// the whole 40 x 40 x 1024 Gram with a warp per pair (round 1) took 2.7 ms for 8 events: pure load latency.
constexpr int GSL = 128;
__global__ void __launch_bounds__(256) gram_partial_kernel(const float* A, const float* B, const float* Cm, int seq,
                                                           int dim, float* part, int stride) {
  extern __shared__ float sm[];
  const int e = blockIdx.x, sl = blockIdx.y, k0 = sl * GSL;
  const int kw = dim - k0 < GSL ? dim - k0 : GSL;
  float* As = sm;
  float* Bs = (B == A) ? As : As + seq * (GSL + 1);
  const float* Ae = A + (int64_t)e * seq * dim;
  const float* Be = B + (int64_t)e * seq * dim;
  for (int idx = threadIdx.x; idx < seq * GSL; idx += 256) {
    const int i = idx / GSL, k = idx - i * GSL;
    As[i * (GSL + 1) + k] = k < kw ? Ae[(int64_t)i * dim + k0 + k] : 0.f;
    if (B != A) Bs[i * (GSL + 1) + k] = k < kw ? Be[(int64_t)i * dim + k0 + k] : 0.f;
  }
  __syncthreads();
  float* out = part + ((int64_t)e * gridDim.y + sl) * stride;
  for (int p = threadIdx.x; p < seq * seq; p += 256) {
    const int i = p / seq, j = p - i * seq;
    const float* a = As + i * (GSL + 1);
    const float* b = Bs + j * (GSL + 1);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < GSL; k += 2) { s0 = fmaf(a[k], b[k], s0); s1 = fmaf(a[k + 1], b[k + 1], s1); }
    out[p] = s0 + s1;
  }
  if (Cm) {
    const float* Ce = Cm + (int64_t)e * seq * dim;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = wid; i < seq; i += 8) {
      float ac = 0.f, cc = 0.f;
      for (int k = lane; k < kw; k += 32) {
        const float c = Ce[(int64_t)i * dim + k0 + k];
        ac = fmaf(As[i * (GSL + 1) + k], c, ac);
        cc = fmaf(c, c, cc);
      }
      ac = warp_sum(ac); cc = warp_sum(cc);
      if (lane == 0) { out[seq * seq + 2 * i] = ac; out[seq * seq + 2 * i + 1] = cc; }
    }
  }
}
// sum of the slices of one event's partial Gram (and row dots) into shared memory
__device__ __forceinline__ void gram_collect(const float* part, int e, int nsl, int stride, int count, float* dst) {
  for (int p = threadIdx.x; p < count; p += blockDim.x) {
    float s = 0.f;
    for (int sl = 0; sl < nsl; ++sl) s += part[((int64_t)e * nsl + sl) * stride + p];
    dst[p] = s;
  }
}

// saved per event: P[seq][seq] (weights exp(.)/den_i, 0 on the diagonal), Cs[seq][seq] (cosines),
// q[seq] (num/den), cp[seq] (cos(e_i,p_i)), ne[seq], np[seq], then the event's loss  -> 2*seq*seq+4*seq+1
__global__ void __launch_bounds__(256) contrastive_fwd_kernel(const float* part, int nsl, int stride, int seq, int dim,
                                                              float temp, float margin, float* saved_all) {
  extern __shared__ float sm[];
  float* S = sm;               // [seq][seq]
  float* ne = S + seq * seq;   // [seq]
  float* np_ = ne + seq;
  float* cp = np_ + seq;
  float* li = cp + seq;        // per-row loss
  __shared__ float red[33];
  const int e = blockIdx.x;
  float* saved = saved_all + (int64_t)e * (2 * seq * seq + 4 * seq + 1);
  gram_collect(part, e, nsl, stride, seq * seq, S);  // E E^T
  for (int i = threadIdx.x; i < seq; i += blockDim.x) {  // e_i . p_i and |p_i|^2
    float ac = 0.f, cc = 0.f;
    for (int sl = 0; sl < nsl; ++sl) {
      const float* o = part + ((int64_t)e * nsl + sl) * stride + seq * seq + 2 * i;
      ac += o[0]; cc += o[1];
    }
    cp[i] = ac; np_[i] = sqrtf(cc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < seq; i += blockDim.x) ne[i] = sqrtf(S[i * seq + i]);
  __syncthreads();
  for (int i = threadIdx.x; i < seq; i += blockDim.x) {
    const float ni = fmaxf(ne[i], 1e-8f);
    float c_p = cp[i] / (ni * fmaxf(np_[i], 1e-8f));
    float num = __expf((c_p - margin) / temp), den = num;
    for (int j = 0; j < seq; ++j)
      if (j != i) {
        float c = S[i * seq + j] / (ni * fmaxf(ne[j], 1e-8f));
        saved[seq * seq + i * seq + j] = c;
        den += __expf((c - margin) / temp);
      }
    saved[seq * seq + i * seq + i] = 1.f;
    for (int j = 0; j < seq; ++j)
      saved[i * seq + j] = (j == i) ? 0.f : __expf((saved[seq * seq + i * seq + j] - margin) / temp) / den;
    saved[2 * seq * seq + i] = num / den;
    saved[2 * seq * seq + seq + i] = c_p;
    saved[2 * seq * seq + 2 * seq + i] = ne[i];
    saved[2 * seq * seq + 3 * seq + i] = np_[i];
    li[i] = -__logf(temp * num / den);
  }
  __syncthreads();
  float t = 0.f;
  for (int i = threadIdx.x; i < seq; i += blockDim.x) t += li[i];
  t = block_sum(t, red);
  if (threadIdx.x == 0) saved[2 * seq * seq + 4 * seq] = t / seq;
}

// grid (events, column slices of SLICE features): the 40 x dim gradient of an event is independent per
// feature column, so the columns are spread over the chip instead of one CTA per event
constexpr int CBWD_SLICE = 64;
__global__ void __launch_bounds__(256) contrastive_bwd_kernel(const float* embed, const float* proxy,
                                                              const float* saved_all, const float* dloss, int events,
                                                              int seq, int dim, float temp, float* dembed,
                                                              float* dproxy) {
  extern __shared__ float sm[];
  const int e = blockIdx.x, k0 = blockIdx.y * CBWD_SLICE;
  const float* saved = saved_all + (int64_t)e * (2 * seq * seq + 4 * seq + 1);
  float* A = sm;                 // [seq][seq] coefficient on cos(e_i, e_j), symmetrised, times 1/(|e_i||e_j|)
  float* A2 = A + seq * seq;     // [seq] sum_j A_ij * cos_ij / |e_i|^2   (coefficient of e_i itself)
  float* B = A2 + seq;           // [seq] coefficient on cos(e_i, p_i)
  float* ine = B + seq;          // [seq] 1 / max(|e_i|, eps)
  float* inp = ine + seq;        // [seq] 1 / max(|p_i|, eps)
  const float* P = saved; const float* Cs = saved + seq * seq;
  const float* q = saved + 2 * seq * seq; const float* cp = q + seq; const float* ne = cp + seq; const float* np_ = ne + seq;
  const float k = dloss[0] / (events * (float)seq * temp);
  for (int i = threadIdx.x; i < seq; i += blockDim.x) {
    ine[i] = 1.f / fmaxf(ne[i], 1e-8f);
    inp[i] = 1.f / fmaxf(np_[i], 1e-8f);
    B[i] = -k * (1.f - q[i]);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < seq * seq; p += blockDim.x) {
    const int i = p / seq, j = p - i * seq;
    A[p] = i == j ? 0.f : k * (P[i * seq + j] + P[j * seq + i]) * ine[i] * ine[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < seq; i += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < seq; ++j) a = fmaf(A[i * seq + j] * fmaxf(ne[j], 1e-8f), Cs[i * seq + j], a);  // A_ij cos_ij / |e_i|
    A2[i] = a * ine[i];
  }
  __syncthreads();
  const float* E = embed + (int64_t)e * seq * dim;
  const float* Pr = proxy + (int64_t)e * seq * dim;
  const int ncol = dim - k0 < CBWD_SLICE ? dim - k0 : CBWD_SLICE;
  for (int idx = threadIdx.x; idx < seq * ncol; idx += blockDim.x) {
    const int i = idx / ncol, kk = k0 + idx - i * ncol;
    const float ei = E[(int64_t)i * dim + kk], pi = Pr[(int64_t)i * dim + kk];
    float g = -A2[i] * ei;
    for (int j = 0; j < seq; ++j) g = fmaf(A[i * seq + j], E[(int64_t)j * dim + kk], g);
    g = fmaf(B[i], pi * ine[i] * inp[i] - cp[i] * ei * ine[i] * ine[i], g);
    dembed[(int64_t)e * seq * dim + (int64_t)i * dim + kk] = g;
    if (dproxy) dproxy[(int64_t)e * seq * dim + (int64_t)i * dim + kk] = B[i] * (ei * ine[i] * inp[i] - cp[i] * pi * inp[i] * inp[i]);
  }
}

// saved per event: D[seq][seq] = (softmax(F F^T) - softmax(R R^T)) / seq, then the event's loss
__global__ void __launch_bounds__(256) iea_fwd_kernel(const float* part_f, const float* part_r, int nsl, int stride,
                                                      int seq, int dim, float* saved_all) {
  extern __shared__ float sm[];
  float* SF = sm; float* SR = SF + seq * seq; float* li = SR + seq * seq;
  __shared__ float red[33];
  const int e = blockIdx.x;
  float* saved = saved_all + (int64_t)e * (seq * seq + 1);
  gram_collect(part_f, e, nsl, stride, seq * seq, SF);
  gram_collect(part_r, e, nsl, stride, seq * seq, SR);
  __syncthreads();
  for (int i = threadIdx.x; i < seq; i += blockDim.x) {
    float mf = -3e38f, mr = -3e38f;
    for (int j = 0; j < seq; ++j) { mf = fmaxf(mf, SF[i * seq + j]); mr = fmaxf(mr, SR[i * seq + j]); }
    float zf = 0.f, zr = 0.f;
    for (int j = 0; j < seq; ++j) { zf += __expf(SF[i * seq + j] - mf); zr += __expf(SR[i * seq + j] - mr); }
    const float lzf = mf + __logf(zf), lzr = mr + __logf(zr);
    float acc = 0.f;
    for (int j = 0; j < seq; ++j) {
      float lq = SF[i * seq + j] - lzf, lp = SR[i * seq + j] - lzr;
      float p = __expf(lp);
      if (p > 0.f) acc += p * (lp - lq);
      saved[i * seq + j] = (__expf(lq) - p) / seq;
    }
    li[i] = acc;
  }
  __syncthreads();
  float t = 0.f;
  for (int i = threadIdx.x; i < seq; i += blockDim.x) t += li[i];
  t = block_sum(t, red);
  if (threadIdx.x == 0) saved[seq * seq] = t / seq;
}
__global__ void __launch_bounds__(256) iea_bwd_kernel(const float* kf, const float* saved_all, const float* dloss,
                                                      int events, int seq, int dim, float* dkf) {
  extern __shared__ float sm[];
  const int e = blockIdx.x;
  const float* D = saved_all + (int64_t)e * (seq * seq + 1);
  const float k = dloss[0] / events;
  for (int p = threadIdx.x; p < seq * seq; p += blockDim.x) {
    int i = p / seq, j = p - i * seq;
    sm[p] = k * (D[i * seq + j] + D[j * seq + i]);
  }
  __syncthreads();
  const float* F = kf + (int64_t)e * seq * dim;
  const int k0 = blockIdx.y * CBWD_SLICE;  // (the gradient is independent per feature column: slices over the chip)
  const int ncol = dim - k0 < CBWD_SLICE ? dim - k0 : CBWD_SLICE;
  for (int idx = threadIdx.x; idx < seq * ncol; idx += blockDim.x) {
    const int i = idx / ncol, kk = k0 + idx - i * ncol;
    float g = 0.f;
    for (int j = 0; j < seq; ++j) g = fmaf(sm[i * seq + j], F[(int64_t)j * dim + kk], g);
    dkf[(int64_t)e * seq * dim + (int64_t)i * dim + kk] = g;
  }
}

// saved per event: Wm[seq][seq] = exp(-t |x_i - x_j|^2) (0 on the diagonal), S = sum_{i<j}, loss
__global__ void __launch_bounds__(256) unif_fwd_kernel(const float* part, int nsl, int stride, int seq, int dim,
                                                       float t, float* saved_all) {
  extern __shared__ float sm[];  // [seq][seq] Gram
  __shared__ float red[33];
  const int e = blockIdx.x;
  float* saved = saved_all + (int64_t)e * (seq * seq + 2);
  gram_collect(part, e, nsl, stride, seq * seq, sm);
  __syncthreads();
  // |x_i - x_j|^2 = |x_i|^2 + |x_j|^2 - 2 x_i . x_j   (torch.pdist(x)^2, loss.py:8-9)
  for (int p = threadIdx.x; p < seq * seq; p += blockDim.x) {
    const int i = p / seq, j = p - i * seq;
    const float d2 = fmaxf(sm[i * seq + i] + sm[j * seq + j] - 2.f * sm[p], 0.f);
    saved[p] = (i == j) ? 0.f : __expf(-t * d2);
  }
  __syncthreads();
  float a = 0.f;
  for (int p = threadIdx.x; p < seq * seq; p += blockDim.x) { int i = p / seq, j = p - i * seq; if (i < j) a += saved[p]; }
  a = block_sum(a, red);
  if (threadIdx.x == 0) {
    saved[seq * seq] = a;
    saved[seq * seq + 1] = __logf(a / (0.5f * seq * (seq - 1)));
  }
}
__global__ void __launch_bounds__(256) unif_bwd_kernel(const float* x, const float* saved_all, const float* dloss,
                                                       int events, int seq, int dim, float t, float* dx) {
  const int e = blockIdx.x;
  const float* Wm = saved_all + (int64_t)e * (seq * seq + 2);
  const float k = dloss[0] / events * (-2.f * t) / Wm[seq * seq];
  const float* X = x + (int64_t)e * seq * dim;
  const int k0 = blockIdx.y * CBWD_SLICE;
  const int ncol = dim - k0 < CBWD_SLICE ? dim - k0 : CBWD_SLICE;
  for (int idx = threadIdx.x; idx < seq * ncol; idx += blockDim.x) {
    const int i = idx / ncol, kk = k0 + idx - i * ncol;
    const float xi = X[(int64_t)i * dim + kk];
    float g = 0.f;
    for (int j = 0; j < seq; ++j) g = fmaf(Wm[i * seq + j], xi - X[(int64_t)j * dim + kk], g);
    dx[(int64_t)e * seq * dim + (int64_t)i * dim + kk] = k * g;
  }
}
}  // namespace

extern "C" {
int iea_diffaug_fwd(const float* x, const iea_aug_draws* d, int64_t n, int h, int w, float* y, float* mean_scratch,
                    iea_stream_t st) {
  Aug a = to_aug(d);
  if (a.ct) img_mean_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)st>>>(x, (int64_t)h * w, mean_scratch);
  diffaug_fwd_kernel<<<ew_blocks(n * h * (int64_t)w), 256, 0, (cudaStream_t)st>>>(x, a, n, h, w, y, mean_scratch);
  return check_launch("iea_diffaug_fwd");
}
int iea_diffaug_bwd(const float* dy, const iea_aug_draws* d, int64_t n, int h, int w, float* dx, float* mean_scratch,
                    iea_stream_t st) {
  Aug a = to_aug(d);
  if (a.ct) diffaug_bwd_mean_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)st>>>(dy, a, h, w, mean_scratch);
  diffaug_bwd_kernel<<<ew_blocks(n * h * (int64_t)w), 256, 0, (cudaStream_t)st>>>(dy, a, n, h, w, dx, mean_scratch);
  return check_launch("iea_diffaug_bwd");
}
int iea_loss_hinge_dis(const float* fake, const float* real, int64_t n, float* out, iea_stream_t st) {
  hinge_dis_kernel<<<1, 256, 0, (cudaStream_t)st>>>(fake, real, n, out);
  return check_launch("iea_loss_hinge_dis");
}
int iea_loss_hinge_dis_bwd(const float* fake, const float* real, const float* dout, int64_t n, float* dfake,
                           float* dreal, iea_stream_t st) {
  hinge_dis_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(fake, real, dout, n, dfake, dreal);
  return check_launch("iea_loss_hinge_dis_bwd");
}
int iea_loss_mean(const float* x, int64_t n, float scale, float* out, iea_stream_t st) {
  mean_kernel<<<1, 256, 0, (cudaStream_t)st>>>(x, n, scale, out);
  return check_launch("iea_loss_mean");
}
int iea_loss_mean_bwd(const float* dout, int64_t n, float scale, float* dx, iea_stream_t st) {
  mean_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(dout, n, scale, dx);
  return check_launch("iea_loss_mean_bwd");
}
int iea_loss_l2(const float* a, const float* b, int64_t n, float* out, iea_stream_t st) {
  l2_kernel<<<1, 256, 0, (cudaStream_t)st>>>(a, b, n, out);
  return check_launch("iea_loss_l2");
}
int iea_loss_l2_bwd(const float* a, const float* b, const float* dout, int64_t n, float* da, float* db,
                    iea_stream_t st) {
  l2_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)st>>>(a, b, dout, n, da, db);
  return check_launch("iea_loss_l2_bwd");
}
int64_t iea_loss_scratch_floats(int events, int seq, int dim) {
  return (int64_t)events * cdiv(dim, GSL) * 2 * ((int64_t)seq * seq + 2 * seq);
}
static int gram_partials(const float* A, const float* B, const float* Cm, int events, int seq, int dim, float* part,
                         cudaStream_t st) {
  const int stride = seq * seq + 2 * seq;
  const size_t smem = (size_t)(B == A ? 1 : 2) * seq * (GSL + 1) * sizeof(float);
  IEA_CHECK_ARG(smem <= 48 * 1024, "loss kernels: %d rows per event exceed the shared-memory Gram tile", seq);
  gram_partial_kernel<<<dim3(events, cdiv(dim, GSL)), 256, smem, st>>>(A, B, Cm, seq, dim, part, stride);
  return 0;
}
int iea_loss_contrastive_fwd(const float* embed, const float* proxy, int events, int seq, int dim, float temperature,
                             float margin, float* loss, float* saved, float* scratch, iea_stream_t st) {
  if (int rc = gram_partials(embed, embed, proxy, events, seq, dim, scratch, (cudaStream_t)st)) return rc;
  size_t smem = (size_t)(seq * seq + 4 * seq) * sizeof(float);
  contrastive_fwd_kernel<<<events, 256, smem, (cudaStream_t)st>>>(scratch, cdiv(dim, GSL), seq * seq + 2 * seq, seq, dim,
                                                                  temperature, margin, saved);
  int stride = 2 * seq * seq + 4 * seq + 1;
  sum_events_kernel<<<1, 32, 0, (cudaStream_t)st>>>(saved + stride - 1, events, stride, loss);
  return check_launch("iea_loss_contrastive_fwd");
}
int iea_loss_contrastive_bwd(const float* embed, const float* proxy, const float* saved, const float* dloss,
                             int events, int seq, int dim, float temperature, float* dembed, float* dproxy,
                             iea_stream_t st) {
  size_t smem = (size_t)(seq * seq + 4 * seq) * sizeof(float);
  contrastive_bwd_kernel<<<dim3(events, cdiv(dim, CBWD_SLICE)), 256, smem, (cudaStream_t)st>>>(embed, proxy, saved, dloss, events, seq, dim,
                                                                  temperature, dembed, dproxy);
  return check_launch("iea_loss_contrastive_bwd");
}
int iea_loss_iea_fwd(const float* kf, const float* kr, int events, int seq, int dim, float* loss, float* saved,
                     float* scratch, iea_stream_t st) {
  const int stride = seq * seq + 2 * seq, nsl = cdiv(dim, GSL);
  float* part_r = scratch + (int64_t)events * nsl * stride;
  if (int rc = gram_partials(kf, kf, nullptr, events, seq, dim, scratch, (cudaStream_t)st)) return rc;
  if (int rc = gram_partials(kr, kr, nullptr, events, seq, dim, part_r, (cudaStream_t)st)) return rc;
  size_t smem = (size_t)(2 * seq * seq + seq) * sizeof(float);
  iea_fwd_kernel<<<events, 256, smem, (cudaStream_t)st>>>(scratch, part_r, nsl, stride, seq, dim, saved);
  int sstride = seq * seq + 1;
  sum_events_kernel<<<1, 32, 0, (cudaStream_t)st>>>(saved + sstride - 1, events, sstride, loss);
  return check_launch("iea_loss_iea_fwd");
}
int iea_loss_iea_bwd(const float* kf, const float* saved, const float* dloss, int events, int seq, int dim,
                     float* dkf, iea_stream_t st) {
  iea_bwd_kernel<<<dim3(events, cdiv(dim, CBWD_SLICE)), 256, (size_t)seq * seq * sizeof(float), (cudaStream_t)st>>>(
      kf, saved, dloss, events, seq, dim, dkf);
  return check_launch("iea_loss_iea_bwd");
}
int iea_loss_unif_fwd(const float* x, int events, int seq, int dim, float t, float* loss, float* saved,
                      float* scratch, iea_stream_t st) {
  if (int rc = gram_partials(x, x, nullptr, events, seq, dim, scratch, (cudaStream_t)st)) return rc;
  unif_fwd_kernel<<<events, 256, (size_t)seq * seq * sizeof(float), (cudaStream_t)st>>>(
      scratch, cdiv(dim, GSL), seq * seq + 2 * seq, seq, dim, t, saved);
  int stride = seq * seq + 2;
  sum_events_kernel<<<1, 32, 0, (cudaStream_t)st>>>(saved + stride - 1, events, stride, loss);
  return check_launch("iea_loss_unif_fwd");
}
int iea_loss_unif_bwd(const float* x, const float* saved, const float* dloss, int events, int seq, int dim, float t,
                      float* dx, iea_stream_t st) {
  unif_bwd_kernel<<<dim3(events, cdiv(dim, CBWD_SLICE)), 256, 0, (cudaStream_t)st>>>(x, saved, dloss, events, seq, dim, t, dx);
  return check_launch("iea_loss_unif_bwd");
}
}
