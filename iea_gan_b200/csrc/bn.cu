// bn.cu -- batch-norm statistics plumbing (per event of 40 images).
// layers.py:656-689 (ccbn.forward), :728-742 (bn.forward), F.batch_norm semantics:
// normalise with the biased batch variance, running_var gets the unbiased one, momentum 0.1.
// The heavy passes (sums in the producer epilogue, apply in the consumer prologue) live in
// the conv kernels; these kernels are the tiny per-(event, channel) glue around them.
#include "common.cuh"
using namespace iea;

namespace {

// running statistics <- one event's batch statistics.  mode bits (iea_bn_finalize): 1 training
// (F.batch_norm: momentum update with the UNBIASED variance, layers.py:664-673); 2 myBN (momentum
// update with the biased variance, layers.py:585-592); 4 myBN standing statistics (plain sums,
// layers.py:579-582).
__device__ __forceinline__ void fold_running(float& rm, float& rv, float mean, double var, double count,
                                             float momentum, int mode) {
  if (mode & 4) { rm += mean; rv += (float)var; return; }
  const double v = (mode & 2) ? var : (count > 1.0 ? var * count / (count - 1.0) : var);
  rm = (1.f - momentum) * rm + momentum * mean;
  rv = (1.f - momentum) * rv + momentum * (float)v;
}

// partials[e][t][c][2]; grid (events*tiles, channel chunks)
__global__ void __launch_bounds__(256) bn_stats_kernel(const void* x, int x_dtype, int x_ld, int rows_per_event, int c,
                                                       int tiles, float* partials, int cb) {
  __shared__ float r1[256], r2[256];
  const int e = blockIdx.x / tiles, t = blockIdx.x % tiles;
  const int rows_per_tile = (rows_per_event + tiles - 1) / tiles;
  const int64_t r0 = (int64_t)e * rows_per_event + (int64_t)t * rows_per_tile;
  int64_t r1e = r0 + rows_per_tile;
  const int64_t rend = (int64_t)(e + 1) * rows_per_event;
  if (r1e > rend) r1e = rend;
  const int lanes = 256 / cb;
  const int cl = threadIdx.x % cb, pl = threadIdx.x / cb;
  const int cc = blockIdx.y * cb + cl;
  float s1 = 0.f, s2 = 0.f;
  if (cc < c && pl < lanes)
    for (int64_t m = r0 + pl; m < r1e; m += lanes) {
      float v = ld_act(x, x_dtype, m * x_ld + cc);
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
  __syncthreads();
  if (pl == 0 && cc < c) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) { a += r1[l * cb + cl]; b += r2[l * cb + cl]; }
    float* o = partials + (((int64_t)e * tiles + t) * c + cc) * 2;
    o[0] = a; o[1] = b;
  }
}

// stage 1: per (event, channel) reduce the tile partials in double (fixed order -> deterministic).
// grid (events, ceil(c/8)); 256 threads = 8 channels x 32 tile lanes.  red[e][c] = (mean, biased var)
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* partials, int tiles, double count, int c,
                                                        float* mean_out, float* var_out) {
  __shared__ double r1[256], r2[256];
  const int e = blockIdx.x, cl = threadIdx.x & 7, tl = threadIdx.x >> 3, cc = blockIdx.y * 8 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (cc < c) {
    const float* p = partials + ((int64_t)e * tiles * c + cc) * 2;
    for (int t = tl; t < tiles; t += 32) {
      const float2 v = *reinterpret_cast<const float2*>(p + (int64_t)t * c * 2);
      s1 += (double)v.x; s2 += (double)v.y;
    }
  }
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
  __syncthreads();
  if (tl == 0 && cc < c) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < 32; ++l) { a += r1[l * 8 + cl]; b += r2[l * 8 + cl]; }
    const double m = a / count;
    double var = b / count - m * m;
    if (var < 0.0) var = 0.0;
    mean_out[e * c + cc] = (float)m;
    var_out[e * c + cc] = (float)var;
  }
}

// stage 2: one thread per (image, channel): scale/shift; the threads of image 0 also fold the events
// into the running statistics in order (momentum update is sequential in the events).
__global__ void bn_affine_kernel(int events, double count, int imgs, int c, const float* gain, int64_t gain_ld,
                                 float gain_add, const float* bias, int64_t bias_ld, float* stored_mean,
                                 float* stored_var, int training, float momentum, float eps, float* mean_io,
                                 float* rstd_io, float* scale, float* shift) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)events * imgs * c;
  if (idx >= total) return;
  const int cc = idx % c;
  const int64_t n = idx / c;
  const int e = (int)(n / imgs);
  float mean, rstd;
  if (training) {
    mean = mean_io[e * c + cc];
    rstd = (float)(1.0 / sqrt((double)rstd_io[e * c + cc] + (double)eps));  // rstd_io holds the variance here
  } else {
    mean = stored_mean[cc];
    rstd = rsqrtf(stored_var[cc] + eps);
  }
  const float gn = gain_add + (gain ? gain[n * gain_ld + cc] : 0.f);
  const float bs = bias ? bias[n * bias_ld + cc] : 0.f;
  const float sc = rstd * gn;
  scale[n * c + cc] = sc;
  shift[n * c + cc] = bs - mean * sc;
}
__global__ void bn_running_kernel(int events, double count, int c, float* stored_mean, float* stored_var,
                                  int training, float momentum, float eps, float* mean_io, float* rstd_io) {
  const int cc = blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= c) return;
  if (!training) {
    for (int e = 0; e < events; ++e) { mean_io[e * c + cc] = stored_mean[cc]; rstd_io[e * c + cc] = rsqrtf(stored_var[cc] + eps); }
    return;
  }
  float rm = stored_mean ? stored_mean[cc] : 0.f, rv = stored_var ? stored_var[cc] : 1.f;
  for (int e = 0; e < events; ++e) {
    const double var = (double)rstd_io[e * c + cc];
    fold_running(rm, rv, mean_io[e * c + cc], var, count, momentum, training);
    rstd_io[e * c + cc] = (float)(1.0 / sqrt(var + (double)eps));  // variance -> rstd, saved for backward
  }
  if (stored_mean) { stored_mean[cc] = rm; stored_var[cc] = rv; }
}

// one block per channel; threads stride over the 40 images of an event (fixed-order smem reduction)
__global__ void __launch_bounds__(64) bn_finalize_bwd_kernel(const float* dscale, const float* dshift, const float* scale,
                                       const float* mean, const float* rstd, int events, int imgs, double count,
                                       int c, const float* gain, int64_t gain_ld, float gain_add, float* dgain,
                                       int64_t dgain_ld, float* dbias, int64_t dbias_ld, int reduce_over_n,
                                       int training, float* ds1, float* ds2) {
  __shared__ float r_m[64], r_r[64], r_g[64], r_b[64];
  const int cc = blockIdx.x, t = threadIdx.x;
  float dg_tot = 0.f, db_tot = 0.f;
  for (int e = 0; e < events; ++e) {
    const float mu = mean[e * c + cc], rs = rstd[e * c + cc];
    float dmean = 0.f, drstd = 0.f, dg = 0.f, db = 0.f;
    for (int i = t; i < imgs; i += 64) {
      const int64_t n = (int64_t)e * imgs + i;
      const float dsc = dscale[n * c + cc], dsh = dshift[n * c + cc];
      const float gn = gain_add + (gain ? gain[n * gain_ld + cc] : 0.f);
      const float dsc_tot = dsc - mu * dsh;  // scale = rstd*gn ; shift = bias - mean*scale
      const float dgn = rs * dsc_tot;
      drstd += gn * dsc_tot;
      dmean -= dsh * scale[n * c + cc];
      if (reduce_over_n) { dg += dgn; db += dsh; }
      else {
        if (dgain) dgain[n * dgain_ld + cc] = dgn;
        if (dbias) dbias[n * dbias_ld + cc] = dsh;
      }
    }
    r_m[t] = dmean; r_r[t] = drstd; r_g[t] = dg; r_b[t] = db;
    __syncthreads();
    if (t == 0) {
      float a_m = 0.f, a_r = 0.f;
      for (int k = 0; k < 64; ++k) { a_m += r_m[k]; a_r += r_r[k]; dg_tot += r_g[k]; db_tot += r_b[k]; }
      if (ds1) {
        float a = 0.f, b2 = 0.f;
        if (training) {  // rstd = (var+eps)^-1/2, var = S2/M - mean^2, mean = S1/M
          const double dvar = -0.5 * (double)a_r * (double)rs * (double)rs * (double)rs;
          const double dm = (double)a_m - 2.0 * (double)mu * dvar;
          a = (float)(dm / count);
          b2 = (float)(dvar / count);
        }
        ds1[e * c + cc] = a;
        ds2[e * c + cc] = b2;
      }
    }
    __syncthreads();
  }
  if (reduce_over_n && t == 0) {
    if (dgain) dgain[cc] = dg_tot;
    if (dbias) dbias[cc] = db_tot;
  }
}

__global__ void affine_act_kernel(const void* x, int x_dtype, const float* scale, const float* shift, int64_t n,
                                  int64_t hw, int c, int relu, void* y, int y_dtype) {
  const int64_t total = n * hw * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cc = i % c;
    int64_t nn = i / ((int64_t)hw * c);
    float v = fmaf(ld_act(x, x_dtype, i), scale[nn * c + cc], shift[nn * c + cc]);
    if (relu) v = fmaxf(v, 0.f);
    st_act(y, y_dtype, i, v);
  }
}

// training-mode finalize in ONE launch: grid (events, ceil(c/8)) blocks reduce the tile partials of their
// (event, 8 channels) in double exactly like bn_reduce_kernel; the LAST block to finish for a channel group
// (ticket counter) then does everything that needs all events of those channels: running-statistics update in
// event order, variance -> rstd, and the per-image scale / shift the consumer convolution's prologue reads.
// Fixed summation orders throughout: deterministic, whichever block ends up last.
// `ticket` is caller-owned (one self-resetting counter per 8-channel group, zero before the first use): every
// batch-norm layer passes its own, so finalize launches in flight on different streams never share counters.
__global__ void __launch_bounds__(256) bn_finalize_fused_kernel(unsigned int* g_bn_ticket, int mode,
                                                                const float* partials, int tiles, double count, int events,
                                                                int imgs, int c, const float* gain, int64_t gain_ld,
                                                                float gain_add, const float* bias, int64_t bias_ld,
                                                                float* stored_mean, float* stored_var, float momentum,
                                                                float eps, float* mean_io, float* rstd_io, float* scale,
                                                                float* shift) {
  __shared__ double r1[256], r2[256];
  __shared__ float mean_s[8], rstd_s[8];
  __shared__ bool last;
  const int e = blockIdx.x, cl = threadIdx.x & 7, tl = threadIdx.x >> 3, cc = blockIdx.y * 8 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (cc < c) {
    const float* p = partials + ((int64_t)e * tiles * c + cc) * 2;
    int t = tl;
    for (; t + 96 < tiles; t += 128) {  // four independent loads in flight
      const float2 v0 = *reinterpret_cast<const float2*>(p + (int64_t)t * c * 2);
      const float2 v1 = *reinterpret_cast<const float2*>(p + (int64_t)(t + 32) * c * 2);
      const float2 v2 = *reinterpret_cast<const float2*>(p + (int64_t)(t + 64) * c * 2);
      const float2 v3 = *reinterpret_cast<const float2*>(p + (int64_t)(t + 96) * c * 2);
      s1 += (double)v0.x; s2 += (double)v0.y; s1 += (double)v1.x; s2 += (double)v1.y;
      s1 += (double)v2.x; s2 += (double)v2.y; s1 += (double)v3.x; s2 += (double)v3.y;
    }
    for (; t < tiles; t += 32) {
      const float2 v = *reinterpret_cast<const float2*>(p + (int64_t)t * c * 2);
      s1 += (double)v.x; s2 += (double)v.y;
    }
  }
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
  __syncthreads();
  if (tl == 0) {
    float mu = 0.f, rs = 0.f;
    if (cc < c) {
      double a = 0.0, b = 0.0;
      for (int l = 0; l < 32; ++l) { a += r1[l * 8 + cl]; b += r2[l * 8 + cl]; }
      const double m = a / count;
      double var = b / count - m * m;
      if (var < 0.0) var = 0.0;
      mu = (float)m;
      rs = (float)(1.0 / sqrt((double)(float)var + (double)eps));  // (from the fp32-rounded variance, like the fold below)
      mean_io[e * c + cc] = mu;
      rstd_io[e * c + cc] = (float)var;  // variance until the last block of the channel group converts it
    }
    mean_s[cl] = mu; rstd_s[cl] = rs;
  }
  __syncthreads();
  // this event's per-image scale / shift for the 8 channels: needs nothing from the other events
  for (int i = threadIdx.x; i < imgs * 8; i += 256) {
    const int ch = blockIdx.y * 8 + (i & 7);
    if (ch >= c) continue;
    const int64_t n = (int64_t)e * imgs + (i >> 3);
    const float gn = gain_add + (gain ? gain[n * gain_ld + ch] : 0.f);
    const float bs = bias ? bias[n * bias_ld + ch] : 0.f;
    const float sc = rstd_s[i & 7] * gn;
    scale[n * c + ch] = sc;
    shift[n * c + ch] = bs - mean_s[i & 7] * sc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&g_bn_ticket[blockIdx.y], 1u);
    last = t == (unsigned int)events - 1;
    if (last) g_bn_ticket[blockIdx.y] = 0;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // last block of this channel group: running statistics in event order, variance -> rstd for the backward pass
  float* m_s = reinterpret_cast<float*>(r1);    // [events][8] mean      (events <= 64)
  float* v_s = reinterpret_cast<float*>(r2);    // [events][8] biased variance
  for (int i = threadIdx.x; i < events * 8; i += 256) {
    const int ev = i >> 3, ch = blockIdx.y * 8 + (i & 7);
    float mu = 0.f, var = 0.f;
    if (ch < c) { mu = __ldcg(mean_io + ev * c + ch); var = __ldcg(rstd_io + ev * c + ch); }
    m_s[i] = mu; v_s[i] = var;
  }
  __syncthreads();
  if (tl == 0 && cc < c) {
    float rm = stored_mean ? stored_mean[cc] : 0.f, rv = stored_var ? stored_var[cc] : 1.f;
    for (int ev = 0; ev < events; ++ev)
      fold_running(rm, rv, m_s[ev * 8 + cl], (double)v_s[ev * 8 + cl], count, momentum, mode);
    if (stored_mean) { stored_mean[cc] = rm; stored_var[cc] = rv; }
  }
  for (int i = threadIdx.x; i < events * 8; i += 256) {
    const int ev = i >> 3, ch = blockIdx.y * 8 + (i & 7);
    if (ch < c) rstd_io[ev * c + ch] = (float)(1.0 / sqrt((double)v_s[i] + (double)eps));
  }
}
}  // namespace

extern "C" int iea_bn_stats(const void* x, int x_dtype, int x_ld, int64_t rows, int rows_per_event, int c,
                            int tiles_per_event, float* partials, iea_stream_t stream) {
  IEA_CHECK_ARG(rows % rows_per_event == 0, "iea_bn_stats: rows (%lld) not a multiple of rows_per_event (%d)",
                (long long)rows, rows_per_event);
  int events = (int)(rows / rows_per_event);
  int cb = 1;
  while (cb < c && cb < 32) cb <<= 1;
  dim3 grid(events * tiles_per_event, cdiv(c, cb));
  bn_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_dtype, x_ld, rows_per_event, c, tiles_per_event,
                                                           partials, cb);
  return check_launch("iea_bn_stats");
}

extern "C" int iea_bn_finalize(const float* partials, int events, int tiles_per_event, int64_t count_per_event,
                               int imgs_per_event, int c, const float* gain, int64_t gain_ld, float gain_add,
                               const float* bias, int64_t bias_ld, float* stored_mean, float* stored_var,
                               int mode, float momentum, float eps, float* mean_out, float* rstd_out,
                               float* scale, float* shift, unsigned int* ticket, iea_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)events * imgs_per_event;
  const int training = mode;  // (bit 0 set <=> batch statistics; the other bits choose the running-statistics rule)
  IEA_CHECK_ARG(!(mode & 6) || (mode & 1), "iea_bn_finalize: myBN mode bits (%d) need the training bit", mode);
  if ((mode & 1) && ticket && events <= 64) {
    bn_finalize_fused_kernel<<<dim3(events, cdiv(c, 8)), 256, 0, st>>>(
        ticket, mode, partials, tiles_per_event, (double)count_per_event, events, imgs_per_event, c, gain, gain_ld, gain_add, bias, bias_ld,
        stored_mean, stored_var, momentum, eps, mean_out, rstd_out, scale, shift);
    return check_launch("iea_bn_finalize(fused)");
  }
  if (mode & 1)
    bn_reduce_kernel<<<dim3(events, cdiv(c, 8)), 256, 0, st>>>(partials, tiles_per_event, (double)count_per_event, c,
                                                                 mean_out, rstd_out);
  bn_affine_kernel<<<cdiv(n * c, 256), 256, 0, st>>>(events, (double)count_per_event, imgs_per_event, c, gain, gain_ld,
                                                      gain_add, bias, bias_ld, stored_mean, stored_var, training,
                                                      momentum, eps, mean_out, rstd_out, scale, shift);
  bn_running_kernel<<<cdiv(c, 64), 64, 0, st>>>(events, (double)count_per_event, c, stored_mean, stored_var, training,
                                                 momentum, eps, mean_out, rstd_out);
  return check_launch("iea_bn_finalize");
}

extern "C" int iea_bn_finalize_bwd(const float* dscale, const float* dshift, const float* scale, const float* mean,
                                   const float* rstd, int events, int imgs_per_event, int64_t count_per_event, int c,
                                   const float* gain, int64_t gain_ld, float gain_add, float* dgain, int64_t dgain_ld,
                                   float* dbias, int64_t dbias_ld, int reduce_over_n, int training, float* ds1,
                                   float* ds2, iea_stream_t stream) {
  bn_finalize_bwd_kernel<<<c, 64, 0, (cudaStream_t)stream>>>(
      dscale, dshift, scale, mean, rstd, events, imgs_per_event, (double)count_per_event, c, gain, gain_ld, gain_add,
      dgain, dgain_ld, dbias, dbias_ld, reduce_over_n, training, ds1, ds2);
  return check_launch("iea_bn_finalize_bwd");
}

extern "C" int iea_affine_act(const void* x, int x_dtype, const float* scale, const float* shift, int64_t n,
                              int64_t hw, int c, int relu, void* y, int y_dtype, iea_stream_t stream) {
  int64_t total = n * hw * c;
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  affine_act_kernel<<<(int)(b < 1 ? 1 : b), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, scale, shift, n, hw, c, relu,
                                                                          y, y_dtype);
  return check_launch("iea_affine_act");
}
