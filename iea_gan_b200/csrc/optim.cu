// optim.cu -- the per-step parameter update of SURVEY section 8(f) N2 as multi-tensor kernels:
// gradient-norm clipping (torch.nn.utils.clip_grad_norm_, train_fns.py:133-136,190-191), Adam
// (torch.optim.Adam as constructed at model.py:410-416,858-864: betas (B1,B2), eps, no weight decay, no amsgrad)
// and the exponential moving average of the generator (utils/__init__.py:825-837) in THREE launches per net,
// whatever the number of tensors (215 in G, 132 in D; the reference's python loops issue ~10 launches per tensor).
// Pure HBM streaming: 4 reads + 3 (4 with EMA) writes of 4 bytes per parameter, 16-byte vector accesses.
// Everything the update needs that changes from step to step lives in device memory (`scalars`, `hyper`), so a
// captured CUDA graph of the step replays correctly.
#include "common.cuh"
using namespace iea;

namespace {

constexpr int MT_THREADS = 256;

// one block per chunk: sum of squares of the gradient slice, fixed order -> deterministic
__global__ void __launch_bounds__(MT_THREADS) mt_sqnorm_kernel(const iea_mt_chunk* chunks, float* partial) {
  __shared__ float red[33];
  const iea_mt_chunk c = chunks[blockIdx.x];
  float s = 0.f;
  const int n4 = (reinterpret_cast<uintptr_t>(c.g) & 15) == 0 ? c.n >> 2 : 0;
  const float4* g4 = reinterpret_cast<const float4*>(c.g);
  for (int i = threadIdx.x; i < n4; i += MT_THREADS) {
    const float4 v = g4[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  for (int i = (n4 << 2) + threadIdx.x; i < c.n; i += MT_THREADS) s = fmaf(c.g[i], c.g[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// scalars: [0] step count (float, exact up to 2^24), [1] clip coefficient, [2] 1-b1^t, [3] sqrt(1-b2^t),
//          [4] total gradient norm (before clipping)
__global__ void __launch_bounds__(MT_THREADS) mt_prepare_kernel(const float* partial, int n_chunks, float max_norm,
                                                                float b1, float b2, float* scalars) {
  __shared__ double red[MT_THREADS];
  double s = 0.0;
  if (partial)
    for (int i = threadIdx.x; i < n_chunks; i += MT_THREADS) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < MT_THREADS; ++i) t += red[i];
    const float norm = (float)sqrt(t);
    float coef = 1.f;
    if (partial && max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));  // torch's clip_coef_clamped
    const float step = scalars[0] + 1.f;
    scalars[0] = step;
    scalars[1] = coef;
    scalars[2] = (float)(1.0 - pow((double)b1, (double)step));
    scalars[3] = (float)sqrt(1.0 - pow((double)b2, (double)step));
    scalars[4] = norm;
  }
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float b1, float b2, float eps,
                                      float step_size, float inv_bc2s) {
  m = b1 * m + (1.f - b1) * g;                      // exp_avg.lerp_(grad, 1 - beta1)
  v = b2 * v + (1.f - b2) * g * g;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) * inv_bc2s + eps;    // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p -= step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// hyper: [0] learning rate, [1] EMA decay (device memory: a schedule or the ema start switch changes them
// without re-capturing a graph)
__global__ void __launch_bounds__(MT_THREADS) mt_adam_kernel(const iea_mt_chunk* chunks, float b1, float b2, float eps,
                                                             const float* scalars, const float* hyper) {
  const iea_mt_chunk c = chunks[blockIdx.x];
  const float coef = scalars[1], step_size = hyper[0] / scalars[2], inv_bc2s = 1.f / scalars[3];
  const float d = hyper[1];
  const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) | reinterpret_cast<uintptr_t>(c.m) |
                     reinterpret_cast<uintptr_t>(c.v) | reinterpret_cast<uintptr_t>(c.ema)) & 15) == 0;
  const int n4 = vec ? c.n >> 2 : 0;
  float4* p4 = reinterpret_cast<float4*>(c.p);
  const float4* g4 = reinterpret_cast<const float4*>(c.g);
  float4* m4 = reinterpret_cast<float4*>(c.m);
  float4* v4 = reinterpret_cast<float4*>(c.v);
  float4* e4 = reinterpret_cast<float4*>(c.ema);
  for (int i = threadIdx.x; i < n4; i += MT_THREADS) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    adam1(p.x, g.x * coef, m.x, v.x, b1, b2, eps, step_size, inv_bc2s);
    adam1(p.y, g.y * coef, m.y, v.y, b1, b2, eps, step_size, inv_bc2s);
    adam1(p.z, g.z * coef, m.z, v.z, b1, b2, eps, step_size, inv_bc2s);
    adam1(p.w, g.w * coef, m.w, v.w, b1, b2, eps, step_size, inv_bc2s);
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (c.ema) {
      float4 e = e4[i];
      e.x = e.x * d + p.x * (1.f - d); e.y = e.y * d + p.y * (1.f - d);
      e.z = e.z * d + p.z * (1.f - d); e.w = e.w * d + p.w * (1.f - d);
      e4[i] = e;
    }
  }
  for (int i = (n4 << 2) + threadIdx.x; i < c.n; i += MT_THREADS) {
    float p = c.p[i], m = c.m[i], v = c.v[i];
    adam1(p, c.g[i] * coef, m, v, b1, b2, eps, step_size, inv_bc2s);
    c.p[i] = p; c.m[i] = m; c.v[i] = v;
    if (c.ema) c.ema[i] = c.ema[i] * d + p * (1.f - d);
  }
}

// dst(p) = d*dst + (1-d)*src(g): the EMA of everything that is not an optimised parameter (u0, sv0, running stats)
__global__ void __launch_bounds__(MT_THREADS) mt_lerp_kernel(const iea_mt_chunk* chunks, const float* hyper) {
  const iea_mt_chunk c = chunks[blockIdx.x];
  const float d = hyper[1];
  for (int i = threadIdx.x; i < c.n; i += MT_THREADS) c.p[i] = c.p[i] * d + c.g[i] * (1.f - d);
}

}  // namespace

extern "C" {

int iea_mt_sqnorm(const iea_mt_chunk* chunks, int n_chunks, float* partial, iea_stream_t stream) {
  IEA_CHECK_ARG(n_chunks > 0, "iea_mt_sqnorm: empty chunk table");
  mt_sqnorm_kernel<<<n_chunks, MT_THREADS, 0, (cudaStream_t)stream>>>(chunks, partial);
  return check_launch("iea_mt_sqnorm");
}

int iea_mt_adam(const iea_mt_chunk* chunks, int n_chunks, const float* partial, float max_norm, float beta1,
                float beta2, float eps, float* scalars, const float* hyper, iea_stream_t stream) {
  IEA_CHECK_ARG(n_chunks > 0, "iea_mt_adam: empty chunk table");
  cudaStream_t st = (cudaStream_t)stream;
  mt_prepare_kernel<<<1, MT_THREADS, 0, st>>>(partial, n_chunks, max_norm, beta1, beta2, scalars);
  mt_adam_kernel<<<n_chunks, MT_THREADS, 0, st>>>(chunks, beta1, beta2, eps, scalars, hyper);
  return check_launch("iea_mt_adam");
}

int iea_mt_lerp(const iea_mt_chunk* chunks, int n_chunks, const float* hyper, iea_stream_t stream) {
  IEA_CHECK_ARG(n_chunks > 0, "iea_mt_lerp: empty chunk table");
  mt_lerp_kernel<<<n_chunks, MT_THREADS, 0, (cudaStream_t)stream>>>(chunks, hyper);
  return check_launch("iea_mt_lerp");
}
}
