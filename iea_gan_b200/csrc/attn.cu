// attn.cu -- BigGAN self-attention core (layers.py:289-299): beta = softmax(theta^T phi) with no
// 1/sqrt(d) scaling, o = g beta^T.  Flash-style: the (hw x hw/4) attention map is never written to
// HBM; forward keeps a running max / sum per query and saves only the log-sum-exp.  Backward is two
// deterministic passes (query-parallel for d theta, key-parallel for d phi / d g), no atomics.
//   theta [n][hw][ck], phi [n][hwk][ck], g [n][hwk][cv]  ->  o [n][hw][cv]
// CUDA-core version for odd shapes / fp32 activations; the shipped shape (ck 32, cv 128, bf16) runs on the
// tcgen05 kernels of attn_tc.cu.
#include "common.cuh"
using namespace iea;

static inline bool tc_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
namespace {
constexpr int QT = 32;   // queries per block
constexpr int KT = 64;   // keys per smem tile (forward / pass A)

__global__ void __launch_bounds__(256) attn_fwd_kernel(const void* theta, const void* phi, const void* g, int dtp,
                                                       int hw, int hwk, int ck, int cv, void* o, float* lse) {
  extern __shared__ float sm[];
  float* th = sm;                         // [QT][ck+1]
  float* ph = th + QT * (ck + 1);         // [KT][ck+1]
  float* gv = ph + KT * (ck + 1);         // [KT][cv]
  const int64_t n = blockIdx.y;
  const int q0 = blockIdx.x * QT;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < QT * ck; i += 256) {
    int q = i / ck, c = i - q * ck;
    th[q * (ck + 1) + c] = (q0 + q < hw) ? ld_act(theta, dtp, (n * hw + q0 + q) * ck + c) : 0.f;
  }
  float m[4], l[4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -3.0e38f; l[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;
  }
  for (int k0 = 0; k0 < hwk; k0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * ck; i += 256) {
      int j = i / ck, c = i - j * ck;
      ph[j * (ck + 1) + c] = (k0 + j < hwk) ? ld_act(phi, dtp, (n * hwk + k0 + j) * ck + c) : 0.f;
    }
    for (int i = threadIdx.x; i < KT * cv; i += 256) {
      int j = i / cv;
      gv[i] = (k0 + j < hwk) ? ld_act(g, dtp, (n * hwk + k0) * cv + i) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const int q = wid + 8 * qi;
      float s0 = 0.f, s1 = 0.f;
      for (int c = 0; c < ck; ++c) {
        float t = th[q * (ck + 1) + c];
        s0 = fmaf(t, ph[lane * (ck + 1) + c], s0);
        s1 = fmaf(t, ph[(lane + 32) * (ck + 1) + c], s1);
      }
      if (k0 + lane >= hwk) s0 = -3.0e38f;
      if (k0 + lane + 32 >= hwk) s1 = -3.0e38f;
      float mx = warp_max(fmaxf(s0, s1));
      float mn = fmaxf(m[qi], mx);
      float corr = __expf(m[qi] - mn);
      float p0 = (k0 + lane < hwk) ? __expf(s0 - mn) : 0.f;
      float p1 = (k0 + lane + 32 < hwk) ? __expf(s1 - mn) : 0.f;
      l[qi] = l[qi] * corr + warp_sum(p0 + p1);
      m[qi] = mn;
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[qi][t] *= corr;
      for (int j = 0; j < 32; ++j) {
        float pj0 = __shfl_sync(0xffffffffu, p0, j), pj1 = __shfl_sync(0xffffffffu, p1, j);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          int c = lane + 32 * t;
          if (c < cv) acc[qi][t] = fmaf(pj0, gv[j * cv + c], fmaf(pj1, gv[(j + 32) * cv + c], acc[qi][t]));
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = q0 + wid + 8 * qi;
    if (q >= hw) continue;
    float inv = 1.f / l[qi];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int c = lane + 32 * t;
      if (c < cv) st_act(o, dtp, (n * hw + q) * cv + c, acc[qi][t] * inv);
    }
    if (lane == 0) lse[n * hw + q] = m[qi] + __logf(l[qi]);
  }
}

// pass A: d theta (one warp per query, lane = key inside a 32-key tile); also writes D_q = dO_q . O_q
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const void* d_o, const void* theta, const void* phi,
                                                         const void* g, const void* o, const float* lse, int dtp,
                                                         int hw, int hwk, int ck, int cv, void* dtheta, float* dq) {
  extern __shared__ float sm[];
  float* th = sm;                        // [QT][ck+1]
  float* dO = th + QT * (ck + 1);        // [QT][cv+1]
  float* ph = dO + QT * (cv + 1);        // [32][ck+1]
  float* gv = ph + 32 * (ck + 1);        // [32][cv+1]
  const int64_t n = blockIdx.y;
  const int q0 = blockIdx.x * QT;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < QT * ck; i += 256) {
    int q = i / ck, c = i - q * ck;
    th[q * (ck + 1) + c] = (q0 + q < hw) ? ld_act(theta, dtp, (n * hw + q0 + q) * ck + c) : 0.f;
  }
  for (int i = threadIdx.x; i < QT * cv; i += 256) {
    int q = i / cv, c = i - q * cv;
    dO[q * (cv + 1) + c] = (q0 + q < hw) ? ld_act(d_o, dtp, (n * hw + q0 + q) * cv + c) : 0.f;
  }
  __syncthreads();
  float D[4], ls[4], dth[4];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = wid + 8 * qi;
    float a = 0.f;
    if (q0 + q < hw)
      for (int c = lane; c < cv; c += 32) a = fmaf(dO[q * (cv + 1) + c], ld_act(o, dtp, (n * hw + q0 + q) * cv + c), a);
    D[qi] = warp_sum(a);
    ls[qi] = (q0 + q < hw) ? lse[n * hw + q0 + q] : 0.f;
    dth[qi] = 0.f;
    if (lane == 0 && q0 + q < hw) dq[n * hw + q0 + q] = D[qi];
  }
  for (int k0 = 0; k0 < hwk; k0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * ck; i += 256) {
      int j = i / ck, c = i - j * ck;
      ph[j * (ck + 1) + c] = (k0 + j < hwk) ? ld_act(phi, dtp, (n * hwk + k0 + j) * ck + c) : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * cv; i += 256) {
      int j = i / cv, c = i - j * cv;
      gv[j * (cv + 1) + c] = (k0 + j < hwk) ? ld_act(g, dtp, (n * hwk + k0 + j) * cv + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const int q = wid + 8 * qi;
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < ck; ++c) s = fmaf(th[q * (ck + 1) + c], ph[lane * (ck + 1) + c], s);
      for (int c = 0; c < cv; ++c) dp = fmaf(dO[q * (cv + 1) + c], gv[lane * (cv + 1) + c], dp);
      float p = (k0 + lane < hwk) ? __expf(s - ls[qi]) : 0.f;
      float dS = p * (dp - D[qi]);
      for (int j = 0; j < 32; ++j) {
        float dsj = __shfl_sync(0xffffffffu, dS, j);
        if (lane < ck) dth[qi] = fmaf(dsj, ph[j * (ck + 1) + lane], dth[qi]);
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int q = q0 + wid + 8 * qi;
    if (q < hw && lane < ck) st_act(dtheta, dtp, (n * hw + q) * ck + lane, dth[qi]);
  }
}

// pass B: d phi, d g (one warp per 4 keys of a 32-key block; lane = query inside a 32-query tile)
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const void* d_o, const void* theta, const void* phi,
                                                          const void* g, const float* lse, const float* dq, int dtp,
                                                          int hw, int hwk, int ck, int cv, void* dphi, void* dg) {
  extern __shared__ float sm[];
  float* ph = sm;                        // [32][ck+1]   this block's keys
  float* gv = ph + 32 * (ck + 1);        // [32][cv+1]
  float* th = gv + 32 * (cv + 1);        // [32][ck+1]   query tile
  float* dO = th + 32 * (ck + 1);        // [32][cv+1]
  float* lq = dO + 32 * (cv + 1);        // [32] lse, [32] D
  const int64_t n = blockIdx.y;
  const int k0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * ck; i += 256) {
    int j = i / ck, c = i - j * ck;
    ph[j * (ck + 1) + c] = (k0 + j < hwk) ? ld_act(phi, dtp, (n * hwk + k0 + j) * ck + c) : 0.f;
  }
  for (int i = threadIdx.x; i < 32 * cv; i += 256) {
    int j = i / cv, c = i - j * cv;
    gv[j * (cv + 1) + c] = (k0 + j < hwk) ? ld_act(g, dtp, (n * hwk + k0 + j) * cv + c) : 0.f;
  }
  float dph[4], dgv[4][4];
#pragma unroll
  for (int ki = 0; ki < 4; ++ki) {
    dph[ki] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) dgv[ki][t] = 0.f;
  }
  for (int q0 = 0; q0 < hw; q0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * ck; i += 256) {
      int q = i / ck, c = i - q * ck;
      th[q * (ck + 1) + c] = (q0 + q < hw) ? ld_act(theta, dtp, (n * hw + q0 + q) * ck + c) : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * cv; i += 256) {
      int q = i / cv, c = i - q * cv;
      dO[q * (cv + 1) + c] = (q0 + q < hw) ? ld_act(d_o, dtp, (n * hw + q0 + q) * cv + c) : 0.f;
    }
    if (threadIdx.x < 32) {
      bool ok = q0 + threadIdx.x < hw;
      lq[threadIdx.x] = ok ? lse[n * hw + q0 + threadIdx.x] : 0.f;
      lq[32 + threadIdx.x] = ok ? dq[n * hw + q0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int ki = 0; ki < 4; ++ki) {
      const int j = wid * 4 + ki;
      float s = 0.f, dp = 0.f;
      for (int c = 0; c < ck; ++c) s = fmaf(th[lane * (ck + 1) + c], ph[j * (ck + 1) + c], s);
      for (int c = 0; c < cv; ++c) dp = fmaf(dO[lane * (cv + 1) + c], gv[j * (cv + 1) + c], dp);
      float p = (q0 + lane < hw && k0 + j < hwk) ? __expf(s - lq[lane]) : 0.f;
      float dS = p * (dp - lq[32 + lane]);
      for (int q = 0; q < 32; ++q) {
        float pq = __shfl_sync(0xffffffffu, p, q), dsq = __shfl_sync(0xffffffffu, dS, q);
        if (lane < ck) dph[ki] = fmaf(dsq, th[q * (ck + 1) + lane], dph[ki]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          int c = lane + 32 * t;
          if (c < cv) dgv[ki][t] = fmaf(pq, dO[q * (cv + 1) + c], dgv[ki][t]);
        }
      }
    }
  }
#pragma unroll
  for (int ki = 0; ki < 4; ++ki) {
    const int j = k0 + wid * 4 + ki;
    if (j >= hwk) continue;
    if (lane < ck) st_act(dphi, dtp, (n * hwk + j) * ck + lane, dph[ki]);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int c = lane + 32 * t;
      if (c < cv) st_act(dg, dtp, (n * hwk + j) * cv + c, dgv[ki][t]);
    }
  }
}
}  // namespace

int iea_attn_tc_ok(int dtype, int64_t n, int hw, int hwk, int ck, int cv, const void* a, const void* b, const void* c,
                   const void* d);
int iea_attn_fwd_tc(const void* theta, const void* phi, const void* g, int64_t n, int hw, int hwk, void* o, float* lse,
                    cudaStream_t s);
int iea_attn_bwd_tc(const void* d_o, const void* theta, const void* phi, const void* g, const void* o, const float* lse,
                    int64_t n, int hw, int hwk, void* dtheta, void* dphi, void* dg, float* dq, cudaStream_t s);

extern "C" int iea_attn_fwd(const void* theta, const void* phi, const void* g, int dtype, int64_t n, int hw, int hwk,
                            int ck, int cv, void* o, float* lse, iea_stream_t stream) {
  IEA_CHECK_ARG(ck <= 32 && cv <= 128 && ck > 0 && cv > 0, "iea_attn_fwd: ck=%d cv=%d outside the built range", ck, cv);
  if (iea_attn_tc_ok(dtype, n, hw, hwk, ck, cv, theta, phi, g, o))
    return iea_attn_fwd_tc(theta, phi, g, n, hw, hwk, o, lse, (cudaStream_t)stream);
  size_t smem = (size_t)(QT * (ck + 1) + KT * (ck + 1) + KT * cv) * sizeof(float);
  IEA_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(hw, QT), (unsigned)n);
  attn_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(theta, phi, g, dtype, hw, hwk, ck, cv, o, lse);
  return check_launch("iea_attn_fwd");
}

extern "C" int iea_attn_bwd(const void* d_o, const void* theta, const void* phi, const void* g, const void* o,
                            const float* lse, int dtype, int64_t n, int hw, int hwk, int ck, int cv, void* dtheta,
                            void* dphi, void* dg, float* dq_scratch, iea_stream_t stream) {
  IEA_CHECK_ARG(ck <= 32 && cv <= 128 && ck > 0 && cv > 0, "iea_attn_bwd: ck=%d cv=%d outside the built range", ck, cv);
  if (iea_attn_tc_ok(dtype, n, hw, hwk, ck, cv, theta, phi, g, o) && tc_al16(d_o) && tc_al16(dtheta) && tc_al16(dphi) && tc_al16(dg))
    return iea_attn_bwd_tc(d_o, theta, phi, g, o, lse, n, hw, hwk, dtheta, dphi, dg, dq_scratch, (cudaStream_t)stream);
  cudaStream_t s = (cudaStream_t)stream;
  size_t sa = (size_t)(QT * (ck + 1) + QT * (cv + 1) + 32 * (ck + 1) + 32 * (cv + 1)) * sizeof(float);
  IEA_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa));
  attn_bwd_q_kernel<<<dim3(cdiv(hw, QT), (unsigned)n), 256, sa, s>>>(d_o, theta, phi, g, o, lse, dtype, hw, hwk, ck,
                                                                      cv, dtheta, dq_scratch);
  size_t sb = (size_t)(2 * 32 * (ck + 1) + 2 * 32 * (cv + 1) + 64) * sizeof(float);
  IEA_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
  attn_bwd_kv_kernel<<<dim3(cdiv(hwk, 32), (unsigned)n), 256, sb, s>>>(d_o, theta, phi, g, lse, dq_scratch, dtype, hw,
                                                                       hwk, ck, cv, dphi, dg);
  return check_launch("iea_attn_bwd");
}
