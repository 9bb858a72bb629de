// sn.cu -- grouped spectral-norm power iteration + weight repack, and its backward.
// Replaces layers.py:89-111 (power_iteration) and layers.py:151-165 (SN.W_) for every
// spectrally-normalised layer of a net in four launches (HBM/latency bound; W is read
// twice, the second time from L2).
#include "common.cuh"
using namespace iea;

// phase A: partial v = sum_{i in chunk} u[i] * W[i, :]
__global__ void __launch_bounds__(256) sn_phase_a(const iea_sn_layer* L, const int32_t* chunks, float* scratch) {
  const int32_t* ch = chunks + 3 * blockIdx.x;
  const iea_sn_layer& l = L[ch[0]];
  if (!l.spectral) return;
  const int cols = l.cin * l.taps, r0 = ch[1], r1 = ch[2];
  float* pv = scratch + l.scratch_off + (int64_t)(blockIdx.x - l.chunk0) * cols;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    float acc = 0.f;
    const float* w = l.w + (int64_t)r0 * cols + j;
    for (int i = r0; i < r1; ++i, w += cols) acc = fmaf(l.u_in[i], *w, acc);
    pv[j] = acc;
  }
}

// phase B: v = normalize(sum of partials)
__global__ void __launch_bounds__(256) sn_phase_b(const iea_sn_layer* L, float* scratch) {
  __shared__ float red[33];
  const iea_sn_layer& l = L[blockIdx.x];
  if (!l.spectral) return;
  const int cols = l.cin * l.taps;
  const float* pv = scratch + l.scratch_off;
  float ss = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < l.nchunks; ++c) acc += pv[(int64_t)c * cols + j];
    l.v_out[j] = acc;
    ss = fmaf(acc, acc, ss);
  }
  float nrm = sqrtf(block_sum(ss, red));
  float inv = 1.f / fmaxf(nrm, l.eps);
  for (int j = threadIdx.x; j < cols; j += blockDim.x) l.v_out[j] *= inv;
}

// phase C: t[i] = W[i,:] . v, and repack W into the conv layouts
__global__ void __launch_bounds__(256) sn_phase_c(const iea_sn_layer* L, const int32_t* chunks, float* scratch) {
  const int32_t* ch = chunks + 3 * blockIdx.x;
  const iea_sn_layer& l = L[ch[0]];
  const int cols = l.cin * l.taps, r0 = ch[1], r1 = ch[2], taps = l.taps, cin = l.cin, rows = l.rows;
  float* t = scratch + l.scratch_off + (int64_t)l.nchunks * cols;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = r0 + wid; i < r1; i += nw) {
    const float* w = l.w + (int64_t)i * cols;
    float acc = 0.f;
    for (int j = lane; j < cols; j += 32) {
      float wv = w[j];
      if (l.spectral) acc = fmaf(wv, l.v_out[j], acc);
      int ci = j / taps, tp = j - ci * taps;
      if (l.pack_fprop) st_act(l.pack_fprop, l.pack_dtype, ((int64_t)i * taps + tp) * cin + ci, wv);
      if (l.pack_dgrad) st_act(l.pack_dgrad, l.pack_dtype, ((int64_t)ci * taps + (taps - 1 - tp)) * l.pack_dgrad_ld + i, wv);
      if (l.pack_tc_fprop) {  // [tap][kb][chunk][co][8], k blocks of min(64, cin) input channels
        const int cin_p = l.pack_tc_cin;  // cin rounded up to 16 (1-channel stem)
        const int KB = cin_p < 64 ? cin_p : 64, kb = ci / KB, c = (ci % KB) >> 3, cpr = KB >> 3, nkb = cin_p / KB;
        ((bf16*)l.pack_tc_fprop)[((((int64_t)tp * nkb + kb) * cpr + c) * l.pack_tc_rows + i) * 8 + (ci & 7)] = __float2bfloat16_rn(wv);
      }
      if (l.pack_tc_dgrad) {  // [tap'][kb][chunk][ci][8], k blocks of min(64, rows) output channels
        const int KB = rows < 64 ? rows : 64, kb = i / KB, c = (i % KB) >> 3, cpr = KB >> 3, nkb = rows / KB;
        ((bf16*)l.pack_tc_dgrad)[((((int64_t)(taps - 1 - tp) * nkb + kb) * cpr + c) * l.pack_tc_cin + ci) * 8 + (i & 7)] = __float2bfloat16_rn(wv);
      }
    }
    if (l.spectral) {
      acc = warp_sum(acc);
      if (lane == 0) t[i] = acc;
    }
  }
}

// phase D: u' = normalize(t), sigma = t . u'
__global__ void __launch_bounds__(256) sn_phase_d(const iea_sn_layer* L, float* scratch) {
  __shared__ float red[33];
  const iea_sn_layer& l = L[blockIdx.x];
  float inv_sigma = 1.f;
  if (l.spectral) {
    const int cols = l.cin * l.taps;
    const float* t = scratch + l.scratch_off + (int64_t)l.nchunks * cols;
    float ss = 0.f;
    for (int i = threadIdx.x; i < l.rows; i += blockDim.x) ss = fmaf(t[i], t[i], ss);
    float n2 = block_sum(ss, red);
    float inv = 1.f / fmaxf(sqrtf(n2), l.eps);
    for (int i = threadIdx.x; i < l.rows; i += blockDim.x) l.u_out[i] = t[i] * inv;
    float sigma = n2 * inv;
    inv_sigma = 1.f / sigma;
    if (threadIdx.x == 0 && l.sigma_out) l.sigma_out[0] = sigma;
  }
  if (threadIdx.x == 0) l.inv_sigma_out[0] = inv_sigma;
  if (l.colscale_out)
    for (int i = threadIdx.x; i < l.colscale_n; i += blockDim.x) l.colscale_out[i] = inv_sigma;
}

extern "C" int iea_sn_power_iter(const iea_sn_layer* layers_dev, int n_layers, const int32_t* chunks_dev,
                                 int n_chunks, float* scratch, int max_cols, iea_stream_t stream) {
  IEA_CHECK_ARG(n_layers > 0 && n_chunks > 0, "iea_sn_power_iter: empty layer list");
  cudaStream_t s = (cudaStream_t)stream;
  (void)max_cols;
  sn_phase_a<<<n_chunks, 256, 0, s>>>(layers_dev, chunks_dev, scratch);
  sn_phase_b<<<n_layers, 256, 0, s>>>(layers_dev, scratch);
  sn_phase_c<<<n_chunks, 256, 0, s>>>(layers_dev, chunks_dev, scratch);
  sn_phase_d<<<n_layers, 256, 0, s>>>(layers_dev, scratch);
  return check_launch("iea_sn_power_iter");
}

// ---------------- backward ----------------
__device__ __forceinline__ float sum_splits(const float* gpart, int nsplit, int64_t total, int64_t idx) {
  float g = 0.f;
  for (int s = 0; s < nsplit; ++s) g += gpart[(int64_t)s * total + idx];
  return g;
}

// partial <G, W> per block; idx runs over the packed layout [rows][taps][cin]
__global__ void __launch_bounds__(256) sn_bwd_dot(const float* gpart, int nsplit, const float* w, int rows,
                                                  int cin, int taps, float* part) {
  __shared__ float red[33];
  const int64_t total = (int64_t)rows * cin * taps;
  float acc = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int ci = idx % cin;
    int64_t r = idx / cin;
    int tp = r % taps;
    int64_t i = r / taps;
    acc = fmaf(sum_splits(gpart, nsplit, total, idx), w[(i * cin + ci) * taps + tp], acc);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256) sn_bwd_apply(const float* gpart, int nsplit, const float* u, const float* v,
                                                    const float* inv_sigma, int spectral, float* dw, float beta,
                                                    int rows, int cin, int taps, const float* part, int nparts) {
  __shared__ float s_dot;
  if (threadIdx.x == 0) {
    float d = 0.f;
    if (spectral) for (int i = 0; i < nparts; ++i) d += part[i];
    s_dot = d;
  }
  __syncthreads();
  const float inv = spectral ? inv_sigma[0] : 1.f;
  const float coef = spectral ? s_dot * inv * inv : 0.f;
  const int64_t total = (int64_t)rows * cin * taps;
  // idx runs over the MASTER layout [rows][cin][taps] so the store is coalesced
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int tp = idx % taps;
    int64_t r = idx / taps;
    int ci = r % cin;
    int64_t i = r / cin;
    float g = sum_splits(gpart, nsplit, total, (i * taps + tp) * cin + ci);
    float val = g * inv;
    if (spectral) val -= coef * u[i] * v[ci * taps + tp];
    dw[idx] = beta != 0.f ? fmaf(beta, dw[idx], val) : val;
  }
}

// ---- grouped form: every layer of a backward pass in two launches (the per-layer calls are ~18 us of launch
// latency each, 278 of them per train step).  Block b serves item i with block0[i] <= b < block0[i] + nblocks[i].
__device__ __forceinline__ int find_item(const iea_sn_bwd_item* items, int n, int b) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[mid].block0 <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__global__ void __launch_bounds__(256) sn_bwd_dot_grouped(const iea_sn_bwd_item* items, int n_items, float* part) {
  __shared__ float red[33];
  const iea_sn_bwd_item it = items[find_item(items, n_items, blockIdx.x)];
  float acc = 0.f;
  if (it.spectral) {
    const int64_t total = (int64_t)it.rows * it.cin * it.taps;
    for (int64_t idx = (int64_t)(blockIdx.x - it.block0) * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)it.nblocks * blockDim.x) {
      const int ci = idx % it.cin;
      const int64_t r = idx / it.cin;
      const int tp = r % it.taps;
      const int64_t i = r / it.taps;
      acc = fmaf(sum_splits(it.gpart, it.nsplit, total, idx), it.w[(i * it.cin + ci) * it.taps + tp], acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) sn_bwd_apply_grouped(const iea_sn_bwd_item* items, int n_items, const float* part) {
  __shared__ float s_dot;
  const iea_sn_bwd_item it = items[find_item(items, n_items, blockIdx.x)];
  if (threadIdx.x == 0) {
    float d = 0.f;
    if (it.spectral) for (int i = 0; i < it.nblocks; ++i) d += part[it.block0 + i];
    s_dot = d;
  }
  __syncthreads();
  const float inv = it.spectral ? it.inv_sigma[0] : 1.f;
  const float coef = it.spectral ? s_dot * inv * inv : 0.f;
  const int64_t total = (int64_t)it.rows * it.cin * it.taps;
  for (int64_t idx = (int64_t)(blockIdx.x - it.block0) * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)it.nblocks * blockDim.x) {
    const int tp = idx % it.taps;
    const int64_t r = idx / it.taps;
    const int ci = r % it.cin;
    const int64_t i = r / it.cin;
    const float g = sum_splits(it.gpart, it.nsplit, total, (i * it.taps + tp) * it.cin + ci);
    float val = g * inv;
    if (it.spectral) val -= coef * it.u[i] * it.v[ci * it.taps + tp];
    it.dw[idx] = it.beta != 0.f ? fmaf(it.beta, it.dw[idx], val) : val;
  }
}

extern "C" int iea_sn_weight_bwd_grouped(const iea_sn_bwd_item* items_dev, int n_items, int total_blocks, float* scratch,
                                         iea_stream_t stream) {
  IEA_CHECK_ARG(items_dev && n_items > 0 && total_blocks > 0 && scratch, "iea_sn_weight_bwd_grouped: empty item table");
  cudaStream_t s = (cudaStream_t)stream;
  sn_bwd_dot_grouped<<<total_blocks, 256, 0, s>>>(items_dev, n_items, scratch);
  sn_bwd_apply_grouped<<<total_blocks, 256, 0, s>>>(items_dev, n_items, scratch);
  return check_launch("iea_sn_weight_bwd_grouped");
}

extern "C" int iea_sn_weight_bwd(const float* gpart, int nsplit, const float* w, const float* u, const float* v,
                                 const float* inv_sigma, int spectral, float* dw, float beta, int rows, int cin,
                                 int taps, float* scratch, iea_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t total = (int64_t)rows * cin * taps;
  int blocks = (int)((total + 1023) / 1024);
  if (blocks > 256) blocks = 256;
  if (blocks < 1) blocks = 1;
  if (spectral) sn_bwd_dot<<<blocks, 256, 0, s>>>(gpart, nsplit, w, rows, cin, taps, scratch);
  sn_bwd_apply<<<blocks, 256, 0, s>>>(gpart, nsplit, u, v, inv_sigma, spectral, dw, beta, rows, cin, taps,
                                      scratch, blocks);
  return check_launch("iea_sn_weight_bwd");
}
