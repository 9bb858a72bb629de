// ortho.cu -- modified orthogonal regularisation added straight to the gradients:
//     grad += strength * 2 * ((W W^T) o (1 - I)) W            utils/__init__.py:843-859, train_fns.py:185-188
// for every >= 2-D parameter of a net, as THREE grouped launches over tile tables (row norms of the tall
// matrices, Gram tiles, product tiles) instead of the reference's 3 matmuls + eye + mul + add per parameter.
// Tall matrices (rows > cols, e.g. G.linear 8192 x 256, 24576 x 256 at H_base 3) use the algebraically
// equal small form  W (W^T W) - diag(|w_i|^2) W, so the rows x rows Gram matrix (2.4 GB at H_base 3) is never
// formed.  fp32 FFMA tiles (64 x 64 x 16, 4 x 4 per thread): ~6 GFLOP per Generator step, nowhere near a roofline
// that matters next to the 1.7 TFLOP step -- the point is launch count and the missing temporaries.
#include "common.cuh"
using namespace iea;

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// C tile (64 x 64) of A (M x K) . B (K x N) with arbitrary element strides; returns the thread's 4 x 4 block
__device__ __forceinline__ void tile_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
                                          int64_t sbn, int M, int N, int K, int m0, int n0, float acc[4][4]) {
  __shared__ float As[TK][TM + 4], Bs[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      // pick the fast index along the contiguous dimension of each operand
      int m, k;
      if (sak == 1) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
      float v = 0.f;
      if (m0 + m < M && k0 + k < K) v = A[(int64_t)(m0 + m) * sam + (int64_t)(k0 + k) * sak];
      As[k][m] = v;
      int n, kb;
      if (sbk == 1) { kb = e % TK; n = e / TK; } else { n = e % TN; kb = e / TN; }
      float u = 0.f;
      if (n0 + n < N && k0 + kb < K) u = B[(int64_t)(k0 + kb) * sbk + (int64_t)(n0 + n) * sbn];
      Bs[kb][n] = u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// row norms |w_r|^2 of the tall items: one warp per row
__global__ void __launch_bounds__(256) ortho_rownorm_kernel(const iea_ortho_item* items, const int2* rows_tab, int n_rows) {
  const int gw = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_rows) return;
  const int2 rt = rows_tab[gw];
  const iea_ortho_item it = items[rt.x];
  const float* w = it.w + (int64_t)rt.y * it.cols;
  float s = 0.f;
  for (int c = lane; c < it.cols; c += 32) s = fmaf(w[c], w[c], s);
  s = warp_sum(s);
  if (lane == 0) it.rownorm[rt.y] = s;
}

// tiles: (item, tile row, tile col)
__global__ void __launch_bounds__(256) ortho_gram_kernel(const iea_ortho_item* items, const int4* tiles) {
  const int4 t = tiles[blockIdx.x];
  const iea_ortho_item it = items[t.x];
  const int dim = it.tall ? it.cols : it.rows, kk = it.tall ? it.rows : it.cols;
  const int m0 = t.y * TM, n0 = t.z * TN;
  float acc[4][4];
  float* dst = it.gram;
  if (it.tall) {  // G = W^T W: A[m][k] = W[k][m], B[k][n] = W[k][n]; the long K (= rows) is split over t.w CTAs
    const int per = (kk + it.ksplits - 1) / it.ksplits, k0 = t.w * per;
    const int kn = kk - k0 < per ? kk - k0 : per;
    tile_gemm(it.w + (int64_t)k0 * it.cols, 1, it.cols, it.w + (int64_t)k0 * it.cols, it.cols, 1, dim, dim, kn > 0 ? kn : 0, m0, n0, acc);
    dst = it.gram_part + (int64_t)t.w * dim * dim;
  } else {        // G = W W^T with the diagonal removed: A[m][k] = W[m][k], B[k][n] = W[n][k]
    tile_gemm(it.w, it.cols, 1, it.w, 1, it.cols, dim, dim, kk, m0, n0, acc);
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < dim && n < dim) dst[(int64_t)m * dim + n] = (!it.tall && m == n) ? 0.f : acc[i][j];
    }
}

// fixed-order sum of the K-split partial Gram matrices of the tall items (deterministic)
__global__ void __launch_bounds__(256) ortho_gram_reduce_kernel(const iea_ortho_item* items, const int2* red_tab) {
  const int2 rt = red_tab[blockIdx.x];  // (item, first element of this block's 256-element span)
  const iea_ortho_item it = items[rt.x];
  const int64_t n = (int64_t)it.cols * it.cols, i = rt.y + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int s = 0; s < it.ksplits; ++s) a += it.gram_part[(int64_t)s * n + i];
  it.gram[i] = a;
}

__global__ void __launch_bounds__(256) ortho_apply_kernel(const iea_ortho_item* items, const int4* tiles) {
  const int4 t = tiles[blockIdx.x];
  const iea_ortho_item it = items[t.x];
  const int m0 = t.y * TM, n0 = t.z * TN;
  float acc[4][4];
  if (it.tall)  // W (W^T W): A = W (rows x cols), B = gram (cols x cols)
    tile_gemm(it.w, it.cols, 1, it.gram, it.cols, 1, it.rows, it.cols, it.cols, m0, n0, acc);
  else          // ((W W^T) o (1-I)) W: A = gram (rows x rows), B = W (rows x cols)
    tile_gemm(it.gram, it.rows, 1, it.w, it.cols, 1, it.rows, it.cols, it.rows, m0, n0, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float a2 = 2.f * it.strength;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < it.rows && n < it.cols) {
        const int64_t o = (int64_t)m * it.cols + n;
        float v = acc[i][j];
        if (it.tall) v -= it.rownorm[m] * it.w[o];
        it.grad[o] += a2 * v;
      }
    }
}

}  // namespace

extern "C" int iea_ortho_grouped(const iea_ortho_item* items, const int32_t* rownorm_rows, int n_rownorm_rows,
                                 const int32_t* gram_tiles, int n_gram_tiles, const int32_t* reduce_blocks,
                                 int n_reduce_blocks, const int32_t* apply_tiles, int n_apply_tiles,
                                 iea_stream_t stream) {
  IEA_CHECK_ARG(n_gram_tiles > 0 && n_apply_tiles > 0, "iea_ortho_grouped: empty tile tables");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rownorm_rows > 0)
    ortho_rownorm_kernel<<<cdiv((int64_t)n_rownorm_rows * 32, 256), 256, 0, st>>>(
        items, reinterpret_cast<const int2*>(rownorm_rows), n_rownorm_rows);
  ortho_gram_kernel<<<n_gram_tiles, 256, 0, st>>>(items, reinterpret_cast<const int4*>(gram_tiles));
  if (n_reduce_blocks > 0)
    ortho_gram_reduce_kernel<<<n_reduce_blocks, 256, 0, st>>>(items, reinterpret_cast<const int2*>(reduce_blocks));
  ortho_apply_kernel<<<n_apply_tiles, 256, 0, st>>>(items, reinterpret_cast<const int4*>(apply_tiles));
  return check_launch("iea_ortho_grouped");
}
