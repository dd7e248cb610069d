// K6 / K8 / T1 build-side kernels: k-means centroid update, residuals, column slicing, row
// gather, normalisation and the inverted-list append.
//
// Replaces, in the reference: the centroid update of faiss::Clustering (driven from
// gamma_index_ivfflat.cc:407 and gamma_index_ivfpq.cc:372), compute_residuals
// (gamma_index_ivfpq.cc:378-391), the training-slab merge (gamma_index_ivfflat.cc:377-405) and
// RTInvertIndex::AddKeys / RealTimeMemData::AddKeys (index/realtime/realtime_mem_data.cc:258-296).
#include "common.cuh"
#include "kernels.h"

namespace gb {

namespace {

// one CTA per cluster; members are summed in the given (point) order so the result equals a
// sequential CPU accumulation bit for bit, then scaled by 1/count (faiss: c *= 1/hassign).
__global__ void __launch_bounds__(128)
    segment_mean_kernel(const float* __restrict__ x, int64_t ldx, int d, const int32_t* __restrict__ perm,
                        const int32_t* __restrict__ off, float* __restrict__ centroids, int64_t ldc) {
  const int c = blockIdx.x;
  const int b = off[c], e = off[c + 1];
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float acc = 0.f;
    int i = b;
    for (; i + 4 <= e; i += 4) {
      float v0 = x[(int64_t)perm[i] * ldx + j], v1 = x[(int64_t)perm[i + 1] * ldx + j];
      float v2 = x[(int64_t)perm[i + 2] * ldx + j], v3 = x[(int64_t)perm[i + 3] * ldx + j];
      acc += v0;
      acc += v1;
      acc += v2;
      acc += v3;
    }
    for (; i < e; i++) acc += x[(int64_t)perm[i] * ldx + j];
    float out = 0.f;
    if (e > b) out = acc * (1.0f / (float)(e - b));
    centroids[(int64_t)c * ldc + j] = out;
  }
}

__global__ void residual_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d,
                                const float* __restrict__ centroids, int64_t ldc, const int32_t* __restrict__ assign,
                                float* __restrict__ out, int64_t ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n * ldo;
  if (i >= total) return;
  int64_t r = i / ldo;
  int c = (int)(i - r * ldo);
  out[i] = c < d ? x[r * ldx + c] - centroids[(int64_t)assign[r] * ldc + c] : 0.f;
}

__global__ void slice_cols_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int col0, int w,
                                  float* __restrict__ out, int64_t ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n * ldo;
  if (i >= total) return;
  int64_t r = i / ldo;
  int c = (int)(i - r * ldo);
  out[i] = c < w ? x[r * ldx + col0 + c] : 0.f;
}

__global__ void gather_rows_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ idx, int64_t n,
                                   int d, float* __restrict__ out, int64_t ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n * ldo;
  if (i >= total) return;
  int64_t r = i / ldo;
  int c = (int)(i - r * ldo);
  out[i] = c < d ? x[(int64_t)idx[r] * ldx + c] : 0.f;
}

// one warp per row; fixed reduction order => deterministic
__global__ void normalize_rows_kernel(float* __restrict__ x, int64_t ldx, int64_t n, int d) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n) return;
  float* row = x + r * ldx;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) s = fmaf(row[j], row[j], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (s > 0.f) {
    float inv = 1.0f / sqrtf(s);
    for (int j = lane; j < d; j += 32) row[j] *= inv;
  }
}

// one warp per appended row (float4 copies)
__global__ void ivf_append_vecs_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d,
                                       const int32_t* __restrict__ list, const int32_t* __restrict__ pos,
                                       float* const* __restrict__ list_vecs, int64_t* const* __restrict__ list_ids,
                                       int64_t vid0) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n) return;
  int l = list[r];
  if (l < 0) return;  // skipped (deleted before indexing, gamma_index_ivfflat.cc:436)
  int p = pos[r];
  const float4* src = reinterpret_cast<const float4*>(x + r * ldx);
  float4* dst = reinterpret_cast<float4*>(list_vecs[l] + (int64_t)p * d);
  for (int c = lane; c < (d >> 2); c += 32) dst[c] = src[c];
  if (lane == 0) list_ids[l][p] = vid0 + r;
}

__global__ void ivf_append_codes_kernel(const uint8_t* __restrict__ codes, int64_t n, int M,
                                        const int32_t* __restrict__ list, const int32_t* __restrict__ pos,
                                        uint8_t* const* __restrict__ list_codes, int64_t* const* __restrict__ list_ids,
                                        int64_t vid0) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int l = list[r];
  if (l < 0) return;
  int p = pos[r];
  const uint8_t* src = codes + r * M;
  uint8_t* dst = list_codes[l] + (int64_t)p * M;
  if ((M & 3) == 0) {
    for (int j = 0; j < M; j += 4) *reinterpret_cast<uint32_t*>(dst + j) = *reinterpret_cast<const uint32_t*>(src + j);
  } else {
    for (int j = 0; j < M; j++) dst[j] = src[j];
  }
  list_ids[l][p] = vid0 + r;
}

inline unsigned blocks_for(int64_t total, int bs) { return (unsigned)((total + bs - 1) / bs); }

// column means, one CTA per column (rows strided over the threads, tree reduction in shared memory)
__global__ void __launch_bounds__(256)
    col_mean_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, float* __restrict__ mean) {
  __shared__ double acc[256];
  const int j = blockIdx.x;
  double a = 0;
  for (int64_t i = threadIdx.x; i < n; i += 256) a += x[i * ldx + j];
  acc[threadIdx.x] = a;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) acc[threadIdx.x] += acc[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[j] = (float)(acc[0] / (double)n);
}

__global__ void center_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d, const float* __restrict__ mean,
                                   float* __restrict__ out, int64_t ldo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * ldo) return;
  const int64_t i = idx / ldo;
  const int j = (int)(idx - i * ldo);
  out[idx] = j < d ? x[i * ldx + j] - mean[j] : 0.f;
}

__global__ void transpose_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d, float* __restrict__ out,
                                 int64_t ldo) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int j0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int j = j0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && j < d) ? x[i * ldx + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r;
    const int64_t i = i0 + threadIdx.x;
    if (j < d && i < n) out[(int64_t)j * ldo + i] = tile[threadIdx.x][r];
  }
}

__global__ void pq_decode_kernel(const uint8_t* __restrict__ codes, int64_t n, const float* __restrict__ pq, int M, int dsub,
                                 float* __restrict__ recon, int64_t ldr) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = M * dsub;
  if (idx >= n * d) return;
  const int64_t i = idx / d;
  const int c = (int)(idx - i * d);
  const int m = c / dsub, j = c - m * dsub;
  recon[i * ldr + c] = pq[((int64_t)m * 256 + codes[i * M + m]) * dsub + j];
}

}  // namespace

cudaError_t launch_center_rows(const float* x, int64_t ldx, int64_t n, int d, float* out, int64_t ldo, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  float* mean = nullptr;
  cudaError_t e = cudaMallocAsync(&mean, sizeof(float) * d, st);
  if (e != cudaSuccess) return e;
  col_mean_kernel<<<d, 256, 0, st>>>(x, ldx, n, mean);
  note_launch();
  center_rows_kernel<<<blocks_for(n * ldo, 256), 256, 0, st>>>(x, ldx, n, d, mean, out, ldo);
  note_launch();
  cudaFreeAsync(mean, st);
  return cudaGetLastError();
}

cudaError_t launch_transpose(const float* x, int64_t ldx, int64_t n, int d, float* out, int64_t ldo, cudaStream_t st) {
  if (n <= 0 || d <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((d + 31) / 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, st>>>(x, ldx, n, d, out, ldo);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pq_decode(const uint8_t* codes, int64_t n, const float* pq_centroids, int M, int dsub, float* recon,
                             int64_t ldr, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  pq_decode_kernel<<<blocks_for(n * M * dsub, 256), 256, 0, st>>>(codes, n, pq_centroids, M, dsub, recon, ldr);
  note_launch();
  return cudaGetLastError();
}


cudaError_t launch_segment_mean(const float* x, int64_t ldx, int d, const int32_t* perm, const int32_t* off, int k,
                                float* centroids, int64_t ldc, cudaStream_t st) {
  if (k <= 0) return cudaSuccess;
  segment_mean_kernel<<<k, 128, 0, st>>>(x, ldx, d, perm, off, centroids, ldc);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_residual(const float* x, int64_t ldx, int64_t n, int d, const float* centroids, int64_t ldc,
                            const int32_t* assign, float* out, int64_t ldo, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  residual_kernel<<<blocks_for(n * ldo, 256), 256, 0, st>>>(x, ldx, n, d, centroids, ldc, assign, out, ldo);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_slice_cols(const float* x, int64_t ldx, int64_t n, int col0, int w, float* out, int64_t ldo,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  slice_cols_kernel<<<blocks_for(n * ldo, 256), 256, 0, st>>>(x, ldx, n, col0, w, out, ldo);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const float* x, int64_t ldx, const int32_t* idx, int64_t n, int d, float* out,
                               int64_t ldo, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  gather_rows_kernel<<<blocks_for(n * ldo, 256), 256, 0, st>>>(x, ldx, idx, n, d, out, ldo);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_normalize_rows(float* x, int64_t ldx, int64_t n, int d, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  normalize_rows_kernel<<<blocks_for(n * 32, 256), 256, 0, st>>>(x, ldx, n, d);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_ivf_append_vecs(const float* x, int64_t ldx, int64_t n, int d, const int32_t* list, const int32_t* pos,
                                   float* const* list_vecs, int64_t* const* list_ids, int64_t vid0, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ivf_append_vecs_kernel<<<blocks_for(n * 32, 256), 256, 0, st>>>(x, ldx, n, d, list, pos, list_vecs, list_ids, vid0);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_ivf_append_codes(const uint8_t* codes, int64_t n, int M, const int32_t* list, const int32_t* pos,
                                    uint8_t* const* list_codes, int64_t* const* list_ids, int64_t vid0,
                                    cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ivf_append_codes_kernel<<<blocks_for(n, 256), 256, 0, st>>>(codes, n, M, list, pos, list_codes, list_ids, vid0);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gb
