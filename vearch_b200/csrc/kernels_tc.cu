// K2 / K6 tensor-core variant: the ONE genuinely dense contraction of the path -- query x centroid
// (coarse quantiser, gamma_index_ivfflat.cc:568 / gamma_index_ivfpq.cc:595) and point x centroid
// (k-means assign, faiss Clustering via gamma_index_ivfflat.cc:407 / gamma_index_ivfpq.cc:372) --
// on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM, read back with
// tcgen05.ld for a fused epilogue (|x|^2 + |c|^2 - 2 x.c, then either the score tile or the row
// arg-min).  faiss itself evaluates this contraction with sgemm + norms (IndexFlat, >= 20 queries).
//
// Precision: every operand is split x = hi + lo with hi = the TF32-representable head (low 13
// mantissa bits cleared) and three MMAs accumulate hi*hi + hi*lo + lo*hi in fp32 (error-compensated
// "3xTF32", relative error ~2^-21).  Integer-valued operands up to 2^11 have lo == 0 and every
// product and partial sum is exact, so on the SIFT-shaped parity data results are bit-equal to the
// exact fp32 kernel (kernels_dist.cu), which remains the path for list assignment at add time.
//
// Operand staging: plain coalesced-enough global loads -> st.shared in the canonical no-swizzle
// K-major UMMA layout (8-row x 16-byte core matrices; LBO = stride between the two K core
// matrices of one MMA, SBO = stride between 8-row groups), fence.proxy.async, one elected thread
// issues the MMAs and commits to an mbarrier.  One 128 x 128 output tile per CTA.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace gb {

namespace {

constexpr int TC_M = 128, TC_N = 128, TC_BK = 32, TC_NT = 128;
constexpr int TC_TILE_BYTES = TC_M * TC_BK * 4;  // 16 KiB per operand tile (hi or lo)

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
  // base_offset 0 | lbo_mode 0 | layout_type SWIZZLE_NONE (0) [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int METRIC, int EPI_ARGMIN>
__global__ void __launch_bounds__(TC_NT)
    dist_tc_kernel(const float* __restrict__ X, int64_t ldx, int n, const float* __restrict__ C, int64_t ldc, int m,
                   int d, const float* __restrict__ xnorm, const float* __restrict__ cnorm, float* __restrict__ out,
                   int64_t ldo, unsigned long long* __restrict__ best) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + TC_TILE_BYTES;
  unsigned char* b_hi = smem + 2 * TC_TILE_BYTES;
  unsigned char* b_lo = smem + 3 * TC_TILE_BYTES;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row0 = blockIdx.y * TC_M, col0 = blockIdx.x * TC_N;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(TC_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10),
  // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  // canonical layout: element (r, k) of a 128 x 32 tile lives at (k/4)*(128*16) + (r/8)*128 + (r%8)*16 + (k%4)*4
  const uint32_t LBO = TC_M * 16, SBO = 128;

  uint32_t phase = 0;
  const int nk = (d + TC_BK - 1) / TC_BK;
  for (int kc = 0; kc < nk; kc++) {
    const int k0 = kc * TC_BK;
    // ---- stage both operand tiles (hi and lo parts) ----
#pragma unroll
    for (int it = 0; it < (TC_M * TC_BK / 4) / TC_NT; it++) {  // 1024 float4 per operand / 128 threads = 8
      const int f = it * TC_NT + tid;
      const int r = f & (TC_M - 1), kb = f >> 7;  // consecutive threads -> consecutive rows: conflict-free st.shared
      const int gk = k0 + kb * 4;
      const uint32_t off = (uint32_t)kb * LBO + (uint32_t)(r >> 3) * SBO + (uint32_t)(r & 7) * 16;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (row0 + r < n && gk < d) va = __ldg(reinterpret_cast<const float4*>(X + (int64_t)(row0 + r) * ldx + gk));
      if (col0 + r < m && gk < d) vb = __ldg(reinterpret_cast<const float4*>(C + (int64_t)(col0 + r) * ldc + gk));
      auto split = [](float v, float& hi, float& lo) {
        hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        lo = v - hi;
      };
      float4 ah, al, bh, bl;
      split(va.x, ah.x, al.x), split(va.y, ah.y, al.y), split(va.z, ah.z, al.z), split(va.w, ah.w, al.w);
      split(vb.x, bh.x, bl.x), split(vb.y, bh.y, bl.y), split(vb.z, bh.z, bl.z), split(vb.w, bh.w, bl.w);
      *reinterpret_cast<float4*>(a_hi + off) = ah;
      *reinterpret_cast<float4*>(a_lo + off) = al;
      *reinterpret_cast<float4*>(b_hi + off) = bh;
      *reinterpret_cast<float4*>(b_lo + off) = bl;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async (tensor) proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ks++) {  // one MMA consumes K = 8 (two 4-wide core matrices)
        const uint32_t koff = (uint32_t)ks * 2 * LBO;
        const uint64_t dah = make_smem_desc(smem_u32(a_hi) + koff, LBO, SBO);
        const uint64_t dal = make_smem_desc(smem_u32(a_lo) + koff, LBO, SBO);
        const uint64_t dbh = make_smem_desc(smem_u32(b_hi) + koff, LBO, SBO);
        const uint64_t dbl = make_smem_desc(smem_u32(b_lo) + koff, LBO, SBO);
        tc_mma_tf32(tmem_d, dah, dbh, idesc, (kc | ks) != 0);
        tc_mma_tf32(tmem_d, dah, dbl, idesc, 1);
        tc_mma_tf32(tmem_d, dal, dbh, idesc, 1);
      }
      // arrives on mma_bar when every MMA issued so far has finished reading smem / writing TMEM
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mma_bar))
                   : "memory");
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();  // smem tiles may be overwritten by the next chunk
  }

  // ---- epilogue: thread t owns output row row0 + t (TMEM lane t) ----
  const int gr = row0 + tid;
  const float xn = (METRIC == kMetricL2 && gr < n) ? xnorm[gr] : 0.f;
  unsigned long long kbest = kKeySentinel;
#pragma unroll 1
  for (int c0 = 0; c0 < TC_N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int gc = col0 + c0 + j;
      float dot = __uint_as_float(v[j]);
      float s = dot;
      if (METRIC == kMetricL2) {
        // faiss IndexFlat clamps the expanded form at 0
        s = fmaxf(0.f, fmaf(-2.0f, dot, xn + __ldg(cnorm + (gc < m ? gc : 0))));
      }
      if (gr < n && gc < m) {
        if (EPI_ARGMIN) {
          unsigned long long key = make_key(score2ord<METRIC>(s), (uint32_t)gc);
          kbest = key < kbest ? key : kbest;
        } else {
          out[(int64_t)gr * ldo + gc] = s;
        }
      }
    }
  }
  if (EPI_ARGMIN && gr < n && kbest != kKeySentinel) atomicMin(best + gr, kbest);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_N) : "memory");
  }
}

// |x|^2 per row, fixed sequential order per lane then a fixed shuffle tree (deterministic)
__global__ void row_norms_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d, float* __restrict__ out) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n) return;
  const float* row = x + r * ldx;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) s = fmaf(row[j], row[j], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[r] = s;
}

template <int METRIC, int EPI>
cudaError_t launch_tc(const float* X, int64_t ldx, int n, const float* C, int64_t ldc, int m, int d, const float* xnorm,
                      const float* cnorm, float* out, int64_t ldo, unsigned long long* best, cudaStream_t st) {
  const size_t smem = 4 * TC_TILE_BYTES;
  cudaError_t e = cudaFuncSetAttribute(dist_tc_kernel<METRIC, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int max_rows = 65535 * TC_M;
  for (int r0 = 0; r0 < n; r0 += max_rows) {
    int nr = n - r0 < max_rows ? n - r0 : max_rows;
    dim3 grid((m + TC_N - 1) / TC_N, (nr + TC_M - 1) / TC_M);
    dist_tc_kernel<METRIC, EPI><<<grid, TC_NT, smem, st>>>(X + (int64_t)r0 * ldx, ldx, nr, C, ldc, m, d,
                                                          xnorm ? xnorm + r0 : nullptr, cnorm,
                                                          out ? out + (int64_t)r0 * ldo : nullptr, ldo,
                                                          best ? best + r0 : nullptr);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace

cudaError_t launch_row_norms(const float* x, int64_t ldx, int64_t n, int d, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  row_norms_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(x, ldx, n, d, out);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_dist_matrix_tc(const float* X, int64_t ldx, int n, const float* C, int64_t ldc, int m, int d,
                                  int metric, const float* xnorm, const float* cnorm, float* out, int64_t ldo,
                                  cudaStream_t st) {
  if ((d & 3) || (ldx & 3) || (ldc & 3)) return cudaErrorInvalidValue;
  if (n <= 0 || m <= 0) return cudaSuccess;
  return metric == kMetricL2 ? launch_tc<kMetricL2, 0>(X, ldx, n, C, ldc, m, d, xnorm, cnorm, out, ldo, nullptr, st)
                             : launch_tc<kMetricIP, 0>(X, ldx, n, C, ldc, m, d, xnorm, cnorm, out, ldo, nullptr, st);
}

cudaError_t launch_dist_argmin_tc(const float* X, int64_t ldx, int n, const float* C, int64_t ldc, int m, int d,
                                  int metric, const float* xnorm, const float* cnorm, unsigned long long* best,
                                  cudaStream_t st) {
  if ((d & 3) || (ldx & 3) || (ldc & 3)) return cudaErrorInvalidValue;
  if (n <= 0 || m <= 0) return cudaSuccess;
  return metric == kMetricL2 ? launch_tc<kMetricL2, 1>(X, ldx, n, C, ldc, m, d, xnorm, cnorm, nullptr, 0, best, st)
                             : launch_tc<kMetricIP, 1>(X, ldx, n, C, ldc, m, d, xnorm, cnorm, nullptr, 0, best, st);
}

}  // namespace gb
