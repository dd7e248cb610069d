// Tensor-core kernels of the path (tcgen05.mma kind::tf32, accumulators in TMEM, tcgen05.ld epilogue):
//
//  * K2 / K6  dist_tc_kernel: query x centroid (coarse quantiser, gamma_index_ivfflat.cc:568 /
//    gamma_index_ivfpq.cc:595) and point x centroid (k-means assign, faiss Clustering via
//    gamma_index_ivfflat.cc:407 / gamma_index_ivfpq.cc:372).  faiss itself evaluates this
//    contraction as sgemm + norms (IndexFlat, >= 20 queries).
//  * K3 list-major: when many queries of a batch probe the same list the IVF-Flat scan
//    (gamma_index_ivfflat.cc:579-787) IS a dense contraction.  The (query, probe) pairs are grouped by
//    list on the device (lm_count / lmk_scan / lmk_assign / lmk_items); a work item multiplies a group
//    of <= 128 pairs by the 128-row tiles of a row segment of the list, so the list is read once per
//    128 queries instead of once per query, and tombstones / bitmaps / score window / top-k are
//    applied as scan_codes does (gamma_index_ivfflat.h:63-91).  Four kernels, one algorithm:
//      ivf_listmajor_tma_kernel   operands arrive by cp.async.bulk from the pre-tiled mirror (default)
//      ivf_listmajor_pipe_kernel  warp-specialised, operands staged through registers (no mirror)
//      ivf_listmajor_topk_kernel  the same without warp specialisation (single-chunk rows)
//      ivf_listmajor_tc_kernel + seg_select_kernel   dense scores + segment select (k > 64)
//
// Precision: every operand is split x = hi + lo, hi = the TF32-representable head (low 13 mantissa
// bits cleared), and three MMAs accumulate hi*hi + hi*lo + lo*hi in fp32 (error-compensated
// "3xTF32", relative error ~2^-21).  Integer-valued operands below 2^11 have lo == 0 and every
// product and partial sum is exact, so on the SIFT-shaped parity data scores are bit-equal to the
// exact fp32 kernels.  L2 uses |x|^2 + |y|^2 - 2 x.y with both norms accumulated in-kernel from the
// staged operands (faiss clamps the expanded form at 0; so do we).
//
// Operand layout: canonical no-swizzle K-major UMMA tiles (8-row x 16-byte core matrices; LBO = stride
// between the two K core matrices of one MMA, SBO = stride between 8-row groups).  Register-staged
// kernels: global loads -> hi/lo split -> st.shared, fence.proxy.async, one elected thread issues the
// MMAs and commits to an mbarrier; thread t stages row t of both operands, so it also owns |x_t|^2
// and |y_t|^2.  TMA kernel: the same tiles already exist in global memory (tc_mirror_build_kernel,
// lm_stage_queries_kernel) and arrive as 16 KiB bulk copies.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace gb {

namespace {

constexpr int TC_M = 128, TC_N = 128, TC_BK = 16, TC_NT = 128, TC_STAGES = 2;
constexpr int TC_TILE_BYTES = TC_M * TC_BK * 4;          // 8 KiB per operand part (hi or lo) per stage
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;        // a_hi, a_lo, b_hi, b_lo
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES;      // 64 KiB



// One 128 x 128 tile: D[i][j] = <A_i, B_j> over d columns.  arow / brow: row pointers of THIS
// thread's row of A and B (nullptr => all zeros).  On return the accumulator sits in TMEM at
// *tmem_out (lane = A row, column = B row), thread t holds |A_t|^2 in *an and sh->cn[t] = |B_t|^2.
// Caller must have 128 threads, TC_SMEM bytes of dynamic smem at `smem`, and must call tc_release().
//
// Pipeline: two smem stages of K = 16.  While the tensor core works on stage s (asynchronously,
// completion signalled by tcgen05.commit on mma_bar[s]) the threads already hold the NEXT chunk in
// registers (global loads issued one chunk ahead) and write it into the other stage, so neither the
// global-load latency nor the MMA latency sits on the critical path.
struct TcShared {
  alignas(16) float cn[TC_N];  // |y|^2 of the tile's rows (read as float4 by the fused epilogues)
  uint64_t mma_bar[TC_STAGES];
  uint32_t tmem_base;
};

// allocate TMEM (TC_N fp32 columns), initialise the stage barriers; returns the TMEM base address
__device__ __forceinline__ uint32_t tc_begin(TcShared* sh) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "n"(TC_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; s++) mbar_init(&sh->mma_bar[s], 1);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return sh->tmem_base;
}

// One tile on an allocated accumulator.  gc = number of K chunks this CTA has pushed through the
// stage ring so far (barrier phases continue across tiles); returns the updated count.  A caller
// that runs several tiles must make every thread execute tcgen05.fence::before_thread_sync after
// its last tcgen05.ld of the previous tile (the first chunk's __syncthreads orders the rest).
__device__ __forceinline__ uint32_t tc_tile_run(unsigned char* smem, TcShared* sh, uint32_t tmem_d, uint32_t gc0,
                                                const float* __restrict__ arow, const float* __restrict__ brow, int d,
                                                float* an) {
  const int tid = threadIdx.x;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10),
  // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  // canonical layout: element (r, k) of a 128 x 16 tile lives at (k/4)*(128*16) + (r/8)*128 + (r%8)*16 + (k%4)*4
  const uint32_t LBO = TC_M * 16, SBO = 128;
  const uint32_t row_off = (uint32_t)(tid >> 3) * SBO + (uint32_t)(tid & 7) * 16;
  constexpr int NV = TC_BK / 4;  // float4 per operand row per chunk

  float an_acc = 0.f, bn_acc = 0.f;
  const int nk = (d + TC_BK - 1) / TC_BK;
  // two register sets: the loads of chunk kc + 2 are issued as soon as chunk kc has been written to
  // shared memory, so a full chunk period (barrier + MMA issue + the other set's stores) plus the
  // MMA time covers the L2 latency
  float4 pa0[NV], pb0[NV], pa1[NV], pb1[NV];
  auto prefetch = [&](int kc, float4 (&pa)[NV], float4 (&pb)[NV]) {
#pragma unroll
    for (int kb = 0; kb < NV; kb++) {
      const int gk = kc * TC_BK + kb * 4;
      pa[kb] = make_float4(0.f, 0.f, 0.f, 0.f);
      pb[kb] = pa[kb];
      if (arow && gk < d) pa[kb] = __ldg(reinterpret_cast<const float4*>(arow + gk));
      if (brow && gk < d) pb[kb] = __ldg(reinterpret_cast<const float4*>(brow + gk));
    }
  };
  auto chunk = [&](int kc, float4 (&pa)[NV], float4 (&pb)[NV]) {
    const uint32_t g = gc0 + (uint32_t)kc;
    const int s = (int)(g & 1u);
    unsigned char* a_hi = smem + (size_t)s * TC_STAGE_BYTES;
    unsigned char* a_lo = a_hi + TC_TILE_BYTES;
    unsigned char* b_hi = a_hi + 2 * TC_TILE_BYTES;
    unsigned char* b_lo = a_hi + 3 * TC_TILE_BYTES;
    if (g >= (uint32_t)TC_STAGES) {  // the MMAs of chunk g-2 must have finished reading this stage
      mbar_wait(&sh->mma_bar[s], ((g >> 1) - 1u) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
#pragma unroll
    for (int kb = 0; kb < NV; kb++) {
      const uint32_t off = (uint32_t)kb * LBO + row_off;  // consecutive threads -> consecutive rows: conflict-free
      const float4 va = pa[kb], vb = pb[kb];
      an_acc = fmaf(va.x, va.x, an_acc), an_acc = fmaf(va.y, va.y, an_acc);
      an_acc = fmaf(va.z, va.z, an_acc), an_acc = fmaf(va.w, va.w, an_acc);
      bn_acc = fmaf(vb.x, vb.x, bn_acc), bn_acc = fmaf(vb.y, vb.y, bn_acc);
      bn_acc = fmaf(vb.z, vb.z, bn_acc), bn_acc = fmaf(vb.w, vb.w, bn_acc);
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
    if (kc + 2 < nk) prefetch(kc + 2, pa, pb);
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
      // arrives on mma_bar[s] when every MMA issued so far has finished reading smem / writing TMEM
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&sh->mma_bar[s]))
                   : "memory");
    }
  };
  prefetch(0, pa0, pb0);
  if (nk > 1) prefetch(1, pa1, pb1);
  for (int kc = 0; kc < nk; kc += 2) {
    chunk(kc, pa0, pb0);
    if (kc + 1 < nk) chunk(kc + 1, pa1, pb1);
  }
  // the last commit covers every MMA of the tile
  const uint32_t gl = gc0 + (uint32_t)nk - 1u;
  mbar_wait(&sh->mma_bar[gl & 1u], (gl >> 1) & 1u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  sh->cn[tid] = bn_acc;
  __syncthreads();
  *an = an_acc;
  return gc0 + (uint32_t)nk;
}

__device__ __forceinline__ void tc_tile(unsigned char* smem, TcShared* sh, const float* __restrict__ arow,
                                        const float* __restrict__ brow, int d, uint32_t* tmem_out, float* an) {
  const uint32_t tmem_d = tc_begin(sh);
  tc_tile_run(smem, sh, tmem_d, 0u, arow, brow, d, an);
  *tmem_out = tmem_d;
}

// 32 accumulator columns [c0, c0+32) of this thread's row (TMEM lane = warp*32 + lane)
__device__ __forceinline__ void tc_load32(uint32_t tmem_d, int c0, uint32_t (&v)[32]) {
  const uint32_t taddr = tmem_d + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)c0;
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
}

__device__ __forceinline__ void tc_release(uint32_t tmem_d) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_N) : "memory");
  }
}

template <int METRIC>
__device__ __forceinline__ float tc_score(float dot, float xn, float cn) {
  // faiss IndexFlat clamps the expanded L2 form at 0
  return METRIC == kMetricL2 ? fmaxf(0.f, fmaf(-2.0f, dot, xn + cn)) : dot;
}

// ---- K2 / K6: dense score tile or fused row arg-min ----------------------------------------
template <int METRIC, int EPI_ARGMIN>
__global__ void __launch_bounds__(TC_NT)
    dist_tc_kernel(const float* __restrict__ X, int64_t ldx, int n, const float* __restrict__ C, int64_t ldc, int m,
                   int d, float* __restrict__ out, int64_t ldo, unsigned long long* __restrict__ best) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ TcShared sh;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.y * TC_M, col0 = blockIdx.x * TC_N;
  const float* arow = row0 + tid < n ? X + (int64_t)(row0 + tid) * ldx : nullptr;
  const float* brow = col0 + tid < m ? C + (int64_t)(col0 + tid) * ldc : nullptr;
  uint32_t tmem_d;
  float xn;
  tc_tile(smem, &sh, arow, brow, d, &tmem_d, &xn);

  const int gr = row0 + tid;
  unsigned long long kbest = kKeySentinel;
#pragma unroll 1
  for (int c0 = 0; c0 < TC_N; c0 += 32) {
    uint32_t v[32];
    tc_load32(tmem_d, c0, v);
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const int gc = col0 + c0 + j;
      const float s = tc_score<METRIC>(__uint_as_float(v[j]), xn, sh.cn[c0 + j]);
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
  tc_release(tmem_d);
}

// ---- K3 list-major: one (list, 128-pair group, 128-row tile) work item per CTA --------------
template <int METRIC>
__global__ void __launch_bounds__(TC_NT)
    ivf_listmajor_tc_kernel(const float* __restrict__ xq, int64_t ldq, int d, const LmTile* __restrict__ tiles,
                            const int32_t* __restrict__ pair_q, const int64_t* __restrict__ pair_off,
                            ListDirectory dir, float* __restrict__ scores) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ TcShared sh;
  const int tid = threadIdx.x;
  const LmTile t = tiles[blockIdx.x];
  const float* arow = nullptr;
  int64_t obase = 0;
  if (tid < t.npairs) {
    arow = xq + (int64_t)pair_q[t.pair0 + tid] * ldq;
    obase = pair_off[t.pair0 + tid] + t.row0;
  }
  const float* brow = tid < t.nrows ? dir.vecs[t.list] + (int64_t)(t.row0 + tid) * d : nullptr;
  uint32_t tmem_d;
  float xn;
  tc_tile(smem, &sh, arow, brow, d, &tmem_d, &xn);
  // Epilogue.  Thread t owns accumulator row t, but its row lands in a score segment far away from
  // its neighbours' rows, so writing straight from registers would touch 32 different cache lines
  // per store.  Each warp therefore transposes its 32 x 32 block through shared memory (operand
  // tiles are free by now; row stride 33 floats => conflict-free both ways) and writes every row
  // as one contiguous 128-byte run.
  float* stile = reinterpret_cast<float*>(smem) + (size_t)(tid >> 5) * 32 * 33;
  const int lane = tid & 31, wrow0 = (tid >> 5) * 32;
#pragma unroll 1
  for (int c0 = 0; c0 < TC_N; c0 += 32) {
    uint32_t v[32];
    tc_load32(tmem_d, c0, v);
#pragma unroll
    for (int j = 0; j < 32; j++) stile[lane * 33 + j] = tc_score<METRIC>(__uint_as_float(v[j]), xn, sh.cn[c0 + j]);
    __syncwarp();
    const bool colok = c0 + lane < t.nrows;
#pragma unroll 4
    for (int rr = 0; rr < 32; rr++) {
      const long long ob = __shfl_sync(0xffffffffu, (long long)obase, rr);
      if (wrow0 + rr < t.npairs && colok) scores[ob + c0 + lane] = stile[rr * 33 + lane];
    }
    __syncwarp();
  }
  tc_release(tmem_d);
}

// ---- K3 list-major, fused top-k epilogue -----------------------------------------------------
// Thread t of the CTA owns pair t of the item for the whole row segment: accumulator row t, a k-slot
// key set in shared memory (slot-major, thread-minor: conflict-free) and a register bound tau.  The
// hot loop is  score -> one float compare against the bound;  everything else (row validity, score
// window, tombstone, bitmaps, exact 64-bit key compare, set update) sits behind that compare in a
// non-inlined function, as rare as in the query-major kernels (expected k ln(n/k) hits per query).
// tau_g[q] carries the best bound any CTA has published for query q: a key is only dropped when it
// is >= the k-th best of k valid keys of the same query, so the union of the per-pair sets always
// contains the query's true top-k, whatever the CTA schedule.
struct LmkState {
  unsigned long long tau;  // admission bound (exclusive)
  int n;                   // keys held, <= k
};
constexpr uint32_t kLmkNoVid = 0xFFFFFFFFu;

template <int METRIC>
__device__ __noinline__ LmkState lmk_consider(unsigned long long* hk, int k, uint32_t vid, float sc, float min_score,
                                              float max_score, LmkState st) {
  if (vid == kLmkNoVid) return st;  // padding row, tombstone, deleted or filtered out
  if (!(sc <= max_score && sc >= min_score)) return st;
  const unsigned long long key = make_key(score2ord<METRIC>(sc), vid);
  if (key >= st.tau) return st;
  if (st.n < k) {
    hk[st.n * TC_NT] = key;
    if (++st.n < k) return st;
    unsigned long long mx = 0;
    for (int i = 0; i < k; i++) {
      const unsigned long long v = hk[i * TC_NT];
      mx = v > mx ? v : mx;
    }
    if (mx < st.tau) st.tau = mx;
    return st;
  }
  unsigned long long mx = 0, mx2 = 0;
  int pos = 0;
  for (int i = 0; i < k; i++) {
    const unsigned long long v = hk[i * TC_NT];
    if (v > mx) {
      mx2 = mx, mx = v, pos = i;
    } else if (v > mx2) {
      mx2 = v;
    }
  }
  unsigned long long nm = mx;
  if (key < mx) {
    hk[pos * TC_NT] = key;
    nm = key > mx2 ? key : mx2;
  }
  if (nm < st.tau) st.tau = nm;
  return st;
}

// 32 accumulator columns of one pair (thread): turn them into scores in place and collect, branch-free, one
// bit per column that beats the bound the block started with; only then visit the hits (the rare path
// re-checks every hit against the live bound).  cn32 / vid32: the columns' |y|^2 (16-byte aligned) and row ids.
template <int METRIC>
__device__ __forceinline__ void lmk_scan_block(uint32_t (&v)[32], float xn, const float* cn32, const uint32_t* vid32,
                                               unsigned long long* hk, int k, float min_score, float max_score,
                                               LmkState& st, float& bound) {
  uint32_t hit = 0;
#pragma unroll
  for (int j4 = 0; j4 < 32; j4 += 4) {
    const float4 c4 = *reinterpret_cast<const float4*>(cn32 + j4);
    const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const float sc = tc_score<METRIC>(__uint_as_float(v[j4 + u]), xn, cc[u]);
      v[j4 + u] = __float_as_uint(sc);
      hit |= (METRIC == kMetricL2 ? sc <= bound : sc >= bound) ? (1u << (j4 + u)) : 0u;
    }
  }
  if (hit) {
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
      if (hit & (1u << jj)) st = lmk_consider<METRIC>(hk, k, vid32[jj], __uint_as_float(v[jj]), min_score, max_score, st);
    }
    bound = key_bound<METRIC>(st.tau);
  }
}

template <int METRIC>
__global__ void __launch_bounds__(TC_NT)
    ivf_listmajor_topk_kernel(const float* __restrict__ xq, int64_t ldq, int d, const LmTile* __restrict__ items,
                              const int64_t* __restrict__ totals, const int64_t* __restrict__ pair_j, int nprobe,
                              ListDirectory dir, int k, int nseg_max, FilterArgs f, unsigned long long* tau_g,
                              unsigned long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ TcShared sh;
  if ((int64_t)blockIdx.x >= totals[1]) return;
  const int tid = threadIdx.x;
  const LmTile t = items[blockIdx.x];
  unsigned long long* hk = reinterpret_cast<unsigned long long*>(smem + TC_SMEM) + tid;
  // the operand stages are idle between a tile's last MMA and the next tile's first store: the
  // epilogue keeps the tile's row ids there (vid, or kLmkNoVid for rows that can never be returned)
  uint32_t* s_vid = reinterpret_cast<uint32_t*>(smem);
  const bool valid = tid < t.npairs;
  int64_t j = 0;
  int q = 0;
  const float* arow = nullptr;
  LmkState st{0ull, 0};
  if (valid) {
    j = pair_j[t.pair0 + tid];
    q = (int)(j / nprobe);
    arow = xq + (int64_t)q * ldq;
    st.tau = __ldcg(tau_g + q);
  }
  const uint32_t tmem_d = tc_begin(&sh);
  uint32_t gc = 0;
  const float* lvecs = dir.vecs[t.list];
  const int64_t* __restrict__ lids = dir.ids[t.list];
  const int row_end = t.row0 + t.nrows;
  float bound = key_bound<METRIC>(st.tau);
  for (int r0 = t.row0; r0 < row_end; r0 += TC_N) {
    const float* brow = nullptr;
    uint32_t myvid = kLmkNoVid;  // validity of row r0 + tid, resolved while the tile is being multiplied
    if (r0 + tid < row_end) {
      brow = lvecs + (int64_t)(r0 + tid) * d;
      const int64_t raw = lids[r0 + tid];
      if (raw >= 0 && ctx_is_valid(f.del_bits, f.filter_bits, (uint32_t)raw)) myvid = (uint32_t)raw;  // tombstone: ivfflat.h:72
    }
    float xn;
    gc = tc_tile_run(smem, &sh, tmem_d, gc, arow, brow, d, &xn);
    s_vid[tid] = myvid;
    __syncthreads();
#pragma unroll 1
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
      uint32_t v[32];
      tc_load32(tmem_d, c0, v);
      lmk_scan_block<METRIC>(v, xn, &sh.cn[c0], &s_vid[c0], hk, k, f.min_score, f.max_score, st, bound);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // next tile's MMAs overwrite the accumulator
    if (valid) {  // exchange bounds with the other CTAs working on this query
      const unsigned long long tg = __ldcg(tau_g + q);
      if (tg < st.tau) {
        st.tau = tg;
        bound = key_bound<METRIC>(st.tau);
      } else if (st.tau < tg) {
        atomicMin(tau_g + q, st.tau);
      }
    }
    __syncthreads();  // s_vid is overwritten by the next tile's operand stores
  }
  if (valid) {
    unsigned long long* o = out + ((int64_t)j * nseg_max + t.seg) * k;
    for (int i = 0; i < k; i++) o[i] = i < st.n ? hk[i * TC_NT] : kKeySentinel;
  }
  tc_release(tmem_d);
}

// ---- K3 list-major, warp-specialised pipeline -------------------------------------------------
// Same work item and the same results as ivf_listmajor_topk_kernel, but the three stages of a tile run
// concurrently in dedicated warps, synchronised by mbarriers only:
//   warps 4-7  producers: thread p stages row p of the query group and of the list tile, one K = 16
//              chunk at a time, into a ring of LW_STAGES operand stages (hi/lo TF32 split, canonical
//              K-major layout), two chunks of global loads in flight per thread;
//   warp  8    one lane issues the tcgen05 MMAs of a chunk as soon as its stage is full; tcgen05.commit
//              frees the stage, and after a tile's last chunk hands the accumulator to the epilogue;
//   warps 0-3  epilogue: tcgen05.ld the 128 x 128 accumulator of tile i (TMEM buffer i & 1) and run
//              the bound test / key-set update while the MMAs of tile i + 1 fill the other buffer.
constexpr int LW_STAGES = 3;
constexpr int LW_NT = 288;
constexpr int LW_RING = LW_STAGES * TC_STAGE_BYTES;

struct LwShared {
  alignas(16) float cn[2][TC_N];  // |y|^2 of the tile's rows (read as float4)
  uint32_t vid[2][TC_N];          // ids of the tile's rows, kLmkNoVid = can never be returned
  float xn[TC_M];                 // |x|^2 of the group's queries
  uint64_t full[LW_STAGES], empty[LW_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};


template <int METRIC>
__global__ void __launch_bounds__(LW_NT, 2)
    ivf_listmajor_pipe_kernel(const float* __restrict__ xq, int64_t ldq, int d, const LmTile* __restrict__ items,
                              const int64_t* __restrict__ totals, const int64_t* __restrict__ pair_j, int nprobe,
                              ListDirectory dir, int k, int nseg_max, FilterArgs f, unsigned long long* tau_g,
                              unsigned long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ LwShared sh;
  if ((int64_t)blockIdx.x >= totals[1]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const LmTile t = items[blockIdx.x];
  const int row_end = t.row0 + t.nrows;
  const int ntiles = (t.nrows + TC_N - 1) / TC_N;
  const int nk = (d + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "n"(2 * TC_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < LW_STAGES; s++) {
      mbar_init(&sh.full[s], TC_M);  // every producer thread arrives
      mbar_init(&sh.empty[s], 1);    // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&sh.acc_full[b], 1);      // tcgen05.commit
      mbar_init(&sh.acc_empty[b], TC_M);  // every epilogue thread arrives
    }
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = sh.tmem_base;
  const uint32_t LBO = TC_M * 16, SBO = 128;

  if (warp < 4) {
    // ======================= epilogue =======================
    unsigned long long* hk = reinterpret_cast<unsigned long long*>(smem + LW_RING) + tid;
    const bool valid = tid < t.npairs;
    int64_t j = 0;
    int q = 0;
    LmkState st{0ull, 0};
    if (valid) {
      j = pair_j[t.pair0 + tid];
      q = (int)(j / nprobe);
      st.tau = __ldcg(tau_g + q);
    }
    float bound = key_bound<METRIC>(st.tau);
    float xn = 0.f;
    for (int i = 0; i < ntiles; i++) {
      const int b = i & 1;
      mbar_wait(&sh.acc_full[b], (uint32_t)((i >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (i == 0) xn = sh.xn[tid];
#pragma unroll 1
      for (int c0 = 0; c0 < TC_N; c0 += 32) {
        uint32_t v[32];
        tc_load32(tmem_d + (uint32_t)(b * TC_N), c0, v);
        lmk_scan_block<METRIC>(v, xn, &sh.cn[b][c0], &sh.vid[b][c0], hk, k, f.min_score, f.max_score, st, bound);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&sh.acc_empty[b]);  // accumulator b, cn[b], vid[b] may be overwritten
      if (valid) {                    // exchange bounds with the other CTAs working on this query
        const unsigned long long tg = __ldcg(tau_g + q);
        if (tg < st.tau) {
          st.tau = tg;
          bound = key_bound<METRIC>(st.tau);
        } else if (st.tau < tg) {
          atomicMin(tau_g + q, st.tau);
        }
      }
    }
    if (valid) {
      unsigned long long* o = out + ((int64_t)j * nseg_max + t.seg) * k;
      for (int i = 0; i < k; i++) o[i] = i < st.n ? hk[i * TC_NT] : kKeySentinel;
    }
  } else if (warp < 8) {
    // ======================= producers =======================
    const int p = tid - TC_M;
    const float* arow = nullptr;
    if (p < t.npairs) arow = xq + (pair_j[t.pair0 + p] / nprobe) * ldq;
    const float* lvecs = dir.vecs[t.list];
    const int64_t* __restrict__ lids = dir.ids[t.list];
    const uint32_t row_off = (uint32_t)(p >> 3) * SBO + (uint32_t)(p & 7) * 16;
    constexpr int NV = TC_BK / 4;
    const int T = ntiles * nk;
    float an_acc = 0.f, bn_acc = 0.f;
    uint32_t myvid = kLmkNoVid;
    float4 pa0[NV], pb0[NV], pa1[NV], pb1[NV];
    auto prefetch = [&](int n, float4 (&pa)[NV], float4 (&pb)[NV]) {
      const int i = n / nk, kc = n - i * nk;
      const int row = t.row0 + i * TC_N + p;
      const float* brow = row < row_end ? lvecs + (int64_t)row * d : nullptr;
#pragma unroll
      for (int kb = 0; kb < NV; kb++) {
        const int gk = kc * TC_BK + kb * 4;
        pa[kb] = make_float4(0.f, 0.f, 0.f, 0.f);
        pb[kb] = pa[kb];
        if (arow && gk < d) pa[kb] = __ldg(reinterpret_cast<const float4*>(arow + gk));
        if (brow && gk < d) pb[kb] = __ldg(reinterpret_cast<const float4*>(brow + gk));
      }
    };
    auto chunk = [&](int n, float4 (&pa)[NV], float4 (&pb)[NV]) {
      const int i = n / nk, kc = n - i * nk;
      const int s = n % LW_STAGES;
      if (kc == 0) {  // new tile: row validity is resolved while the tile's chunks stream through
        bn_acc = 0.f;
        myvid = kLmkNoVid;
        const int row = t.row0 + i * TC_N + p;
        if (row < row_end) {
          const int64_t raw = lids[row];  // tombstone: gamma_index_ivfflat.h:72
          if (raw >= 0 && ctx_is_valid(f.del_bits, f.filter_bits, (uint32_t)raw)) myvid = (uint32_t)raw;
        }
      }
      mbar_wait(&sh.empty[s], (uint32_t)(((n / LW_STAGES) & 1) ^ 1));
      unsigned char* a_hi = smem + (size_t)s * TC_STAGE_BYTES;
      unsigned char* a_lo = a_hi + TC_TILE_BYTES;
      unsigned char* b_hi = a_hi + 2 * TC_TILE_BYTES;
      unsigned char* b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
      for (int kb = 0; kb < NV; kb++) {
        const uint32_t off = (uint32_t)kb * LBO + row_off;
        const float4 va = pa[kb], vb = pb[kb];
        an_acc = fmaf(va.x, va.x, an_acc), an_acc = fmaf(va.y, va.y, an_acc);
        an_acc = fmaf(va.z, va.z, an_acc), an_acc = fmaf(va.w, va.w, an_acc);
        bn_acc = fmaf(vb.x, vb.x, bn_acc), bn_acc = fmaf(vb.y, vb.y, bn_acc);
        bn_acc = fmaf(vb.z, vb.z, bn_acc), bn_acc = fmaf(vb.w, vb.w, bn_acc);
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
      if (kc == nk - 1) {  // tile complete: publish its per-row scalars for the epilogue
        const int b = i & 1;
        mbar_wait(&sh.acc_empty[b], (uint32_t)(((i >> 1) & 1) ^ 1));  // epilogue of tile i-2 is done with them
        sh.cn[b][p] = bn_acc;
        sh.vid[b][p] = myvid;
        if (i == 0) sh.xn[p] = an_acc;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core proxy
      mbar_arrive(&sh.full[s]);
      // after the arrive: a release-arrive waits for the thread's outstanding loads
      if (n + 2 < T) prefetch(n + 2, pa, pb);
    };
    if (T > 0) prefetch(0, pa0, pb0);
    if (T > 1) prefetch(1, pa1, pb1);
    for (int n = 0; n < T; n += 2) {
      chunk(n, pa0, pb0);
      if (n + 1 < T) chunk(n + 1, pa1, pb1);
    }
  } else if (lane == 0) {
    // ======================= MMA issuer =======================
    const uint32_t idesc =
        (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    for (int i = 0; i < ntiles; i++) {
      const int b = i & 1;
      mbar_wait(&sh.acc_empty[b], (uint32_t)(((i >> 1) & 1) ^ 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_d + (uint32_t)(b * TC_N);
      for (int kc = 0; kc < nk; kc++) {
        const int n = i * nk + kc;
        const int s = n % LW_STAGES;
        mbar_wait(&sh.full[s], (uint32_t)((n / LW_STAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + (size_t)s * TC_STAGE_BYTES);
        const uint32_t a_lo = a_hi + TC_TILE_BYTES, b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ks++) {
          const uint32_t koff = (uint32_t)ks * 2 * LBO;
          const uint64_t dah = make_smem_desc(a_hi + koff, LBO, SBO), dal = make_smem_desc(a_lo + koff, LBO, SBO);
          const uint64_t dbh = make_smem_desc(b_hi + koff, LBO, SBO), dbl = make_smem_desc(b_lo + koff, LBO, SBO);
          tc_mma_tf32(acc, dah, dbh, idesc, (kc | ks) != 0);
          tc_mma_tf32(acc, dah, dbl, idesc, 1);
          tc_mma_tf32(acc, dal, dbh, idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&sh.empty[s]))
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&sh.acc_full[b]))
                   : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * TC_N) : "memory");
  }
}

// ---- K3 list-major, TMA-fed pipeline over the pre-tiled mirror ---------------------------------
// Same work items, same epilogue and the same results as ivf_listmajor_pipe_kernel; what changes is
// who feeds the tensor core.  Both operands already sit in global memory in the canonical operand
// layout (lists: the mirror built by tc_mirror_build_kernel; queries: staged per pair group by
// lm_stage_queries_kernel), so one thread issues two 16 KiB cp.async.bulk copies per K chunk into
// a ring of stages and the copies complete on the stage's mbarrier (complete_tx).  No producer warps, no
// registers or scoreboards tied up by loads, as many chunks in flight as the ring is deep.
//   warps 0-3  epilogue (thread = pair): fetch the tile's row ids / norms, tcgen05.ld, bound test
//   warp  4    lane 0: TMA producer          warp 5    lane 0: MMA issuer
constexpr int LT_STAGES = 3;  // 32 KiB stages (query chunk + list chunk), 2 CTAs/SM
constexpr int LT_NT = 192;
constexpr int LT_RING = LT_STAGES * TC_STAGE_BYTES;
constexpr int LT_HALF = TC_STAGE_BYTES / 2;  // one operand's chunk: hi + lo tile, 16 KiB
// (Keeping the group's staged queries resident in shared memory and streaming only the list chunks
// halves the L2 -> SM traffic but fits one CTA per SM only: measured 3.8 ms against 2.3 ms on C2.)

struct LtShared {
  alignas(16) float cn[2][TC_N];
  uint32_t vid[2][TC_N];
  uint64_t full[LT_STAGES], empty[LT_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};



template <int METRIC>
__global__ void __launch_bounds__(LT_NT, 2)
    ivf_listmajor_tma_kernel(const float* __restrict__ a_scratch, const float* __restrict__ a_norms, TcMirrorView mv,
                             const LmTile* __restrict__ items, const int64_t* __restrict__ totals,
                             const int64_t* __restrict__ pair_j, int nprobe, ListDirectory dir, int k, int nseg_max,
                             FilterArgs f, unsigned long long* tau_g, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ LtShared sh;
  if ((int64_t)blockIdx.x >= totals[1]) return;
  constexpr int NST = LT_STAGES;
  constexpr int STAGE = TC_STAGE_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5;
  const LmTile t = items[blockIdx.x];
  const int row_end = t.row0 + t.nrows;
  const int ntiles = (t.nrows + TC_N - 1) / TC_N;
  const int nk = mv.k16 / TC_BK;
  unsigned char* ring = smem;  // [ring: NST * STAGE] [key sets: k * 1 KiB]

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "n"(2 * TC_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < NST; s++) {
      mbar_init(&sh.full[s], 1);   // the producer's arrive.expect_tx; the copies complete the phase
      mbar_init(&sh.empty[s], 1);  // tcgen05.commit
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&sh.acc_full[b], 1);
      mbar_init(&sh.acc_empty[b], TC_M);
    }
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = sh.tmem_base;
  const int64_t tile_floats = (int64_t)TC_N * mv.k16 * 2;
  const int64_t ltile0 = mv.tile0[t.list] + t.row0 / TC_N;  // first mirror tile of this item

  if (warp < 4) {
    // ======================= epilogue =======================
    unsigned long long* hk = reinterpret_cast<unsigned long long*>(ring + (size_t)NST * STAGE) + tid;
    const bool valid = tid < t.npairs;
    int64_t j = 0;
    int q = 0;
    LmkState st{0ull, 0};
    if (valid) {
      j = pair_j[t.pair0 + tid];
      q = (int)(j / nprobe);
      st.tau = __ldcg(tau_g + q);
    }
    float bound = key_bound<METRIC>(st.tau);
    const float xn = a_norms[(int64_t)t.grp * TC_M + tid];
    const int64_t* __restrict__ lids = dir.ids[t.list];
    for (int i = 0; i < ntiles; i++) {
      const int b = i & 1;
      {  // this tile's per-row scalars, fetched while its MMAs run
        const int row = t.row0 + i * TC_N + tid;
        uint32_t myvid = kLmkNoVid;
        float cn = 0.f;
        if (row < row_end) {
          const int64_t raw = lids[row];  // tombstone: gamma_index_ivfflat.h:72
          if (raw >= 0 && ctx_is_valid(f.del_bits, f.filter_bits, (uint32_t)raw)) myvid = (uint32_t)raw;
          cn = mv.norms[(ltile0 + i) * TC_N + tid];
        }
        sh.cn[b][tid] = cn;
        sh.vid[b][tid] = myvid;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      }
      mbar_wait(&sh.acc_full[b], (uint32_t)((i >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c0 = 0; c0 < TC_N; c0 += 32) {
        uint32_t v[32];
        tc_load32(tmem_d + (uint32_t)(b * TC_N), c0, v);
        lmk_scan_block<METRIC>(v, xn, &sh.cn[b][c0], &sh.vid[b][c0], hk, k, f.min_score, f.max_score, st, bound);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&sh.acc_empty[b]);
      if (valid) {
        const unsigned long long tg = __ldcg(tau_g + q);
        if (tg < st.tau) {
          st.tau = tg;
          bound = key_bound<METRIC>(st.tau);
        } else if (st.tau < tg) {
          atomicMin(tau_g + q, st.tau);
        }
      }
    }
    if (valid) {
      unsigned long long* o = out + ((int64_t)j * nseg_max + t.seg) * k;
      for (int i = 0; i < k; i++) o[i] = i < st.n ? hk[i * TC_NT] : kKeySentinel;
    }
  } else if (warp == 4) {
    // ======================= TMA producer (whole warp loops, one elected lane issues) =======================
    const char* asrc = reinterpret_cast<const char*>(a_scratch) + (int64_t)t.grp * nk * LT_HALF;
    const char* bsrc = reinterpret_cast<const char*>(mv.base + ltile0 * tile_floats);
    int n = 0;
    for (int i = 0; i < ntiles; i++) {
      for (int kc = 0; kc < nk; kc++, n++) {
        const int s = n % NST;
        mbar_wait(&sh.empty[s], (uint32_t)(((n / NST) & 1) ^ 1));
        if (elect_one()) {
          unsigned char* stage = ring + (size_t)s * STAGE;
          mbar_arrive_expect_tx(&sh.full[s], STAGE);
          bulk_g2s(stage, asrc + (int64_t)kc * LT_HALF, LT_HALF, &sh.full[s]);
          bulk_g2s(stage + LT_HALF, bsrc + ((int64_t)i * nk + kc) * LT_HALF, LT_HALF, &sh.full[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================= MMA issuer (whole warp loops, one elected lane issues) =======================
    const uint32_t idesc =
        (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t r_base = smem_u32(ring);
    for (int i = 0; i < ntiles; i++) {
      const int b = i & 1;
      mbar_wait(&sh.acc_empty[b], (uint32_t)(((i >> 1) & 1) ^ 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_d + (uint32_t)(b * TC_N);
      for (int kc = 0; kc < nk; kc++) {
        const int n = i * nk + kc;
        const int s = n % NST;
        mbar_wait(&sh.full[s], (uint32_t)((n / NST) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a_hi = r_base + (uint32_t)s * STAGE, b_hi = a_hi + LT_HALF;
          const uint64_t dah = lt_desc(a_hi), dal = lt_desc(a_hi + TC_TILE_BYTES);
          const uint64_t dbh = lt_desc(b_hi), dbl = lt_desc(b_hi + TC_TILE_BYTES);
          constexpr uint64_t KS = (2u * TC_M * 16u) >> 4;  // second K = 8 step: two core-matrix columns further
          tc_mma_tf32(acc, dah, dbh, idesc, kc != 0);
          tc_mma_tf32(acc, dah, dbl, idesc, 1);
          tc_mma_tf32(acc, dal, dbh, idesc, 1);
          tc_mma_tf32(acc, dah + KS, dbh + KS, idesc, 1);
          tc_mma_tf32(acc, dah + KS, dbl + KS, idesc, 1);
          tc_mma_tf32(acc, dal + KS, dbh + KS, idesc, 1);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&sh.empty[s]))
                       : "memory");
          if (kc == nk - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&sh.acc_full[b]))
                         : "memory");
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * TC_N) : "memory");
  }
}

// one operand row -> its 16-byte units of every K chunk, TF32 head and fp32 remainder, canonical
// K-major layout; returns |row|^2 accumulated in index order (the order the staging loops use)
__device__ __forceinline__ float tc_split_row(const float* __restrict__ src, int d, int k16, int r, float* tile) {
  float nrm = 0.f;
  for (int kc = 0; kc < k16 / TC_BK; kc++) {
    float* hi = tile + (int64_t)kc * (TC_STAGE_BYTES / 8);  // 16 KiB per chunk = 4096 floats: hi then lo
    float* lo = hi + TC_TILE_BYTES / 4;
#pragma unroll
    for (int k4 = 0; k4 < 4; k4++) {
      const int gk = kc * TC_BK + k4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src && gk < d) v = __ldg(reinterpret_cast<const float4*>(src + gk));
      nrm = fmaf(v.x, v.x, nrm), nrm = fmaf(v.y, v.y, nrm), nrm = fmaf(v.z, v.z, nrm), nrm = fmaf(v.w, v.w, nrm);
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u), l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u), l.w = v.w - h.w;
      const int off = k4 * (TC_M * 4) + r * 4;  // unit (r, k4): k4 * LBO + r * 16 bytes
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  }
  return nrm;
}

__global__ void __launch_bounds__(TC_N)
    tc_mirror_build_kernel(ListDirectory dir, int d, int k16, const int64_t* __restrict__ tile0, float* __restrict__ mirror,
                           float* __restrict__ norms) {
  const int64_t gt = blockIdx.x;
  int lo = 0, hi = dir.nlist;  // last list whose first tile is <= gt
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile0[mid] <= gt) lo = mid; else hi = mid;
  }
  const int l = lo, r = threadIdx.x;
  const int row = (int)(gt - tile0[l]) * TC_N + r;
  const float* src = row < dir.len[l] ? dir.vecs[l] + (int64_t)row * d : nullptr;
  norms[gt * TC_N + r] = tc_split_row(src, d, k16, r, mirror + gt * ((int64_t)TC_N * k16 * 2));
}

__global__ void tc_mirror_append_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int d, int k16,
                                        const int32_t* __restrict__ list, const int32_t* __restrict__ pos,
                                        const int64_t* __restrict__ tile0, float* __restrict__ mirror,
                                        float* __restrict__ norms) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = list[i];
  if (l < 0) return;
  const int p = pos[i];
  const int64_t gt = tile0[l] + p / TC_N;
  norms[gt * TC_N + p % TC_N] = tc_split_row(x + i * ldx, d, k16, p % TC_N, mirror + gt * ((int64_t)TC_N * k16 * 2));
}

__global__ void __launch_bounds__(TC_M)
    lm_stage_queries_kernel(const float* __restrict__ xq, int64_t ldq, int d, int k16, const LmTile* __restrict__ items,
                            const int64_t* __restrict__ totals, const int64_t* __restrict__ pair_j, int nprobe,
                            float* __restrict__ a_scratch, float* __restrict__ a_norms) {
  if ((int64_t)blockIdx.x >= totals[1]) return;
  const LmTile t = items[blockIdx.x];
  if (t.seg != 0) return;  // one staging per pair group: the first item of the group owns the slot
  const int r = threadIdx.x;
  const float* src = r < t.npairs ? xq + (pair_j[t.pair0 + r] / nprobe) * ldq : nullptr;
  a_norms[(int64_t)t.grp * TC_M + r] = tc_split_row(src, d, k16, r, a_scratch + (int64_t)t.grp * ((int64_t)TC_M * k16 * 2));
}

// ---- grouping for the fused kernel: histogram (lm_count_kernel) -> scan -> slots -> items ----
__device__ __forceinline__ int lmk_nseg(int len, int nseg_max) {
  int n = (len + kLmkSegRows - 1) / kLmkSegRows;
  return n < 1 ? 1 : (n > nseg_max ? nseg_max : n);
}

__global__ void __launch_bounds__(1024)
    lmk_scan_kernel(const int32_t* __restrict__ cnt, ListDirectory dir, int nseg_max, int32_t* __restrict__ start,
                    int32_t* __restrict__ item_start, int32_t* __restrict__ grp_start, int64_t* __restrict__ totals) {
  __shared__ long long s_pairs[1024], s_items[1024], s_grps[1024];
  __shared__ long long carry[3];
  const int tid = threadIdx.x;
  if (tid == 0) carry[0] = carry[1] = carry[2] = 0;
  __syncthreads();
  for (int base = 0; base < dir.nlist; base += 1024) {
    const int l = base + tid;
    long long c = 0, it = 0, g = 0;
    if (l < dir.nlist) {
      c = cnt[l];
      const int len = dir.len[l];
      if (c > 0 && len > 0) {
        g = (c + TC_M - 1) / TC_M;
        it = g * lmk_nseg(len, nseg_max);
      }
    }
    s_pairs[tid] = c, s_items[tid] = it, s_grps[tid] = g;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      long long a = 0, b = 0, e = 0;
      if (tid >= off) a = s_pairs[tid - off], b = s_items[tid - off], e = s_grps[tid - off];
      __syncthreads();
      s_pairs[tid] += a, s_items[tid] += b, s_grps[tid] += e;
      __syncthreads();
    }
    if (l < dir.nlist) {
      start[l] = (int32_t)(carry[0] + s_pairs[tid] - c);
      item_start[l] = (int32_t)(carry[1] + s_items[tid] - it);
      grp_start[l] = (int32_t)(carry[2] + s_grps[tid] - g);
    }
    __syncthreads();
    if (tid == 1023) carry[0] += s_pairs[1023], carry[1] += s_items[1023], carry[2] += s_grps[1023];
    __syncthreads();
  }
  if (tid == 0) totals[0] = carry[2], totals[1] = carry[1], totals[2] = carry[0];  // groups, items, pairs
}

__global__ void lmk_assign_kernel(const int32_t* __restrict__ probe_ids, int64_t npairs, ListDirectory dir,
                                  const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                                  int64_t* __restrict__ pair_j) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npairs) return;
  const int l = probe_ids[j];
  if (l < 0 || l >= dir.nlist || dir.len[l] <= 0) return;
  pair_j[start[l] + atomicAdd(cursor + l, 1)] = j;  // slot order inside a list does not affect results
}

__global__ void lmk_items_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ start,
                                 const int32_t* __restrict__ item_start, const int32_t* __restrict__ grp_start,
                                 ListDirectory dir, int nseg_max, LmTile* __restrict__ items) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= dir.nlist) return;
  const int c = cnt[l], len = dir.len[l];
  if (c <= 0 || len <= 0) return;
  const int nseg = lmk_nseg(len, nseg_max);
  const int seglen = (((len + nseg - 1) / nseg) + TC_N - 1) / TC_N * TC_N;
  int it = item_start[l];
  for (int p0 = 0; p0 < c; p0 += TC_M)
    for (int sg = 0; sg < nseg; sg++) {
      LmTile tl;
      tl.list = l, tl.pair0 = start[l] + p0, tl.npairs = min(TC_M, c - p0);
      tl.row0 = sg * seglen, tl.nrows = max(0, min(seglen, len - tl.row0)), tl.seg = sg;
      tl.grp = grp_start[l] + p0 / TC_M;
      items[it++] = tl;
    }
}

// ---- segment select: per query, stream its P score segments, filter, keep the top-k ---------
constexpr int SEG_NT = 256;
constexpr int SEG_ITEMS = 4;

template <int METRIC>
__global__ void __launch_bounds__(SEG_NT)
    seg_select_kernel(const float* __restrict__ scores, const int64_t* __restrict__ seg_off,
                      const int32_t* __restrict__ probe_ids, int nprobe, ListDirectory dir, int k, int KP, int SORTN,
                      FilterArgs f, unsigned long long* __restrict__ out_keys) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(sel_smem);
  __shared__ int s_cnt;
  __shared__ unsigned long long s_tau;
  CandQueue cq{buf, &s_cnt, &s_tau, k, KP, SORTN};
  cq.init();
  const int q = blockIdx.x;
  int est = 0;
  const int flush_at = max(SEG_NT * SEG_ITEMS - KP, 4 * KP);
  for (int p = 0; p < nprobe; p++) {
    const int list = probe_ids[(int64_t)q * nprobe + p];
    const int64_t off = seg_off[(int64_t)q * nprobe + p];
    if (list < 0 || list >= dir.nlist || off < 0) continue;  // CTA-uniform
    const int len = dir.len[list];
    const float* __restrict__ seg = scores + off;
    const int64_t* __restrict__ lids = dir.ids[list];
    for (int base = 0; base < len; base += SEG_NT * SEG_ITEMS) {
      const unsigned long long tau = s_tau;
      float sv[SEG_ITEMS];
#pragma unroll
      for (int u = 0; u < SEG_ITEMS; u++) {
        const int j = base + u * SEG_NT + threadIdx.x;
        sv[u] = j < len ? seg[j] : 0.f;
      }
      int pushed = 0;
#pragma unroll
      for (int u = 0; u < SEG_ITEMS; u++) {
        const int j = base + u * SEG_NT + threadIdx.x;
        const float sc = sv[u];
        bool pred = j < len && sc <= f.max_score && sc >= f.min_score;
        unsigned long long key = kKeySentinel;
        const uint32_t ord = score2ord<METRIC>(sc);
        pred = pred && ord <= (uint32_t)(tau >> 32);
        if (pred) {
          const int64_t raw = lids[j];
          pred = raw >= 0;  // tombstone (gamma_index_ivfflat.h:72)
          const uint32_t vid = (uint32_t)raw;
          if (pred) pred = ctx_is_valid(f.del_bits, f.filter_bits, vid);
          key = make_key(ord, vid);
          pred = pred && key < tau;
        }
        cq.push_warp(pred, key);
        pushed |= pred ? 1 : 0;
      }
      // exact fill after a barrier pair (one more barrier than an upper bound, but flushes -- block
      // bitonic sorts -- are what this kernel spends its instructions on): flush early so tau
      // tightens after the first ~1k candidates, and whenever the next round might overflow
      (void)pushed;
      __syncthreads();
      est = s_cnt;
      __syncthreads();
      if (est >= flush_at || est + SEG_NT * SEG_ITEMS > cq.cap()) cq.flush();
    }
  }
  __syncthreads();
  cq.flush(true);
  for (int i = threadIdx.x; i < k; i += SEG_NT) out_keys[(int64_t)q * k + i] = buf[i];
}

// ---- list-major grouping on device: histogram -> scan -> slot assignment -> tile table ------
__global__ void lm_count_kernel(const int32_t* __restrict__ probe_ids, int64_t npairs, ListDirectory dir,
                                int32_t* __restrict__ cnt) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npairs) return;
  int l = probe_ids[j];
  if (l >= 0 && l < dir.nlist && dir.len[l] > 0) atomicAdd(cnt + l, 1);
}

// single CTA: exclusive scans over the lists of (pairs, score floats, tiles); totals[0] = floats,
// totals[1] = tiles, totals[2] = valid pairs
__global__ void __launch_bounds__(1024)
    lm_scan_kernel(const int32_t* __restrict__ cnt, ListDirectory dir, int32_t* __restrict__ start,
                   int64_t* __restrict__ base_off, int32_t* __restrict__ tile_start, int64_t* __restrict__ totals) {
  __shared__ long long s_pairs[1024], s_floats[1024], s_tiles[1024];
  __shared__ long long carry[3];
  const int tid = threadIdx.x;
  if (tid == 0) carry[0] = carry[1] = carry[2] = 0;
  __syncthreads();
  for (int base = 0; base < dir.nlist; base += 1024) {
    const int l = base + tid;
    long long c = 0, fl = 0, tl = 0;
    if (l < dir.nlist) {
      c = cnt[l];
      const long long len = dir.len[l];
      fl = c * len;
      tl = ((c + 127) / 128) * ((len + 127) / 128);
    }
    s_pairs[tid] = c, s_floats[tid] = fl, s_tiles[tid] = tl;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      long long a = 0, b = 0, e = 0;
      if (tid >= off) a = s_pairs[tid - off], b = s_floats[tid - off], e = s_tiles[tid - off];
      __syncthreads();
      s_pairs[tid] += a, s_floats[tid] += b, s_tiles[tid] += e;
      __syncthreads();
    }
    if (l < dir.nlist) {
      start[l] = (int32_t)(carry[0] + s_pairs[tid] - c);
      base_off[l] = carry[1] + s_floats[tid] - fl;
      tile_start[l] = (int32_t)(carry[2] + s_tiles[tid] - tl);
    }
    __syncthreads();
    if (tid == 1023) carry[0] += s_pairs[1023], carry[1] += s_floats[1023], carry[2] += s_tiles[1023];
    __syncthreads();
  }
  if (tid == 0) totals[0] = carry[1], totals[1] = carry[2], totals[2] = carry[0];
}

__global__ void lm_assign_kernel(const int32_t* __restrict__ probe_ids, int64_t npairs, int nprobe, ListDirectory dir,
                                 const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                                 const int64_t* __restrict__ base_off, int32_t* __restrict__ pair_q,
                                 int64_t* __restrict__ pair_off, int64_t* __restrict__ seg_off) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npairs) return;
  int l = probe_ids[j];
  if (l < 0 || l >= dir.nlist || dir.len[l] <= 0) {
    seg_off[j] = -1;
    return;
  }
  const int within = atomicAdd(cursor + l, 1);  // slot order inside a list does not affect results
  const int slot = start[l] + within;
  const int64_t off = base_off[l] + (int64_t)within * dir.len[l];
  pair_q[slot] = (int32_t)(j / nprobe);
  pair_off[slot] = off;
  seg_off[j] = off;
}

__global__ void lm_tiles_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ start,
                                const int32_t* __restrict__ tile_start, ListDirectory dir, LmTile* __restrict__ tiles) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= dir.nlist) return;
  const int c = cnt[l], len = dir.len[l];
  if (c <= 0 || len <= 0) return;
  int t = tile_start[l];
  for (int p0 = 0; p0 < c; p0 += 128)
    for (int r0 = 0; r0 < len; r0 += 128) {
      LmTile tl;
      tl.list = l, tl.pair0 = start[l] + p0, tl.npairs = min(128, c - p0), tl.row0 = r0, tl.nrows = min(128, len - r0), tl.seg = 0, tl.grp = 0;
      tiles[t++] = tl;
    }
}

template <int METRIC, int EPI>
cudaError_t launch_tc(const float* X, int64_t ldx, int n, const float* C, int64_t ldc, int m, int d, float* out,
                      int64_t ldo, unsigned long long* best, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(dist_tc_kernel<METRIC, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e != cudaSuccess) return e;
  const int max_rows = 65535 * TC_M;
  for (int r0 = 0; r0 < n; r0 += max_rows) {
    int nr = n - r0 < max_rows ? n - r0 : max_rows;
    dim3 grid((m + TC_N - 1) / TC_N, (nr + TC_M - 1) / TC_M);
    dist_tc_kernel<METRIC, EPI><<<grid, TC_NT, TC_SMEM, st>>>(X + (int64_t)r0 * ldx, ldx, nr, C, ldc, m, d,
                                                             out ? out + (int64_t)r0 * ldo : nullptr, ldo,
                                                             best ? best + r0 : nullptr);
    note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace

cudaError_t launch_dist_matrix_tc(const float* X, int64_t ldx, int n, const float* C, int64_t ldc, int m, int d,
                                  int metric, float* out, int64_t ldo, cudaStream_t st) {
  if ((d & 3) || (ldx & 3) || (ldc & 3)) return cudaErrorInvalidValue;
  if (n <= 0 || m <= 0) return cudaSuccess;
  return metric == kMetricL2 ? launch_tc<kMetricL2, 0>(X, ldx, n, C, ldc, m, d, out, ldo, nullptr, st)
                             : launch_tc<kMetricIP, 0>(X, ldx, n, C, ldc, m, d, out, ldo, nullptr, st);
}

cudaError_t launch_dist_argmin_tc(const float* X, int64_t ldx, int n, const float* C, int64_t ldc, int m, int d,
                                  int metric, unsigned long long* best, cudaStream_t st) {
  if ((d & 3) || (ldx & 3) || (ldc & 3)) return cudaErrorInvalidValue;
  if (n <= 0 || m <= 0) return cudaSuccess;
  return metric == kMetricL2 ? launch_tc<kMetricL2, 1>(X, ldx, n, C, ldc, m, d, nullptr, 0, best, st)
                             : launch_tc<kMetricIP, 1>(X, ldx, n, C, ldc, m, d, nullptr, 0, best, st);
}

cudaError_t launch_ivf_listmajor_tc(const float* xq, int64_t ldq, int d, const LmTile* tiles, int ntiles,
                                    const int32_t* pair_q, const int64_t* pair_off, ListDirectory dir, int metric,
                                    float* scores, cudaStream_t st) {
  if ((d & 3) || (ldq & 3)) return cudaErrorInvalidValue;
  if (ntiles <= 0) return cudaSuccess;
  cudaError_t e;
  if (metric == kMetricL2) {
    e = cudaFuncSetAttribute(ivf_listmajor_tc_kernel<kMetricL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    ivf_listmajor_tc_kernel<kMetricL2><<<ntiles, TC_NT, TC_SMEM, st>>>(xq, ldq, d, tiles, pair_q, pair_off, dir, scores);
  } else {
    e = cudaFuncSetAttribute(ivf_listmajor_tc_kernel<kMetricIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    ivf_listmajor_tc_kernel<kMetricIP><<<ntiles, TC_NT, TC_SMEM, st>>>(xq, ldq, d, tiles, pair_q, pair_off, dir, scores);
  }
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_lm_count_scan(const int32_t* probe_ids, int64_t npairs, ListDirectory dir, int32_t* cnt,
                                 int32_t* start, int64_t* base_off, int32_t* tile_start, int64_t* totals,
                                 cudaStream_t st) {
  if (npairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int32_t) * dir.nlist, st);
  if (e != cudaSuccess) return e;
  lm_count_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(probe_ids, npairs, dir, cnt);
  note_launch();
  lm_scan_kernel<<<1, 1024, 0, st>>>(cnt, dir, start, base_off, tile_start, totals);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_lm_assign_tiles(const int32_t* probe_ids, int64_t npairs, int nprobe, ListDirectory dir,
                                   const int32_t* cnt, const int32_t* start, int32_t* cursor, const int64_t* base_off,
                                   const int32_t* tile_start, int32_t* pair_q, int64_t* pair_off, int64_t* seg_off,
                                   LmTile* tiles, cudaStream_t st) {
  if (npairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int32_t) * dir.nlist, st);
  if (e != cudaSuccess) return e;
  lm_assign_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(probe_ids, npairs, nprobe, dir, start, cursor, base_off,
                                                                    pair_q, pair_off, seg_off);
  note_launch();
  lm_tiles_kernel<<<(dir.nlist + 255) / 256, 256, 0, st>>>(cnt, start, tile_start, dir, tiles);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_lmk_group(const int32_t* probe_ids, int64_t npairs, ListDirectory dir, int nseg_max, int32_t* cnt,
                             int32_t* start, int32_t* cursor, int32_t* item_start, int32_t* grp_start, int64_t* totals,
                             int64_t* pair_j, LmTile* items, cudaStream_t st) {
  if (npairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int32_t) * dir.nlist, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(cursor, 0, sizeof(int32_t) * dir.nlist, st);
  if (e != cudaSuccess) return e;
  const unsigned nb = (unsigned)((npairs + 255) / 256);
  lm_count_kernel<<<nb, 256, 0, st>>>(probe_ids, npairs, dir, cnt);
  note_launch();
  lmk_scan_kernel<<<1, 1024, 0, st>>>(cnt, dir, nseg_max, start, item_start, grp_start, totals);
  note_launch();
  lmk_assign_kernel<<<nb, 256, 0, st>>>(probe_ids, npairs, dir, start, cursor, pair_j);
  note_launch();
  lmk_items_kernel<<<(dir.nlist + 255) / 256, 256, 0, st>>>(cnt, start, item_start, grp_start, dir, nseg_max, items);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_tc_mirror_build(ListDirectory dir, int d, int k16, const int64_t* tile0, int64_t total_tiles,
                                   float* mirror, float* norms, cudaStream_t st) {
  if (total_tiles <= 0) return cudaSuccess;
  if (total_tiles > INT32_MAX || (d & 3) || (k16 % TC_BK)) return cudaErrorInvalidValue;
  tc_mirror_build_kernel<<<(unsigned)total_tiles, TC_N, 0, st>>>(dir, d, k16, tile0, mirror, norms);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_tc_mirror_append(const float* x, int64_t ldx, int64_t n, int d, int k16, const int32_t* list,
                                    const int32_t* pos, const int64_t* tile0, float* mirror, float* norms, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  tc_mirror_append_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(x, ldx, n, d, k16, list, pos, tile0, mirror, norms);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_lm_stage_queries(const float* xq, int64_t ldq, int d, int k16, const LmTile* items, int max_items,
                                    const int64_t* totals, const int64_t* pair_j, int nprobe, float* a_scratch,
                                    float* a_norms, cudaStream_t st) {
  if (max_items <= 0) return cudaSuccess;
  lm_stage_queries_kernel<<<max_items, TC_M, 0, st>>>(xq, ldq, d, k16, items, totals, pair_j, nprobe, a_scratch, a_norms);
  note_launch();
  return cudaGetLastError();
}

template <int METRIC>
static cudaError_t launch_tma_t(const float* a_scratch, const float* a_norms, TcMirrorView mv, const LmTile* items,
                                int max_items, const int64_t* totals, const int64_t* pair_j, int nprobe, ListDirectory dir,
                                int k, int nseg_max, FilterArgs f, unsigned long long* tau_g, unsigned long long* out,
                                cudaStream_t st) {
  const size_t smem = (size_t)LT_RING + (size_t)k * TC_NT * 8;
  cudaError_t e = cudaFuncSetAttribute(ivf_listmajor_tma_kernel<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  ivf_listmajor_tma_kernel<METRIC><<<max_items, LT_NT, smem, st>>>(a_scratch, a_norms, mv, items, totals, pair_j, nprobe, dir,
                                                                  k, nseg_max, f, tau_g, out);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_ivf_listmajor_tma(const float* a_scratch, const float* a_norms, TcMirrorView mv, const LmTile* items,
                                     int max_items, const int64_t* totals, const int64_t* pair_j, int nprobe,
                                     ListDirectory dir, int k, int nseg_max, int metric, FilterArgs f,
                                     unsigned long long* tau_g, unsigned long long* out, cudaStream_t st) {
  if (k <= 0 || k > kLmkMaxK || mv.k16 <= TC_BK || (mv.k16 % TC_BK)) return cudaErrorInvalidValue;
  if (max_items <= 0) return cudaSuccess;
  if (metric == kMetricL2)
    return launch_tma_t<kMetricL2>(a_scratch, a_norms, mv, items, max_items, totals, pair_j, nprobe, dir, k, nseg_max, f,
                                   tau_g, out, st);
  return launch_tma_t<kMetricIP>(a_scratch, a_norms, mv, items, max_items, totals, pair_j, nprobe, dir, k, nseg_max, f, tau_g,
                                 out, st);
}

cudaError_t launch_ivf_listmajor_topk(const float* xq, int64_t ldq, int d, const LmTile* items, int max_items,
                                      const int64_t* totals, const int64_t* pair_j, int nprobe, ListDirectory dir, int k,
                                      int nseg_max, int metric, FilterArgs f, unsigned long long* tau_g,
                                      unsigned long long* out, cudaStream_t st) {
  if ((d & 3) || (ldq & 3) || k <= 0 || k > kLmkMaxK) return cudaErrorInvalidValue;
  if (max_items <= 0) return cudaSuccess;
  static const int pipe = [] {
    const char* e = getenv("GB_LM_PIPE");
    return e ? atoi(e) : 1;
  }();
  if (pipe && d > TC_BK) {  // the ring's phase arithmetic assumes >= 2 chunks per tile
    const size_t smem = (size_t)LW_RING + (size_t)k * TC_NT * 8;
    cudaError_t e;
    if (metric == kMetricL2) {
      e = cudaFuncSetAttribute(ivf_listmajor_pipe_kernel<kMetricL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      ivf_listmajor_pipe_kernel<kMetricL2><<<max_items, LW_NT, smem, st>>>(xq, ldq, d, items, totals, pair_j, nprobe, dir,
                                                                          k, nseg_max, f, tau_g, out);
    } else {
      e = cudaFuncSetAttribute(ivf_listmajor_pipe_kernel<kMetricIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      ivf_listmajor_pipe_kernel<kMetricIP><<<max_items, LW_NT, smem, st>>>(xq, ldq, d, items, totals, pair_j, nprobe, dir,
                                                                          k, nseg_max, f, tau_g, out);
    }
    note_launch();
    return cudaGetLastError();
  }
  const size_t smem = (size_t)TC_SMEM + (size_t)k * TC_NT * 8;
  cudaError_t e;
  if (metric == kMetricL2) {
    e = cudaFuncSetAttribute(ivf_listmajor_topk_kernel<kMetricL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ivf_listmajor_topk_kernel<kMetricL2><<<max_items, TC_NT, smem, st>>>(xq, ldq, d, items, totals, pair_j, nprobe, dir, k,
                                                                        nseg_max, f, tau_g, out);
  } else {
    e = cudaFuncSetAttribute(ivf_listmajor_topk_kernel<kMetricIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ivf_listmajor_topk_kernel<kMetricIP><<<max_items, TC_NT, smem, st>>>(xq, ldq, d, items, totals, pair_j, nprobe, dir, k,
                                                                        nseg_max, f, tau_g, out);
  }
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_seg_select(const float* scores, const int64_t* seg_off, const int32_t* probe_ids, int nq, int nprobe,
                              ListDirectory dir, int k, int metric, FilterArgs f, unsigned long long* out_keys,
                              cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  if (k <= 0 || k > 4096) return cudaErrorInvalidValue;
  const int KP = next_pow2(k < 16 ? 16 : k);
  const int SORTN = next_pow2(KP + 2 * SEG_NT * SEG_ITEMS);
  const size_t smem = (size_t)SORTN * 8;
  cudaError_t e;
  if (metric == kMetricL2) {
    e = cudaFuncSetAttribute(seg_select_kernel<kMetricL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    seg_select_kernel<kMetricL2><<<nq, SEG_NT, smem, st>>>(scores, seg_off, probe_ids, nprobe, dir, k, KP, SORTN, f, out_keys);
  } else {
    e = cudaFuncSetAttribute(seg_select_kernel<kMetricIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    seg_select_kernel<kMetricIP><<<nq, SEG_NT, smem, st>>>(scores, seg_off, probe_ids, nprobe, dir, k, KP, SORTN, f, out_keys);
  }
  note_launch();
  return cudaGetLastError();
}

}  // namespace gb
