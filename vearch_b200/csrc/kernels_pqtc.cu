// K5, list-major: the IVF-PQ ADC scan of a large query batch as a filtered tensor-core contraction.
//
// Reference semantics: GammaIVFPQScanner::scan_list_with_table (index/impl/gamma_index_ivfpq.h:923-953)
// driven by search_preassigned (gamma_index_ivfpq.cc:730-947): per (query, probed list, entry)
//     dis = dis0 + sum_m tab[m][code[m]],   tab = T[list] - 2 ip[query]   (L2, precomputed-table form)
//     dis = <x, centroid> + sum_m ip[query][m][code[m]]                    (inner product)
// and the k' best (k' = recall_num or k) survive.  ivfpq_scan_kernel (kernels_ivfpq.cu) evaluates that
// sum literally, 16 shared-memory gathers per (query, entry): it is bound by instruction issue and
// LDS wavefronts, not by HBM.  When every list is probed by dozens of queries of the same batch the
// same numbers are a dense contraction:
//     sum_m tab[m][code[m]] = |r_e|^2 - 2 <x - c_list, r_e>,   r_e = the entry's PQ reconstruction,
// i.e. (query, list) pairs x decoded entries.  This file computes THAT on tcgen05 in fp16 (power-of-two scaled) and uses it
// only as a FILTER; every number that leaves the pipeline is recomputed with the reference's own
// arithmetic, so results stay bit-identical to the LUT kernel and to the oracle:
//
//   phase A  the exact LUT kernel scans each query's leading probes in full (the fewest whose lists hold
//            >= 4 k' entries, pqtc_plan_phase_a_kernel) and yields k' exact keys, unordered, the largest --
//            B_q, an upper bound of the query's final k'-th score -- in slot k' - 1;
//   phase B  (this file) for the remaining probes: pairs grouped by list (lmk_group), per group a
//            fp16 operand tile of sa * (x - c_list) rows staged once (pq_stage_pairs_kernel), then a
//            persistent warp-specialised kernel (pqtc_scan_kernel) decodes 128 entries at a time
//            from their codes into the second operand (codebook in shared memory, pre-scaled by -2),
//            multiplies 128 pairs x 128 entries x d on the tensor core and compares
//            acc + |r_e|^2 against  B_q - dis0 + eps(pair).  eps bounds |approximate - reference fp32|
//            rigorously (fp16 rounding of both operands: 2^-10 |a||r| ; fp32 noise of both evaluations),
//            so every entry whose reference score is <= B_q passes.  |r_e|^2 comes from a per-entry cache
//            in HBM (pq_entry_norms_kernel), streamed with the codes.  Passing entries (a few per query)
//            go through a shared-memory queue to a drain warp that appends (probe, position) to the
//            query's candidate list, so no global atomic sits in the epilogue;
//   phase C  pq_rescore_kernel re-evaluates the candidates with the reference arithmetic (same
//            fmaf / add order as the LUT kernel), merges them with phase A's keys and keeps k'.
//   A query whose candidate list overflows (no usable bound, adversarial data) is redone by the
//   exact kernel over all its probes (flag-gated launch), so the result never depends on the filter.
//
// Operand layout: canonical no-swizzle K-major UMMA tiles of fp16 (8-row x 16-byte core matrices,
// LBO 2048 = next core matrix along K, SBO 128 = next 8 rows).  With dsub = 8 a sub-quantiser's
// centroid IS one 16-byte core-matrix row: decoding an entry is M x (LDS.128 from the codebook,
// STS.128 into the tile), conflict-free on the store side (consecutive entries -> consecutive rows).
//
// Algorithmic bytes: (M + 8) per scanned entry per 128-pair group instead of per pair.
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace gb {

namespace {

constexpr int PT_M = 128;          // pairs per group  (MMA M, TMEM lanes)
constexpr int PT_N = 128;          // entries per tile (MMA N, accumulator columns)
constexpr int PT_NT = 352;         // warps 0-3 epilogue, 4-7 decode, 8 TMA producer, 9 MMA issuer, 10 candidate drain
constexpr int PT_RQ = 512;         // records of the CTA's candidate queue (shared memory)
constexpr unsigned long long kRecEmpty = ~0ull;
constexpr int PT_CS = 4;           // code stages
constexpr int PT_KSUB = 256;

__device__ unsigned long long g_pqtc_dbg[4];  // debugging counters (GB_PQTC_DBG & 4): [0] stale code words seen by decode

struct PtShared {
  uint64_t code_full[PT_CS], code_empty[PT_CS];
  uint64_t b_full[2], b_empty[2], a_full[2], a_empty[2], acc_full[2], acc_empty[2];
  alignas(16) float ne[2][PT_N];  // |r_e|^2 of the tile's entries, +inf = can never be returned
  uint32_t tmem_base;
  // candidate queue: the epilogue warps append (query, probe, position) records with shared-memory atomics and move on;
  // the drain warp turns them into the global per-query lists.  A global atomicAdd-with-return costs ~1 us of latency:
  // issued from the epilogue it stalled the whole pipeline as soon as a loose bound let a percent of the entries through.
  unsigned int q_head, q_tail;
  int q_done;
  unsigned long long q_rec[PT_RQ];  // (q << 48 | p << 32 | pos), kRecEmpty = free slot
};

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}
// power of two that brings a magnitude `mx` into [2^13, 2^14): fp16 keeps 11 significant bits there and a further
// factor 2 (the -2 of the L2 form) still fits below 65504
__device__ __forceinline__ float pow2_scale_for(float mx) { return mx > 0.f && mx < INFINITY ? ldexpf(1.0f, 13 - ilogbf(mx)) : 1.0f; }

// ---- tables derived from the PQ codebook ---------------------------------------------------------
// sb = power of two scaling the largest codebook component into [2^13, 2^14) (exact in fp32 and fp16);
// cb[m][c][0..dsub) = fp16(scale * sb * pq[m][c][.]),  nrm[m][c] = |pq[m][c]|^2 (fp32, L2 only),
// rmax2[0] = sum_m max_c |pq[m][c]|^2,  rmax2[1] = sb   (single CTA of 256 threads)
__global__ void __launch_bounds__(256)
    pqtc_tables_kernel(const float* __restrict__ pq, int M, int dsub, float scale, uint16_t* __restrict__ cb,
                       float* __restrict__ nrm, float* __restrict__ rmax2) {
  __shared__ float s_max[256];
  const int c = threadIdx.x;
  float mx = 0.f;
  for (int i = c; i < M * PT_KSUB * dsub; i += 256) mx = fmaxf(mx, fabsf(pq[i]));
  s_max[c] = mx;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (c < off) s_max[c] = fmaxf(s_max[c], s_max[c + off]);
    __syncthreads();
  }
  const float sb = pow2_scale_for(s_max[0]);
  __syncthreads();
  float total = 0.f;
  for (int m = 0; m < M; m++) {
    const float* p = pq + ((int64_t)m * PT_KSUB + c) * dsub;
    float n2 = 0.f;
    for (int j = 0; j < dsub; j++) {
      const float v = p[j];
      n2 = fmaf(v, v, n2);
      const uint32_t b = pack_f16x2(scale * sb * v, 0.f);
      cb[((int64_t)m * PT_KSUB + c) * dsub + j] = (uint16_t)(b & 0xFFFFu);
    }
    nrm[m * PT_KSUB + c] = n2;
    s_max[c] = n2;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (c < off) s_max[c] = fmaxf(s_max[c], s_max[c + off]);
      __syncthreads();
    }
    total += s_max[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) rmax2[0] = total, rmax2[1] = sb;
}

// ---- phase A's plan ---------------------------------------------------------------------------------------------------
// Query q's phase A scans its first P_q probes IN FULL: the fewest whose lists hold >= target entries together, at most
// pa_max.  probes_a / probes_b = probe_ids with the other phase's probes set to -1 ("no list", skipped by both scans);
// row_limit[q][p] = rows of probe p phase A scores exactly (the whole list, or 0).  A tight bound matters far more than
// a cheap phase A: bounding from a 3 200-row PREFIX of the nearest list instead of the whole list was measured at
// 60x the candidates and a 2.8x slower filter kernel at C3, 4x slower at 100 M vectors (profiles/README.md, r2 notes).
__global__ void pqtc_plan_phase_a_kernel(const int32_t* __restrict__ probe_ids, int64_t npairs, int nprobe, int pa_max,
                                         long long target, const int* __restrict__ list_len, int32_t* __restrict__ probes_a,
                                         int32_t* __restrict__ probes_b, int* __restrict__ row_limit) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npairs) return;
  const int64_t q = j / nprobe;
  const int p = (int)(j - q * nprobe);
  const int32_t* row = probe_ids + q * nprobe;
  int P = pa_max;
  long long cum = 0;
  for (int i = 0; i < pa_max; i++) {
    const int l = row[i];
    if (l >= 0) cum += list_len[l];
    if (cum >= target) {
      P = i + 1;
      break;
    }
  }
  const int32_t id = row[p];
  const bool in_a = p < P;
  probes_a[j] = in_a ? id : -1;
  probes_b[j] = in_a ? -1 : id;
  row_limit[j] = in_a && id >= 0 ? list_len[id] : 0;
}

// ---- per pair group: the A operand tile and the pairs' thresholds ---------------------------------
struct PairMeta {
  float thr;  // candidate  <=>  acc + c * ne <= thr   (thr, lo already multiplied by c)
  float lo;   // ... and acc + c * ne >= lo  (the other side of the score window, checked on the rare path)
  float c;    // sa * sb: the operands' power-of-two scales (accumulator = c * (-2 <a, r>))
  int q, p;
  int lim;    // rows [0, lim) of this pair's list were scored exactly by phase A: never candidates
};

template <int METRIC>
__global__ void __launch_bounds__(PT_M)
    pq_stage_pairs_kernel(const float* __restrict__ xq, int64_t ldq, int d, const float* __restrict__ coarse, int64_t ldc,
                          const LmTile* __restrict__ items, const int64_t* __restrict__ totals,
                          const int64_t* __restrict__ pair_j, int nprobe, const float* __restrict__ coarse_dis,
                          const unsigned long long* __restrict__ bound_keys, int64_t bound_stride, int kprime,
                          const float* __restrict__ rmax2, float min_score, float max_score, float eps_scale,
                          unsigned char* __restrict__ a_scratch, PairMeta* __restrict__ meta, int* __restrict__ cand_cnt,
                          int cap, const int* __restrict__ row_limit) {
  if ((int64_t)blockIdx.x >= totals[1]) return;
  const LmTile t = items[blockIdx.x];
  if (t.seg != 0) return;  // one staging per pair group: the first item of the group owns the slot
  const int r = threadIdx.x;
  const int kc_n = d / 8;
  unsigned char* tile = a_scratch + (int64_t)t.grp * ((int64_t)kc_n * 2048);
  PairMeta pm{-INFINITY, INFINITY, 1.0f, -1, 0, 0};
  const float* x = nullptr;
  const float* c = nullptr;
  if (r < t.npairs) {
    const int64_t j = pair_j[t.pair0 + r];
    pm.q = (int)(j / nprobe);
    pm.p = (int)(j - (int64_t)pm.q * nprobe);
    x = xq + (int64_t)pm.q * ldq;
    c = coarse + (int64_t)t.list * ldc;
  }
  // pass 1: norms and the largest component of the row a = x - c (L2) / x (IP)
  float na2 = 0.f, nx2 = 0.f, nc2 = 0.f, amax = 0.f;
  if (x) {
    for (int kc = 0; kc < kc_n; kc++) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + kc * 8)), x1 = __ldg(reinterpret_cast<const float4*>(x + kc * 8 + 4));
      float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      if (METRIC == kMetricL2) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(c + kc * 8)), c1 = __ldg(reinterpret_cast<const float4*>(c + kc * 8 + 4));
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int u = 0; u < 8; u++) {
          nx2 = fmaf(a[u], a[u], nx2);
          nc2 = fmaf(cc[u], cc[u], nc2);
          a[u] -= cc[u];
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) na2 = fmaf(a[u], a[u], na2), amax = fmaxf(amax, fabsf(a[u]));
    }
  }
  // pass 2: the row as fp16(sa * a), sa a power of two (exact), in the canonical K-major operand layout
  const float sa = pow2_scale_for(amax);
  for (int kc = 0; kc < kc_n; kc++) {
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (x) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + kc * 8)), x1 = __ldg(reinterpret_cast<const float4*>(x + kc * 8 + 4));
      float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      if (METRIC == kMetricL2) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(c + kc * 8)), c1 = __ldg(reinterpret_cast<const float4*>(c + kc * 8 + 4));
        a[0] -= c0.x, a[1] -= c0.y, a[2] -= c0.z, a[3] -= c0.w, a[4] -= c1.x, a[5] -= c1.y, a[6] -= c1.z, a[7] -= c1.w;
      }
      out.x = pack_f16x2(sa * a[0], sa * a[1]), out.y = pack_f16x2(sa * a[2], sa * a[3]);
      out.z = pack_f16x2(sa * a[4], sa * a[5]), out.w = pack_f16x2(sa * a[6], sa * a[7]);
    }
    *reinterpret_cast<uint4*>(tile + (int64_t)kc * 2048 + r * 16) = out;
  }
  if (x) {
    const float dis0 = coarse_dis[(int64_t)pm.q * nprobe + pm.p];
    pm.lim = row_limit ? row_limit[(int64_t)pm.q * nprobe + pm.p] : 0;
    const unsigned long long bk = bound_keys[(int64_t)pm.q * bound_stride + kprime - 1];
    const float R2 = rmax2[0], R = sqrtf(R2);
    pm.c = sa * rmax2[1];
    if (bk == kKeySentinel) {
      // phase A found fewer than k' entries: no bound -> the exact kernel redoes this query
      atomicMax(cand_cnt + pm.q, cap + 1);
    } else {
      const float B = ord2score((uint32_t)(bk >> 32), METRIC);
      float thr, lo;
      if (METRIC == kMetricL2) {
        // |approx - reference| <= 2^-10 (1 + 2^-12) |a| |r|: fp16 rounding (unit roundoff 2^-12) of both operands of
        //   -2 <a, r>; the power-of-two scales are exact, components below fp16's normal range are >= 2^27 times
        //   smaller than the row's largest and fall inside the 5 % slack, as does the fp32 accumulation of the MMA
        //   (<= 2^-15 |a| |r| for d <= 4096 even if every add truncated)
        //   + fp32 noise of the reference evaluation, bounded by 2^-17 of the magnitudes that enter it
        const float Z = fabsf(dis0) + R2 + 2.f * (sqrtf(nx2) + sqrtf(nc2)) * R;
        const float eps = eps_scale * (1.05f * 0.0009765625f * sqrtf(na2) * R + 7.62939453125e-6f * Z);
        thr = fminf(B, max_score) - dis0 + eps;
        lo = min_score - dis0 - eps;
      } else {
        // score = dis0 - acc / c (the codebook is staged negated); better = larger
        const float Z = fabsf(dis0) + sqrtf(na2) * R;
        const float eps = eps_scale * (1.05f * 0.00048828125f * sqrtf(na2) * R + 7.62939453125e-6f * Z);
        thr = dis0 - fmaxf(B, min_score) + eps;
        lo = dis0 - max_score - eps;
      }
      // compared against acc + c * ne; FLT_MAX keeps masked entries (ne = +inf) out even when the product overflows
      pm.thr = fminf(thr * pm.c, FLT_MAX);
      pm.lo = lo * pm.c;
    }
  }
  meta[(int64_t)t.grp * PT_M + r] = pm;
}

// ---- the scan ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pt_load32(uint32_t taddr, uint32_t (&v)[32]) {
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

__device__ __forceinline__ void pt_load32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
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
}

// rare path of the epilogue: the 32 columns of one pair contain at least one candidate.
// Called by the WHOLE warp (tcgen05.ld is .sync.aligned: every lane must execute it, convergently); `active` = this
// lane's pair has a candidate among the 32 columns.  The columns are read from TMEM again here: handing the caller's
// register block over by reference would force a local-memory copy of every accumulator block in the hot loop.
__device__ __noinline__ void pt_push_hits(uint32_t taddr, const float* ne32, float c, float thr, float lo, int q, int p, uint32_t pos0,
                                          int lim, PtShared* sh, bool active) {
  uint32_t v[32];
  pt_load32(taddr, v);
  uint32_t mask = 0;
  if (active) {
#pragma unroll
    for (int j = 0; j < 32; j++) {
      const float s = fmaf(ne32[j], c, __uint_as_float(v[j]));
      if (s <= thr && s >= lo) mask |= 1u << j;
    }
    // rows phase A has scored exactly stay out
    if ((int)pos0 < lim) mask = (int)pos0 + 32 <= lim ? 0u : mask & ~((1u << (lim - (int)pos0)) - 1u);
  }
  if (mask) {
    unsigned int idx = atomicAdd(&sh->q_head, (unsigned int)__popc(mask));
    const unsigned long long hi = ((unsigned long long)(uint32_t)q << 48) | ((unsigned long long)(uint32_t)p << 32);
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      // the slot's previous occupant (idx - PT_RQ) has been consumed once the tail has passed it
      while (idx - *reinterpret_cast<volatile unsigned int*>(&sh->q_tail) >= (unsigned int)PT_RQ) __nanosleep(32);
      *reinterpret_cast<volatile unsigned long long*>(&sh->q_rec[idx % PT_RQ]) = hi | (unsigned long long)(pos0 + (uint32_t)j);
      idx++;
    }
  }
  __syncwarp();
}

// EAGER: tombstones and the docid predicate are resolved in the decode warps (one id + bitmap lookup per entry per
// 128 pairs) -- needed when a deletion / filter bitmap is present, because the bound B_q was taken over valid entries
// only and a selective filter would otherwise flood the candidate lists.  Without bitmaps the ids are not touched
// here at all (their DRAM latency sat on the decode warps' critical path: profiles/r2_ncu_pqtc_scan.txt);
// tombstoned entries that pass the filter are dropped by pq_rescore_kernel, which reads the id anyway.
template <int M, int DSUB, bool HAS_NORM, bool EAGER>
__global__ void __launch_bounds__(PT_NT, 1)
    pqtc_scan_kernel(const unsigned char* __restrict__ a_scratch, const PairMeta* __restrict__ meta,
                     const uint16_t* __restrict__ cb_g, const float* __restrict__ pqnorm,
                     const int64_t* __restrict__ pqnorm_off, const LmTile* __restrict__ items,
                     const int64_t* __restrict__ totals, ListDirectory dir, FilterArgs f, int* __restrict__ cand_cnt,
                     unsigned long long* __restrict__ cand, int cap, int tile_stride, int dbg) {
  constexpr int D = M * DSUB;
  constexpr int KC = D / 8;                // core matrices along K
  constexpr int TILE = KC * 2048;          // one 128-row operand tile, bytes
  constexpr int CB_BYTES = M * PT_KSUB * DSUB * 2;
  // |r_e|^2 of every list entry is cached in HBM beside the lists (IVFPQIndex::ensure_pq_norms) and arrives with the
  // codes: 512 bytes per tile by TMA instead of 16 table gathers per entry in the decode warps (a sixth of the kernel's
  // shared-memory wavefronts, profiles/r2_ncu_pqtc_scan.txt)
  constexpr int NRM_BYTES = HAS_NORM ? PT_CS * PT_N * 4 : 0;
  constexpr int CODE_STAGE = PT_N * M;
  constexpr int UNIT = DSUB * 2;           // bytes of one centroid in the fp16 codebook
  static_assert(D % 16 == 0 && M % 4 == 0, "shape");
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ PtShared sh;
  unsigned char* cb = smem;
  float* nstage = reinterpret_cast<float*>(smem + CB_BYTES);  // [PT_CS][PT_N] norms of the staged tiles
  // Operand tiles start on 4 KiB boundaries, so the 4 KiB block one MMA reads (two core-matrix columns, LBO apart)
  // never straddles a 128 KiB line of the shared-memory window.  Measured on B200: with a tile at 0x1e800 the MMA
  // whose second core-matrix column began exactly at 0x20000 intermittently produced wrong accumulator columns;
  // every other placement tried was clean (GB_PQTC_DBG bit 5 + bits 8.. place the tiles at a chosen offset to reproduce).
  unsigned char* a_buf = smem + CB_BYTES + NRM_BYTES;
  if (!(dbg & 32)) a_buf += (4096u - (smem_u32(a_buf) & 4095u)) & 4095u;
  else a_buf += (dbg >> 8) & 0xFFF0;  // debugging: place the tiles at a chosen (mis)alignment
  unsigned char* b_buf = a_buf + 2 * tile_stride;
  unsigned char* code_buf = b_buf + 2 * tile_stride;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t n_items = totals[1];

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "n"(2 * PT_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < PT_CS; s++) {
      mbar_init(&sh.code_full[s], 1);       // producer's arrive.expect_tx; the bulk copy completes the phase
      mbar_init(&sh.code_empty[s], PT_N);   // every decode thread has read its code
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&sh.b_full[b], PT_N);       // every decode thread has written its row
      mbar_init(&sh.b_empty[b], 1);         // tcgen05.commit
      mbar_init(&sh.a_full[b], 1);          // bulk copy
      mbar_init(&sh.a_empty[b], 1);         // tcgen05.commit after the item's last tile
      mbar_init(&sh.acc_full[b], 1);        // tcgen05.commit
      mbar_init(&sh.acc_empty[b], PT_M);    // every epilogue thread
    }
    mbar_fence_init();
    sh.q_head = 0, sh.q_tail = 0, sh.q_done = 0;
  }
  for (int i = tid; i < PT_RQ; i += PT_NT) sh.q_rec[i] = kRecEmpty;
  // codebook (fp16, pre-scaled) and centroid norms: resident for the CTA's lifetime
  for (int i = tid; i < CB_BYTES / 16; i += PT_NT)
    reinterpret_cast<uint4*>(cb)[i] = __ldg(reinterpret_cast<const uint4*>(cb_g) + i);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = sh.tmem_base;
  if ((dbg & 64) && blockIdx.x == 0 && tid == 0)
    printf("[pqtc] smem: dyn base 0x%x a_buf 0x%x b_buf 0x%x B[1] 0x%x code 0x%x\n", smem_u32(smem), smem_u32(a_buf), smem_u32(b_buf), smem_u32(b_buf) + tile_stride, smem_u32(code_buf));

  if (warp < 4) {
    // ======================= epilogue: thread = pair =======================
    uint32_t bn = 0;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const LmTile t = items[it];
      PairMeta pm = meta[(int64_t)t.grp * PT_M + tid];
      const int ntiles = (t.nrows + PT_N - 1) / PT_N;
      for (int i = 0; i < ntiles; i++, bn++) {
        const int b = bn & 1;
        mbar_wait(&sh.acc_full[b], (bn >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg & 16) __nanosleep(1000);
#pragma unroll 1
        for (int cc = 0; cc < PT_N; cc += 64) {
          // two 32-column blocks per wait: the second load's latency hides behind the first block's arithmetic
          uint32_t v0[32], v1[32];
          const uint32_t ta = tmem_d + lane_base + (uint32_t)(b * PT_N + cc);
          pt_load32_nowait(ta, v0);
          pt_load32_nowait(ta + 32, v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float* ne0 = &sh.ne[b][cc];
          float mn0 = INFINITY, mn1 = INFINITY;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            const float4 n4 = *reinterpret_cast<const float4*>(ne0 + j4);
            const float4 m4 = *reinterpret_cast<const float4*>(ne0 + 32 + j4);
            const float s0 = fmaf(n4.x, pm.c, __uint_as_float(v0[j4])), s1 = fmaf(n4.y, pm.c, __uint_as_float(v0[j4 + 1]));
            const float s2 = fmaf(n4.z, pm.c, __uint_as_float(v0[j4 + 2])), s3 = fmaf(n4.w, pm.c, __uint_as_float(v0[j4 + 3]));
            const float u0 = fmaf(m4.x, pm.c, __uint_as_float(v1[j4])), u1 = fmaf(m4.y, pm.c, __uint_as_float(v1[j4 + 1]));
            const float u2 = fmaf(m4.z, pm.c, __uint_as_float(v1[j4 + 2])), u3 = fmaf(m4.w, pm.c, __uint_as_float(v1[j4 + 3]));
            mn0 = fminf(mn0, fminf(fminf(s0, s1), fminf(s2, s3)));
            mn1 = fminf(mn1, fminf(fminf(u0, u1), fminf(u2, u3)));
          }
          const uint32_t pos = (uint32_t)(t.row0 + i * PT_N + cc);
          const bool h0 = mn0 <= pm.thr, h1 = mn1 <= pm.thr;
          if (__any_sync(0xffffffffu, h0)) pt_push_hits(ta, ne0, pm.c, pm.thr, pm.lo, pm.q, pm.p, pos, pm.lim, &sh, h0);
          if (__any_sync(0xffffffffu, h1))
            pt_push_hits(ta + 32, ne0 + 32, pm.c, pm.thr, pm.lo, pm.q, pm.p, pos + 32, pm.lim, &sh, h1);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&sh.acc_empty[b]);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) {
      __threadfence_block();
      atomicAdd(&sh.q_done, 1);  // this warp's records are all in the queue
    }
  } else if (warp < 8) {
    // ======================= decode: thread = entry of the tile =======================
    const int e = tid - 4 * 32;
    uint32_t cn = 0, bn = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const LmTile t = items[it];
      const int64_t* __restrict__ lids = dir.ids[t.list];
      const int row_end = t.row0 + t.nrows;
      const int ntiles = (t.nrows + PT_N - 1) / PT_N;
      for (int i = 0; i < ntiles; i++, cn++, bn++) {
        const int s = cn % PT_CS;
        mbar_wait(&sh.code_full[s], (cn / PT_CS) & 1);
        uint32_t w[M / 4];
        {
          const unsigned char* cp = code_buf + s * CODE_STAGE + e * M;
          if (M % 16 == 0) {
#pragma unroll
            for (int u = 0; u < M / 16; u++) {
              const uint4 q4 = reinterpret_cast<const uint4*>(cp)[u];
              w[4 * u] = q4.x, w[4 * u + 1] = q4.y, w[4 * u + 2] = q4.z, w[4 * u + 3] = q4.w;
            }
          } else if (M % 8 == 0) {
#pragma unroll
            for (int u = 0; u < M / 8; u++) {
              const uint2 q2 = reinterpret_cast<const uint2*>(cp)[u];
              w[2 * u] = q2.x, w[2 * u + 1] = q2.y;
            }
          } else {
#pragma unroll
            for (int u = 0; u < M / 4; u++) w[u] = reinterpret_cast<const uint32_t*>(cp)[u];
          }
        }
        if (dbg & 4) {  // compare what the stage held with the list in global memory
          const int rowc = t.row0 + i * PT_N + e;
          if (rowc < t.row0 + t.nrows) {
            const uint32_t* g = reinterpret_cast<const uint32_t*>(dir.codes[t.list] + (int64_t)rowc * M);
            bool bad = false;
#pragma unroll
            for (int u = 0; u < M / 4; u++) bad |= (w[u] != g[u]);
            if (bad) atomicAdd(&g_pqtc_dbg[0], 1ull);
          }
        }
        const float nstaged = HAS_NORM ? nstage[s * PT_N + e] : 0.f;
        mbar_arrive(&sh.code_empty[s]);
        // entries past the end of the segment decode whatever bytes the stage holds: finite values,
        // masked by ne = +inf.  Tombstones (gamma_index_ivfpq.h:930) and the docid predicate
        // (IsValid, :934-939) are resolved here, once per entry per 128 pairs.
        const int row = t.row0 + i * PT_N + e;
        bool valid = row < row_end;
        if (EAGER && valid) {
          const int64_t raw = lids[row];
          valid = raw >= 0 && ctx_is_valid(f.del_bits, f.filter_bits, (uint32_t)raw);
        }
        const int b = bn & 1;
        mbar_wait(&sh.b_empty[b], ((bn >> 1) & 1) ^ 1);
        if (dbg & 1) __nanosleep(3000);
        unsigned char* brow = b_buf + b * tile_stride + e * 16;
        const float nsum = nstaged;
#pragma unroll
        for (int m = 0; m < M; m++) {
          const uint32_t c = (w[m >> 2] >> (8 * (m & 3))) & 0xFFu;
          const unsigned char* src = cb + (m * PT_KSUB + c) * UNIT;
          if (DSUB == 4) {  // two sub-quantisers share a core-matrix row
            const uint2 val = *reinterpret_cast<const uint2*>(src);
            *reinterpret_cast<uint2*>(brow + (m >> 1) * 2048 + (m & 1) * 8) = val;
          } else {
#pragma unroll
            for (int u = 0; u < UNIT / 16; u++) {
              const uint4 val = reinterpret_cast<const uint4*>(src)[u];
              *reinterpret_cast<uint4*>(brow + (m * (UNIT / 16) + u) * 2048) = val;
            }
          }
        }
        // ne[b] is read by the epilogue of tile bn - 2: wait until it has released the buffer
        mbar_wait(&sh.acc_empty[b], ((bn >> 1) & 1) ^ 1);
        sh.ne[b][e] = valid ? nsum : INFINITY;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core proxy
        mbar_arrive(&sh.b_full[b]);
      }
    }
  } else if (warp == 8) {
    // ======================= TMA producer (whole warp loops, one elected lane issues) =======================
    uint32_t an = 0, cn = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const LmTile t = items[it];
      if (t.nrows <= 0) continue;  // every role skips empty segments the same way
      const int ab = an & 1;
      mbar_wait(&sh.a_empty[ab], ((an >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&sh.a_full[ab], TILE);
        bulk_g2s(a_buf + ab * tile_stride, a_scratch + (int64_t)t.grp * TILE, TILE, &sh.a_full[ab]);
      }
      __syncwarp();
      const unsigned char* lcodes = dir.codes[t.list];
      const float* lnorms = HAS_NORM ? pqnorm + pqnorm_off[t.list] : nullptr;  // list segments start 128-byte aligned
      const int ntiles = (t.nrows + PT_N - 1) / PT_N;
      for (int i = 0; i < ntiles; i++, cn++) {
        const int s = cn % PT_CS;
        mbar_wait(&sh.code_empty[s], ((cn / PT_CS) & 1) ^ 1);
        if (elect_one()) {
          const int n_e = min(PT_N, t.nrows - i * PT_N);
          const uint32_t bytes = ((uint32_t)n_e * M + 15u) & ~15u;
          const uint32_t nbytes = HAS_NORM ? ((uint32_t)n_e * 4u + 15u) & ~15u : 0u;
          mbar_arrive_expect_tx(&sh.code_full[s], bytes + nbytes);
          bulk_g2s(code_buf + s * CODE_STAGE, lcodes + (int64_t)(t.row0 + i * PT_N) * M, bytes, &sh.code_full[s]);
          if (HAS_NORM) bulk_g2s(nstage + s * PT_N, lnorms + (t.row0 + i * PT_N), nbytes, &sh.code_full[s]);
        }
        __syncwarp();
      }
      an++;
    }
  } else if (warp == 10) {
    // ======================= candidate drain: shared queue -> global per-query lists =======================
    const int lane = tid & 31;
    unsigned int t = 0;
    for (;;) {
      unsigned long long r[4];
      int n = 0;
      bool open = true;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        r[u] = *reinterpret_cast<volatile unsigned long long*>(&sh.q_rec[(t + (unsigned)(lane + 32 * u)) % PT_RQ]);
        const unsigned int m = __ballot_sync(0xffffffffu, r[u] != kRecEmpty);
        if (open) {  // records are consumed as a contiguous prefix, so the tail stays a single counter
          if (m == 0xffffffffu) n += 32;
          else n += __ffs(~m) - 1, open = false;
        }
      }
      if (n == 0) {
        if (*reinterpret_cast<volatile int*>(&sh.q_done) == 4 && *reinterpret_cast<volatile unsigned int*>(&sh.q_head) == t) break;
        __nanosleep(64);
        continue;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (lane + 32 * u < n) {
          const int q = (int)(r[u] >> 48), p = (int)((r[u] >> 32) & 0xFFFFu);
          const int slot = atomicAdd(cand_cnt + q, 1);  // a count above cap flags the query for the exact kernel
          if (slot < cap) cand[(int64_t)q * cap + slot] = ((unsigned long long)(uint32_t)p << 32) | (r[u] & 0xFFFFFFFFull);
          *reinterpret_cast<volatile unsigned long long*>(&sh.q_rec[(t + (unsigned)(lane + 32 * u)) % PT_RQ]) = kRecEmpty;
        }
      }
      __threadfence_block();
      __syncwarp();
      t += (unsigned int)n;
      if (lane == 0) *reinterpret_cast<volatile unsigned int*>(&sh.q_tail) = t;
    }
  } else {
    // ======================= MMA issuer (whole warp loops, one elected lane issues) =======================
    // instruction descriptor (cute::UMMA::InstrDescriptor): c = F32 (1 << 4), a = b = F16 (format 0 at [7, 10) and
    // [10, 13)), K-major A and B, N >> 3 at [17, 23), M >> 4 at [24, 29)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(PT_N >> 3) << 17) | ((uint32_t)(PT_M >> 4) << 24);
    const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
    uint32_t an = 0, bn = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const LmTile t = items[it];
      if (t.nrows <= 0) continue;
      const int ab = an & 1;
      mbar_wait(&sh.a_full[ab], (an >> 1) & 1);
      const int ntiles = (t.nrows + PT_N - 1) / PT_N;
      for (int i = 0; i < ntiles; i++, bn++) {
        const int b = bn & 1;
        mbar_wait(&sh.b_full[b], (bn >> 1) & 1);
        mbar_wait(&sh.acc_empty[b], ((bn >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t da = lt_desc(a_base + (uint32_t)(ab * tile_stride)), db = lt_desc(b_base + (uint32_t)(b * tile_stride));
          const uint32_t acc = tmem_d + (uint32_t)(b * PT_N);
#pragma unroll
          for (int kk = 0; kk < KC / 2; kk++) {  // K = 16 per instruction: two core matrices = 4096 bytes further
            const uint64_t adv = (uint64_t)((kk * 4096) >> 4);
            tc_mma_f16(acc, da + adv, db + adv, idesc, kk != 0);
          }
          tc_commit(&sh.b_empty[b]);
          tc_commit(&sh.acc_full[b]);
          if (i == ntiles - 1) tc_commit(&sh.a_empty[ab]);
        }
        __syncwarp();
      }
      an++;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * PT_N) : "memory");
  }
}

// ---- phase C: candidates -> reference arithmetic -> merge with phase A's keys ---------------------
constexpr int RS_NT = 256;

template <int METRIC>
__global__ void __launch_bounds__(RS_NT)
    pq_rescore_kernel(const float* __restrict__ ip_table, const int32_t* __restrict__ probe_ids,
                      const float* __restrict__ coarse_dis, int nprobe, ListDirectory dir, int M, const float* __restrict__ T,
                      const int* __restrict__ cand_cnt, const unsigned long long* __restrict__ cand, int cap,
                      const unsigned long long* __restrict__ keys_a, int64_t keys_a_stride, int kprime, int NP,
                      const int* __restrict__ row_limit, FilterArgs f, int sorted_out, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(rs_smem);  // [NP]
  float* ips = reinterpret_cast<float*>(buf + NP);                            // [M][256]
  const int q = blockIdx.x, tid = threadIdx.x;
  const int cnt = cand_cnt[q];
  if (cnt > cap) return;  // overflow: the exact kernel redoes this query (pq_fallback_merge_kernel writes out)
  if (cnt == 0 && !sorted_out) {  // nothing passed the filter: phase A's (unordered) keys are the answer -- the usual case
    for (int i = tid; i < kprime; i += RS_NT) out[(int64_t)q * kprime + i] = keys_a[(int64_t)q * keys_a_stride + i];
    return;
  }
  // a handful of candidates (the usual case) read the query's inner-product table straight from L2; staging the
  // 4 M bytes-per-sub-quantiser table in shared memory pays only for long candidate lists
  const float* ipq = ip_table + (int64_t)q * M * PT_KSUB;
  const bool staged = cnt > 64;
  if (staged) {
    const float4* ipq4 = reinterpret_cast<const float4*>(ipq);
    for (int i = tid; i < M * PT_KSUB / 4; i += RS_NT) reinterpret_cast<float4*>(ips)[i] = __ldg(ipq4 + i);
  }
  // CandQueue layout: buf[0, KP) = the current best (phase A's keys, sentinel padded), buf[KP, KP + cnt) = candidates;
  // its flush keeps the k' best -- a bitonic sort of the occupied prefix when short, an MSB radix select (linear in the
  // fill) when a loose bound let thousands of candidates through
  __shared__ int s_cnt;
  __shared__ unsigned long long s_tau;
  int KP = 16;
  while (KP < kprime) KP <<= 1;
  CandQueue cq{buf, &s_cnt, &s_tau, kprime, KP, NP};
  for (int i = tid; i < KP; i += RS_NT) buf[i] = i < kprime ? keys_a[(int64_t)q * keys_a_stride + i] : kKeySentinel;
  if (tid == 0) s_cnt = cnt, s_tau = kKeySentinel;
  __syncthreads();
  for (int i = tid; i < cnt; i += RS_NT) {
    const unsigned long long rec = cand[(int64_t)q * cap + i];
    const int p = (int)(rec >> 32);
    const uint32_t pos = (uint32_t)rec;
    const int l = probe_ids[(int64_t)q * nprobe + p];
    float dis = coarse_dis[(int64_t)q * nprobe + p];
    const unsigned char* code = dir.codes[l] + (int64_t)pos * M;
    const float* Tl = METRIC == kMetricL2 ? T + (int64_t)l * M * PT_KSUB : nullptr;
    // the reference's order: dis = dis0; dis += tab[m][code[m]], tab = T - 2 ip as ivfpq_scan_kernel builds it
    for (int m = 0; m < M; m++) {
      const int c = code[m];
      const float a = staged ? ips[m * PT_KSUB + c] : __ldg(ipq + m * PT_KSUB + c);
      dis += METRIC == kMetricL2 ? fmaf(-2.0f, a, __ldg(Tl + m * PT_KSUB + c)) : a;
    }
    const int64_t raw = dir.ids[l][pos];
    // positions below row_limit were scored exactly by phase A: they are in keys_a already (or lost there for good)
    const bool ok = (int)pos >= row_limit[(int64_t)q * nprobe + p] && raw >= 0 && ctx_is_valid(f.del_bits, f.filter_bits, (uint32_t)raw) && dis <= f.max_score && dis >= f.min_score;
    buf[KP + i] = ok ? make_key(score2ord<METRIC>(dis), (uint32_t)raw) : kKeySentinel;
  }
  __syncthreads();
  cq.flush(true);
  for (int i = tid; i < kprime; i += RS_NT) out[(int64_t)q * kprime + i] = buf[i];
}

// overflowed queries: merge the exact kernel's per-group partials (sorted runs of k') into out[q]
__global__ void __launch_bounds__(RS_NT)
    pq_fallback_merge_kernel(const int* __restrict__ cand_cnt, int cap, const unsigned long long* __restrict__ partial,
                             int ngroups, int kprime, int NP, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(rs_smem);
  const int q = blockIdx.x, tid = threadIdx.x;
  if (cand_cnt[q] <= cap) return;
  const int n = ngroups * kprime;
  for (int i = tid; i < NP; i += RS_NT) buf[i] = i < n ? partial[(int64_t)q * n + i] : kKeySentinel;
  __syncthreads();
  block_bitonic_sort(buf, NP);
  for (int i = tid; i < kprime; i += RS_NT) out[(int64_t)q * kprime + i] = buf[i];
}

template <int M, int DSUB>
cudaError_t launch_scan_shape(const unsigned char* a_scratch, const PairMeta* meta, const uint16_t* cb, const float* pqnorm,
                              const int64_t* pqnorm_off,
                              const LmTile* items, const int64_t* totals, ListDirectory dir, FilterArgs f, int metric,
                              int* cand_cnt, unsigned long long* cand, int cap, int grid, cudaStream_t st) {
  constexpr int D = M * DSUB;
  int tile_stride = (D / 8) * 2048, dbg = 0;
  if (const char* ev = getenv("GB_PQTC_DBG")) dbg = atoi(ev);
  if (dbg & 2) tile_stride = next_pow2(tile_stride);
  const size_t base = (size_t)M * PT_KSUB * DSUB * 2 + 4 * (size_t)tile_stride + (size_t)PT_CS * PT_N * M + 4096;
  const bool l2 = metric == kMetricL2;
  const bool eager = f.del_bits != nullptr || f.filter_bits != nullptr;
  if (l2 && (!pqnorm || !pqnorm_off)) return cudaErrorInvalidValue;
  const size_t smem = base + (l2 ? (size_t)PT_CS * PT_N * 4 : 0);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, PT_NT, smem, st>>>(a_scratch, meta, cb, pqnorm, pqnorm_off, items, totals, dir, f, cand_cnt, cand, cap, tile_stride, dbg);
    note_launch();
    return cudaGetLastError();
  };
  if (l2) return eager ? go(pqtc_scan_kernel<M, DSUB, true, true>) : go(pqtc_scan_kernel<M, DSUB, true, false>);
  return eager ? go(pqtc_scan_kernel<M, DSUB, false, true>) : go(pqtc_scan_kernel<M, DSUB, false, false>);
}

}  // namespace

// ---- |r_e|^2 of every list entry (L2): out[off[l] + pos] = sum_m nrm[m][code[m]] ------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
    pq_entry_norms_kernel(ListDirectory dir, int M, const float* __restrict__ nrm, const int64_t* __restrict__ off,
                          float* __restrict__ out) {
  extern __shared__ float s_nrm[];  // [M][256]
  const int l = blockIdx.y;
  const int len = dir.len[l];
  const int row0 = blockIdx.x * 1024;
  if (row0 >= len) return;
  for (int i = threadIdx.x; i < M * PT_KSUB; i += 256) s_nrm[i] = nrm[i];
  __syncthreads();
  const unsigned char* codes = dir.codes[l];
  float* dst = out + off[l];
  for (int r = row0 + threadIdx.x; r < min(len, row0 + 1024); r += 256) {
    const unsigned char* c = codes + (int64_t)r * M;
    float sum = 0.f;
    for (int m = 0; m < M; m++) sum += s_nrm[m * PT_KSUB + c[m]];
    dst[r] = sum;
  }
}
}  // namespace

cudaError_t launch_pq_entry_norms(ListDirectory dir, int nlist, int max_len, int M, const float* nrm, const int64_t* off,
                                  float* out, cudaStream_t st) {
  if (nlist <= 0 || max_len <= 0) return cudaSuccess;
  const size_t smem = (size_t)M * PT_KSUB * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pq_entry_norms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)((max_len + 1023) / 1024), (unsigned)nlist);
  pq_entry_norms_kernel<<<grid, 256, smem, st>>>(dir, M, nrm, off, out);
  note_launch();
  return cudaGetLastError();
}

void pqtc_debug_counters(unsigned long long out[4], bool reset) {
  cudaMemcpyFromSymbol(out, g_pqtc_dbg, sizeof(unsigned long long) * 4);
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_pqtc_dbg, z, sizeof(z));
  }
}

bool pqtc_supported(int M, int dsub) {
  // power-of-two shapes only.  (M = 12, dsub = 8 -- 1536-byte code tiles, 24 KiB operand tiles -- compiled and ran, but
  // intermittently (about 1 search in 100, on some boxes more) lost the rows of one decode warp of a tile; the cause
  // was not found in the time available, so that shape stays on the LUT kernel.)
  return (M == 16 && dsub == 8) || (M == 8 && dsub == 16) || (M == 8 && dsub == 8) || (M == 16 && dsub == 4) ||
         (M == 4 && dsub == 16) || (M == 4 && dsub == 32) || (M == 4 && dsub == 8) || (M == 8 && dsub == 4);
}

size_t pqtc_pair_meta_bytes() { return sizeof(PairMeta); }

cudaError_t launch_pqtc_tables(const float* pq, int M, int dsub, int metric, uint16_t* cb, float* nrm, float* rmax2,
                               cudaStream_t st) {
  pqtc_tables_kernel<<<1, 256, 0, st>>>(pq, M, dsub, metric == kMetricL2 ? -2.0f : -1.0f, cb, nrm, rmax2);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pqtc_plan_phase_a(const int32_t* probe_ids, int64_t npairs, int nprobe, int pa_max, long long target,
                                     const int* list_len, int32_t* probes_a, int32_t* probes_b, int* row_limit,
                                     cudaStream_t st) {
  if (npairs <= 0) return cudaSuccess;
  pqtc_plan_phase_a_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(probe_ids, npairs, nprobe, pa_max, target,
                                                                            list_len, probes_a, probes_b, row_limit);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pq_stage_pairs(const float* xq, int64_t ldq, int d, const float* coarse, int64_t ldc, const LmTile* items,
                                  int max_items, const int64_t* totals, const int64_t* pair_j, int nprobe,
                                  const float* coarse_dis, const unsigned long long* bound_keys, int64_t bound_stride,
                                  int kprime, const float* rmax2, FilterArgs f, int metric, float eps_scale,
                                  unsigned char* a_scratch, void* meta, int* cand_cnt, int cap, const int* row_limit,
                                  cudaStream_t st) {
  if (max_items <= 0) return cudaSuccess;
  if (d % 16) return cudaErrorInvalidValue;
  if (metric == kMetricL2)
    pq_stage_pairs_kernel<kMetricL2><<<max_items, PT_M, 0, st>>>(xq, ldq, d, coarse, ldc, items, totals, pair_j, nprobe,
                                                                coarse_dis, bound_keys, bound_stride, kprime, rmax2,
                                                                f.min_score, f.max_score, eps_scale, a_scratch,
                                                                static_cast<PairMeta*>(meta), cand_cnt, cap, row_limit);
  else
    pq_stage_pairs_kernel<kMetricIP><<<max_items, PT_M, 0, st>>>(xq, ldq, d, coarse, ldc, items, totals, pair_j, nprobe,
                                                                coarse_dis, bound_keys, bound_stride, kprime, rmax2,
                                                                f.min_score, f.max_score, eps_scale, a_scratch,
                                                                static_cast<PairMeta*>(meta), cand_cnt, cap, row_limit);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pqtc_scan(const unsigned char* a_scratch, const void* meta, const uint16_t* cb, const float* pqnorm,
                             const int64_t* pqnorm_off,
                             const LmTile* items, int max_items, const int64_t* totals, ListDirectory dir, int M, int dsub,
                             FilterArgs f, int metric, int* cand_cnt, unsigned long long* cand, int cap, int num_sms,
                             cudaStream_t st) {
  if (max_items <= 0) return cudaSuccess;
  const int grid = max_items < num_sms ? max_items : num_sms;
  const PairMeta* pm = static_cast<const PairMeta*>(meta);
#define GB_PT(MM, DS) \
  if (M == MM && dsub == DS) return launch_scan_shape<MM, DS>(a_scratch, pm, cb, pqnorm, pqnorm_off, items, totals, dir, f, metric, cand_cnt, cand, cap, grid, st)
  GB_PT(16, 8);
  GB_PT(8, 16);
  GB_PT(8, 8);
  GB_PT(16, 4);
  GB_PT(4, 16);
  GB_PT(4, 32);
  GB_PT(4, 8);
  GB_PT(8, 4);
#undef GB_PT
  return cudaErrorInvalidValue;
}

cudaError_t launch_pq_rescore(const float* ip_table, int nq, const int32_t* probe_ids, const float* coarse_dis, int nprobe,
                              ListDirectory dir, int M, const float* T, const int* cand_cnt, const unsigned long long* cand,
                              int cap, const unsigned long long* keys_a, int64_t keys_a_stride, int kprime, int metric,
                              const int* row_limit, FilterArgs f, bool sorted_out, unsigned long long* out, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const int NP = next_pow2(next_pow2(kprime < 16 ? 16 : kprime) + cap);  // CandQueue: KP best + cap candidates
  const size_t smem = (size_t)NP * 8 + (size_t)M * PT_KSUB * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e;
  if (metric == kMetricL2) {
    e = cudaFuncSetAttribute(pq_rescore_kernel<kMetricL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pq_rescore_kernel<kMetricL2><<<nq, RS_NT, smem, st>>>(ip_table, probe_ids, coarse_dis, nprobe, dir, M, T, cand_cnt, cand,
                                                         cap, keys_a, keys_a_stride, kprime, NP, row_limit, f, sorted_out ? 1 : 0, out);
  } else {
    e = cudaFuncSetAttribute(pq_rescore_kernel<kMetricIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pq_rescore_kernel<kMetricIP><<<nq, RS_NT, smem, st>>>(ip_table, probe_ids, coarse_dis, nprobe, dir, M, T, cand_cnt, cand,
                                                         cap, keys_a, keys_a_stride, kprime, NP, row_limit, f, sorted_out ? 1 : 0, out);
  }
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pq_fallback_merge(const int* cand_cnt, int cap, int nq, const unsigned long long* partial, int ngroups,
                                     int kprime, unsigned long long* out, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const int NP = next_pow2(ngroups * kprime < 16 ? 16 : ngroups * kprime);
  const size_t smem = (size_t)NP * 8;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(pq_fallback_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  pq_fallback_merge_kernel<<<nq, RS_NT, smem, st>>>(cand_cnt, cap, partial, ngroups, kprime, NP, out);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gb
