/*
 * cpu_gamma.c -- the CPU arm of bench.py: gamma's CPU search path restated for SPEED on the
 * GPU box's host cores.  The real engine cannot be built
 * here (faiss v1.14.1, rocksdb, tbb, protobuf, openblas are absent), so this file
 * plays its role in the "x CPU" ratio.  It is NOT the correctness checker (that is oracle/, built
 * with -ffp-contract=off) and nothing in the product path links it.
 *
 * What it keeps from the reference, loop for loop:
 *   - OpenMP over queries, schedule(dynamic), one scanner + one heap per thread
 *     (gamma_index_ivfflat.cc:695-733, gamma_index_ivfpq.cc:811-848, gamma_index_flat.cc:286-302);
 *   - the ADC inner loop is gamma's own scalar loop  dis = dis0; dis += tab[m][code[m]]
 *     (gamma_index_ivfpq.h:923-953) -- gamma does not use a SIMD fast-scan here;
 *   - LUT = T[list] - 2 ip(x) per (query, list) (fvec_madd, gamma_index_ivfpq.h:254-262);
 *   - exact re-rank of recall_num candidates from the raw vectors (gamma_index_ivfpq.cc:675-726);
 *   - filters (tombstone bit, docid bitmaps, score window) before the heap.
 * What it does the way faiss does it, i.e. as fast as the host allows:
 *   - fvec_L2sqr / fvec_inner_product / fvec_madd with FMA, AVX-512 when the CPU has it (runtime
 *     dispatch through target_clones: the .so is built in a container and runs on another box);
 *   - coarse quantiser for a batch = blocked sgemm (|x|^2 + |c|^2 - 2 x.c, the IndexFlat path for
 *     >= 20 queries) + per-row heap;
 *   - the per-query inner-product table as a small sgemm.
 * Results: the ids it returns are checked against oracle/ by tests/test_cpu_baseline.py (same
 * candidates; scores may differ in the last bits because of FMA contraction).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t idx_t;
#define CG_IP 0
#define CG_L2 1
#define CG_DEL_MASK ((idx_t)1 << 63)
#define CG_KSUB 256
#define CG_SIMD __attribute__((target_clones("avx512f", "avx2,fma", "default")))

/* ---- heaps (faiss CMax / CMin with id tie-break, restated as in oracle/gamma_oracle.c) ---------- */
static inline int h_cmp(int is_max, float a, float b) { return is_max ? (a > b) : (a < b); }
static inline int h_cmp2(int is_max, float a1, float b1, idx_t a2, idx_t b2) {
  return is_max ? ((a1 > b1) || (a1 == b1 && a2 > b2)) : ((a1 < b1) || (a1 == b1 && a2 < b2));
}
static inline float h_neutral(int is_max) { return is_max ? FLT_MAX : -FLT_MAX; }
static void heap_init(int is_max, int k, float *v, idx_t *ids) {
  for (int i = 0; i < k; i++) v[i] = h_neutral(is_max), ids[i] = -1;
}
static inline void heap_replace_top(int is_max, int k, float *bh_val, idx_t *bh_ids, float val, idx_t id) {
  bh_val--, bh_ids--;
  size_t i = 1, i1, i2;
  for (;;) {
    i1 = i << 1, i2 = i1 + 1;
    if (i1 > (size_t)k) break;
    if (i2 == (size_t)k + 1 || h_cmp2(is_max, bh_val[i1], bh_val[i2], bh_ids[i1], bh_ids[i2])) {
      if (h_cmp2(is_max, val, bh_val[i1], id, bh_ids[i1])) break;
      bh_val[i] = bh_val[i1], bh_ids[i] = bh_ids[i1], i = i1;
    } else {
      if (h_cmp2(is_max, val, bh_val[i2], id, bh_ids[i2])) break;
      bh_val[i] = bh_val[i2], bh_ids[i] = bh_ids[i2], i = i2;
    }
  }
  bh_val[i] = val, bh_ids[i] = id;
}
/* best first; empty slots (id -1) to the tail.  Same output order as faiss heap_reorder: L2 ascending
 * (score, id); IP descending score, larger id first among equal scores. */
typedef struct {
  float v;
  idx_t id;
} cg_pair;
static int cmp_l2(const void *a, const void *b) {
  const cg_pair *x = (const cg_pair *)a, *y = (const cg_pair *)b;
  if (x->v != y->v) return x->v < y->v ? -1 : 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}
static int cmp_ip(const void *a, const void *b) {
  const cg_pair *x = (const cg_pair *)a, *y = (const cg_pair *)b;
  if (x->v != y->v) return x->v > y->v ? -1 : 1;
  return x->id > y->id ? -1 : (x->id < y->id);
}
static void heap_reorder(int is_max, int k, float *v, idx_t *ids) {
  cg_pair stack_buf[512];
  cg_pair *p = k <= 512 ? stack_buf : (cg_pair *)malloc(sizeof(cg_pair) * k);
  int n = 0;
  for (int i = 0; i < k; i++)
    if (ids[i] != -1) p[n].v = v[i], p[n].id = ids[i], n++;
  qsort(p, n, sizeof(cg_pair), is_max ? cmp_l2 : cmp_ip);
  for (int i = 0; i < n; i++) v[i] = p[i].v, ids[i] = p[i].id;
  for (int i = n; i < k; i++) v[i] = h_neutral(is_max), ids[i] = -1;
  if (p != stack_buf) free(p);
}

/* ---- SIMD primitives ------------------------------------------------------------------------- */
CG_SIMD static float fvec_l2sqr(const float *x, const float *y, int d) {
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int i = 0;
  float acc[16] = {0};
  for (; i + 16 <= d; i += 16)
    for (int j = 0; j < 16; j++) {
      const float t = x[i + j] - y[i + j];
      acc[j] += t * t;
    }
  for (; i < d; i++) {
    const float t = x[i] - y[i];
    a0 += t * t;
  }
  for (int j = 0; j < 16; j += 4) a0 += acc[j], a1 += acc[j + 1], a2 += acc[j + 2], a3 += acc[j + 3];
  return (a0 + a1) + (a2 + a3);
}
CG_SIMD static float fvec_ip(const float *x, const float *y, int d) {
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  float acc[16] = {0};
  int i = 0;
  for (; i + 16 <= d; i += 16)
    for (int j = 0; j < 16; j++) acc[j] += x[i + j] * y[i + j];
  for (; i < d; i++) a0 += x[i] * y[i];
  for (int j = 0; j < 16; j += 4) a0 += acc[j], a1 += acc[j + 1], a2 += acc[j + 2], a3 += acc[j + 3];
  return (a0 + a1) + (a2 + a3);
}
/* c = a + bf * b  (faiss fvec_madd) */
CG_SIMD static void fvec_madd(int n, const float *a, float bf, const float *b, float *c) {
  for (int i = 0; i < n; i++) c[i] = a[i] + bf * b[i];
}
/* C[i][j] = sum_k A[i][k] * Bt[k][j]   (A: m x d row-major; Bt: d x n, i.e. the second operand packed
 * K-major once and reused for every row block).  Register blocking 4 rows x 32 columns: the j loop is the
 * vector dimension (two zmm / four ymm accumulators per row), a[i][k] is broadcast. */
#define CG_NB 32
CG_SIMD static void sgemm_packed(const float *A, int m, int d, const float *Bt, int n, float *C, int ldc) {
  for (int i0 = 0; i0 < m; i0 += 4) {
    const int mr = m - i0 < 4 ? m - i0 : 4;
    const float *a0 = A + (size_t)i0 * d, *a1 = a0 + (mr > 1 ? d : 0), *a2 = a0 + (mr > 2 ? 2 * d : 0),
                *a3 = a0 + (mr > 3 ? 3 * d : 0);
    for (int j0 = 0; j0 < n; j0 += CG_NB) {
      const int nb = n - j0 < CG_NB ? n - j0 : CG_NB;
      float c0[CG_NB] = {0}, c1[CG_NB] = {0}, c2[CG_NB] = {0}, c3[CG_NB] = {0};
      if (nb == CG_NB) {
        for (int k = 0; k < d; k++) {
          const float *b = Bt + (size_t)k * n + j0;
          const float x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
          for (int j = 0; j < CG_NB; j++) c0[j] += x0 * b[j], c1[j] += x1 * b[j], c2[j] += x2 * b[j], c3[j] += x3 * b[j];
        }
      } else {
        for (int k = 0; k < d; k++) {
          const float *b = Bt + (size_t)k * n + j0;
          const float x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
          for (int j = 0; j < nb; j++) c0[j] += x0 * b[j], c1[j] += x1 * b[j], c2[j] += x2 * b[j], c3[j] += x3 * b[j];
        }
      }
      for (int j = 0; j < nb; j++) {
        C[(size_t)i0 * ldc + j0 + j] = c0[j];
        if (mr > 1) C[(size_t)(i0 + 1) * ldc + j0 + j] = c1[j];
        if (mr > 2) C[(size_t)(i0 + 2) * ldc + j0 + j] = c2[j];
        if (mr > 3) C[(size_t)(i0 + 3) * ldc + j0 + j] = c3[j];
      }
    }
  }
}
/* Bt[k][j] = B[j][k] */
static float *pack_transposed(const float *B, int n, int d) {
  float *Bt = (float *)aligned_alloc(64, ((sizeof(float) * (size_t)n * d + 63) / 64) * 64);
  for (int j = 0; j < n; j++)
    for (int k = 0; k < d; k++) Bt[(size_t)k * n + j] = B[(size_t)j * d + k];
  return Bt;
}

/* ---- predicate --------------------------------------------------------------------------------- */
typedef struct {
  const uint8_t *del_bitmap, *filter_bitmap;
  float min_score, max_score;
} cg_ctx;
static inline int bit_test(const uint8_t *bm, idx_t id) { return (bm[id >> 3] >> (id & 7)) & 1; }
static inline int ctx_is_valid(const cg_ctx *c, idx_t id) {
  if (c->filter_bitmap && !bit_test(c->filter_bitmap, id)) return 0;
  if (c->del_bitmap && bit_test(c->del_bitmap, id)) return 0;
  return 1;
}
static inline int ctx_score_ok(const cg_ctx *c, float s) { return s <= c->max_score && s >= c->min_score; }

/* ---- K1 FLAT (gamma_index_flat.cc:224-302) --------------------------------------------------------- */
int cg_flat_search(const float *db, int64_t ld, int64_t n, int d, const float *xq, int nq, int k, int metric,
                   const uint8_t *del_bitmap, const uint8_t *filter_bitmap, float min_score, float max_score,
                   float *out_dis, int64_t *out_ids) {
  cg_ctx ctx = {del_bitmap, filter_bitmap, min_score, max_score};
  const int is_max = metric == CG_L2;
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < nq; i++) {
    const float *xi = xq + (size_t)i * d;
    float *simi = out_dis + (size_t)i * k;
    idx_t *idxi = out_ids + (size_t)i * k;
    heap_init(is_max, k, simi, idxi);
    for (int64_t vid = 0; vid < n; vid++) {
      if (!ctx_is_valid(&ctx, vid)) continue;
      const float *y = db + vid * ld;
      const float dis = is_max ? fvec_l2sqr(xi, y, d) : fvec_ip(xi, y, d);
      if (ctx_score_ok(&ctx, dis) && h_cmp(is_max, simi[0], dis)) heap_replace_top(is_max, k, simi, idxi, dis, vid);
    }
    heap_reorder(is_max, k, simi, idxi);
  }
  return 0;
}

/* ---- K2 coarse quantiser: IndexFlat::search, the blocked-sgemm path faiss takes for >= 20 queries -------- */
int cg_coarse_search(const float *cent, int nlist, int d, const float *xq, int nq, int nprobe, int metric, float *out_dis,
                     int64_t *out_ids) {
  const int is_max = metric == CG_L2;
  float *cn = (float *)malloc(sizeof(float) * nlist);
  for (int c = 0; c < nlist; c++) cn[c] = fvec_ip(cent + (size_t)c * d, cent + (size_t)c * d, d);
  float *cent_t = pack_transposed(cent, nlist, d);
  const int QB = 32;
#pragma omp parallel
  {
    float *ipb = (float *)malloc(sizeof(float) * QB * nlist);
#pragma omp for schedule(dynamic)
    for (int q0 = 0; q0 < nq; q0 += QB) {
      const int qb = nq - q0 < QB ? nq - q0 : QB;
      sgemm_packed(xq + (size_t)q0 * d, qb, d, cent_t, nlist, ipb, nlist);
      for (int i = 0; i < qb; i++) {
        const float *xi = xq + (size_t)(q0 + i) * d;
        float *simi = out_dis + (size_t)(q0 + i) * nprobe;
        idx_t *idxi = out_ids + (size_t)(q0 + i) * nprobe;
        heap_init(is_max, nprobe, simi, idxi);
        const float xn = is_max ? fvec_ip(xi, xi, d) : 0.f;
        const float *ip = ipb + (size_t)i * nlist;
        for (int c = 0; c < nlist; c++) {
          float dis = ip[c];
          if (is_max) {
            dis = xn + cn[c] - 2.f * dis;
            if (dis < 0) dis = 0;
          }
          if (h_cmp(is_max, simi[0], dis)) heap_replace_top(is_max, nprobe, simi, idxi, dis, c);
        }
        heap_reorder(is_max, nprobe, simi, idxi);
      }
    }
    free(ipb);
  }
  free(cn);
  free(cent_t);
  return 0;
}

/* ---- K3 IVF-Flat (gamma_index_ivfflat.cc:695-733 + gamma_index_ivfflat.h:63-91) ------------------------ */
int cg_ivfflat_search_preassigned(const int64_t *list_off, const float *list_vecs, const int64_t *list_ids, int nlist, int d,
                                  const float *xq, int nq, int k, const int64_t *keys, int nprobe, int metric,
                                  const uint8_t *del_bitmap, const uint8_t *filter_bitmap, float min_score, float max_score,
                                  float *out_dis, int64_t *out_ids) {
  cg_ctx ctx = {del_bitmap, filter_bitmap, min_score, max_score};
  const int is_max = metric == CG_L2;
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < nq; i++) {
    const float *xi = xq + (size_t)i * d;
    float *simi = out_dis + (size_t)i * k;
    idx_t *idxi = out_ids + (size_t)i * k;
    heap_init(is_max, k, simi, idxi);
    for (int ik = 0; ik < nprobe; ik++) {
      const idx_t key = keys[(size_t)i * nprobe + ik];
      if (key < 0 || key >= nlist) continue;
      for (int64_t j = list_off[key]; j < list_off[key + 1]; j++) {
        if (list_ids[j] & CG_DEL_MASK) continue;
        if (!ctx_is_valid(&ctx, list_ids[j])) continue;
        const float *y = list_vecs + (size_t)j * d;
        const float dis = is_max ? fvec_l2sqr(xi, y, d) : fvec_ip(xi, y, d);
        if (ctx_score_ok(&ctx, dis) && h_cmp(is_max, simi[0], dis)) heap_replace_top(is_max, k, simi, idxi, dis, list_ids[j]);
      }
    }
    heap_reorder(is_max, k, simi, idxi);
  }
  return 0;
}

/* ---- K4 + K5 + K5r IVF-PQ (gamma_index_ivfpq.cc:730-947, gamma_index_ivfpq.h:154-309, 923-953) ---------- */
int cg_ivfpq_search_preassigned(const int64_t *list_off, const uint8_t *list_codes, const int64_t *list_ids, int nlist, int d,
                                int M, const float *coarse, const float *pq, const float *T, const float *xq, int nq, int k,
                                const int64_t *keys, const float *coarse_dis, int nprobe, int metric, int recall_num_param,
                                const float *raw, int64_t raw_ld, const uint8_t *del_bitmap, const uint8_t *filter_bitmap,
                                float min_score, float max_score, float *out_dis, int64_t *out_ids) {
  const int dsub = d / M;
  cg_ctx ctx = {del_bitmap, filter_bitmap, min_score, max_score};
  const int is_max = metric == CG_L2;
  if (is_max && !T) return -2; /* the bench always runs the precomputed-table mode gamma ends up in */
  int recall_num = k;
  const int rerank = recall_num_param > 0;
  if (recall_num_param > k) recall_num = recall_num_param;
  if (rerank && !raw) return -1;
  float *pqt = (float *)aligned_alloc(64, sizeof(float) * (size_t)M * CG_KSUB * dsub);  /* [m][dsub][256] */
  for (int m = 0; m < M; m++)
    for (int c = 0; c < CG_KSUB; c++)
      for (int j = 0; j < dsub; j++) pqt[((size_t)m * dsub + j) * CG_KSUB + c] = pq[((size_t)m * CG_KSUB + c) * dsub + j];
#pragma omp parallel
  {
    float *tab = (float *)aligned_alloc(64, sizeof(float) * M * CG_KSUB);
    float *ipt = (float *)aligned_alloc(64, sizeof(float) * M * CG_KSUB);
    float *rsimi = (float *)malloc(sizeof(float) * recall_num);
    idx_t *ridx = (idx_t *)malloc(sizeof(idx_t) * recall_num);
#pragma omp for schedule(dynamic)
    for (int i = 0; i < nq; i++) {
      const float *xi = xq + (size_t)i * d;
      float *simi = out_dis + (size_t)i * k;
      idx_t *idxi = out_ids + (size_t)i * k;
      heap_init(is_max, k, simi, idxi);
      float *rs = rerank ? rsimi : simi;
      idx_t *ri = rerank ? ridx : idxi;
      if (rerank) heap_init(is_max, recall_num, rs, ri);
      /* init_query: inner-product table, one 256 x dsub panel per sub-quantiser */
      for (int m = 0; m < M; m++) sgemm_packed(xi + m * dsub, 1, dsub, pqt + (size_t)m * dsub * CG_KSUB, CG_KSUB, ipt + m * CG_KSUB, CG_KSUB);
      for (int ik = 0; ik < nprobe; ik++) {
        const idx_t key = keys[(size_t)i * nprobe + ik];
        if (key < 0 || key >= nlist) continue;
        const int64_t b = list_off[key], e = list_off[key + 1];
        if (b == e) continue;
        float dis0;
        const float *lut;
        if (is_max) {
          dis0 = coarse_dis[(size_t)i * nprobe + ik];
          fvec_madd(M * CG_KSUB, T + (size_t)key * M * CG_KSUB, -2.0f, ipt, tab);
          lut = tab;
        } else {
          dis0 = fvec_ip(xi, coarse + (size_t)key * d, d);
          lut = ipt;
        }
        const uint8_t *codes = list_codes + (size_t)b * M;
        for (int64_t j = b; j < e; j++, codes += M) { /* scan_list_with_table: gamma's own scalar loop */
          if (list_ids[j] & CG_DEL_MASK) continue;
          if (!ctx_is_valid(&ctx, list_ids[j])) continue;
          float dis = dis0;
          const float *t = lut;
          for (int m = 0; m < M; m++, t += CG_KSUB) dis += t[codes[m]];
          if (ctx_score_ok(&ctx, dis) && h_cmp(is_max, rs[0], dis)) heap_replace_top(is_max, recall_num, rs, ri, dis, list_ids[j]);
        }
      }
      if (rerank) { /* compute_dis: exact distances of the ADC candidates from the raw vectors */
        for (int j = 0; j < recall_num; j++) {
          if (ri[j] < 0) continue;
          const float *v = raw + (size_t)ri[j] * raw_ld;
          const float dis = is_max ? fvec_l2sqr(xi, v, d) : fvec_ip(xi, v, d);
          if (ctx_score_ok(&ctx, dis) && h_cmp(is_max, simi[0], dis)) heap_replace_top(is_max, k, simi, idxi, dis, ri[j]);
        }
        heap_reorder(is_max, k, simi, idxi);
      } else {
        heap_reorder(is_max, recall_num, rs, ri);
      }
    }
    free(tab), free(ipt), free(rsimi), free(ridx);
  }
  free(pqt);
  return 0;
}

/* threads the loops above will use, and the cores the cgroup actually grants (cpu.max quota) */
int cg_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void cg_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
const char *cg_isa(void) {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return "avx512f";
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return "avx2+fma";
  return "scalar";
}
