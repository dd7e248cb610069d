// K4 / K5 / K5r / K8(encode): IVF-PQ look-up tables, asymmetric-distance scan, exact re-rank,
// PQ encoding.
//
// Replaces, in the reference:
//   QueryTables::init_query_{L2,IP} / precompute_list_tables_{L2,IP}
//       (index/impl/gamma_index_ivfpq.h:154-175, 223-309)             -> pq_ip_table + in-CTA LUT
//   faiss IndexIVFPQ::precompute_table (called at gamma_index_ivfpq.cc:1099)  -> pq_precompute_table
//   GammaIVFPQScanner::scan_list_with_table (gamma_index_ivfpq.h:923-953),
//       scan_one_list / search_preassigned (gamma_index_ivfpq.cc:635-673, 730-947) -> ivfpq_scan
//   compute_dis exact re-rank (gamma_index_ivfpq.cc:675-726)            -> rerank
//   pq.compute_codes on residuals in GammaIVFPQIndex::Add (gamma_index_ivfpq.cc:478-494) -> pq_encode
//
// Arithmetic order is the reference's: tab = T[list] + (-2)*ip ; dis = dis0 ; dis += tab[m][code[m]]
// for m = 0..M-1 (one fp32 add at a time).  Table kernels use unfused mul/add so that tables are
// bit-identical to a scalar CPU evaluation; given the same (keys, coarse_dis) the ADC distances
// are bit-equal to the oracle's.
//
// Algorithmic bytes of the scan: (M + 8) per entry + M*256*4 per (query, list) for the T row.
#include <float.h>

#include "common.cuh"
#include "kernels.h"

namespace gb {

namespace {

constexpr int KSUB = 256;

// ---------------- per-query inner-product table ------------------------------------------
constexpr int IPT_QB = 32;

__global__ void __launch_bounds__(KSUB)
    pq_ip_table_kernel(const float* __restrict__ xq, int64_t ldq, int nq, const float* __restrict__ pq, int M, int dsub,
                       float* __restrict__ ip) {
  extern __shared__ float sm[];
  float* ps = sm;                // [dsub][256]
  float* xs = sm + dsub * KSUB;  // [IPT_QB][dsub]
  const int m = blockIdx.x;
  const int q0 = blockIdx.y * IPT_QB;
  const int c = threadIdx.x;
  for (int idx = threadIdx.x; idx < dsub * KSUB; idx += KSUB) {
    int cc = idx / dsub, j = idx - cc * dsub;
    ps[j * KSUB + cc] = pq[((int64_t)m * KSUB) * dsub + idx];
  }
  for (int idx = threadIdx.x; idx < IPT_QB * dsub; idx += KSUB) {
    int qq = idx / dsub, j = idx - qq * dsub;
    xs[idx] = (q0 + qq < nq) ? xq[(int64_t)(q0 + qq) * ldq + m * dsub + j] : 0.f;
  }
  __syncthreads();
  for (int qq = 0; qq < IPT_QB && q0 + qq < nq; qq++) {
    float acc = 0.f;
    for (int j = 0; j < dsub; j++) acc = __fadd_rn(acc, __fmul_rn(xs[qq * dsub + j], ps[j * KSUB + c]));
    ip[((int64_t)(q0 + qq) * M + m) * KSUB + c] = acc;
  }
}

// ---------------- precomputed table T[l][m][c] --------------------------------------------
__global__ void __launch_bounds__(KSUB)
    pq_precompute_table_kernel(const float* __restrict__ coarse, int64_t ldc, const float* __restrict__ pq, int M,
                               int dsub, float* __restrict__ T) {
  const int l = blockIdx.x, c = threadIdx.x;
  for (int m = 0; m < M; m++) {
    const float* cl = coarse + (int64_t)l * ldc + m * dsub;
    const float* p = pq + ((int64_t)m * KSUB + c) * dsub;
    float nrm = 0.f, ipv = 0.f;
    for (int j = 0; j < dsub; j++) {
      float pj = p[j];
      nrm = __fadd_rn(nrm, __fmul_rn(pj, pj));
      ipv = __fadd_rn(ipv, __fmul_rn(__ldg(cl + j), pj));
    }
    T[((int64_t)l * M + m) * KSUB + c] = __fadd_rn(nrm, __fmul_rn(2.0f, ipv));
  }
}

// ---------------- ADC scan ------------------------------------------------------------------
constexpr int PQ_NT = 256;
constexpr int PQ_NST = 2;
constexpr int PQ_MAX_PG = 32;  // probes per CTA

// entries per thread per tile for the vector-load code layouts (MW = M / 4 code words per entry): per-tile
// bookkeeping (barrier, tau reload, fill estimate, refill) is paid once per PT entries, while LUT + queue +
// two code stages must stay within ~52 KiB so that four CTAs share an SM (M = 16: 2 x 256 entries, 8 KiB;
// 3 x 256 and 4 x 256 entries per tile measured slower: 20.2 / 21.1 ms against 19.7 ms)
__host__ __device__ constexpr int pq_pt(int MW) { return MW == 2 ? 4 : (MW == 4 ? 2 : (MW == 8 ? 2 : 1)); }

__host__ __device__ inline int pq_tile_entries(int M) {
  if (M == 8 || M == 16 || M == 32 || M == 64) return pq_pt(M / 4) * PQ_NT;
  int e = 16384 / M;
  e = (e / PQ_NT) * PQ_NT;
  if (e < PQ_NT) e = PQ_NT;
  if (e > 1024) e = 1024;
  return e;
}

template <int MW>
__device__ __forceinline__ float adc_distance(const unsigned char* __restrict__ code, const float* __restrict__ lut,
                                              float dis0, int M) {
  float dis = dis0;
  if (MW > 0) {
    uint32_t w[MW > 0 ? MW : 1];
    if (MW % 4 == 0) {
#pragma unroll
      for (int i = 0; i < MW / 4; i++) {
        uint4 v = reinterpret_cast<const uint4*>(code)[i];
        w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
      }
    } else if (MW % 2 == 0) {
#pragma unroll
      for (int i = 0; i < MW / 2; i++) {
        uint2 v = reinterpret_cast<const uint2*>(code)[i];
        w[2 * i] = v.x, w[2 * i + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < MW; i++) w[i] = reinterpret_cast<const uint32_t*>(code)[i];
    }
#pragma unroll
    for (int i = 0; i < MW; i++) {
      uint32_t x = w[i];
      dis += lut[(4 * i + 0) * KSUB + (x & 0xffu)];
      dis += lut[(4 * i + 1) * KSUB + ((x >> 8) & 0xffu)];
      dis += lut[(4 * i + 2) * KSUB + ((x >> 16) & 0xffu)];
      dis += lut[(4 * i + 3) * KSUB + (x >> 24)];
    }
  } else {
    for (int m = 0; m < M; m++) dis += lut[m * KSUB + code[m]];
  }
  return dis;
}

template <int METRIC, int MW>
__global__ void __launch_bounds__(PQ_NT)
    ivfpq_scan_kernel(const float* __restrict__ ip_table, const int32_t* __restrict__ probe_ids,
                      const float* __restrict__ coarse_dis, int nprobe, int ld_probe, int pg, ListDirectory dir, int M,
                      const float* __restrict__ T, int tile_e, int k, int KP, int SORTN, FilterArgs f,
                      const int* __restrict__ gate_cnt, int gate_cap, const int* __restrict__ row_limit, int sorted_out,
                      unsigned long long* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tile_bytes = tile_e * M;
  // the LUT sits at the start of the dynamic window: its shared address is a link-time constant, so
  // every gather is  LDS [byte_offset + imm]  with no base add
  float* lut = reinterpret_cast<float*>(smem_raw);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(lut + M * KSUB);
  unsigned char* stages = reinterpret_cast<unsigned char*>(buf + SORTN);
  __shared__ __align__(8) uint64_t full_bar[PQ_NST];
  __shared__ int s_cnt;
  __shared__ unsigned long long s_tau;
  __shared__ int g_list[PQ_MAX_PG], g_len[PQ_MAX_PG], g_tile0[PQ_MAX_PG + 1];
  __shared__ float g_dis0[PQ_MAX_PG];

  const int tid = threadIdx.x;
  const int q = blockIdx.y, grp = blockIdx.x;
  // flag-gated launch (fallback of the tensor-core filter, kernels_pqtc.cu): only queries whose
  // candidate list overflowed are scanned
  if (gate_cnt && gate_cnt[q] <= gate_cap) return;
  unsigned long long* out = partial + ((int64_t)q * gridDim.x + grp) * k;
  const int p0 = grp * pg;
  const int np = min(pg, nprobe - p0);

  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < np; i++) {
      int l = probe_ids[(int64_t)q * ld_probe + p0 + i];
      int len = (l >= 0 && l < dir.nlist) ? dir.len[l] : 0;  // key < 0: "not enough centroids" (ivfpq.cc:640)
      // phase A of the tensor-core filter scans only a prefix of its lists (kernels_pqtc.cu)
      if (row_limit) len = min(len, row_limit[(int64_t)q * ld_probe + p0 + i]);
      g_list[i] = l;
      g_len[i] = len;
      g_dis0[i] = coarse_dis[(int64_t)q * ld_probe + p0 + i];
      g_tile0[i] = acc;
      acc += (len + tile_e - 1) / tile_e;
    }
    g_tile0[np > 0 ? np : 0] = acc;
    for (int s = 0; s < PQ_NST; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  CandQueue cq{buf, &s_cnt, &s_tau, k, KP, SORTN};
  cq.init();  // __syncthreads inside: g_* and barriers visible
  const int total_tiles = np > 0 ? g_tile0[np] : 0;
  if (total_tiles == 0) {
    for (int i = tid; i < k; i += PQ_NT) out[i] = kKeySentinel;
    return;
  }

  // per-query inner-product table (L2-resident): IP => it IS the LUT (dis0 = <x, centroid>);
  // L2 => combined with the list's precomputed row when a new list starts
  const float4* ipq4 = reinterpret_cast<const float4*>(ip_table + (int64_t)q * M * KSUB);
  if (METRIC != kMetricL2) {
    float4* dst = reinterpret_cast<float4*>(lut);
    for (int i = tid; i < M * KSUB / 4; i += PQ_NT) dst[i] = __ldg(ipq4 + i);
  }

  // producer cursor (thread 0 only)
  int pr_pi = 0;
  auto issue = [&](int gt) {
    while (gt >= g_tile0[pr_pi + 1]) pr_pi++;
    int ti = gt - g_tile0[pr_pi];
    int n_e = min(tile_e, g_len[pr_pi] - ti * tile_e);
    uint32_t bytes = ((uint32_t)n_e * M + 15u) & ~15u;
    int s = gt % PQ_NST;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(stages + (size_t)s * tile_bytes, dir.codes[g_list[pr_pi]] + (int64_t)ti * tile_e * M, bytes,
             &full_bar[s]);
  };
  if (tid == 0)
    for (int gt = 0; gt < PQ_NST && gt < total_tiles; gt++) issue(gt);

  int pi = -1;  // consumer cursor
  float dis0 = 0.f;
  const int64_t* __restrict__ lids = nullptr;
  // entries per thread per tile: compile-time for the vector-load code layouts (gives the
  // compiler PT independent ADC chains to interleave), runtime for the generic layout
  constexpr int PT = pq_pt(MW);
  const int per_thread = MW > 0 ? PT : tile_e / PQ_NT;
  int est = 0;  // upper bound of the queue fill, identical in every thread

  for (int gt = 0; gt < total_tiles; gt++) {
    int npi = pi < 0 ? 0 : pi;
    while (gt >= g_tile0[npi + 1]) npi++;
    if (npi != pi) {  // first tile of a new list: build its LUT (precompute_list_tables)
      pi = npi;
      dis0 = g_dis0[pi];
      lids = dir.ids[g_list[pi]];
      if (METRIC == kMetricL2) {
        const float4* Tl = reinterpret_cast<const float4*>(T + (int64_t)g_list[pi] * M * KSUB);
        float4* lut4 = reinterpret_cast<float4*>(lut);
        for (int i = tid; i < M * KSUB / 4; i += PQ_NT) {
          float4 t = __ldg(Tl + i), a = __ldg(ipq4 + i);
          lut4[i] = make_float4(fmaf(-2.0f, a.x, t.x), fmaf(-2.0f, a.y, t.y), fmaf(-2.0f, a.z, t.z),
                                fmaf(-2.0f, a.w, t.w));
        }
      }
      __syncthreads();
    }
    const int ti = gt - g_tile0[pi];
    const int n_e = min(tile_e, g_len[pi] - ti * tile_e);
    const int s = gt % PQ_NST;
    mbar_wait(&full_bar[s], (gt / PQ_NST) & 1);
    const unsigned long long tau = s_tau;
    // score window and the queue's bound folded into one float interval: two compares per entry, the
    // order-preserving key is only built for entries inside it
    const float tb = key_bound<METRIC>(tau);
    const float lo_b = METRIC == kMetricL2 ? f.min_score : fmaxf(tb, f.min_score);
    const float hi_b = METRIC == kMetricL2 ? fminf(tb, f.max_score) : f.max_score;
    const unsigned char* st = stages + (size_t)s * tile_bytes;
    int pushed = 0;

    auto consider = [&](int e, float dis) {
      bool pred = e < n_e && dis <= hi_b && dis >= lo_b;
      unsigned long long key = kKeySentinel;
      if (pred) {
        int64_t raw = lids[(int64_t)ti * tile_e + e];
        pred = raw >= 0;  // tombstone (gamma_index_ivfpq.h:930)
        uint32_t vid = (uint32_t)raw;
        if (pred) pred = ctx_is_valid(f.del_bits, f.filter_bits, vid);
        key = make_key(score2ord<METRIC>(dis), vid);
        pred = pred && key < tau;
      }
      cq.push_warp(pred, key);
      pushed |= pred ? 1 : 0;
    };
    if (MW > 0) {
      float dis[PT];
#pragma unroll
      for (int u = 0; u < PT; u++) {
        const int e = u * PQ_NT + tid;
        // entries past n_e read stale-but-in-bounds stage bytes; they are masked in consider()
        dis[u] = adc_distance<MW>(st + (size_t)e * M, lut, dis0, M);
      }
#pragma unroll
      for (int u = 0; u < PT; u++) consider(u * PQ_NT + tid, dis[u]);
    } else {
      for (int u = 0; u < per_thread; u++) {
        const int e = u * PQ_NT + tid;
        float dis = e < n_e ? adc_distance<MW>(st + (size_t)e * M, lut, dis0, M) : 0.f;
        consider(e, dis);
      }
    }
    // stage consumed + pushes done; the count is identical in every thread (no 2nd barrier needed)
    est += __syncthreads_count(pushed) * per_thread;
    if (tid == 0 && gt + PQ_NST < total_tiles) issue(gt + PQ_NST);
    if (gt + 1 < total_tiles && est + tile_e > cq.cap()) {
      cq.flush();
      est = 0;
    }
  }
  __syncthreads();
  if (sorted_out) {
    cq.flush(true);
  } else {
    // The caller wants the k best and their bound, not their order (phase A of the tensor-core filter): exact radix
    // select, no sort; the largest key -- the sentinel when fewer than k entries qualified -- goes to slot k - 1.
    cq.flush(false, true);
    __shared__ unsigned long long s_wmax[PQ_NT / 32];
    __shared__ int s_wpos[PQ_NT / 32];
    unsigned long long mx = 0;
    int pos = -1;
    for (int i = tid; i < k; i += PQ_NT) {
      const unsigned long long v = buf[i];
      if (pos < 0 || v > mx) mx = v, pos = i;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, off);
      const int op = __shfl_xor_sync(0xffffffffu, pos, off);
      if (op >= 0 && (pos < 0 || om > mx || (om == mx && op > pos))) mx = om, pos = op;
    }
    if ((tid & 31) == 0) s_wmax[tid >> 5] = mx, s_wpos[tid >> 5] = pos;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < PQ_NT / 32; w++)
        if (s_wpos[w] >= 0 && (pos < 0 || s_wmax[w] > mx || (s_wmax[w] == mx && s_wpos[w] > pos))) mx = s_wmax[w], pos = s_wpos[w];
      if (pos >= 0 && pos != k - 1) {
        buf[pos] = buf[k - 1];
        buf[k - 1] = mx;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < k; i += PQ_NT) out[i] = buf[i];
}

// ---------------- exact re-rank -------------------------------------------------------------
constexpr int RR_NT = 256;

__global__ void __launch_bounds__(RR_NT)
    rerank_kernel(const unsigned long long* __restrict__ cand_keys, int ncand, int NP, const float* __restrict__ xq,
                  int64_t ldq, int d, const float* const* __restrict__ raw_segments, int seg_shift, int64_t ld_raw,
                  int k, int metric, FilterArgs f, unsigned long long* __restrict__ out_keys) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(rr_smem);  // [NP]
  float* qs = reinterpret_cast<float*>(buf + NP);                              // [d]
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < d; i += RR_NT) qs[i] = xq[(int64_t)q * ldq + i];
  for (int i = threadIdx.x; i < NP; i += RR_NT) buf[i] = kKeySentinel;
  __syncthreads();
  const float4* q4 = reinterpret_cast<const float4*>(qs);
  const uint32_t seg_mask = (1u << seg_shift) - 1u;
  // four candidates per warp iteration: 2 KiB of independent row loads in flight per warp (eight: 0.61 -> 1.04 ms; the gather is pure HBM
  // latency: recall_num random 4d-byte rows per query)
  constexpr int RR_U = 4;
  for (int j0 = warp * RR_U; j0 < ncand; j0 += (RR_NT / 32) * RR_U) {
    uint32_t vid[RR_U];
    const float4* row[RR_U];
    bool live[RR_U];
#pragma unroll
    for (int u = 0; u < RR_U; u++) {
      const int j = j0 + u;
      const unsigned long long ck = j < ncand ? cand_keys[(int64_t)q * ncand + j] : kKeySentinel;
      live[u] = ck != kKeySentinel;  // recall_idxi[j] < 0 (ivfpq.cc:691)
      vid[u] = live[u] ? (uint32_t)ck : 0u;
      row[u] = live[u] ? reinterpret_cast<const float4*>(raw_segments[vid[u] >> seg_shift] + (int64_t)(vid[u] & seg_mask) * ld_raw)
                       : nullptr;
    }
    float acc[RR_U][4];
#pragma unroll
    for (int u = 0; u < RR_U; u++) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    for (int c = lane; c < (d >> 2); c += 32) {
      const float4 w = q4[c];
      float4 v[RR_U];
#pragma unroll
      for (int u = 0; u < RR_U; u++) v[u] = live[u] ? __ldg(row[u] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < RR_U; u++) {
        if (metric == kMetricL2) {
          const float t0 = v[u].x - w.x, t1 = v[u].y - w.y, t2 = v[u].z - w.z, t3 = v[u].w - w.w;
          acc[u][0] = fmaf(t0, t0, acc[u][0]), acc[u][1] = fmaf(t1, t1, acc[u][1]);
          acc[u][2] = fmaf(t2, t2, acc[u][2]), acc[u][3] = fmaf(t3, t3, acc[u][3]);
        } else {
          acc[u][0] = fmaf(v[u].x, w.x, acc[u][0]), acc[u][1] = fmaf(v[u].y, w.y, acc[u][1]);
          acc[u][2] = fmaf(v[u].z, w.z, acc[u][2]), acc[u][3] = fmaf(v[u].w, w.w, acc[u][3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RR_U; u++) {
      float dis = (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) dis += __shfl_xor_sync(0xffffffffu, dis, off);
      if (lane == 0 && live[u] && dis <= f.max_score && dis >= f.min_score) buf[j0 + u] = make_key(score2ord(dis, metric), vid[u]);
    }
  }
  __syncthreads();
  block_bitonic_sort(buf, NP);
  for (int i = threadIdx.x; i < k; i += RR_NT) out_keys[(int64_t)q * k + i] = i < NP ? buf[i] : kKeySentinel;
}

// ---------------- PQ encode -------------------------------------------------------------------
constexpr int ENC_NT = 256;
constexpr int ENC_MAX_DSUB = 128;

template <int DSUB>
__global__ void __launch_bounds__(ENC_NT)
    pq_encode_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, const float* __restrict__ coarse, int64_t ldc,
                     const int32_t* __restrict__ assign, const float* __restrict__ pq, int M, int dsub_rt,
                     uint8_t* __restrict__ codes) {
  extern __shared__ float ps[];  // [256][dsub]
  const int dsub = DSUB > 0 ? DSUB : dsub_rt;
  const int m = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * ENC_NT + threadIdx.x;
  for (int idx = threadIdx.x; idx < KSUB * dsub; idx += ENC_NT) ps[idx] = pq[(int64_t)m * KSUB * dsub + idx];
  __syncthreads();
  if (i >= n) return;
  float r[DSUB > 0 ? DSUB : ENC_MAX_DSUB];
  const float* xi = x + i * ldx + m * dsub;
  const float* ci = (coarse && assign[i] >= 0) ? coarse + (int64_t)assign[i] * ldc + m * dsub : nullptr;
#pragma unroll
  for (int j = 0; j < dsub; j++) r[j] = ci ? __fsub_rn(xi[j], ci[j]) : xi[j];
  float best = FLT_MAX;
  int bi = 0;
  for (int c = 0; c < KSUB; c++) {
    float dis = 0.f;
#pragma unroll
    for (int j = 0; j < dsub; j++) {
      float t = __fsub_rn(r[j], ps[c * dsub + j]);
      dis = __fadd_rn(dis, __fmul_rn(t, t));
    }
    if (dis < best) {
      best = dis;
      bi = c;
    }
  }
  codes[i * M + m] = (uint8_t)bi;
}

void pq_cq_geometry(int k, int tile_e, int* KP, int* SORTN) {
  *KP = next_pow2(k < 16 ? 16 : k);
  *SORTN = next_pow2(*KP + tile_e + tile_e / 2);  // cap >= tile_e: one tile always fits after a flush
}

template <int METRIC, int MW>
cudaError_t launch_scan_t(const float* ip_table, int nq, const int32_t* probe_ids, const float* coarse_dis, int nprobe,
                          int ld_probe, int pg, ListDirectory dir, int M, const float* T, int k, FilterArgs f,
                          const int* gate_cnt, int gate_cap, const int* row_limit, int sorted_out, unsigned long long* partial,
                          cudaStream_t st) {
  int tile_e = pq_tile_entries(M);
  int KP, SORTN;
  pq_cq_geometry(k, tile_e, &KP, &SORTN);
  size_t smem = (size_t)PQ_NST * tile_e * M + (size_t)M * KSUB * 4 + (size_t)SORTN * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(ivfpq_scan_kernel<METRIC, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  int ngroups = (nprobe + pg - 1) / pg;
  dim3 grid(ngroups, nq);
  ivfpq_scan_kernel<METRIC, MW><<<grid, PQ_NT, smem, st>>>(ip_table, probe_ids, coarse_dis, nprobe, ld_probe, pg, dir, M, T,
                                                          tile_e, k, KP, SORTN, f, gate_cnt, gate_cap, row_limit, sorted_out, partial);
                                                          note_launch();
  return cudaGetLastError();
}

template <int METRIC>
cudaError_t launch_scan_m(const float* ip_table, int nq, const int32_t* probe_ids, const float* coarse_dis, int nprobe,
                          int ld_probe, int pg, ListDirectory dir, int M, const float* T, int k, FilterArgs f,
                          const int* gate_cnt, int gate_cap, const int* row_limit, int sorted_out, unsigned long long* partial,
                          cudaStream_t st) {
#define GB_SCAN(MW) \
  return launch_scan_t<METRIC, MW>(ip_table, nq, probe_ids, coarse_dis, nprobe, ld_probe, pg, dir, M, T, k, f, gate_cnt, \
                                   gate_cap, row_limit, sorted_out, partial, st)
  switch (M) {
    case 8: GB_SCAN(2);
    case 16: GB_SCAN(4);
    case 32: GB_SCAN(8);
    case 64: GB_SCAN(16);
    default: GB_SCAN(0);
  }
#undef GB_SCAN
}

}  // namespace

cudaError_t launch_pq_ip_table(const float* xq, int64_t ldq, int nq, const float* pq_centroids, int M, int dsub,
                               float* ip, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  size_t smem = ((size_t)dsub * KSUB + (size_t)IPT_QB * dsub) * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pq_ip_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(M, (nq + IPT_QB - 1) / IPT_QB);
  pq_ip_table_kernel<<<grid, KSUB, smem, st>>>(xq, ldq, nq, pq_centroids, M, dsub, ip);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pq_precompute_table(const float* coarse, int64_t ldc, int nlist, const float* pq_centroids, int M,
                                       int dsub, float* T, cudaStream_t st) {
  if (nlist <= 0) return cudaSuccess;
  pq_precompute_table_kernel<<<nlist, KSUB, 0, st>>>(coarse, ldc, pq_centroids, M, dsub, T);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_ivfpq_scan(const float* ip_table, int nq, const int32_t* probe_ids, const float* coarse_dis,
                              int nprobe, int pg, ListDirectory dir, int M, const float* T, int k, int metric,
                              FilterArgs f, unsigned long long* partial, cudaStream_t st, int ld_probe,
                              const int* gate_cnt, int gate_cap, const int* row_limit, bool sorted_out) {
  if (nq <= 0 || nprobe <= 0) return cudaSuccess;
  if (ld_probe <= 0) ld_probe = nprobe;
  if (k <= 0 || k > 4096 || nq > 65535 || pg < 1 || pg > PQ_MAX_PG) return cudaErrorInvalidValue;
  if (metric == kMetricL2) {
    if (!T) return cudaErrorInvalidValue;
    return launch_scan_m<kMetricL2>(ip_table, nq, probe_ids, coarse_dis, nprobe, ld_probe, pg, dir, M, T, k, f, gate_cnt, gate_cap, row_limit, sorted_out ? 1 : 0, partial, st);
  }
  return launch_scan_m<kMetricIP>(ip_table, nq, probe_ids, coarse_dis, nprobe, ld_probe, pg, dir, M, T, k, f, gate_cnt, gate_cap, row_limit, sorted_out ? 1 : 0, partial, st);
}

cudaError_t launch_rerank(const unsigned long long* cand_keys, int ncand, int nq, const float* xq, int64_t ldq, int d,
                          const float* const* raw_segments, int seg_shift, int64_t ld_raw, int k, int metric,
                          FilterArgs f, unsigned long long* out_keys, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  if (ncand <= 0 || ncand > 8192 || (d & 3)) return cudaErrorInvalidValue;
  int NP = next_pow2(ncand < 2 ? 2 : ncand);  // >= 16 bytes of keys: the query row behind them is read as float4
  size_t smem = (size_t)NP * 8 + (size_t)d * 4;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  rerank_kernel<<<nq, RR_NT, smem, st>>>(cand_keys, ncand, NP, xq, ldq, d, raw_segments, seg_shift, ld_raw, k, metric,
                                         f, out_keys);
                                         note_launch();
  return cudaGetLastError();
}

cudaError_t launch_pq_encode(const float* x, int64_t ldx, int64_t n, const float* coarse, int64_t ldc,
                             const int32_t* assign, const float* pq_centroids, int M, int dsub, uint8_t* codes,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (dsub > ENC_MAX_DSUB) return cudaErrorInvalidValue;
  size_t smem = (size_t)KSUB * dsub * 4;
  dim3 grid((unsigned)((n + ENC_NT - 1) / ENC_NT), M);
#define GB_ENC(DS)                                                                                               \
  {                                                                                                              \
    if (smem > 48 * 1024) {                                                                                      \
      cudaError_t e =                                                                                            \
          cudaFuncSetAttribute(pq_encode_kernel<DS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
      if (e != cudaSuccess) return e;                                                                            \
    }                                                                                                            \
    pq_encode_kernel<DS><<<grid, ENC_NT, smem, st>>>(x, ldx, n, coarse, ldc, assign, pq_centroids, M, dsub, codes); \
    note_launch(); \
  }
  switch (dsub) {
    case 2: GB_ENC(2) break;
    case 4: GB_ENC(4) break;
    case 8: GB_ENC(8) break;
    case 16: GB_ENC(16) break;
    case 32: GB_ENC(32) break;
    default: GB_ENC(0) break;
  }
#undef GB_ENC
  return cudaGetLastError();
}

}  // namespace gb
