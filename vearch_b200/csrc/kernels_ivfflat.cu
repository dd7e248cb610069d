// K3: IVF-Flat inverted-list scan (query-major), the HBM-roofline kernel of the path.
//
// Replaces GammaIVFFlatIndex::search_preassigned (index/impl/gamma_index_ivfflat.cc:579-787,
// pmode-0 loop :695-733) and GammaIVFFlatScanner::scan_codes (gamma_index_ivfflat.h:63-91):
//   for each query, for each probed list, for each entry j:
//     skip if ids[j] & kDelIdxMask, skip if !IsValid(vid), dis = L2sqr/IP(x, vec_j),
//     keep if IsSimilarScoreValid(dis) and better than the current k-th best.
//
// One CTA per (query, probed list, list split).  The list's rows stream HBM -> shared memory
// through a 4-stage ring of TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), so the
// copy engine keeps 4 x 16 KB in flight per CTA with no register staging.  LPR lanes share one
// row (float4 chunks, rotated start so both the row and the query reads are bank-conflict-free),
// partial sums meet through warp shuffles, and candidates that beat the running k-th best go to
// the CandQueue (common.cuh).  List ids are only read for candidates that beat the threshold.
//
// Algorithmic bytes: (4*d + 8) per scanned entry; roofline: HBM.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace gb {

namespace {

constexpr int IVF_NT = 128;
constexpr int IVF_NST = 4;                 // pipeline stages
constexpr int IVF_STAGE_BYTES = 16 * 1024; // target bytes per stage
constexpr int IVF_CHUNK_ROWS = 8192;       // rows per CTA (long lists are split)

struct IvfGeom {
  int LPR, G, R, T, J, S;
  int stage_bytes;
};

__host__ __device__ inline IvfGeom ivf_geom(int d) {
  IvfGeom g;
  g.S = d >> 2;
  int lpr = 1;
  while (lpr < 32 && g.S / (lpr * 2) >= 8) lpr <<= 1;  // aim for >= 8 float4 chunks per lane
  g.LPR = lpr;
  g.G = IVF_NT / lpr;
  g.J = (g.S + lpr - 1) / lpr;
  int row_bytes = d * 4;
  int r = IVF_STAGE_BYTES / (g.G * row_bytes);
  g.R = r < 1 ? 1 : r;
  g.T = g.G * g.R;
  g.stage_bytes = g.T * row_bytes;
  return g;
}

template <int METRIC>
__global__ void __launch_bounds__(IVF_NT)
    ivfflat_scan_kernel(const float* __restrict__ xq, int64_t ldq, int d, const int32_t* __restrict__ probe_ids,
                        int nprobe, int nsplit, ListDirectory dir, IvfGeom g, int k, int KP, int SORTN, FilterArgs f,
                        unsigned long long* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stages = reinterpret_cast<float*>(smem_raw);
  float* qs = reinterpret_cast<float*>(smem_raw + (size_t)IVF_NST * g.stage_bytes);
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(qs + d);
  __shared__ __align__(8) uint64_t full_bar[IVF_NST];
  __shared__ int s_cnt;
  __shared__ unsigned long long s_tau;

  const int tid = threadIdx.x;
  const int q = blockIdx.y;
  const int part = blockIdx.x;
  const int probe = part / nsplit, split = part - probe * nsplit;
  unsigned long long* out = partial + ((int64_t)q * gridDim.x + part) * k;

  const int list = probe_ids[(int64_t)q * nprobe + probe];
  int len = 0;
  if (list >= 0 && list < dir.nlist) len = dir.len[list];
  const int r0 = split * IVF_CHUNK_ROWS;
  const int r1 = min(len, r0 + IVF_CHUNK_ROWS);
  if (r0 >= r1) {  // nothing to scan: "not enough centroids" (ivfflat.cc:653) or empty split
    for (int i = tid; i < k; i += IVF_NT) out[i] = kKeySentinel;
    return;
  }
  const float* __restrict__ lvecs = dir.vecs[list];
  const int64_t* __restrict__ lids = dir.ids[list];

  CandQueue cq{buf, &s_cnt, &s_tau, k, KP, SORTN};
  for (int i = tid; i < d; i += IVF_NT) qs[i] = xq[(int64_t)q * ldq + i];
  if (tid == 0) {
    for (int s = 0; s < IVF_NST; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  cq.init();  // includes __syncthreads()

  const int ntiles = (r1 - r0 + g.T - 1) / g.T;
  const int row_bytes = d * 4;
  auto issue = [&](int t) {
    int s = t % IVF_NST;
    int rows = min(g.T, r1 - (r0 + t * g.T));
    uint32_t bytes = (uint32_t)rows * row_bytes;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(reinterpret_cast<unsigned char*>(stages) + (size_t)s * g.stage_bytes,
             lvecs + (int64_t)(r0 + t * g.T) * d, bytes, &full_bar[s]);
  };
  if (tid == 0)
    for (int t = 0; t < IVF_NST && t < ntiles; t++) issue(t);

  const int grp = tid / g.LPR;      // row slot inside a round
  const int p = tid - grp * g.LPR;  // lane inside the row group
  const float4* qs4 = reinterpret_cast<const float4*>(qs);

  for (int t = 0; t < ntiles; t++) {
    const int s = t % IVF_NST;
    mbar_wait(&full_bar[s], (t / IVF_NST) & 1);
    const int tile_rows = min(g.T, r1 - (r0 + t * g.T));
    const unsigned long long tau = s_tau;
    const uint32_t tau_hi = (uint32_t)(tau >> 32);
    const float4* st4 = reinterpret_cast<const float4*>(reinterpret_cast<unsigned char*>(stages) + (size_t)s * g.stage_bytes);

    for (int r = 0; r < g.R; r++) {
      const int rit = r * g.G + grp;  // row in tile
      const bool valid = rit < tile_rows;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (valid) {
        const float4* rowp = st4 + (size_t)rit * g.S;
        int jj = rit % g.J;  // rotated start: bank-conflict-free for S % 8 == 0
#pragma unroll 4
        for (int j = 0; j < g.J; j++) {
          int c = p + g.LPR * jj;
          if (c < g.S) {
            float4 v = rowp[c];
            float4 w = qs4[c];
            if (METRIC == kMetricL2) {
              float t0 = v.x - w.x, t1 = v.y - w.y, t2 = v.z - w.z, t3 = v.w - w.w;
              a0 = fmaf(t0, t0, a0), a1 = fmaf(t1, t1, a1), a2 = fmaf(t2, t2, a2), a3 = fmaf(t3, t3, a3);
            } else {
              a0 = fmaf(v.x, w.x, a0), a1 = fmaf(v.y, w.y, a1), a2 = fmaf(v.z, w.z, a2), a3 = fmaf(v.w, w.w, a3);
            }
          }
          jj = (jj + 1 == g.J) ? 0 : jj + 1;
        }
      }
      float dis = (a0 + a1) + (a2 + a3);
      for (int off = g.LPR >> 1; off > 0; off >>= 1) dis += __shfl_xor_sync(0xffffffffu, dis, off);

      bool pred = valid && p == 0 && dis <= f.max_score && dis >= f.min_score;
      unsigned long long key = kKeySentinel;
      if (pred) {
        uint32_t ord = score2ord<METRIC>(dis);
        pred = ord <= tau_hi;
        if (pred) {
          int64_t raw = lids[r0 + t * g.T + rit];
          pred = raw >= 0;  // top bit set => tombstone (gamma_index_ivfflat.h:72)
          uint32_t vid = (uint32_t)raw;
          if (pred) pred = ctx_is_valid(f.del_bits, f.filter_bits, vid);
          key = make_key(ord, vid);
          pred = pred && key < tau;
        }
      }
      cq.push_warp(pred, key);
    }
    __syncthreads();  // stage s fully consumed, all pushes of this tile done
    const int c_now = s_cnt;
    if (tid == 0 && t + IVF_NST < ntiles) issue(t + IVF_NST);
    __syncthreads();  // everyone holds the same c_now before any warp pushes again
    if (t + 1 < ntiles && c_now + g.T > cq.cap()) cq.flush();
  }
  cq.flush(true);
  for (int i = tid; i < k; i += IVF_NT) out[i] = buf[i];
}

void ivf_cq_geometry(int k, int T, int* KP, int* SORTN) {
  *KP = next_pow2(k < 16 ? 16 : k);
  *SORTN = next_pow2(*KP + 2 * T);
}

// ---- fast path: compile-time geometry ---------------------------------------------------
// LPR >= 8 lanes share a row, lane p owns float4 chunks {p + LPR*j, j < J} (S = LPR*J).  A
// quarter-warp then reads 8 consecutive float4 of ONE row, so shared-memory reads are
// conflict-free without rotation and the query chunks sit in registers (4*J floats per lane).
// One __syncthreads_count per tile both releases the stage and gives every thread the same
// upper bound on the queue fill, so the flush decision needs no second barrier.
constexpr int IVF_MAX_PG = 32;  // work items (probe, split) per CTA

template <int METRIC, int LPR, int J>
__global__ void __launch_bounds__(IVF_NT)
    ivfflat_scan_fast_kernel(const float* __restrict__ xq, int64_t ldq, const int32_t* __restrict__ probe_ids,
                             int nprobe, int nsplit, int pg, ListDirectory dir, int R, int nst, int k, int KP, int SORTN,
                             FilterArgs f, unsigned long long* __restrict__ partial) {
  constexpr int S = LPR * J;        // float4 per row
  constexpr int G = IVF_NT / LPR;   // rows per round
  constexpr int ROW_BYTES = S * 16;
  const int T = G * R;
  const int stage_bytes = T * ROW_BYTES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* buf = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)nst * stage_bytes);
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ int s_cnt;
  __shared__ unsigned long long s_tau;
  __shared__ int g_list[IVF_MAX_PG], g_r0[IVF_MAX_PG], g_r1[IVF_MAX_PG], g_tile0[IVF_MAX_PG + 1];

  const int tid = threadIdx.x;
  const int q = blockIdx.y, grp_id = blockIdx.x;
  unsigned long long* out = partial + ((int64_t)q * gridDim.x + grp_id) * k;
  const int nitems = nprobe * nsplit;
  const int i0 = grp_id * pg;
  const int ni = min(pg, nitems - i0);

  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < ni; i++) {
      const int item = i0 + i;
      const int probe = item / nsplit, split = item - probe * nsplit;
      const int list = probe_ids[(int64_t)q * nprobe + probe];
      int len = 0;
      if (list >= 0 && list < dir.nlist) len = dir.len[list];  // key < 0: "not enough centroids" (ivfflat.cc:653)
      const int r0 = split * IVF_CHUNK_ROWS;
      const int r1 = min(len, r0 + IVF_CHUNK_ROWS);
      g_list[i] = list;
      g_r0[i] = r0;
      g_r1[i] = r1 > r0 ? r1 : r0;
      g_tile0[i] = acc;
      acc += r1 > r0 ? (r1 - r0 + T - 1) / T : 0;
    }
    g_tile0[ni > 0 ? ni : 0] = acc;
    for (int s = 0; s < nst; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  CandQueue cq{buf, &s_cnt, &s_tau, k, KP, SORTN};
  cq.init();  // __syncthreads inside
  const int total_tiles = ni > 0 ? g_tile0[ni] : 0;
  if (total_tiles == 0) {
    for (int i = tid; i < k; i += IVF_NT) out[i] = kKeySentinel;
    return;
  }

  const int grp = tid / LPR, p = tid % LPR;
  float4 qreg[J];
  {
    const float4* q4 = reinterpret_cast<const float4*>(xq + (int64_t)q * ldq);
#pragma unroll
    for (int j = 0; j < J; j++) qreg[j] = __ldg(q4 + p + LPR * j);
  }

  // producer (thread 0): walks the same (item, tile) sequence NST tiles ahead of the consumers,
  // so the ring keeps streaming across list boundaries
  int pr_pi = 0;
  auto issue = [&](int gt) {
    while (gt >= g_tile0[pr_pi + 1]) pr_pi++;
    const int ti = gt - g_tile0[pr_pi];
    const int row0 = g_r0[pr_pi] + ti * T;
    const int rows = min(T, g_r1[pr_pi] - row0);
    const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
    const int s = gt % nst;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(smem_raw + (size_t)s * stage_bytes, dir.vecs[g_list[pr_pi]] + (int64_t)row0 * (S * 4), bytes, &full_bar[s]);
  };
  if (tid == 0)
    for (int gt = 0; gt < nst && gt < total_tiles; gt++) issue(gt);

  int pi = 0;
  const int64_t* __restrict__ lids = nullptr;
  int cur = -1;
  int est = 0;  // upper bound of the queue fill, identical in every thread
  for (int gt = 0; gt < total_tiles; gt++) {
    while (gt >= g_tile0[pi + 1]) pi++;
    if (pi != cur) {
      cur = pi;
      lids = dir.ids[g_list[pi]];
    }
    const int ti = gt - g_tile0[pi];
    const int row0 = g_r0[pi] + ti * T;
    const int tile_rows = min(T, g_r1[pi] - row0);
    const int s = gt % nst;
    mbar_wait(&full_bar[s], (gt / nst) & 1);
    const unsigned long long tau = s_tau;
    const uint32_t tau_hi = (uint32_t)(tau >> 32);
    const float4* st4 = reinterpret_cast<const float4*>(smem_raw + (size_t)s * stage_bytes);
    int pushed = 0;
    for (int r = 0; r < R; r++) {
      const int rit = r * G + grp;
      const bool valid = rit < tile_rows;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (valid) {
        const float4* rowp = st4 + (size_t)rit * S + p;
#pragma unroll
        for (int j = 0; j < J; j++) {
          float4 v = rowp[LPR * j];
          float4 w = qreg[j];
          if (METRIC == kMetricL2) {
            float t0 = v.x - w.x, t1 = v.y - w.y, t2 = v.z - w.z, t3 = v.w - w.w;
            a0 = fmaf(t0, t0, a0), a1 = fmaf(t1, t1, a1), a2 = fmaf(t2, t2, a2), a3 = fmaf(t3, t3, a3);
          } else {
            a0 = fmaf(v.x, w.x, a0), a1 = fmaf(v.y, w.y, a1), a2 = fmaf(v.z, w.z, a2), a3 = fmaf(v.w, w.w, a3);
          }
        }
      }
      float dis = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int off = LPR >> 1; off > 0; off >>= 1) dis += __shfl_xor_sync(0xffffffffu, dis, off);

      bool pred = valid && p == 0 && dis <= f.max_score && dis >= f.min_score;
      unsigned long long key = kKeySentinel;
      if (pred) {
        uint32_t ord = score2ord<METRIC>(dis);
        pred = ord <= tau_hi;
        if (pred) {
          int64_t raw = lids[row0 + rit];
          pred = raw >= 0;  // top bit set => tombstone (gamma_index_ivfflat.h:72)
          uint32_t vid = (uint32_t)raw;
          if (pred) pred = ctx_is_valid(f.del_bits, f.filter_bits, vid);
          key = make_key(ord, vid);
          pred = pred && key < tau;
        }
      }
      cq.push_warp(pred, key);
      pushed |= pred ? 1 : 0;
    }
    // stage s consumed + pushes done; the count is the same value in every thread
    est += __syncthreads_count(pushed) * R;
    if (tid == 0 && gt + nst < total_tiles) issue(gt + nst);
    if (gt + 1 < total_tiles && est + T > cq.cap()) {
      cq.flush();
      est = 0;
    }
  }
  __syncthreads();
  cq.flush(true);
  for (int i = tid; i < k; i += IVF_NT) out[i] = buf[i];
}

struct FastCfg {
  int pg, nst, stage_bytes;
  int warp_mode;  // 1: warp-autonomous kernel (per-warp TMA rings), 0: block-synchronous ring
};
inline FastCfg fast_cfg(int nprobe, int nsplit, int avg_len) {
  FastCfg c;
  // measured on B200 (gpurun_out/cfg_sweep.log): resident CTAs matter more than ring depth --
  // 8 KiB x 2 stages: 59 ms, x3: 62 ms, x4: 74 ms, x6: 100 ms, 16 KiB x 4: 91 ms (C2, nq = 10 k)
  // warp-autonomous kernel, 4 KiB per warp per ring slot, 2 slots: 51 ms on C2 (10.4 TB/s
  // algorithmic) vs 58 ms for the best block-synchronous configuration (gpurun_out/cfg_sweep2.log)
  c.stage_bytes = 16 * 1024;
  c.nst = 2;
  c.warp_mode = 1;
  if (const char* e = getenv("GB_IVF_MODE")) c.warp_mode = !strcmp(e, "warp");
  if (!c.warp_mode) c.stage_bytes = 8 * 1024;
  // ~16k rows of work per CTA: enough tiles to amortise start-up, enough CTAs to balance
  int per = avg_len > 0 ? 16384 / avg_len : 1;
  c.pg = per < 1 ? 1 : (per > IVF_MAX_PG ? IVF_MAX_PG : per);
  if (nsplit > 1) c.pg = 1;
  if (const char* e = getenv("GB_IVF_STAGE_KB")) c.stage_bytes = atoi(e) * 1024;
  if (const char* e = getenv("GB_IVF_NST")) c.nst = atoi(e);
  if (const char* e = getenv("GB_IVF_PG")) c.pg = atoi(e);
  if (c.nst < 2) c.nst = 2;
  if (c.nst > 8) c.nst = 8;
  if (c.pg < 1) c.pg = 1;
  if (c.pg > IVF_MAX_PG) c.pg = IVF_MAX_PG;
  if (c.pg > nprobe * nsplit) c.pg = nprobe * nsplit;
  return c;
}

template <int METRIC, int LPR, int J>
cudaError_t launch_fast(const float* xq, int64_t ldq, int nq, const int32_t* probe_ids, int nprobe, int nsplit,
                        FastCfg c, ListDirectory dir, int k, FilterArgs f, unsigned long long* partial,
                        cudaStream_t st) {
  constexpr int S = LPR * J, G = IVF_NT / LPR, ROW_BYTES = S * 16;
  int R = c.stage_bytes / (G * ROW_BYTES);
  if (R < 1) R = 1;
  const int T = G * R;
  int KP, SORTN;
  ivf_cq_geometry(k, T, &KP, &SORTN);
  size_t smem = (size_t)c.nst * T * ROW_BYTES + (size_t)SORTN * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(ivfflat_scan_fast_kernel<METRIC, LPR, J>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int nitems = nprobe * nsplit;
  dim3 grid((nitems + c.pg - 1) / c.pg, nq);
  ivfflat_scan_fast_kernel<METRIC, LPR, J><<<grid, IVF_NT, smem, st>>>(xq, ldq, probe_ids, nprobe, nsplit, c.pg, dir, R,
                                                                       c.nst, k, KP, SORTN, f, partial);
  note_launch();
  return cudaGetLastError();
}

// ---- warp-autonomous variant ---------------------------------------------------------------
// Same work decomposition as the fast kernel (one CTA per query x group of (probe, split) items,
// flat tile sequence), but every WARP owns its slice of the tile sequence (tiles w, w+4, ...), its
// own ring of TMA bulk copies with its own mbarriers (lane 0 produces), its own threshold and its
// own candidate queue (count and tau live in registers, pushes are ballot-compacted, flushes are
// warp-level bitonic sorts).  No block-wide barrier exists until the final 4-way merge, so a warp
// never waits for another warp's memory latency.
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncwarp();
    }
  }
}

constexpr int IVF_NWARP = IVF_NT / 32;
constexpr int IVF_MAX_NST = 8;

template <int METRIC, int LPR, int J>
__global__ void __launch_bounds__(IVF_NT)
    ivfflat_scan_warp_kernel(const float* __restrict__ xq, int64_t ldq, const int32_t* __restrict__ probe_ids,
                             int nprobe, int nsplit, int pg, ListDirectory dir, int R, int nst, int k, int KP, int SORTNW,
                             FilterArgs f, unsigned long long* __restrict__ partial) {
  constexpr int S = LPR * J;
  constexpr int RPW = 32 / LPR;  // rows per warp round
  constexpr int ROW_BYTES = S * 16;
  const int TW = RPW * R;  // rows per warp tile
  const int stage_bytes = TW * ROW_BYTES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)IVF_NWARP * nst * stage_bytes);
  unsigned long long* fin = keys + (size_t)IVF_NWARP * SORTNW;  // [IVF_NWARP * KP]
  __shared__ __align__(8) uint64_t bars[IVF_NWARP][IVF_MAX_NST];
  __shared__ int g_list[IVF_MAX_PG], g_r0[IVF_MAX_PG], g_r1[IVF_MAX_PG], g_tile0[IVF_MAX_PG + 1];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int q = blockIdx.y, grp_id = blockIdx.x;
  unsigned long long* out = partial + ((int64_t)q * gridDim.x + grp_id) * k;
  const int nitems = nprobe * nsplit;
  const int i0 = grp_id * pg;
  const int ni = min(pg, nitems - i0);

  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < ni; i++) {
      const int item = i0 + i;
      const int probe = item / nsplit, split = item - probe * nsplit;
      const int list = probe_ids[(int64_t)q * nprobe + probe];
      int len = 0;
      if (list >= 0 && list < dir.nlist) len = dir.len[list];
      const int r0 = split * IVF_CHUNK_ROWS;
      const int r1 = min(len, r0 + IVF_CHUNK_ROWS);
      g_list[i] = list;
      g_r0[i] = r0;
      g_r1[i] = r1 > r0 ? r1 : r0;
      g_tile0[i] = acc;
      acc += r1 > r0 ? (r1 - r0 + TW - 1) / TW : 0;
    }
    g_tile0[ni > 0 ? ni : 0] = acc;
    for (int ww = 0; ww < IVF_NWARP; ww++)
      for (int s = 0; s < nst; s++) mbar_init(&bars[ww][s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int total_tiles = ni > 0 ? g_tile0[ni] : 0;
  if (total_tiles == 0) {
    for (int i = tid; i < k; i += IVF_NT) out[i] = kKeySentinel;
    return;
  }

  unsigned long long* wbuf = keys + (size_t)w * SORTNW;
  for (int i = lane; i < SORTNW; i += 32) wbuf[i] = kKeySentinel;
  __syncwarp();
  const int capw = SORTNW - KP;
  unsigned long long tau = kKeySentinel;
  int cnt = 0;
  const int rowslot = lane / LPR, p = lane % LPR;
  float4 qreg[J];
  {
    const float4* q4 = reinterpret_cast<const float4*>(xq + (int64_t)q * ldq);
#pragma unroll
    for (int j = 0; j < J; j++) qreg[j] = __ldg(q4 + p + LPR * j);
  }
  unsigned char* wstage = smem_raw + (size_t)w * nst * stage_bytes;
  const int n_my = total_tiles > w ? (total_tiles - w + IVF_NWARP - 1) / IVF_NWARP : 0;

  auto warp_flush = [&]() {
    for (int i = KP + cnt + lane; i < SORTNW; i += 32) wbuf[i] = kKeySentinel;
    __syncwarp();
    warp_bitonic_sort(wbuf, SORTNW, lane);
    tau = wbuf[k - 1];
    cnt = 0;
    __syncwarp();
  };

  int pr_pi = 0;  // producer cursor (lane 0)
  auto issue = [&](int li) {
    const int gt = w + IVF_NWARP * li;
    while (gt >= g_tile0[pr_pi + 1]) pr_pi++;
    const int ti = gt - g_tile0[pr_pi];
    const int row0 = g_r0[pr_pi] + ti * TW;
    const int rows = min(TW, g_r1[pr_pi] - row0);
    const uint32_t bytes = (uint32_t)rows * ROW_BYTES;
    const int s = li % nst;
    mbar_arrive_expect_tx(&bars[w][s], bytes);
    bulk_g2s(wstage + (size_t)s * stage_bytes, dir.vecs[g_list[pr_pi]] + (int64_t)row0 * (S * 4), bytes, &bars[w][s]);
  };
  if (lane == 0)
    for (int li = 0; li < nst && li < n_my; li++) issue(li);

  int pi = 0;
  for (int li = 0; li < n_my; li++) {
    const int gt = w + IVF_NWARP * li;
    while (gt >= g_tile0[pi + 1]) pi++;
    const int64_t* __restrict__ lids = dir.ids[g_list[pi]];
    const int ti = gt - g_tile0[pi];
    const int row0 = g_r0[pi] + ti * TW;
    const int tile_rows = min(TW, g_r1[pi] - row0);
    const int s = li % nst;
    mbar_wait(&bars[w][s], (li / nst) & 1);
    const float4* st4 = reinterpret_cast<const float4*>(wstage + (size_t)s * stage_bytes);
    for (int r = 0; r < R; r++) {
      const int rit = r * RPW + rowslot;
      const bool valid = rit < tile_rows;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (valid) {
        const float4* rowp = st4 + (size_t)rit * S + p;
#pragma unroll
        for (int j = 0; j < J; j++) {
          float4 v = rowp[LPR * j];
          float4 wq = qreg[j];
          if (METRIC == kMetricL2) {
            float t0 = v.x - wq.x, t1 = v.y - wq.y, t2 = v.z - wq.z, t3 = v.w - wq.w;
            a0 = fmaf(t0, t0, a0), a1 = fmaf(t1, t1, a1), a2 = fmaf(t2, t2, a2), a3 = fmaf(t3, t3, a3);
          } else {
            a0 = fmaf(v.x, wq.x, a0), a1 = fmaf(v.y, wq.y, a1), a2 = fmaf(v.z, wq.z, a2), a3 = fmaf(v.w, wq.w, a3);
          }
        }
      }
      float dis = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int off = LPR >> 1; off > 0; off >>= 1) dis += __shfl_xor_sync(0xffffffffu, dis, off);

      bool pred = valid && p == 0 && dis <= f.max_score && dis >= f.min_score;
      unsigned long long key = kKeySentinel;
      if (pred) {
        uint32_t ord = score2ord<METRIC>(dis);
        pred = ord <= (uint32_t)(tau >> 32);
        if (pred) {
          int64_t raw = lids[row0 + rit];
          pred = raw >= 0;  // top bit set => tombstone (gamma_index_ivfflat.h:72)
          uint32_t vid = (uint32_t)raw;
          if (pred) pred = ctx_is_valid(f.del_bits, f.filter_bits, vid);
          key = make_key(ord, vid);
          pred = pred && key < tau;
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, pred);
      if (mask) {
        if (pred) wbuf[KP + cnt + __popc(mask & ((1u << lane) - 1u))] = key;
        cnt += __popc(mask);
        if (cnt + RPW > capw) warp_flush();  // warp-uniform
      }
    }
    __syncwarp();  // every lane is done with stage s
    if (lane == 0 && li + nst < n_my) issue(li + nst);
  }
  warp_flush();
  __syncthreads();
  // merge the four per-warp top-KP lists
  const int NF = IVF_NWARP * KP;
  for (int i = tid; i < NF; i += IVF_NT) fin[i] = keys[(size_t)(i / KP) * SORTNW + (i % KP)];
  __syncthreads();
  block_bitonic_sort(fin, NF);
  for (int i = tid; i < k; i += IVF_NT) out[i] = fin[i];
}

template <int METRIC, int LPR, int J>
cudaError_t launch_warp(const float* xq, int64_t ldq, int nq, const int32_t* probe_ids, int nprobe, int nsplit,
                        FastCfg c, ListDirectory dir, int k, FilterArgs f, unsigned long long* partial,
                        cudaStream_t st) {
  constexpr int S = LPR * J, RPW = 32 / LPR, ROW_BYTES = S * 16;
  int R = c.stage_bytes / IVF_NWARP / (RPW * ROW_BYTES);  // stage_bytes = per-CTA bytes per ring slot
  if (R < 1) R = 1;
  const int TW = RPW * R;
  const int KP = next_pow2(k < 16 ? 16 : k);
  const int SORTNW = next_pow2(KP + 2 * RPW + TW);
  size_t smem = (size_t)IVF_NWARP * c.nst * TW * ROW_BYTES + (size_t)IVF_NWARP * SORTNW * 8 + (size_t)IVF_NWARP * KP * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(ivfflat_scan_warp_kernel<METRIC, LPR, J>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int nitems = nprobe * nsplit;
  dim3 grid((nitems + c.pg - 1) / c.pg, nq);
  ivfflat_scan_warp_kernel<METRIC, LPR, J><<<grid, IVF_NT, smem, st>>>(xq, ldq, probe_ids, nprobe, nsplit, c.pg, dir, R,
                                                                       c.nst, k, KP, SORTNW, f, partial);
  note_launch();
  return cudaGetLastError();
}

// returns cudaErrorNotSupported when no instantiation fits (caller falls back to the generic kernel)
template <int METRIC>
cudaError_t dispatch_fast(int S, const float* xq, int64_t ldq, int nq, const int32_t* probe_ids, int nprobe,
                          int nsplit, FastCfg c, ListDirectory dir, int k, FilterArgs f, unsigned long long* partial,
                          cudaStream_t st) {
#define GB_FAST(LPR, J)                                                                                              \
  if (S == (LPR) * (J))                                                                                            \
    return c.warp_mode ? launch_warp<METRIC, LPR, J>(xq, ldq, nq, probe_ids, nprobe, nsplit, c, dir, k, f, partial, st) \
                       : launch_fast<METRIC, LPR, J>(xq, ldq, nq, probe_ids, nprobe, nsplit, c, dir, k, f, partial, st)
  GB_FAST(8, 1);   // d = 32
  GB_FAST(8, 2);   // d = 64
  GB_FAST(8, 3);   // d = 96
  GB_FAST(8, 4);   // d = 128
  GB_FAST(8, 5);
  GB_FAST(8, 6);   // d = 192
  GB_FAST(8, 7);
  GB_FAST(8, 8);   // d = 256
  GB_FAST(16, 5);
  GB_FAST(16, 6);  // d = 384
  GB_FAST(16, 7);
  GB_FAST(16, 8);  // d = 512
  GB_FAST(32, 5);
  GB_FAST(32, 6);  // d = 768
  GB_FAST(32, 7);
  GB_FAST(32, 8);  // d = 1024
  GB_FAST(32, 12); // d = 1536
#undef GB_FAST
  return cudaErrorNotSupported;
}

}  // namespace

int ivfflat_scan_nparts(int nprobe, int max_list_len) {
  int nsplit = (max_list_len + IVF_CHUNK_ROWS - 1) / IVF_CHUNK_ROWS;
  if (nsplit < 1) nsplit = 1;
  return nprobe * nsplit;
}

cudaError_t launch_ivfflat_scan(const float* xq, int64_t ldq, int nq, int d, const int32_t* probe_ids, int nprobe,
                                ListDirectory dir, int max_list_len, int avg_list_len, int k, int metric, FilterArgs f,
                                unsigned long long* partial, int* nparts_out, cudaStream_t st) {
  if (nq <= 0 || nprobe <= 0) return cudaSuccess;
  if ((d & 3) || k <= 0 || k > 4096 || nq > 65535) return cudaErrorInvalidValue;
  IvfGeom g = ivf_geom(d);
  int nparts = ivfflat_scan_nparts(nprobe, max_list_len);
  int nsplit = nparts / nprobe;
  if (nparts_out) *nparts_out = nparts;
  if (ldq % 4 == 0 && !getenv("GB_IVF_GENERIC")) {
    FastCfg c = fast_cfg(nprobe, nsplit, avg_list_len);
    cudaError_t fe = metric == kMetricL2 ? dispatch_fast<kMetricL2>(d >> 2, xq, ldq, nq, probe_ids, nprobe, nsplit, c,
                                                                    dir, k, f, partial, st)
                                         : dispatch_fast<kMetricIP>(d >> 2, xq, ldq, nq, probe_ids, nprobe, nsplit, c,
                                                                    dir, k, f, partial, st);
    if (fe != cudaErrorNotSupported) {
      if (nparts_out) *nparts_out = (nparts + c.pg - 1) / c.pg;  // partial[q][group][k]
      return fe;
    }
  }
  int KP, SORTN;
  ivf_cq_geometry(k, g.T, &KP, &SORTN);
  size_t smem = (size_t)IVF_NST * g.stage_bytes + (size_t)d * 4 + (size_t)SORTN * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  dim3 grid(nparts, nq);
  cudaError_t e;
  if (metric == kMetricL2) {
    e = cudaFuncSetAttribute(ivfflat_scan_kernel<kMetricL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ivfflat_scan_kernel<kMetricL2><<<grid, IVF_NT, smem, st>>>(xq, ldq, d, probe_ids, nprobe, nsplit, dir, g, k, KP,
                                                               SORTN, f, partial);
                                                               note_launch();
  } else {
    e = cudaFuncSetAttribute(ivfflat_scan_kernel<kMetricIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ivfflat_scan_kernel<kMetricIP><<<grid, IVF_NT, smem, st>>>(xq, ldq, d, probe_ids, nprobe, nsplit, dir, g, k, KP,
                                                               SORTN, f, partial);
                                                               note_launch();
  }
  return cudaGetLastError();
}

}  // namespace gb
