// Per-query merge of the N-split candidate lists (shared by merge_splits_kernel and the fused readout).
#pragma once
#include "common.cuh"

namespace vosmem {

constexpr int MERGE_MAX_SPLITS = 16;
constexpr int MERGE_ROWS = 16;                    // 32-candidate rows merge_query_small holds in shared memory
constexpr int MERGE_BUF = (MERGE_ROWS + 1) * 32;  // warp-private scratch entries (scores and indices each)

struct SplitLists {
  const CandEntry *cand;    // [splits][hw_pad][slots]
  const int *cand_count;    // [splits][hw_pad]
  const PubEntry *pub;      // [pub_rows][hw_pad] published lower bounds per (virtual split, query) of the selection launch
  int splits, pub_rows, hw_pad;   // pub_rows == 0: the selection published nothing (SIMT path), no threshold
  const WsControl *ctl;     // ctl->last = the epoch the selection launch that filled the lists ran under
  // Group mode (tcgen05 selection, select_tc.cu): gscore != NULL, `splits` counts the VIRTUAL splits, a list has
  // GSLOTS entries {group maximum, locator} and gscore holds the 8 scores of every entry.  merge_query then returns
  // the best 32 GROUPS (index = list << 6 | slot) and expand_groups turns them into the best 32 keys.
  const float *gscore;
  int seg_begin[2], len0;   // candidate index of key n of segment s = n - seg_begin[s] + (s ? len0 : 0)
};

// host side: the lists a selection launch left in its workspace
inline SplitLists split_lists_of(const Workspace &ws, int n_lists, int n_pub, int hw, bool groups,
                                 const vosmem_select_desc *d) {
  SplitLists L{};
  L.cand = ws.cand;
  L.cand_count = ws.cand_count;
  L.pub = ws.pub;
  L.splits = n_lists;
  L.pub_rows = n_pub;
  L.hw_pad = (int)round_up64(hw, TQ);
  L.ctl = ws.ctl;
  L.gscore = groups ? ws.gscore : nullptr;
  if (groups && d != nullptr) {
    L.seg_begin[0] = (int)d->seg[0].begin;
    L.seg_begin[1] = d->n_segments > 1 ? (int)d->seg[1].begin : 0;
    L.len0 = (int)(d->seg[0].end - d->seg[0].begin);
  }
  return L;
}

// Folds `buffered` parked candidates into the running best 32.  Out of line and by value: the sorting network is
// long, and merge_query has many call sites for it (instruction-cache footprint of the readout kernel).
static __device__ __noinline__ WarpTop32 drain_buffer(WarpTop32 top, const float *buf_s, const int *buf_i, int buffered, int lane) {
  __syncwarp();
  for (int off = 0; off < buffered; off += 32) {
    const bool ok = off + lane < buffered;
    top.push(ok ? buf_s[off + lane] : -INFINITY, ok ? buf_i[off + lane] : 0x7fffffff, lane);
  }
  __syncwarp();
  return top;
}

// One warp folds the lists of query q into the exact best 32 (best first, one per lane).
// All loads (<= MERGE_MAX_SPLITS x 64 slots per batch) are issued up front; candidates below the shared threshold
// (min over splits of the published r-th best: a lower bound of the true 32nd best, see select_tc.cu) are dropped
// before the sorting network sees them, which usually leaves 32-64 survivors.  buf_s / buf_i: MERGE_BUF entries of
// warp-private shared memory.
// MB: lists whose first 32 slots are loaded together (registers: 2 * MB per lane).
// GR: group mode (entries per row and the index payload differ; a compile-time switch so that the single-key path
// keeps its constant row stride).
template <int MB = MERGE_MAX_SPLITS, bool GR = false>
__device__ __forceinline__ WarpTop32 merge_query(const SplitLists &L, int q, float *buf_s, int *buf_i, int lane) {
  constexpr int SLOTS = GR ? GSLOTS : CAND_SLOTS;
  WarpTop32 top;
  top.init();
  int buffered = 0;
  auto drain = [&]() {
    top = drain_buffer(top, buf_s, buf_i, buffered, lane);
    buffered = 0;
  };
  float tau = INFINITY;
  bool tau_ready = false;
  for (int y0 = 0; y0 < L.splits; y0 += MB) {
    // first 32 slots of every list of the batch + the counts, all in flight together; lists longer than 32
    // entries (rare once the shared thresholds work) get a second, conditional read
    CandEntry c[MB];
    int my_cnt = 0;
    if (y0 + lane < L.splits && lane < MB) my_cnt = L.cand_count[(int64_t)(y0 + lane) * L.hw_pad + q];
#pragma unroll
    for (int y = 0; y < MB; ++y) {
      if (y0 + y < L.splits) {
        const int64_t row = ((int64_t)(y0 + y) * L.hw_pad + q) * SLOTS;
        c[y] = L.cand[row + lane];
        if (GR) c[y].index = ((y0 + y) << 6) | lane;   // group mode: where the group's record sits
      }
    }
    if (!tau_ready) {  // lane y owns list y
      if (L.pub_rows > 0) {
        const uint32_t epoch = __ldcg(&L.ctl->last);
        for (int y = lane; y < L.pub_rows; y += 32) tau = fminf(tau, pub_load(L.pub + (int64_t)y * L.hw_pad + q, epoch));
        tau = warp_min(tau);
      } else {
        tau = -INFINITY;
      }
      tau_ready = true;
    }
    auto offer = [&](bool keep, float sc, int ix) {
      const unsigned m = __ballot_sync(FULL, keep);
      if (m == 0) return;
      const int add = __popc(m);
      if (buffered + add > MERGE_BUF) drain();
      if (keep) {
        const int pos = buffered + __popc(m & ((1u << lane) - 1));
        buf_s[pos] = sc;
        buf_i[pos] = ix;
      }
      buffered += add;
    };
    // A cut above the shared threshold first: the lists typically hold 200-300 candidates at or above tau, of which
    // 32 are wanted; every 32 that reach the sorting network cost a full bitonic pass.  Bisect (on the value axis,
    // counting with one warp-wide add per step) for a cut T with 32 <= #{score >= T} <= 40 among the entries in
    // registers.  #{score >= T} >= 32 is kept invariant, so no member of the best 32 is lost.
    int n_mine = 0;
    float s_hi = -INFINITY, s_lo = INFINITY;
#pragma unroll
    for (int y = 0; y < MB; ++y) {
      const int cnt = __shfl_sync(FULL, my_cnt, y);
      const bool ok = y0 + y < L.splits && lane < cnt && c[y].score >= tau && c[y].index != 0x7fffffff;
      if (!ok) c[y].score = -INFINITY;
      else { ++n_mine; s_hi = fmaxf(s_hi, c[y].score); s_lo = fminf(s_lo, c[y].score); }
    }
    float cut = tau;
    if (__reduce_add_sync(FULL, n_mine) > 40) {
      float lo = warp_min(s_lo), hi = warp_max(s_hi);      // #{score >= lo} = all of them >= 32
      for (int it = 0; it < 16 && lo < hi; ++it) {
        const float mid = 0.5f * (lo + hi);
        int n = 0;
#pragma unroll
        for (int y = 0; y < MB; ++y) n += c[y].score >= mid ? 1 : 0;
        n = __reduce_add_sync(FULL, n);
        if (n >= 32) {
          lo = mid;
          if (n <= 40) break;
        } else {
          hi = mid;
        }
      }
      cut = fmaxf(cut, lo);
    }
#pragma unroll
    for (int y = 0; y < MB; ++y) {
      if (y0 + y >= L.splits) break;
      const int cnt = __shfl_sync(FULL, my_cnt, y);
      offer(c[y].score >= cut && c[y].score != -INFINITY, c[y].score, c[y].index);
      for (int off = 32; off < cnt; off += 32) {   // rare: a list longer than 32 entries
        const int64_t row = ((int64_t)(y0 + y) * L.hw_pad + q) * SLOTS;
        CandEntry c2 = L.cand[row + (off + lane < SLOTS ? off + lane : 0)];
        if (GR) c2.index = ((y0 + y) << 6) | (off + lane);
        offer(off + lane < cnt && c2.score >= cut && c2.index != 0x7fffffff, c2.score, c2.index);
      }
    }
  }
  drain();
  return top;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same merge for the common case -- at most 12 lists, at most MERGE_ROWS rows of 32 candidates in all -- written for
// latency and a small instruction footprint: merge_query above executes ~1 200 instructions ONCE per query out of ~6 000
// (every phase unrolled over the lists, the sorting networks inlined), and a third of its stall samples in the fused
// readout were instruction fetch (profiles/r2_davis5_ncu_full_summary.txt).  Here the candidates sit in the warp's
// shared-memory scratch (row = list or list continuation, column = lane) and every phase is a short rolled loop:
//   1. all loads up front (counts, thresholds, the first 32 slots of every list); longer lists get continuation rows
//   2. drop what is below the shared threshold; bisect on the value axis (one warp-wide add per step) for a cut that
//      leaves between top_k and 32 candidates -- no sorting network
//   3. compact the survivors to one per lane (ballot + popc), rank them against each other with 32 shuffles, and
//      scatter by rank: the result is the exact best <= 32, best first, like merge_query's.
// Returns false (nothing else touched but the scratch) when the case does not fit or the bisection cannot separate the
// candidates (exact ties around the cut): the caller then runs merge_query.
__device__ __forceinline__ bool merge_query_small(const SplitLists &L, int q, int top_k, float *bs, int *bi, int lane,
                                                  WarpTop32 &out) {
  const int n_l = L.splits;
  if (n_l > 12) return false;
  // ---- 1. loads ----
  int my_cnt = 0;
  if (lane < n_l) my_cnt = L.cand_count[(int64_t)lane * L.hw_pad + q];
  float tau = INFINITY;
  if (L.pub_rows > 0) {
    const uint32_t epoch = __ldcg(&L.ctl->last);
    for (int y = lane; y < L.pub_rows; y += 32) tau = fminf(tau, pub_load(L.pub + (int64_t)y * L.hw_pad + q, epoch));
  } else {
    tau = -INFINITY;
  }
#pragma unroll
  for (int y = 0; y < 12; ++y) {
    if (y < n_l) {
      const CandEntry e = L.cand[((int64_t)y * L.hw_pad + q) * CAND_SLOTS + lane];
      bs[y * 32 + lane] = e.score;
      bi[y * 32 + lane] = e.index;
    }
  }
  tau = warp_min(tau);
  int row_cnt = my_cnt < 32 ? my_cnt : 32;      // lane r: valid entries of row r
  int rows = n_l;
  if (__any_sync(FULL, my_cnt > 32)) {          // lists longer than 32 entries: continuation rows (second round trip)
    for (int y = 0; y < n_l; ++y) {
      const int cnt = __shfl_sync(FULL, my_cnt, y);
      for (int off = 32; off < cnt; off += 32) {
        if (rows == MERGE_ROWS) return false;
        const CandEntry e = L.cand[((int64_t)y * L.hw_pad + q) * CAND_SLOTS + (off + lane < CAND_SLOTS ? off + lane : 0)];
        bs[rows * 32 + lane] = e.score;
        bi[rows * 32 + lane] = e.index;
        if (lane == rows) row_cnt = cnt - off < 32 ? cnt - off : 32;
        ++rows;
      }
    }
  }
  // ---- 2. threshold, then a cut with top_k <= #{score >= cut} <= 32 (columns are lane-private up to the compaction) ----
  int n_mine = 0;
  float s_hi = -INFINITY, s_lo = INFINITY;
  for (int r = 0; r < rows; ++r) {
    const int c = __shfl_sync(FULL, row_cnt, r);
    float v = bs[r * 32 + lane];
    if (lane >= c || !(v >= tau) || bi[r * 32 + lane] == 0x7fffffff) v = -INFINITY;
    bs[r * 32 + lane] = v;
    if (v > -INFINITY) { ++n_mine; s_hi = fmaxf(s_hi, v); s_lo = fminf(s_lo, v); }
  }
  int n = __reduce_add_sync(FULL, n_mine);
  float cut = -INFINITY;       // every finite candidate
  if (n > 32) {
    float lo = warp_min(s_lo), hi = warp_max(s_hi);   // #{score >= lo} = n >= top_k is kept invariant
    bool found = false;
    for (int it = 0; it < 26 && lo < hi; ++it) {
      const float mid = 0.5f * (lo + hi);
      int c = 0;
      for (int r = 0; r < rows; ++r) c += bs[r * 32 + lane] >= mid ? 1 : 0;
      c = __reduce_add_sync(FULL, c);
      if (c >= top_k) {
        lo = mid;
        if (c <= 32) { found = true; n = c; break; }
      } else {
        hi = mid;
      }
    }
    if (!found) return false;
    cut = lo;
  }
  // ---- 3. compaction (one survivor per lane), ranks, scatter by rank ----
  float *cs = bs + MERGE_ROWS * 32;
  int *ci = bi + MERGE_ROWS * 32;
  int pos = 0;
  for (int r = 0; r < rows; ++r) {
    const float v = bs[r * 32 + lane];
    const bool keep = v >= cut && v > -INFINITY;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) {
      const int dst = pos + __popc(m & ((1u << lane) - 1));
      cs[dst] = v;
      ci[dst] = bi[r * 32 + lane];
    }
    pos += __popc(m);
  }
  __syncwarp();
  float s = lane < pos ? cs[lane] : -INFINITY;
  int i = lane < pos ? ci[lane] : 0x7fffffff;
  int rank = 0;
#pragma unroll 4
  for (int t = 0; t < 32; ++t) {
    const float os = __shfl_sync(FULL, s, t);
    const int oi = __shfl_sync(FULL, i, t);
    rank += better(os, oi, s, i) ? 1 : 0;
  }
  if (lane >= pos) rank = lane;   // empty lanes keep their place behind the `pos` survivors: the ranks are a permutation
  __syncwarp();
  cs[rank] = s;
  ci[rank] = i;
  __syncwarp();
  out.s = cs[lane];
  out.i = ci[lane];
  __syncwarp();
  return true;
}

// Group mode: lane l holds the l-th best GROUP of query q (top.s = its maximum, top.i = list << 6 | slot, or
// 0x7fffffff for none).  The best k keys of the query all sit in its best k groups (a key's group scores at least the
// key), so: read the 8 scores of the lane's group, take the exact best 32 of the 32 group maxima plus every other
// score that reaches the k-th best maximum (a lower bound of the k-th best key).  Returns keys: (score, candidate index).
// (out of line and by value, like drain_buffer: keeps the register budget of the readout kernel's gather loop intact)
static __device__ __noinline__ WarpTop32 expand_groups(const SplitLists L, int q, const WarpTop32 g, int top_k, int lane) {
  float sc[GROUP_KEYS];
  int idx0 = 0;
  const bool have = g.i != 0x7fffffff;
  if (have) {
    const int y = g.i >> 6, slot = g.i & 63;
    const int64_t at = ((int64_t)y * L.hw_pad + q) * GSLOTS + slot;
    const uint32_t loc = (uint32_t)__ldcg(&L.cand[at].index);
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(L.gscore + at * GROUP_KEYS));
    const float4 b = __ldcg(reinterpret_cast<const float4 *>(L.gscore + at * GROUP_KEYS) + 1);
    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w;
    sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
    const int s = loc >> 31;
    idx0 = (int)((loc & ~GROUP_SEG_BIT) * GROUP_KEYS) - L.seg_begin[s] + (s ? L.len0 : 0);
  } else {
#pragma unroll
    for (int j = 0; j < GROUP_KEYS; ++j) sc[j] = -INFINITY;
  }
  // the lane's best key first
  float m1 = sc[0];
  int j1 = 0;
#pragma unroll
  for (int j = 1; j < GROUP_KEYS; ++j)
    if (sc[j] > m1) { m1 = sc[j]; j1 = j; }
  WarpTop32 top;
  top.init();
  top.push(m1, m1 > -INFINITY ? idx0 + j1 : 0x7fffffff, lane);
  const float floor_k = __shfl_sync(FULL, top.s, top_k - 1);
  bool more = false;
#pragma unroll
  for (int j = 0; j < GROUP_KEYS; ++j) more = more || (j != j1 && sc[j] >= floor_k && sc[j] > -INFINITY);
  if (__any_sync(FULL, more)) {
#pragma unroll 1
    for (int j = 0; j < GROUP_KEYS; ++j) {
      float v = -INFINITY;
#pragma unroll
      for (int u = 0; u < GROUP_KEYS; ++u) v = (u == j) ? sc[u] : v;   // (no dynamic register indexing)
      const bool ok = j != j1 && v >= floor_k && v > -INFINITY;
      if (__any_sync(FULL, ok)) top.push(ok ? v : -INFINITY, ok ? idx0 + j : 0x7fffffff, lane);
    }
  }
  return top;
}

// group mode, out of line: the single-key path of the callers keeps the registers and code layout it had without it
template <int MB>
static __device__ __noinline__ WarpTop32 merge_groups_of_query(const SplitLists L, int q, int top_k, float *buf_s, int *buf_i,
                                                               int lane) {
  return expand_groups(L, q, merge_query<MB, true>(L, q, buf_s, buf_i, lane), top_k, lane);
}

// the general merge, out of line: what merge_query_small hands over to
template <int MB>
static __device__ __noinline__ WarpTop32 merge_keys_of_query(const SplitLists L, int q, float *buf_s, int *buf_i, int lane) {
  return merge_query<MB, false>(L, q, buf_s, buf_i, lane);
}

// merge + (group mode) expansion: what every consumer of the selection's lists calls
template <int MB = MERGE_MAX_SPLITS>
__device__ __forceinline__ WarpTop32 merge_lists_of_query(const SplitLists &L, int q, int top_k, float *buf_s, int *buf_i,
                                                          int lane) {
  if (L.gscore == nullptr) {
    WarpTop32 top;
    if (merge_query_small(L, q, top_k, buf_s, buf_i, lane, top)) return top;
    return merge_keys_of_query<MB>(L, q, buf_s, buf_i, lane);       // many lists, or exact ties around the cut
  }
  return merge_groups_of_query<MB>(L, q, top_k, buf_s, buf_i, lane);
}

}  // namespace vosmem
