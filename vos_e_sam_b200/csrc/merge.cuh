// Per-query merge of the N-split candidate lists (shared by merge_splits_kernel and the fused readout).
#pragma once
#include "common.cuh"

namespace vosmem {

constexpr int MERGE_MAX_SPLITS = 16;
constexpr int MERGE_BUF = 128;

struct SplitLists {
  const CandEntry *cand;    // [splits][hw_pad][CAND_SLOTS]
  const int *cand_count;    // [splits][hw_pad]
  const PubEntry *pub;      // [pub_rows][hw_pad] published lower bounds per (virtual split, query) of the selection launch
  int splits, pub_rows, hw_pad;   // pub_rows == 0: the selection published nothing (SIMT path), no threshold
  const WsControl *ctl;     // ctl->last = the epoch the selection launch that filled the lists ran under
};

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
template <int MB = MERGE_MAX_SPLITS>
__device__ __forceinline__ WarpTop32 merge_query(const SplitLists &L, int q, float *buf_s, int *buf_i, int lane) {
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
        const int64_t row = ((int64_t)(y0 + y) * L.hw_pad + q) * CAND_SLOTS;
        c[y] = L.cand[row + lane];
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
        const int64_t row = ((int64_t)(y0 + y) * L.hw_pad + q) * CAND_SLOTS;
        const CandEntry c2 = L.cand[row + off + lane];
        offer(off + lane < cnt && c2.score >= cut && c2.index != 0x7fffffff, c2.score, c2.index);
      }
    }
  }
  drain();
  return top;
}

}  // namespace vosmem
