// Shared device helpers for the memory-readout kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vosmem.h"

namespace vosmem {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CK_TC = 64;                       // channel count the tensor-core image is built for
constexpr int TK = VOSMEM_KEY_TILE;             // keys per image tile
constexpr int TQ = VOSMEM_QUERY_TILE;           // queries per image tile
constexpr int K8 = 34;                          // 16-byte K chunks per operand row: 16 hi + 16 lo + 2 tail
constexpr int KEY_TILE_BYTES = K8 * (TK / 8) * 128;    // 34816
constexpr int QUERY_TILE_BYTES = K8 * (TQ / 8) * 128;  // 69632
constexpr int CAND_SLOTS = 120;                 // per (split, query) candidate slots in the exchange buffer (2 x 60)
constexpr int MAX_SPLITS = 32;
constexpr int MAX_BATCH = VOSMEM_MAX_BATCH;      // independent problems (sequences) one launch can carry (blockIdx.z)
constexpr int LISTS_PER_SPLIT = 2;              // the tcgen05 kernel publishes two thresholds per (split, query)
// Group candidates (select_tc.cu, GROUPS mode): a candidate is a group of 8 consecutive keys of the packed image,
// scored by the maximum of its 8 similarities; all 8 are kept next to the entry so that the consumer can expand the
// best 32 groups of a query into its best k keys without touching the keys again (merge.cuh, expand_groups).
constexpr int GROUP_KEYS = 8;
constexpr int GSLOTS = 60;                      // group entries per (virtual split, query)
static_assert(LISTS_PER_SPLIT * GSLOTS == CAND_SLOTS, "group lists reuse the exchange rows of the element lists");
#define VOSMEM_GROUP_LOG 192                    // records of the per-list score log of the selection kernel (two halves)

void set_error(const char *fmt, ...);
#define VOSMEM_CHECK_ARG(cond, ...)              \
  do {                                           \
    if (!(cond)) {                               \
      ::vosmem::set_error(__VA_ARGS__);          \
      return VOSMEM_EINVAL;                      \
    }                                            \
  } while (0)
#define VOSMEM_CUDA(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      ::vosmem::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return VOSMEM_ECUDA;                                                                  \
    }                                                                                       \
  } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

// ------------------------------------------------------------------------------------------------
// order-preserving float <-> uint map (for atomicMax on thresholds)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
constexpr unsigned ORD_NEG_INF = 0x007fffffu;   // f2ord(-inf)

// ------------------------------------------------------------------------------------------------
// A candidate is (score, index); "better" = higher score, ties to the lower index (total order, so
// every sort below is deterministic).
__device__ __forceinline__ bool better(float s, int i, float os, int oi) { return s > os || (s == os && i < oi); }

// compare-exchange with lane ^ j; keep_best: this lane keeps the better of the pair
__device__ __forceinline__ void cmpx(float &s, int &i, int j, bool keep_best) {
  float os = __shfl_xor_sync(FULL, s, j);
  int oi = __shfl_xor_sync(FULL, i, j);
  bool mine = better(s, i, os, oi);
  if (mine != keep_best) { s = os; i = oi; }
}

// Sort 32 candidates (one per lane) best-first.
__device__ __forceinline__ void warp_sort_desc(float &s, int &i, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      bool desc = (k == 32) || ((lane & k) == 0);
      bool low = (lane & j) == 0;
      cmpx(s, i, j, low == desc);
    }
  }
}

// Bitonic sequence across lanes -> best-first.
__device__ __forceinline__ void warp_bitonic_merge_desc(float &s, int &i, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) cmpx(s, i, j, (lane & j) == 0);
}

// Running best-32 of a stream, one survivor per lane, best-first.  `push` takes one candidate per
// lane; it costs one vote when nothing beats the current 32nd.
struct WarpTop32 {
  float s;
  int i;
  __device__ __forceinline__ void init() { s = -INFINITY; i = 0x7fffffff; }
  __device__ __forceinline__ float floor_score() const { return __shfl_sync(FULL, s, 31); }
  __device__ __forceinline__ void push(float cs, int ci, int lane) {
    float tau = __shfl_sync(FULL, s, 31);
    int ti = __shfl_sync(FULL, i, 31);
    if (!__any_sync(FULL, better(cs, ci, tau, ti))) return;
    warp_sort_desc(cs, ci, lane);
    // reverse the batch so list (descending) ++ batch (ascending) is bitonic, keep the better half
    float rs = __shfl_sync(FULL, cs, 31 - lane);
    int ri = __shfl_sync(FULL, ci, 31 - lane);
    if (better(rs, ri, s, i)) { s = rs; i = ri; }
    warp_bitonic_merge_desc(s, i, lane);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) v += __shfl_xor_sync(FULL, v, j);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, j));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, j));
  return v;
}

// ------------------------------------------------------------------------------------------------
// bf16 hi/lo split: v ~= hi + lo with ~16 mantissa bits
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// byte offset of (row r, 16-byte chunk k8) inside an operand tile image of ROWS rows
template <int ROWS>
__host__ __device__ inline int image_offset(int r, int k8) {
  return (k8 * (ROWS / 8) + (r >> 3)) * 128 + (r & 7) * 16;
}

// ------------------------------------------------------------------------------------------------
// Workspace carving shared by host and kernels.
struct CandEntry {   // one exchanged candidate: score + index on the candidate axis (0x7fffffff = none)
  float score;
  int index;
};
static_assert(sizeof(CandEntry) == 8, "exchange entries are read / written as 8-byte words");
// Group mode: a list entry is CandEntry{maximum of the group's 8 scores, locator} with locator bit 31 = segment,
// bits 0-30 = group index inside the bank (first key / 8); the 8 scores (-inf for keys outside the segment's candidate
// range) sit at the same (list, query, slot) position of a parallel array of 32-byte records.
constexpr uint32_t GROUP_SEG_BIT = 0x80000000u;

// A published threshold (select_tc.cu, "Thresholds") is only meaningful for the launch that wrote it: the entry
// carries that launch's epoch and reads of any other epoch see -inf, so the workspace needs no reset between
// calls.  The epoch is a DEVICE-side word of the workspace (WsControl): every CTA of the selection kernel reads
// it at entry, the last CTA to leave advances it (and records the tag it ran under in WsControl::last, which is
// what the consumer -- merge / fused readout -- validates against).  A captured CUDA graph can therefore be replayed any number of times with new query or memory
// contents: every replay runs under a fresh epoch (a host-drawn epoch would be baked into the graph).
struct PubEntry {
  float value;
  uint32_t epoch;
};
static_assert(sizeof(PubEntry) == 8, "published thresholds are read / written as 8-byte words");
__device__ __forceinline__ float pub_load(const PubEntry *p, uint32_t epoch) {
  const uint2 v = __ldcg(reinterpret_cast<const uint2 *>(p));
  return v.y == epoch ? __uint_as_float(v.x) : -INFINITY;
}
__device__ __forceinline__ void pub_store(PubEntry *p, float value, uint32_t epoch) {
  __stcg(reinterpret_cast<uint2 *>(p), make_uint2(__float_as_uint(value), epoch));
}

// Control words at the head of a workspace (zeroed once by vosmem_workspace_init).
struct WsControl {
  uint32_t epoch;     // tag of the NEXT tcgen05 selection launch on this workspace (first word: see vosmem_workspace_init)
  uint32_t departed;  // CTAs of the running selection launch that have left (reset by the last one)
  uint32_t error;     // sticky device-side error flags (WS_ERR_*), read back by vosmem_workspace_status
  uint32_t last;      // tag the most recent selection launch ran under: what its consumer validates against
};
constexpr uint32_t WS_ERR_TMEM_BASE = 1u;   // tensor-memory allocation did not start at column 0

struct Workspace {
  WsControl *ctl;              // control words (first 256 bytes)
  unsigned char *query_image;  // n_qtiles * QUERY_TILE_BYTES
  PubEntry *pub;               // pub_rows * hw_pad: lower bound published per (virtual split, query) by the current launch
  PubEntry *pub2;              // the same for the (smaller) rank used when thresholds are shared across ranks
  CandEntry *cand;             // splits * hw_pad * CAND_SLOTS
  int *cand_count;             // splits * LISTS_PER_SPLIT * hw_pad
  float *glog;                 // splits * LISTS_PER_SPLIT * hw_pad * VOSMEM_GROUP_LOG x 8 scores: the selection kernel's
                               // per-list score logs (scratch of that kernel only)
  float *gscore;               // splits * LISTS_PER_SPLIT * hw_pad * GSLOTS x 8 scores (tcgen05 kernel, group mode; the
                               // entries themselves then use `cand` as [virtual split][hw_pad][GSLOTS])
  int pub_rows;                // rows of `pub` (= splits_cap * LISTS_PER_SPLIT)
  float *qvec;                 // SIMT path: (2*ck + 1) x hw_pad, c-major  [-e | 2*q*e | -sum e q^2]
  int64_t bytes;
};

// Upper bound on the N-splits either kernel path will ever use for `hw` queries (sizes the exchange buffers).
inline int splits_cap(int hw) {
  int64_t n_qtiles = ceil_div64(hw, TQ);
  int64_t a = 148 / n_qtiles, b = ceil_div64(4736, hw);
  int64_t s = a > b ? a : b;
  if (s < 4) s = 4;  // room for >65536 keys per launch (16-bit in-CTA candidate index)
  if (s > MAX_SPLITS) s = MAX_SPLITS;
  return (int)s;
}

// groups: room for the group-candidate mode of the tcgen05 selection (score logs + exchanged scores; ~140 KB per query)
inline Workspace carve_workspace(void *base, int ck, int hw, bool groups) {
  Workspace w;
  const int64_t cap = (int64_t)splits_cap(hw) * LISTS_PER_SPLIT;   // candidate lists per query
  int64_t n_qtiles = ceil_div64(hw, TQ);
  int64_t hw_pad = n_qtiles * TQ;
  unsigned char *p = static_cast<unsigned char *>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    unsigned char *r = p ? p + off : nullptr;
    off += round_up64(bytes, 256);
    return r;
  };
  w.ctl = reinterpret_cast<WsControl *>(take(sizeof(WsControl)));
  w.query_image = take(n_qtiles * QUERY_TILE_BYTES);
  w.pub = reinterpret_cast<PubEntry *>(take(cap * hw_pad * 8));
  w.pub2 = reinterpret_cast<PubEntry *>(take(cap * hw_pad * 8));
  w.pub_rows = (int)cap;
  w.cand = reinterpret_cast<CandEntry *>(take((int64_t)splits_cap(hw) * hw_pad * CAND_SLOTS * 8));
  w.cand_count = reinterpret_cast<int *>(take(cap * hw_pad * 4));
  w.gscore = reinterpret_cast<float *>(take(groups ? cap * hw_pad * GSLOTS * GROUP_KEYS * 4 : 0));
  w.glog = reinterpret_cast<float *>(take(groups ? cap * hw_pad * VOSMEM_GROUP_LOG * GROUP_KEYS * 4 : 0));
  w.qvec = reinterpret_cast<float *>(take((int64_t)hw_pad * (2 * ck + 1) * 4));
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------------------------------------
// launchers implemented in the .cu files
struct SelectPlan {
  int splits;          // N-splits (gridDim.y)
  int64_t n_total;     // candidates over all segments
};

int launch_pack_query(const float *qk, const float *qe, int ck, int hw, const Workspace &ws, cudaStream_t st);
int launch_select_simt(const vosmem_select_desc &d, const Workspace &ws, int splits, cudaStream_t st);
// Thresholds shared ACROSS ranks of an N-sharded bank (select_tc.cu, "Thresholds across ranks"): rank_pub[d] is the
// [world][hw_pad] array of rank summaries in rank d's (peer-mapped) memory; row `rank` of every array is written by
// this rank, the rows of the local array (rank_pub[rank]) are read.
struct PeerThresholds {
  int world = 1, rank = 0;
  PubEntry *rank_pub[VOSMEM_MAX_RANKS] = {};
};
int launch_select_tc(const vosmem_select_desc *d, const Workspace *ws, int n, int splits, bool groups, cudaStream_t st,
                     const PeerThresholds *peers = nullptr);
struct SplitLists;
int launch_merge_splits(const SplitLists &lists, int hw, int top_k, int64_t index_base, float *out_score,
                        int64_t *out_index, cudaStream_t st);
int choose_splits(int path, int hw, int64_t n_total, int batch = 1);

}  // namespace vosmem
