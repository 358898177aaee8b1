// Exact fp32 similarity + per-query top-k on CUDA cores, and the split/rank merge kernels.
//
// This is the any-CK path (and the on-device cross-check for the tcgen05 path): one warp per query,
// lanes stride over memory keys, the warp keeps a running best-32 (WarpTop32).  The N x HW similarity
// matrix of the reference (memory_util.py:7-39) is never written.
#include "common.cuh"
#include "merge.cuh"

namespace vosmem {

namespace {

struct SimtSeg {
  const float *key;
  int64_t key_ld;
  const float *shrinkage;
  int64_t begin;  // first candidate key of the bank
};

struct SimtArgs {
  SimtSeg seg[2];
  int64_t len0, n_total;  // candidates in seg0, in total
  int ck, hw, hw_pad, splits;
  float inv_sqrt_ck;
  const float *qvec;
  CandEntry *cand;
  int *cand_count;
};

constexpr int SIMT_WARPS = 8;

__global__ void __launch_bounds__(SIMT_WARPS * 32) select_simt_kernel(SimtArgs a) {
  extern __shared__ float smem[];  // SIMT_WARPS * (2*ck+1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * SIMT_WARPS + warp;
  const int vlen = 2 * a.ck + 1;
  float *qv = smem + warp * vlen;
  if (q < a.hw)
    for (int c = lane; c < vlen; c += 32) qv[c] = a.qvec[(int64_t)c * a.hw_pad + q];
  __syncwarp();
  if (q >= a.hw) return;

  // candidate range of this split, in whole warps of 32
  const int64_t per = ceil_div64(ceil_div64(a.n_total, a.splits), 32) * 32;
  const int64_t lo = per * blockIdx.y;
  const int64_t hi = lo + per < a.n_total ? lo + per : a.n_total;

  WarpTop32 top;
  top.init();
  const float y3 = qv[2 * a.ck];
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t cand = base + lane;
    const bool valid = cand < hi;
    float score = -INFINITY;
    if (valid) {
      const int s = cand >= a.len0;
      const SimtSeg &sg = a.seg[s];
      const int64_t n = sg.begin + (s ? cand - a.len0 : cand);
      const float *kp = sg.key + n;
      float acc1 = 0.f, acc2 = 0.f;  // two chains for ILP
#pragma unroll 4
      for (int c = 0; c < a.ck; ++c) {
        float m = kp[(int64_t)c * sg.key_ld];
        acc1 = fmaf(m * m, qv[c], acc1);
        acc2 = fmaf(m, qv[a.ck + c], acc2);
      }
      const float scale = (sg.shrinkage ? sg.shrinkage[n] : 1.0f) * a.inv_sqrt_ck;
      score = (acc1 + acc2 + y3) * scale;
    }
    top.push(score, valid ? (int)cand : 0x7fffffff, lane);
  }
  const int64_t slot = ((int64_t)blockIdx.y * a.hw_pad + q) * CAND_SLOTS;
  a.cand[slot + lane] = CandEntry{top.s, top.i};
  if (lane == 0) a.cand_count[(int64_t)blockIdx.y * a.hw_pad + q] = 32;
}

// One warp per query: fold the per-split candidate lists into the exact top-k, best first (merge.cuh).
template <int MB>   // split lists merged per batch (see merge_push_kernel)
__global__ void __launch_bounds__(256, MB <= 4 ? 4 : 2) merge_splits_kernel(SplitLists L, int hw, int top_k, int64_t index_base,
                                                           float *__restrict__ out_score,
                                                           int64_t *__restrict__ out_index) {
  __shared__ float buf_s[8][MERGE_BUF];
  __shared__ int buf_i[8][MERGE_BUF];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  if (q >= hw) return;
  const WarpTop32 top = merge_lists_of_query<MB>(L, q, top_k, buf_s[warp], buf_i[warp], lane);
  if (lane < top_k) {
    const bool have = top.i != 0x7fffffff;
    out_score[(int64_t)q * top_k + lane] = have ? top.s : -INFINITY;
    out_index[(int64_t)q * top_k + lane] = have ? (int64_t)top.i + index_base : -1;
  }
}

// Merge `n_lists` already-selected lists (e.g. one per rank), each [HW][top_k] behind its own pointer: the lists may
// live in different allocations -- in particular in the peer-mapped memory of other GPUs (NVLink loads), which
// fuses the candidate all-gather of the sharded long-term readout into this kernel.
struct ListPtrs {
  const float *score[VOSMEM_MAX_LISTS];
  const int64_t *index[VOSMEM_MAX_LISTS];
};

__global__ void merge_lists_kernel(const __grid_constant__ ListPtrs lists, int n_lists, int hw, int top_k,
                                   float *__restrict__ out_score, int64_t *__restrict__ out_index) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= hw) return;
  // every list's entry for this lane first (independent loads: remote lists cost an NVLink round trip each)
  float s[VOSMEM_MAX_LISTS];
  int i[VOSMEM_MAX_LISTS];
  const int64_t at = (int64_t)q * top_k + lane;
#pragma unroll
  for (int l = 0; l < VOSMEM_MAX_LISTS; ++l) {
    s[l] = -INFINITY;
    i[l] = 0x7fffffff;
    if (l < n_lists && lane < top_k) {
      const int64_t gi = __ldcg(lists.index[l] + at);
      const float sc = __ldcg(lists.score[l] + at);
      if (gi >= 0) { s[l] = sc; i[l] = (int)gi; }
    }
  }
  WarpTop32 top;
  top.init();
#pragma unroll
  for (int l = 0; l < VOSMEM_MAX_LISTS; ++l)
    if (l < n_lists) top.push(s[l], i[l], lane);
  if (lane < top_k) {
    const bool have = top.i != 0x7fffffff;
    out_score[(int64_t)q * top_k + lane] = have ? top.s : -INFINITY;
    out_index[(int64_t)q * top_k + lane] = have ? (int64_t)top.i : -1;
  }
}

}  // namespace

int launch_select_simt(const vosmem_select_desc &d, const Workspace &ws, int splits, cudaStream_t st) {
  SimtArgs a{};
  int64_t total = 0;
  for (int s = 0; s < 2; ++s) {
    if (s < d.n_segments) {
      const vosmem_segment &g = d.seg[s];
      VOSMEM_CHECK_ARG(g.key != nullptr || g.end == g.begin, "select(SIMT): segment %d has no fp32 keys", s);
      a.seg[s] = SimtSeg{g.key, g.key_ld, g.shrinkage, g.begin};
      if (s == 0) a.len0 = g.end - g.begin;
      total += g.end - g.begin;
    }
  }
  if (d.n_segments == 1) a.seg[1] = a.seg[0];
  a.n_total = total;
  a.ck = d.ck;
  a.hw = d.hw;
  a.hw_pad = (int)round_up64(d.hw, TQ);
  a.splits = splits;
  a.inv_sqrt_ck = 1.0f / sqrtf((float)d.ck);
  a.qvec = ws.qvec;
  a.cand = ws.cand;
  a.cand_count = ws.cand_count;
  dim3 grid((d.hw + SIMT_WARPS - 1) / SIMT_WARPS, splits);
  size_t smem = (size_t)SIMT_WARPS * (2 * d.ck + 1) * sizeof(float);
  select_simt_kernel<<<grid, SIMT_WARPS * 32, smem, st>>>(a);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

int launch_merge_splits(const SplitLists &L, int hw, int top_k, int64_t index_base, float *out_score,
                        int64_t *out_index, cudaStream_t st) {
  if (L.splits <= 4) merge_splits_kernel<4><<<(hw + 7) / 8, 256, 0, st>>>(L, hw, top_k, index_base, out_score, out_index);
  else merge_splits_kernel<MERGE_MAX_SPLITS><<<(hw + 7) / 8, 256, 0, st>>>(L, hw, top_k, index_base, out_score, out_index);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

}  // namespace vosmem

using namespace vosmem;

static int launch_merge_lists(const ListPtrs &lists, int n_lists, int hw, int top_k, float *out_score, int64_t *out_index,
                              cudaStream_t st) {
  VOSMEM_CHECK_ARG(out_score && out_index, "merge_topk: null output");
  VOSMEM_CHECK_ARG(n_lists >= 1 && n_lists <= VOSMEM_MAX_LISTS && hw >= 1, "merge_topk: n_lists=%d (max %d) hw=%d", n_lists,
                   VOSMEM_MAX_LISTS, hw);
  VOSMEM_CHECK_ARG(top_k >= 1 && top_k <= VOSMEM_MAX_TOPK, "merge_topk: top_k=%d outside [1, %d]", top_k, VOSMEM_MAX_TOPK);
  merge_lists_kernel<<<(hw + 7) / 8, 256, 0, st>>>(lists, n_lists, hw, top_k, out_score, out_index);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_merge_topk(const float *scores, const int64_t *indices, int n_lists, int hw, int top_k,
                                 float *out_score, int64_t *out_index, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(scores && indices, "vosmem_merge_topk: null pointer");
  VOSMEM_CHECK_ARG(n_lists >= 1 && n_lists <= VOSMEM_MAX_LISTS, "vosmem_merge_topk: n_lists=%d outside [1, %d]", n_lists,
                   VOSMEM_MAX_LISTS);
  ListPtrs lists{};
  for (int l = 0; l < n_lists; ++l) {
    lists.score[l] = scores + (int64_t)l * hw * top_k;
    lists.index[l] = indices + (int64_t)l * hw * top_k;
  }
  return launch_merge_lists(lists, n_lists, hw, top_k, out_score, out_index, (cudaStream_t)stream);
}

extern "C" int vosmem_merge_topk_ptrs(const float *const *scores, const int64_t *const *indices, int n_lists, int hw,
                                      int top_k, float *out_score, int64_t *out_index, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(scores && indices, "vosmem_merge_topk_ptrs: null pointer array");
  VOSMEM_CHECK_ARG(n_lists >= 1 && n_lists <= VOSMEM_MAX_LISTS, "vosmem_merge_topk_ptrs: n_lists=%d outside [1, %d]", n_lists,
                   VOSMEM_MAX_LISTS);
  ListPtrs lists{};
  for (int l = 0; l < n_lists; ++l) {
    VOSMEM_CHECK_ARG(scores[l] && indices[l], "vosmem_merge_topk_ptrs: list %d is null", l);
    lists.score[l] = scores[l];
    lists.index[l] = indices[l];
  }
  return launch_merge_lists(lists, n_lists, hw, top_k, out_score, out_index, (cudaStream_t)stream);
}
