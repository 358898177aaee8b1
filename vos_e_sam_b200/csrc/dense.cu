// Dense twins of tracker/model/memory_util.py for API parity: get_similarity (:7-39), do_softmax
// (:41-65) and readout / MemoryManager._readout (:73-80, memory_manager.py:53-55).  These materialise
// the N x HW matrix exactly like the reference; the per-frame hot path does not use them (it goes
// through vosmem_select_topk / vosmem_softmax_readout) -- consolidation and training-time reads do.
#include "common.cuh"

namespace vosmem {
namespace {

constexpr int DT = 64;  // output tile edge
constexpr int DK = 16;  // channel chunk

// out[n, q] = scale[n] * ( sum_c k[c,n] * (2 q[c,q] e[c,q] - k[c,n] e[c,q]) - sum_c e[c,q] q[c,q]^2 )
__global__ void __launch_bounds__(256) similarity_dense_kernel(const float *__restrict__ key, int64_t key_ld,
                                                               const float *__restrict__ shrinkage,
                                                               const float *__restrict__ qk,
                                                               const float *__restrict__ qe, int ck, int64_t n,
                                                               int hw, float *__restrict__ out) {
  __shared__ float sk[DK][DT + 1], sq[DK][DT + 1], se[DK][DT + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t n0 = (int64_t)blockIdx.y * DT;
  const int q0 = blockIdx.x * DT;
  float acc[4][4] = {};
  float bsq[4] = {};
  for (int c0 = 0; c0 < ck; c0 += DK) {
    for (int e = threadIdx.x; e < DK * DT; e += 256) {
      int c = e / DT, i = e % DT;
      bool cok = c0 + c < ck;
      sk[c][i] = (cok && n0 + i < n) ? key[(int64_t)(c0 + c) * key_ld + n0 + i] : 0.f;
      bool qok = cok && q0 + i < hw;
      sq[c][i] = qok ? qk[(int64_t)(c0 + c) * hw + q0 + i] : 0.f;
      se[c][i] = qok ? (qe ? qe[(int64_t)(c0 + c) * hw + q0 + i] : 1.0f) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < DK; ++c) {
      float m[4], e[4], t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = sk[c][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        e[j] = se[c][tx * 4 + j];
        float qv = sq[c][tx * 4 + j];
        t[j] = 2.0f * (qv * e[j]);
        bsq[j] = fmaf(e[j], qv * qv, bsq[j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(m[i], t[j] - m[i] * e[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float inv = 1.0f / sqrtf((float)ck);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t nn = n0 + ty * 4 + i;
    if (nn >= n) continue;
    float scale = (shrinkage ? shrinkage[nn] : 1.0f) * inv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int q = q0 + tx * 4 + j;
      if (q < hw) out[nn * hw + q] = (acc[i][j] - (qe ? bsq[j] : 0.f)) * scale;
    }
  }
}

// One warp per query column; 32 warps of a CTA share cache lines of 32 adjacent columns.
__global__ void __launch_bounds__(1024) softmax_topk_dense_kernel(const float *similarity, int64_t sim_ld, int64_t n,
                                                                  int hw, int top_k, float *affinity, int64_t aff_ld,
                                                                  float *usage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 32 + warp;
  if (q >= hw) return;
  WarpTop32 top;
  top.init();
  for (int64_t base = 0; base < n; base += 32) {
    int64_t i = base + lane;
    bool ok = i < n;
    top.push(ok ? similarity[i * sim_ld + q] : -INFINITY, ok ? (int)i : 0x7fffffff, lane);
  }
  // the column is fully consumed: safe to overwrite even when affinity aliases similarity
  const bool live = lane < top_k && top.i != 0x7fffffff;
  // reference top-k branch: exp(v) / sum exp(v) (memory_util.py:48-49); max-subtracted here
  const float m = __shfl_sync(FULL, top.s, 0);
  const float e = live ? expf(top.s - m) : 0.f;
  const float sum = warp_sum(e);
  __syncwarp();
  for (int64_t base = 0; base < n; base += 32) {
    int64_t i = base + lane;
    if (i < n) affinity[i * aff_ld + q] = 0.f;
  }
  __syncwarp();
  if (live) {
    float w = e / sum;
    affinity[(int64_t)top.i * aff_ld + q] = w;
    if (usage) atomicAdd(usage + top.i, w);
  }
}

// Any top_k up to DENSE_MAX_TOPK (the fused per-frame path keeps k <= 32, one survivor per lane; this twin serves
// memory_util.do_softmax callers with larger k): one warp per column extracts the next-best 32 in ceil(k / 32) passes
// ("strictly worse than the last survivor" filters what earlier passes took), parks the survivors in shared memory
// and only then overwrites the column -- so affinity may alias similarity.
constexpr int DENSE_MAX_TOPK = 512;
__global__ void __launch_bounds__(256) softmax_topk_any_kernel(const float *similarity, int64_t sim_ld, int64_t n, int hw,
                                                               int top_k, float *affinity, int64_t aff_ld, float *usage) {
  __shared__ float s_s[8][DENSE_MAX_TOPK];
  __shared__ int s_i[8][DENSE_MAX_TOPK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  if (q >= hw) return;
  float bound_s = INFINITY;   // last survivor of the previous pass: later passes only take candidates worse than it
  int bound_i = -1;
  int kept = 0;
  while (kept < top_k) {
    WarpTop32 top;
    top.init();
    for (int64_t base = 0; base < n; base += 32) {
      const int64_t i = base + lane;
      float s = i < n ? similarity[i * sim_ld + q] : -INFINITY;
      int idx = i < n ? (int)i : 0x7fffffff;
      if (!better(bound_s, bound_i, s, idx)) { s = -INFINITY; idx = 0x7fffffff; }   // taken by an earlier pass
      top.push(s, idx, lane);
    }
    const int take = min(32, top_k - kept);
    if (lane < take) { s_s[warp][kept + lane] = top.s; s_i[warp][kept + lane] = top.i; }
    bound_s = __shfl_sync(FULL, top.s, take - 1);
    bound_i = __shfl_sync(FULL, top.i, take - 1);
    kept += take;
  }
  __syncwarp();
  const float m = s_s[warp][0];
  float sum = 0.f;
  for (int j = lane; j < top_k; j += 32) sum += s_i[warp][j] != 0x7fffffff ? expf(s_s[warp][j] - m) : 0.f;
  sum = warp_sum(sum);
  for (int64_t i = lane; i < n; i += 32) affinity[i * aff_ld + q] = 0.f;
  __syncwarp();
  for (int j = lane; j < top_k; j += 32) {
    if (s_i[warp][j] == 0x7fffffff) continue;
    const float w = expf(s_s[warp][j] - m) / sum;
    affinity[(int64_t)s_i[warp][j] * aff_ld + q] = w;
    if (usage) atomicAdd(usage + s_i[warp][j], w);
  }
}

__global__ void __launch_bounds__(1024) softmax_full_dense_kernel(const float *similarity, int64_t sim_ld, int64_t n,
                                                                  int hw, float *affinity, int64_t aff_ld,
                                                                  float *usage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 32 + warp;
  if (q >= hw) return;
  float m = -INFINITY;
  for (int64_t i = lane; i < n; i += 32) m = fmaxf(m, similarity[i * sim_ld + q]);
  m = warp_max(m);
  float sum = 0.f;
  for (int64_t i = lane; i < n; i += 32) sum += expf(similarity[i * sim_ld + q] - m);
  sum = warp_sum(sum);
  for (int64_t i = lane; i < n; i += 32) {
    float w = expf(similarity[i * sim_ld + q] - m) / sum;
    affinity[i * aff_ld + q] = w;
    if (usage) atomicAdd(usage + i, w);
  }
}

// out[rows x hw] = value[rows x n] @ affinity[n x hw]
__global__ void __launch_bounds__(256) readout_dense_kernel(const float *__restrict__ value, int64_t value_ld,
                                                            const float *__restrict__ aff, int64_t aff_ld, int rows,
                                                            int64_t n, int hw, float *__restrict__ out,
                                                            int64_t out_ld) {
  __shared__ float sv[DT][DK + 1], sa[DK][DT + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int r0 = blockIdx.y * DT, q0 = blockIdx.x * DT;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < n; k0 += DK) {
    for (int e = threadIdx.x; e < DT * DK; e += 256) {
      int r = e / DK, k = e % DK;
      sv[r][k] = (r0 + r < rows && k0 + k < n) ? value[(int64_t)(r0 + r) * value_ld + k0 + k] : 0.f;
      int kk = e / DT, q = e % DT;
      sa[kk][q] = (k0 + kk < n && q0 + q < hw) ? aff[(k0 + kk) * aff_ld + q0 + q] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      float v[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = sv[ty * 4 + i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = sa[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(v[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int q = q0 + tx * 4 + j;
      if (q < hw) out[(int64_t)r * out_ld + q] = acc[i][j];
    }
  }
}

}  // namespace
}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_similarity_dense(const float *key, int64_t key_ld, const float *shrinkage,
                                       const float *query_key, const float *query_selection, int ck, int64_t n, int hw,
                                       float *out, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(key && query_key && out, "vosmem_similarity_dense: null pointer");
  VOSMEM_CHECK_ARG(ck >= 1 && n >= 0 && hw >= 0, "vosmem_similarity_dense: ck=%d n=%lld hw=%d", ck, (long long)n, hw);
  if (n == 0 || hw == 0) return VOSMEM_OK;
  dim3 grid((hw + DT - 1) / DT, (unsigned)ceil_div64(n, DT));
  similarity_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(key, key_ld, shrinkage, query_key, query_selection,
                                                                 ck, n, hw, out);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_softmax_dense(const float *similarity, int64_t sim_ld, int64_t n, int hw, int top_k,
                                    float *affinity, int64_t aff_ld, float *usage, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(similarity && affinity, "vosmem_softmax_dense: null pointer");
  VOSMEM_CHECK_ARG(n >= 1 && hw >= 1, "vosmem_softmax_dense: n=%lld hw=%d", (long long)n, hw);
  VOSMEM_CHECK_ARG(top_k <= DENSE_MAX_TOPK, "vosmem_softmax_dense: top_k=%d above %d", top_k, DENSE_MAX_TOPK);
  VOSMEM_CHECK_ARG(top_k <= 0 || top_k <= n, "vosmem_softmax_dense: top_k=%d exceeds the %lld memory elements", top_k,
                   (long long)n);  // torch.topk raises here too (memory_util.py:46)
  cudaStream_t st = (cudaStream_t)stream;
  if (usage) VOSMEM_CUDA(cudaMemsetAsync(usage, 0, sizeof(float) * n, st));
  dim3 grid((hw + 31) / 32);
  if (top_k > VOSMEM_MAX_TOPK)
    softmax_topk_any_kernel<<<(hw + 7) / 8, 256, 0, st>>>(similarity, sim_ld, n, hw, top_k, affinity, aff_ld, usage);
  else if (top_k > 0)
    softmax_topk_dense_kernel<<<grid, 1024, 0, st>>>(similarity, sim_ld, n, hw, top_k, affinity, aff_ld, usage);
  else
    softmax_full_dense_kernel<<<grid, 1024, 0, st>>>(similarity, sim_ld, n, hw, affinity, aff_ld, usage);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_readout_dense(const float *value, int64_t value_ld, const float *affinity, int64_t aff_ld,
                                    int rows, int64_t n, int hw, float *out, int64_t out_ld, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(value && affinity && out, "vosmem_readout_dense: null pointer");
  VOSMEM_CHECK_ARG(rows >= 1 && n >= 1 && hw >= 1, "vosmem_readout_dense: rows=%d n=%lld hw=%d", rows, (long long)n, hw);
  dim3 grid((hw + DT - 1) / DT, (rows + DT - 1) / DT);
  readout_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(value, value_ld, affinity, aff_ld, rows, n, hw, out,
                                                              out_ld);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}
