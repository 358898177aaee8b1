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

// Dense softmax over the key axis in two launches that are parallel over BOTH axes (the one-warp-per-column kernel above
// leaves all but hw / 32 SMs idle: 630 us for the 8 100 x 128 matrix of a consolidation).  Rows are read as 128-byte
// segments (lane = column of a group of 32).
//   stats : per (row chunk, column) running max m and sum of exp(s - m)
//   apply : combine the chunks' statistics per column, write exp(s - M) / S, add the row sums to `usage`
constexpr int SM_ROWS = 128;    // rows per chunk
__global__ void __launch_bounds__(256) softmax_stats_kernel(const float *__restrict__ sim, int64_t sim_ld, int64_t n, int hw,
                                                            float2 *__restrict__ stats) {
  __shared__ float2 part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * SM_ROWS;
  float m = -INFINITY, sum = 0.f;
  if (q < hw) {
    for (int64_t i = r0 + warp; i < r0 + SM_ROWS && i < n; i += 8) {
      const float v = sim[i * sim_ld + q];
      const float nm = fmaxf(m, v);
      sum = sum * expf(m - nm) + expf(v - nm);
      m = nm;
    }
  }
  part[warp][lane] = make_float2(m, sum);
  __syncthreads();
  if (warp == 0 && q < hw) {
    float M = -INFINITY;
    for (int w8 = 0; w8 < 8; ++w8) M = fmaxf(M, part[w8][lane].x);
    float S = 0.f;
    for (int w8 = 0; w8 < 8; ++w8)
      if (part[w8][lane].y > 0.f) S += part[w8][lane].y * expf(part[w8][lane].x - M);
    stats[(int64_t)blockIdx.y * hw + q] = make_float2(M, S);
  }
}
__global__ void __launch_bounds__(256) softmax_apply_kernel(const float *sim, int64_t sim_ld, int64_t n, int hw,
                                                            const float2 *__restrict__ stats, int chunks, float *affinity,
                                                            int64_t aff_ld, float *usage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * SM_ROWS;
  float M = -INFINITY, S = 0.f;
  if (q < hw) {
    for (int c = 0; c < chunks; ++c) M = fmaxf(M, stats[(int64_t)c * hw + q].x);
    for (int c = 0; c < chunks; ++c) {
      const float2 st = stats[(int64_t)c * hw + q];
      if (st.y > 0.f) S += st.y * expf(st.x - M);
    }
  }
  for (int64_t i = r0 + warp; i < r0 + SM_ROWS && i < n; i += 8) {
    float w = 0.f;
    if (q < hw) {
      w = expf(sim[i * sim_ld + q] - M) / S;
      affinity[i * aff_ld + q] = w;
    }
    if (usage) {
      const float row = warp_sum(w);
      if (lane == 0) atomicAdd(usage + i, row);
    }
  }
}

// out[rows x hw] = value[rows x n] @ affinity[n x hw] for a handful of rows (the shrinkage row of a consolidation,
// memory_manager.py:285): the 64 x 64 tiles of the kernel below would run on hw / 64 CTAs.  grid = (column groups of 32,
// row chunks of the key axis); partial sums are added into the zero-initialised output.
__global__ void __launch_bounds__(256) readout_few_rows_kernel(const float *__restrict__ value, int64_t value_ld,
                                                               const float *__restrict__ aff, int64_t aff_ld, int rows,
                                                               int64_t n, int hw, float *__restrict__ out, int64_t out_ld) {
  __shared__ float part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * SM_ROWS;
  for (int r = 0; r < rows; ++r) {
    float acc = 0.f;
    if (q < hw)
      for (int64_t i = r0 + warp; i < r0 + SM_ROWS && i < n; i += 8) acc = fmaf(value[(int64_t)r * value_ld + i], aff[i * aff_ld + q], acc);
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && q < hw) {
      float t = 0.f;
      for (int w8 = 0; w8 < 8; ++w8) t += part[w8][lane];
      atomicAdd(out + (int64_t)r * out_ld + q, t);
    }
    __syncthreads();
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

extern "C" int64_t vosmem_softmax_dense_scratch_bytes(int64_t n, int hw);

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

// scratch (optional, vosmem_softmax_dense_scratch_bytes): lets the top_k <= 0 (dense) softmax run parallel over rows too
extern "C" int vosmem_softmax_dense_ws(const float *similarity, int64_t sim_ld, int64_t n, int hw, int top_k,
                                       float *affinity, int64_t aff_ld, float *usage, void *stats_scratch,
                                       int64_t scratch_bytes, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(similarity && affinity, "vosmem_softmax_dense: null pointer");
  if (stats_scratch != nullptr && scratch_bytes < vosmem_softmax_dense_scratch_bytes(n, hw)) stats_scratch = nullptr;
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
  else if (stats_scratch != nullptr) {
    const int chunks = (int)ceil_div64(n, SM_ROWS);
    dim3 g2((hw + 31) / 32, chunks);
    softmax_stats_kernel<<<g2, 256, 0, st>>>(similarity, sim_ld, n, hw, static_cast<float2 *>(stats_scratch));
    softmax_apply_kernel<<<g2, 256, 0, st>>>(similarity, sim_ld, n, hw, static_cast<const float2 *>(stats_scratch), chunks,
                                             affinity, aff_ld, usage);
  } else
    softmax_full_dense_kernel<<<grid, 1024, 0, st>>>(similarity, sim_ld, n, hw, affinity, aff_ld, usage);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_softmax_dense(const float *similarity, int64_t sim_ld, int64_t n, int hw, int top_k,
                                    float *affinity, int64_t aff_ld, float *usage, vosmem_stream_t stream) {
  return vosmem_softmax_dense_ws(similarity, sim_ld, n, hw, top_k, affinity, aff_ld, usage, nullptr, 0, stream);
}

extern "C" int64_t vosmem_softmax_dense_scratch_bytes(int64_t n, int hw) {
  return n >= 1 && hw >= 1 ? ceil_div64(n, SM_ROWS) * hw * (int64_t)sizeof(float2) : 0;
}

extern "C" int vosmem_readout_dense(const float *value, int64_t value_ld, const float *affinity, int64_t aff_ld,
                                    int rows, int64_t n, int hw, float *out, int64_t out_ld, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(value && affinity && out, "vosmem_readout_dense: null pointer");
  VOSMEM_CHECK_ARG(rows >= 1 && n >= 1 && hw >= 1, "vosmem_readout_dense: rows=%d n=%lld hw=%d", rows, (long long)n, hw);
  if (rows <= 8) {
    VOSMEM_CUDA(cudaMemset2DAsync(out, out_ld * sizeof(float), 0, hw * sizeof(float), rows, (cudaStream_t)stream));
    readout_few_rows_kernel<<<dim3((hw + 31) / 32, (unsigned)ceil_div64(n, SM_ROWS)), 256, 0, (cudaStream_t)stream>>>(
        value, value_ld, affinity, aff_ld, rows, n, hw, out, out_ld);
    VOSMEM_CUDA(cudaGetLastError());
    return VOSMEM_OK;
  }
  dim3 grid((hw + DT - 1) / DT, (rows + DT - 1) / DT);
  readout_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(value, value_ld, affinity, aff_ld, rows, n, hw, out,
                                                              out_ld);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}
