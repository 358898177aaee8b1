// Softmax over the top-k survivors, usage scatter-add and the sparse affinity x value readout.
//
// Reference: do_softmax top-k branch (tracker/model/memory_util.py:45-54,62-63) builds a dense N x HW
// affinity with k non-zeros per column and MemoryManager._readout multiplies it densely
// (tracker/inference/memory_manager.py:53-55).  Here only the k survivors are touched:
//   out[r, q] = sum_j w[q, j] * V[index[q, j], r]
// with V held one memory element per row (vosmem_pack_values), so each survivor is one contiguous
// row read.  HBM/L2-bound: algorithmic bytes = rows * min(N, HW*k) * sizeof(value) + rows * HW * 4.
#include "common.cuh"

namespace vosmem {

namespace {

constexpr int RQ = 16;        // queries per CTA
constexpr int RTHREADS = 256;
constexpr int RCH = 512;      // value rows (channels) per CTA

struct ReadoutArgs {
  const void *shadow[2];
  int64_t shadow_ld[2];
  int64_t first[2], count[2];
  float *use_count[2];
  int n_segments;
  int hw, top_k, rows;
  const float *score;
  const int64_t *index;
  float *out;
  int64_t out_ld;
  float *out_weight;
};

template <typename T, int VEC>
struct Loader;
template <>
struct Loader<float, 4> {
  static __device__ __forceinline__ void fma(const float *p, float w, float (&acc)[4]) {
    float4 v = *reinterpret_cast<const float4 *>(p);
    acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
    acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
  }
};
template <>
struct Loader<float, 1> {
  static __device__ __forceinline__ void fma(const float *p, float w, float (&acc)[1]) { acc[0] = fmaf(w, *p, acc[0]); }
};
template <>
struct Loader<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void fma(const __nv_bfloat16 *p, float w, float (&acc)[8]) {
    uint4 raw = *reinterpret_cast<const uint4 *>(p);
    const unsigned u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fmaf(w, __uint_as_float(u[i] << 16), acc[2 * i]);
      acc[2 * i + 1] = fmaf(w, __uint_as_float(u[i] & 0xffff0000u), acc[2 * i + 1]);
    }
  }
};
template <>
struct Loader<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void fma(const __nv_bfloat16 *p, float w, float (&acc)[1]) {
    acc[0] = fmaf(w, __bfloat162float(*p), acc[0]);
  }
};

// grid: (ceil(hw / RQ), ceil(rows / RCH)); block RTHREADS.
template <typename T, int VEC>
__global__ void __launch_bounds__(RTHREADS) softmax_readout_kernel(ReadoutArgs a) {
  constexpr int TPQ = RCH / VEC >= RTHREADS ? RTHREADS : RCH / VEC;  // threads covering the channel tile
  constexpr int GROUPS = RTHREADS / TPQ;                             // query groups working concurrently
  constexpr int CH_PER_PASS = TPQ * VEC;                             // <= RCH
  __shared__ float s_w[RQ][32];
  __shared__ const T *s_row[RQ][32];
  __shared__ float s_out[RCH][RQ + 1];

  const int q0 = blockIdx.x * RQ;
  const int ch0 = blockIdx.y * RCH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // --- softmax of the survivors, one warp per query (memory_util.py:48-49, max-subtracted) ---
  for (int qq = warp; qq < RQ; qq += RTHREADS / 32) {
    const int q = q0 + qq;
    float s = -INFINITY;
    int64_t gi = -1;
    if (q < a.hw && lane < a.top_k) {
      s = a.score[(int64_t)q * a.top_k + lane];
      gi = a.index[(int64_t)q * a.top_k + lane];
      if (gi < 0) s = -INFINITY;
    }
    const float m = warp_max(s);
    const float e = (s == -INFINITY) ? 0.f : expf(s - m);
    const float sum = warp_sum(e);
    float w = sum > 0.f ? e / sum : 0.f;
    // resolve the value row
    const T *row = nullptr;
    float *use = nullptr;
#pragma unroll
    for (int sgi = 0; sgi < 2; ++sgi) {
      if (sgi < a.n_segments && gi >= a.first[sgi] && gi < a.first[sgi] + a.count[sgi]) {
        const int64_t r = gi - a.first[sgi];
        row = static_cast<const T *>(a.shadow[sgi]) + r * a.shadow_ld[sgi];
        if (a.use_count[sgi]) use = a.use_count[sgi] + r;
      }
    }
    if (blockIdx.y == 0 && q < a.hw && lane < a.top_k) {
      if (a.out_weight) a.out_weight[(int64_t)q * a.top_k + lane] = w;
      if (use && w > 0.f) atomicAdd(use, w);  // usage = affinity row sums (memory_util.py:63)
    }
    if (row == nullptr) {  // value lives elsewhere (another rank) or padding: contributes nothing here
      w = 0.f;
      row = static_cast<const T *>(a.shadow[0]);
    }
    s_w[qq][lane] = w;
    s_row[qq][lane] = row;
  }
  __syncthreads();

  // --- gather: thread owns VEC consecutive channels of one query at a time ---
  const int g = threadIdx.x / TPQ, t = threadIdx.x % TPQ;
  for (int cbase = 0; cbase < RCH; cbase += CH_PER_PASS) {
    const int c_local = cbase + t * VEC;
    const int ch = ch0 + c_local;
    const bool ch_ok = ch < a.rows;
    for (int qq = g; qq < RQ; qq += GROUPS) {
      float acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
      if (ch_ok && q0 + qq < a.hw) {
#pragma unroll 6
        for (int j = 0; j < a.top_k; ++j) {
          const float w = s_w[qq][j];
          if (w != 0.f) Loader<T, VEC>::fma(s_row[qq][j] + ch, w, acc);  // w == 0: padding / remote value
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_out[c_local + v][qq] = acc[v];
    }
  }
  __syncthreads();

  // --- write rows x HW with HW contiguous: RQ consecutive queries per row segment ---
  for (int e = threadIdx.x; e < RCH * RQ; e += RTHREADS) {
    const int c_local = e / RQ, qq = e % RQ;
    const int ch = ch0 + c_local, q = q0 + qq;
    if (ch < a.rows && q < a.hw) a.out[(int64_t)ch * a.out_ld + q] = s_out[c_local][qq];
  }
}

}  // namespace
}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_softmax_readout(const vosmem_readout_desc *d, const float *score, const int64_t *index,
                                      vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(d && score && index, "vosmem_softmax_readout: null pointer");
  VOSMEM_CHECK_ARG(d->hw >= 1 && d->rows >= 1 && d->out, "vosmem_softmax_readout: hw=%d rows=%d", d->hw, d->rows);
  VOSMEM_CHECK_ARG(d->top_k >= 1 && d->top_k <= VOSMEM_MAX_TOPK, "vosmem_softmax_readout: top_k=%d outside [1, %d]",
                   d->top_k, VOSMEM_MAX_TOPK);
  VOSMEM_CHECK_ARG(d->n_segments >= 1 && d->n_segments <= 2, "vosmem_softmax_readout: n_segments=%d", d->n_segments);
  VOSMEM_CHECK_ARG(d->value_dtype == VOSMEM_F32 || d->value_dtype == VOSMEM_BF16, "vosmem_softmax_readout: dtype %d",
                   d->value_dtype);
  ReadoutArgs a{};
  bool vec_ok = true;
  const int vec = d->value_dtype == VOSMEM_F32 ? 4 : 8;
  for (int s = 0; s < d->n_segments; ++s) {
    const vosmem_value_segment &g = d->seg[s];
    VOSMEM_CHECK_ARG(g.shadow && g.shadow_ld >= d->rows && g.count >= 0, "vosmem_softmax_readout: bad value segment %d", s);
    a.shadow[s] = g.shadow;
    a.shadow_ld[s] = g.shadow_ld;
    a.first[s] = g.first;
    a.count[s] = g.count;
    a.use_count[s] = g.use_count;
    vec_ok = vec_ok && (g.shadow_ld % vec == 0) && (reinterpret_cast<uintptr_t>(g.shadow) % 16 == 0);
  }
  vec_ok = vec_ok && (d->rows % vec == 0);
  a.n_segments = d->n_segments;
  a.hw = d->hw;
  a.top_k = d->top_k;
  a.rows = d->rows;
  a.score = score;
  a.index = index;
  a.out = d->out;
  a.out_ld = d->out_ld;
  a.out_weight = d->out_weight;
  dim3 grid((d->hw + RQ - 1) / RQ, (d->rows + RCH - 1) / RCH);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->value_dtype == VOSMEM_F32) {
    if (vec_ok) softmax_readout_kernel<float, 4><<<grid, RTHREADS, 0, st>>>(a);
    else softmax_readout_kernel<float, 1><<<grid, RTHREADS, 0, st>>>(a);
  } else {
    if (vec_ok) softmax_readout_kernel<__nv_bfloat16, 8><<<grid, RTHREADS, 0, st>>>(a);
    else softmax_readout_kernel<__nv_bfloat16, 1><<<grid, RTHREADS, 0, st>>>(a);
  }
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}
