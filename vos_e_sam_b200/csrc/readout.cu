// Softmax over the top-k survivors, usage scatter-add and the sparse affinity x value readout.
//
// Reference: do_softmax top-k branch (tracker/model/memory_util.py:45-54,62-63) builds a dense N x HW
// affinity with k non-zeros per column and MemoryManager._readout multiplies it densely
// (tracker/inference/memory_manager.py:53-55).  Here only the k survivors are touched:
//   out[r, q] = sum_j w[q, j] * V[index[q, j], r]
// with V held one memory element per row (vosmem_pack_values), so each survivor is one contiguous
// row read.  Bound by L2 -> SM gather bandwidth / latency: algorithmic bytes =
// rows * min(N, HW*k) * sizeof(value) + rows * HW * 4 (+ the HW*k candidate list).
//
// Three front ends share the gather:
//   staged   : score / index lists (HW x k) come from vosmem_select_topk or a cross-rank merge;
//   fused    : the kernel merges the per-split candidate lists of the selection kernel itself (merge.cuh), one
//              warp per query, which removes a kernel and a round trip from vosmem_match;
//   exchange : N-sharded bank -- the kernel waits for the flags of the source ranks, then merges the `world`
//              exchange lists that the ranks pushed into this rank's memory for its own query slice.
#include <cstdlib>

#include "common.cuh"
#include "merge.cuh"

namespace vosmem {

namespace {

constexpr int RTHREADS = 256;
constexpr int RCH = 512;      // value rows (channels) per pass
constexpr int FRONT_STAGED = 0, FRONT_FUSED = 1, FRONT_EXCHANGE = 2;
constexpr int EXCH_K = VOSMEM_EXCH_K;

struct ExchEntry {   // one exchanged candidate: score + GLOBAL key index (-1 = none)
  float score;
  int index;
};
static_assert(sizeof(ExchEntry) == 8, "exchange entries are read / written as 8-byte words");

struct ReadoutArgs {
  const void *shadow[2];
  int64_t shadow_ld[2];
  int64_t first[2], count[2];
  float *use_count[2];
  float *life_count[2];
  int n_segments;
  int hw, top_k, rows;
  const float *score;      // staged front end
  const int64_t *index;
  SplitLists lists;        // fused front end
  const ExchEntry *exch;   // exchange front end: [n_lists][exch_stride] entries, query q of the slice at exch_first + q * EXCH_K
  int exch_lists;
  int64_t exch_stride, exch_first;
  const uint32_t *exch_flags;
  uint32_t exch_seq;
  uint32_t *exch_status;
  float *out;
  int64_t out_ld;
  float *out_weight;
  int out_group_rows;        // 0 = plain rows x HW
  int64_t out_group_stride;
};

struct ReadoutBatch {   // kernel parameter: one ReadoutArgs per problem (blockIdx.z)
  ReadoutArgs p[MAX_BATCH];
};
static_assert(sizeof(ReadoutBatch) <= 4000, "kernel parameter space");

// (a0, a1) += w * (x0, x1) as one packed fp32 FMA (fma.rn.f32x2, sm_100): the same two fmaf results, half the issue slots
__device__ __forceinline__ void fma2(float &a0, float &a1, float x0, float x1, float w) {
  asm("{\n\t.reg .b64 v, ww, a;\n\t"
      "mov.b64 v, {%2, %3};\n\t"
      "mov.b64 ww, {%4, %4};\n\t"
      "mov.b64 a, {%0, %1};\n\t"
      "fma.rn.f32x2 a, v, ww, a;\n\t"
      "mov.b64 {%0, %1}, a;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(x0), "f"(x1), "f"(w));
}

template <typename T, int VEC>
struct Loader;
template <>
struct Loader<float, 4> {
  using Raw = float4;
  static __device__ __forceinline__ Raw load(const float *p) { return *reinterpret_cast<const float4 *>(p); }
  static __device__ __forceinline__ void fma(const Raw &v, float w, float (&acc)[4]) {
    fma2(acc[0], acc[1], v.x, v.y, w);
    fma2(acc[2], acc[3], v.z, v.w, w);
  }
};
template <>
struct Loader<float, 1> {
  using Raw = float;
  static __device__ __forceinline__ Raw load(const float *p) { return *p; }
  static __device__ __forceinline__ void fma(const Raw &v, float w, float (&acc)[1]) { acc[0] = fmaf(w, v, acc[0]); }
};
template <>
struct Loader<__nv_bfloat16, 8> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint4 *>(p); }
  static __device__ __forceinline__ void fma(const Raw &raw, float w, float (&acc)[8]) {
    const unsigned u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      fma2(acc[2 * i], acc[2 * i + 1], __uint_as_float(u[i] << 16), __uint_as_float(u[i] & 0xffff0000u), w);
  }
};
template <>
struct Loader<__nv_bfloat16, 1> {
  using Raw = __nv_bfloat16;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16 *p) { return *p; }
  static __device__ __forceinline__ void fma(const Raw &v, float w, float (&acc)[1]) {
    acc[0] = fmaf(w, __bfloat162float(v), acc[0]);
  }
};

// One warp turns query q's survivors into softmax weights and value-row pointers (32 slots; unused slots and
// survivors without a local row get weight 0 and point at a real row of the query, so the gather needs no branch
// on validity and never touches uninitialised memory; `any` = the query has at least one local row).
// Also the usage scatter-add and the optional weight output (only when `side_effects`).
template <typename T, int FRONT>
__device__ __forceinline__ void resolve_query(const ReadoutArgs &a, int q, bool side_effects, float *m_s, int *m_i, int lane,
                                              float *w_out, const T **row_out, int &any) {
  float s = -INFINITY;
  int64_t gi = -1;
  if (q < a.hw) {
    if (FRONT == FRONT_FUSED) {
      const WarpTop32 top = merge_lists_of_query<12>(a.lists, q, a.top_k, m_s, m_i, lane);
      if (lane < a.top_k && top.i != 0x7fffffff) { s = top.s; gi = top.i; }
    } else if (FRONT == FRONT_EXCHANGE) {
      // one entry per lane from every source rank's list, eight lists in flight at a time, folded into the best 32
      WarpTop32 top;
      top.init();
      for (int l0 = 0; l0 < a.exch_lists; l0 += 8) {
        uint2 raw[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          const int ll = min(l0 + l, a.exch_lists - 1);
          raw[l] = __ldcg(reinterpret_cast<const uint2 *>(a.exch + ll * a.exch_stride + a.exch_first + (int64_t)q * EXCH_K + lane));
        }
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          const bool ok = l0 + l < a.exch_lists && (int)raw[l].y >= 0;
          top.push(ok ? __uint_as_float(raw[l].x) : -INFINITY, ok ? (int)raw[l].y : 0x7fffffff, lane);
        }
      }
      if (lane < a.top_k && top.i != 0x7fffffff) { s = top.s; gi = top.i; }
    } else if (lane < a.top_k) {
      s = a.score[(int64_t)q * a.top_k + lane];
      gi = a.index[(int64_t)q * a.top_k + lane];
      if (gi < 0) s = -INFINITY;
    }
  }
  // softmax over the survivors (memory_util.py:48-49, max-subtracted)
  const float m = warp_max(s);
  const float e = (s == -INFINITY) ? 0.f : expf(s - m);
  const float sum = warp_sum(e);
  float w = sum > 0.f ? e / sum : 0.f;
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
  if (side_effects && q < a.hw && lane < a.top_k) {
    if (a.out_weight) a.out_weight[(int64_t)q * a.top_k + lane] = w;
    if (use && w > 0.f) atomicAdd(use, w);  // usage = affinity row sums (memory_util.py:63)
  }
  const unsigned have = __ballot_sync(FULL, row != nullptr);
  if (row == nullptr) w = 0.f;
  if (have) {
    const int donor = __ffs(have) - 1;
    const unsigned long long dp = __shfl_sync(FULL, reinterpret_cast<unsigned long long>(row), donor);
    if (row == nullptr) row = reinterpret_cast<const T *>(dp);
  }
  w_out[lane] = w;
  row_out[lane] = row;
  if (lane == 0) any = have != 0;
}

// grid: (ceil(hw / RQ), chunks of RCH rows [1 when FUSED: the CTA walks all chunks]); block RTHREADS.
// GROUPED: output rows addressed through out_group_rows / out_group_stride (see vosmem_readout_desc); kept out of the
// plain instantiation, whose write-out loop is measurably faster without the extra address arithmetic.
template <typename T, int VEC, int RQ, int FRONT, bool GROUPED>
__global__ void __launch_bounds__(RTHREADS, 3) softmax_readout_kernel(const __grid_constant__ ReadoutBatch batch) {
  constexpr bool FUSED = FRONT == FRONT_FUSED;
  const ReadoutArgs &a = batch.p[blockIdx.z];
  constexpr int TPQ = RCH / VEC >= RTHREADS ? RTHREADS : RCH / VEC;  // threads covering one pass of channels
  constexpr int GROUPS = RTHREADS / TPQ;                             // query groups working concurrently
  constexpr int CH_PER_PASS = TPQ * VEC;                             // <= RCH
  __shared__ float s_w[RQ][32];
  __shared__ const T *s_row[RQ][32];
  __shared__ int s_any[RQ];
  __shared__ __align__(16) float s_out[RQ][RCH + 8];   // query-major, pitch chosen so that both the vector stores of the
                                                      // gather and the transposed reads of the write-out are conflict-free
  __shared__ float m_s[FUSED ? RQ : 1][FUSED ? MERGE_BUF : 1];
  __shared__ int m_i[FUSED ? RQ : 1][FUSED ? MERGE_BUF : 1];

  const int q0 = blockIdx.x * RQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // --- life_count += 1 over the segments' elements (kv_memory_store.py:99), spread over the CTAs of the first chunk.
  //     With 4 queries per CTA only warps 0-3 resolve queries below: the other warps do the ageing meanwhile, so its
  //     load -> add -> store round trip is off every query's critical path. ---
  constexpr int AGE_WARP0 = RQ < RTHREADS / 32 ? RQ : 0;               // first warp that ages
  constexpr int AGE_THREADS = RTHREADS - AGE_WARP0 * 32;
  if (blockIdx.y == 0 && warp >= AGE_WARP0) {
#pragma unroll
    for (int sgi = 0; sgi < 2; ++sgi)
      if (sgi < a.n_segments && a.life_count[sgi])
        for (int64_t i = (int64_t)blockIdx.x * AGE_THREADS + (threadIdx.x - AGE_WARP0 * 32); i < a.count[sgi];
             i += (int64_t)gridDim.x * AGE_THREADS)
          a.life_count[sgi][i] += 1.0f;
  }

  // --- exchange front end: wait until every source rank's lists for this rank's queries have landed (bounded spin:
  //     a rank that never arrives flags the status word instead of hanging the GPU) ---
  if (FRONT == FRONT_EXCHANGE && a.exch_flags != nullptr) {
    if (threadIdx.x < a.exch_lists) {
      const uint32_t *f = a.exch_flags + threadIdx.x;
      uint32_t v;
      long long spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v == a.exch_seq) break;
        __nanosleep(100);
      } while (++spins < (1ll << 24));
      if (v != a.exch_seq && a.exch_status) *a.exch_status = 1u;
    }
    __syncthreads();
  }

  // --- per query (one warp each): survivors -> softmax weights -> value rows ---
  for (int qq = warp; qq < RQ; qq += RTHREADS / 32)
    resolve_query<T, FRONT>(a, q0 + qq, blockIdx.y == 0, m_s[FUSED ? qq : 0], m_i[FUSED ? qq : 0], lane, s_w[qq], s_row[qq],
                            s_any[qq]);
  __syncthreads();

  const int g = threadIdx.x / TPQ, t = threadIdx.x % TPQ;
  for (int ch0 = blockIdx.y * RCH; ch0 < a.rows; ch0 += gridDim.y * RCH) {
    // --- gather: thread owns VEC consecutive channels of one query at a time ---
    for (int cbase = 0; cbase < RCH; cbase += CH_PER_PASS) {
      const int c_local = cbase + t * VEC;
      const int ch = ch0 + c_local;
      const bool ch_ok = ch < a.rows;
      for (int qq = g; qq < RQ; qq += GROUPS) {
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        if (ch_ok && q0 + qq < a.hw && s_any[qq]) {
          // eight independent row reads in flight per thread before the first use (latency-bound gather)
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 8) {
            if (j0 < a.top_k) {
              typename Loader<T, VEC>::Raw raw[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) raw[u] = Loader<T, VEC>::load(s_row[qq][j0 + u] + ch);
#pragma unroll
              for (int u = 0; u < 8; ++u) Loader<T, VEC>::fma(raw[u], s_w[qq][j0 + u], acc);
            }
          }
        }
        if (VEC % 4 == 0) {
#pragma unroll
          for (int v = 0; v < VEC; v += 4)
            *reinterpret_cast<float4 *>(&s_out[qq][c_local + v]) = make_float4(acc[v], acc[v + 1], acc[v + 2], acc[v + 3]);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) s_out[qq][c_local + v] = acc[v];
        }
      }
    }
    __syncthreads();
    // --- write rows x HW with HW contiguous: RQ consecutive queries per row segment ---
    if (!GROUPED) {
      for (int e = threadIdx.x; e < RCH * RQ; e += RTHREADS) {
        const int c_local = e / RQ, qq = e % RQ;
        const int ch = ch0 + c_local, q = q0 + qq;
        if (ch < a.rows && q < a.hw) a.out[(int64_t)ch * a.out_ld + q] = s_out[qq][c_local];
      }
    } else {
      // a chunk lies inside one group whenever the group size is a multiple of the chunk (e.g. CV = 512): its base
      // is then computed once; other group sizes divide per element
      const int ogr = a.out_group_rows;
      const bool per_element = ogr % RCH != 0;
      const int64_t chunk_base = (int64_t)(ch0 / ogr) * a.out_group_stride + (int64_t)(ch0 % ogr) * a.out_ld;
      for (int e = threadIdx.x; e < RCH * RQ; e += RTHREADS) {
        const int c_local = e / RQ, qq = e % RQ;
        const int ch = ch0 + c_local, q = q0 + qq;
        if (ch < a.rows && q < a.hw) {
          const int64_t at = per_element ? (int64_t)(ch / ogr) * a.out_group_stride + (int64_t)(ch % ogr) * a.out_ld
                                         : chunk_base + (int64_t)c_local * a.out_ld;
          a.out[at + q] = s_out[qq][c_local];
        }
      }
    }
    __syncthreads();
  }
}


template <int RQ, int FRONT>
int launch(const ReadoutBatch &b, int n, int value_dtype, bool vec_ok, cudaStream_t st) {
  constexpr bool FUSED = FRONT != FRONT_STAGED;   // the CTA walks all row chunks itself (front end done once per query)
  int hw = 0, rows = 0;
  for (int i = 0; i < n; ++i) {
    hw = b.p[i].hw > hw ? b.p[i].hw : hw;
    rows = b.p[i].rows > rows ? b.p[i].rows : rows;
  }
  dim3 grid((hw + RQ - 1) / RQ, FUSED ? 1 : (rows + RCH - 1) / RCH, n);
  bool grouped = false;
  for (int i = 0; i < n; ++i) grouped = grouped || b.p[i].out_group_rows != 0;
  for (int i = 0; i < n; ++i)
    VOSMEM_CHECK_ARG(!grouped || b.p[i].out_group_rows != 0, "readout: grouped and plain output rows in one batch");
  static const int carveout = [] { const char *e = getenv("VOSMEM_RD_CARVEOUT"); return e ? atoi(e) : -1; }();   // experiment
#define VOSMEM_LAUNCH_RD_AS(T, VEC, GR)                                                                           \
  do {                                                                                                            \
    if (carveout >= 0)                                                                                            \
      VOSMEM_CUDA(cudaFuncSetAttribute(softmax_readout_kernel<T, VEC, RQ, FRONT, GR>,                             \
                                       cudaFuncAttributePreferredSharedMemoryCarveout, carveout));                \
    softmax_readout_kernel<T, VEC, RQ, FRONT, GR><<<grid, RTHREADS, 0, st>>>(b);                                  \
  } while (0)
#define VOSMEM_LAUNCH_RD(T, VEC)                                                                    \
  do {                                                                                              \
    if (grouped) VOSMEM_LAUNCH_RD_AS(T, VEC, true);                                                 \
    else VOSMEM_LAUNCH_RD_AS(T, VEC, false);                                                        \
  } while (0)
  if (value_dtype == VOSMEM_F32) {
    if (vec_ok) VOSMEM_LAUNCH_RD(float, 4);
    else VOSMEM_LAUNCH_RD(float, 1);
  } else {
    if (vec_ok) VOSMEM_LAUNCH_RD(__nv_bfloat16, 8);
    else VOSMEM_LAUNCH_RD(__nv_bfloat16, 1);
  }
#undef VOSMEM_LAUNCH_RD
#undef VOSMEM_LAUNCH_RD_AS
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

int fill_args(const vosmem_readout_desc *d, ReadoutArgs &a, bool &vec_ok) {
  VOSMEM_CHECK_ARG(d != nullptr, "readout: null descriptor");
  VOSMEM_CHECK_ARG(d->hw >= 1 && d->rows >= 1 && d->out, "readout: hw=%d rows=%d", d->hw, d->rows);
  VOSMEM_CHECK_ARG(d->top_k >= 1 && d->top_k <= VOSMEM_MAX_TOPK, "readout: top_k=%d outside [1, %d]", d->top_k,
                   VOSMEM_MAX_TOPK);
  VOSMEM_CHECK_ARG(d->n_segments >= 1 && d->n_segments <= 2, "readout: n_segments=%d", d->n_segments);
  VOSMEM_CHECK_ARG(d->value_dtype == VOSMEM_F32 || d->value_dtype == VOSMEM_BF16, "readout: dtype %d", d->value_dtype);
  vec_ok = true;
  const int vec = d->value_dtype == VOSMEM_F32 ? 4 : 8;
  for (int s = 0; s < d->n_segments; ++s) {
    const vosmem_value_segment &g = d->seg[s];
    VOSMEM_CHECK_ARG(g.shadow && g.shadow_ld >= d->rows && g.count >= 0, "readout: bad value segment %d", s);
    a.shadow[s] = g.shadow;
    a.shadow_ld[s] = g.shadow_ld;
    a.first[s] = g.first;
    a.count[s] = g.count;
    a.use_count[s] = g.use_count;
    a.life_count[s] = g.life_count;
    vec_ok = vec_ok && (g.shadow_ld % vec == 0) && (reinterpret_cast<uintptr_t>(g.shadow) % 16 == 0);
  }
  vec_ok = vec_ok && (d->rows % vec == 0);
  a.n_segments = d->n_segments;
  a.hw = d->hw;
  a.top_k = d->top_k;
  a.rows = d->rows;
  a.out = d->out;
  a.out_ld = d->out_ld;
  a.out_weight = d->out_weight;
  VOSMEM_CHECK_ARG(d->out_group_rows >= 0 && (d->out_group_rows == 0 || d->rows % d->out_group_rows == 0),
                   "readout: out_group_rows=%d does not divide rows=%d", d->out_group_rows, d->rows);
  a.out_group_rows = d->out_group_rows;
  a.out_group_stride = d->out_group_stride;
  return VOSMEM_OK;
}

}  // namespace

// fused front end, used by vosmem_match / vosmem_match_batch after the selection kernel
int launch_fused_readout(const vosmem_readout_desc *d, const SplitLists *lists, int n, cudaStream_t st) {
  VOSMEM_CHECK_ARG(n >= 1 && n <= MAX_BATCH, "readout: batch of %d problems outside [1, %d]", n, MAX_BATCH);
  ReadoutBatch b{};
  bool vec_all = true;
  for (int i = 0; i < n; ++i) {
    bool vec_ok;
    int rc = fill_args(d + i, b.p[i], vec_ok);
    if (rc != VOSMEM_OK) return rc;
    VOSMEM_CHECK_ARG(d[i].value_dtype == d[0].value_dtype, "readout: problems of one batch must share the value storage type");
    vec_all = vec_all && vec_ok;
    b.p[i].lists = lists[i];
  }
  return launch<4, FRONT_FUSED>(b, n, d[0].value_dtype, vec_all, st);
}

}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_softmax_readout(const vosmem_readout_desc *d, const float *score, const int64_t *index,
                                      vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(score && index, "vosmem_softmax_readout: null candidate lists");
  ReadoutBatch b{};
  ReadoutArgs &a = b.p[0];
  bool vec_ok;
  int rc = fill_args(d, a, vec_ok);
  if (rc != VOSMEM_OK) return rc;
  a.score = score;
  a.index = index;
  return launch<8, FRONT_STAGED>(b, 1, d->value_dtype, vec_ok, (cudaStream_t)stream);
}

extern "C" int vosmem_exchange_readout(const vosmem_readout_desc *d, const vosmem_exchange_desc *x, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(x != nullptr && x->lists != nullptr, "vosmem_exchange_readout: null exchange lists");
  VOSMEM_CHECK_ARG(x->n_lists >= 1 && x->n_lists <= VOSMEM_MAX_RANKS, "vosmem_exchange_readout: n_lists=%d outside [1, %d]",
                   x->n_lists, VOSMEM_MAX_RANKS);
  ReadoutBatch b{};
  ReadoutArgs &a = b.p[0];
  bool vec_ok;
  int rc = fill_args(d, a, vec_ok);
  if (rc != VOSMEM_OK) return rc;
  a.exch = static_cast<const ExchEntry *>(x->lists);
  a.exch_lists = x->n_lists;
  a.exch_stride = x->list_stride;
  a.exch_first = x->first_entry;
  a.exch_flags = x->flags;
  a.exch_seq = x->seq;
  a.exch_status = x->status;
  return launch<4, FRONT_EXCHANGE>(b, 1, d->value_dtype, vec_ok, (cudaStream_t)stream);
}
