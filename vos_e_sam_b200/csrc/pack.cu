// Operand packing: keys / queries -> tensor-core images, values -> gather-friendly shadow.
//
// Reference math (tracker/model/memory_util.py:20-27,35):
//   S[n,q] = s[n]/sqrt(CK) * ( -sum_c k[c,n]^2 e[c,q] + 2 sum_c k[c,n] q[c,q] e[c,q] - sum_c e[c,q] q[c,q]^2 )
// is one inner product of length 2*CK+1 between
//   key row   x[n] = s[n]/sqrt(CK) * [ k^2 , k , 1 ]          (packed once per memory frame)
//   query row y[q] =                 [ -e  , 2 q e , -sum e q^2 ]   (packed once per query frame)
// Each fp32 entry is stored as a bf16 (hi, lo) pair; the kernel accumulates hi*hi + hi*lo + lo*hi
// in fp32, which keeps ~16 mantissa bits per factor (measured score error ~1e-4 on N(0,1) inputs).
#include "common.cuh"

namespace vosmem {

namespace {

struct Chunk8 {
  __nv_bfloat16 v[8];
};
static_assert(sizeof(Chunk8) == 16, "chunk must be one 16-byte core-matrix row");

__device__ __forceinline__ void store_chunk(unsigned char *tile, int off, const Chunk8 &c) {
  *reinterpret_cast<uint4 *>(tile + off) = *reinterpret_cast<const uint4 *>(&c);
}

// One thread per key.  Image chunk order per row: [x1_hi 0-7 | x2_hi 8-15 | x1_lo 16-23 | x2_lo 24-31 | tail 32 | 0 33]
// with x1 = c*k^2, x2 = c*k, tail = (c_hi, c_hi, c_lo, 0...).
__global__ void pack_keys_kernel(const float *__restrict__ key, int64_t key_ld, const float *__restrict__ shrinkage,
                                 int64_t begin, int64_t end, unsigned char *__restrict__ image) {
  int64_t n = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= end) return;
  const float scale = (shrinkage ? shrinkage[n] : 1.0f) * 0.125f;  // 1/sqrt(64)
  unsigned char *tile = image + (n / TK) * (int64_t)KEY_TILE_BYTES;
  const int r = (int)(n % TK);
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    Chunk8 h1, l1, h2, l2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float m = key[(int64_t)(g * 8 + j) * key_ld + n];
      float x2 = scale * m;
      float x1 = x2 * m;
      split_bf16(x1, h1.v[j], l1.v[j]);
      split_bf16(x2, h2.v[j], l2.v[j]);
    }
    store_chunk(tile, image_offset<TK>(r, g), h1);
    store_chunk(tile, image_offset<TK>(r, 8 + g), h2);
    store_chunk(tile, image_offset<TK>(r, 16 + g), l1);
    store_chunk(tile, image_offset<TK>(r, 24 + g), l2);
  }
  Chunk8 t;
  __nv_bfloat16 hi, lo;
  split_bf16(scale, hi, lo);
#pragma unroll
  for (int j = 0; j < 8; ++j) t.v[j] = __float2bfloat16_rn(0.f);
  t.v[0] = hi; t.v[1] = hi; t.v[2] = lo;
  store_chunk(tile, image_offset<TK>(r, 32), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) t.v[j] = __float2bfloat16_rn(0.f);
  store_chunk(tile, image_offset<TK>(r, 33), t);
}

// Block = 32 queries x 8 channel groups (coalesced 128-byte reads of qk / qe).  Produces the fp32 vector for the
// SIMT path (c-major: qvec[c * hw_pad + q]) and the bf16 image (the tcgen05 kernel packs its own query tile in its
// prologue; the image is kept for the tile self-test).  Image chunk order: [y1_hi | y2_hi | y1_lo | y2_lo | tail | 0] with y1 = -e, y2 = 2 q e,
// tail = (y3_hi, y3_lo, y3_hi, 0...), y3 = -sum e q^2.
__global__ void __launch_bounds__(256) pack_query_kernel(const float *__restrict__ qk, const float *__restrict__ qe,
                                                         int ck, int hw, int hw_pad,
                                                         float *__restrict__ qvec, unsigned char *__restrict__ image) {
  __shared__ float red[8][33];
  const int ql = threadIdx.x, gy = threadIdx.y;
  const int q = blockIdx.x * 32 + ql;
  const bool live = q < hw;
  const bool img = (ck == CK_TC);
  unsigned char *tile = image + (int64_t)(q / TQ) * QUERY_TILE_BYTES;
  const int r = q % TQ;
  float y3 = 0.f;
  for (int g = gy; g * 8 < ck; g += 8) {
    Chunk8 h1, l1, h2, l2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float y1 = 0.f, y2 = 0.f;
      if (live && c < ck) {
        const float k = qk[(int64_t)c * hw + q];
        const float e = qe ? qe[(int64_t)c * hw + q] : 1.0f;
        y1 = -e;
        y2 = 2.0f * (k * e);
        if (qe) y3 -= e * (k * k);
      }
      if (c < ck) {
        qvec[(int64_t)c * hw_pad + q] = y1;
        qvec[(int64_t)(ck + c) * hw_pad + q] = y2;
      }
      split_bf16(y1, h1.v[j], l1.v[j]);
      split_bf16(y2, h2.v[j], l2.v[j]);
    }
    if (img) {
      store_chunk(tile, image_offset<TQ>(r, g), h1);
      store_chunk(tile, image_offset<TQ>(r, 8 + g), h2);
      store_chunk(tile, image_offset<TQ>(r, 16 + g), l1);
      store_chunk(tile, image_offset<TQ>(r, 24 + g), l2);
    }
  }
  red[gy][ql] = y3;
  __syncthreads();
  if (gy != 0) return;
#pragma unroll
  for (int g = 1; g < 8; ++g) y3 += red[g][ql];
  qvec[(int64_t)(2 * ck) * hw_pad + q] = y3;
  if (img) {
    Chunk8 t;
    __nv_bfloat16 hi, lo;
    split_bf16(y3, hi, lo);
#pragma unroll
    for (int j = 0; j < 8; ++j) t.v[j] = __float2bfloat16_rn(0.f);
    t.v[0] = hi; t.v[1] = lo; t.v[2] = hi;
    store_chunk(tile, image_offset<TQ>(r, 32), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) t.v[j] = __float2bfloat16_rn(0.f);
    store_chunk(tile, image_offset<TQ>(r, 33), t);
  }
}

template <typename T>
__device__ __forceinline__ T to_store(float v);
template <>
__device__ __forceinline__ float to_store<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_store<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 32 x 32 transpose through shared memory: rows x N (N contiguous) -> N x rows (rows contiguous)
template <typename T>
__global__ void pack_values_kernel(const float *__restrict__ value, int64_t value_ld, int rows, int64_t src_begin,
                                   int64_t n, T *__restrict__ shadow, int64_t shadow_ld, int64_t dst_begin) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    int r = r0 + dy;
    int64_t i = i0 + threadIdx.x;
    tile[dy][threadIdx.x] = (r < rows && i < n) ? value[(int64_t)r * value_ld + src_begin + i] : 0.f;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    int64_t i = i0 + dy;
    int r = r0 + threadIdx.x;
    if (r < rows && i < n) shadow[(dst_begin + i) * shadow_ld + r] = to_store<T>(tile[threadIdx.x][dy]);
  }
}

// ---- vosmem_store_append -------------------------------------------------------------------------------------------
struct AppendKeys {
  int ck, m;
  int64_t n, key_ld, selection_ld, bank_ld;
  const float *key, *shrinkage, *selection;
  float *bank_key, *bank_shrinkage, *bank_selection, *bank_use, *bank_life;
  unsigned char *image;
};
// thread = new element: copy its key / shrinkage / selection column into the bank, initialise the counters, pack its
// row of the tensor-core image (from the bank entries this thread has just written)
__global__ void __launch_bounds__(128) append_keys_kernel(const AppendKeys a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.m) return;
  const int64_t at = a.n + i;
  for (int c = 0; c < a.ck; ++c) a.bank_key[(int64_t)c * a.bank_ld + at] = a.key[(int64_t)c * a.key_ld + i];
  if (a.bank_shrinkage) a.bank_shrinkage[at] = a.shrinkage[i];
  if (a.bank_selection)
    for (int c = 0; c < a.ck; ++c) a.bank_selection[(int64_t)c * a.bank_ld + at] = a.selection[(int64_t)c * a.selection_ld + i];
  if (a.bank_use) a.bank_use[at] = 0.f;          // kv_memory_store.py:37
  if (a.bank_life) a.bank_life[at] = 1e-7f;      // kv_memory_store.py:38
  if (a.image == nullptr) return;
  const float scale = (a.bank_shrinkage ? a.shrinkage[i] : 1.0f) * 0.125f;   // 1 / sqrt(64)
  unsigned char *tile = a.image + (at / TK) * (int64_t)KEY_TILE_BYTES;
  const int r = (int)(at % TK);
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    Chunk8 h1, l1, h2, l2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = a.key[(int64_t)(g * 8 + j) * a.key_ld + i];
      const float x2 = scale * m;
      const float x1 = x2 * m;
      split_bf16(x1, h1.v[j], l1.v[j]);
      split_bf16(x2, h2.v[j], l2.v[j]);
    }
    store_chunk(tile, image_offset<TK>(r, g), h1);
    store_chunk(tile, image_offset<TK>(r, 8 + g), h2);
    store_chunk(tile, image_offset<TK>(r, 16 + g), l1);
    store_chunk(tile, image_offset<TK>(r, 24 + g), l2);
  }
  Chunk8 t;
  __nv_bfloat16 hi, lo;
  split_bf16(scale, hi, lo);
#pragma unroll
  for (int j = 0; j < 8; ++j) t.v[j] = __float2bfloat16_rn(0.f);
  t.v[0] = hi; t.v[1] = hi; t.v[2] = lo;
  store_chunk(tile, image_offset<TK>(r, 32), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) t.v[j] = __float2bfloat16_rn(0.f);
  store_chunk(tile, image_offset<TK>(r, 33), t);
}

// values of one group: one read of the new rows x m block, written to the reference layout AND, transposed through
// shared memory, to the shadow
template <typename T>
__global__ void append_values_kernel(const float *__restrict__ value, int64_t value_ld, int rows, int64_t m,
                                     float *__restrict__ ref, int64_t ref_ld, T *__restrict__ shadow, int64_t shadow_ld,
                                     int64_t n) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int r = r0 + dy;
    const int64_t i = i0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && i < m) {
      v = value[(int64_t)r * value_ld + i];
      ref[(int64_t)r * ref_ld + n + i] = v;
    }
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int64_t i = i0 + dy;
    const int r = r0 + threadIdx.x;
    if (r < rows && i < m) shadow[(n + i) * shadow_ld + r] = to_store<T>(tile[threadIdx.x][dy]);
  }
}

__global__ void age_kernel(float *__restrict__ life, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) life[i] += 1.0f;
}

}  // namespace

int launch_pack_query(const float *qk, const float *qe, int ck, int hw, const Workspace &ws, cudaStream_t st) {
  int hw_pad = (int)round_up64(hw, TQ);
  pack_query_kernel<<<hw_pad / 32, dim3(32, 8), 0, st>>>(qk, qe, ck, hw, hw_pad, ws.qvec, ws.query_image);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_pack_keys(const float *key, int64_t key_ld, const float *shrinkage, int ck, int64_t begin,
                                int64_t end, void *image, int64_t capacity, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(ck == CK_TC, "vosmem_pack_keys: the tensor-core image exists for CK == 64 only (got %d)", ck);
  VOSMEM_CHECK_ARG(key && image, "vosmem_pack_keys: null pointer");
  VOSMEM_CHECK_ARG(0 <= begin && begin <= end && end <= round_up64(capacity, TK),
                   "vosmem_pack_keys: range [%lld, %lld) outside capacity %lld", (long long)begin, (long long)end,
                   (long long)capacity);
  if (begin == end) return VOSMEM_OK;
  int64_t n = end - begin;
  pack_keys_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, (cudaStream_t)stream>>>(
      key, key_ld, shrinkage, begin, end, static_cast<unsigned char *>(image));
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_pack_values(const float *value, int64_t value_ld, int rows, int64_t src_begin, int64_t n,
                                  void *shadow, int64_t shadow_ld, int64_t dst_begin, int shadow_dtype,
                                  vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(value && shadow, "vosmem_pack_values: null pointer");
  VOSMEM_CHECK_ARG(rows > 0 && n >= 0 && shadow_ld >= rows, "vosmem_pack_values: bad shape rows=%d n=%lld ld=%lld",
                   rows, (long long)n, (long long)shadow_ld);
  VOSMEM_CHECK_ARG(shadow_dtype == VOSMEM_F32 || shadow_dtype == VOSMEM_BF16, "vosmem_pack_values: bad dtype %d",
                   shadow_dtype);
  if (n == 0) return VOSMEM_OK;
  dim3 grid((unsigned)ceil_div64(n, 32), (unsigned)((rows + 31) / 32)), block(32, 8);
  if (shadow_dtype == VOSMEM_F32)
    pack_values_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(value, value_ld, rows, src_begin, n,
                                                                        static_cast<float *>(shadow), shadow_ld, dst_begin);
  else
    pack_values_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(
        value, value_ld, rows, src_begin, n, static_cast<__nv_bfloat16 *>(shadow), shadow_ld, dst_begin);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_store_append(const vosmem_append_desc *d, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(d != nullptr, "vosmem_store_append: null descriptor");
  VOSMEM_CHECK_ARG(d->ck >= 1 && d->m >= 1 && d->n >= 0, "vosmem_store_append: ck=%d m=%d n=%lld", d->ck, d->m, (long long)d->n);
  VOSMEM_CHECK_ARG(d->key && d->bank_key && d->n + d->m <= d->capacity, "vosmem_store_append: %lld + %d elements exceed the capacity %lld",
                   (long long)d->n, d->m, (long long)d->capacity);
  VOSMEM_CHECK_ARG((d->bank_shrinkage == nullptr) == (d->shrinkage == nullptr) && (d->bank_selection == nullptr) == (d->selection == nullptr),
                   "vosmem_store_append: shrinkage / selection must be given exactly when the bank holds them");
  VOSMEM_CHECK_ARG(d->key_image == nullptr || d->ck == CK_TC, "vosmem_store_append: the key image exists for CK == 64 only");
  VOSMEM_CHECK_ARG(d->n_groups >= 0 && d->n_groups <= VOSMEM_MAX_GROUPS, "vosmem_store_append: n_groups=%d", d->n_groups);
  VOSMEM_CHECK_ARG(d->value_dtype == VOSMEM_F32 || d->value_dtype == VOSMEM_BF16, "vosmem_store_append: bad dtype %d", d->value_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  AppendKeys a{d->ck, d->m, d->n, d->key_ld, d->selection_ld, d->bank_ld, d->key, d->shrinkage, d->selection, d->bank_key,
               d->bank_shrinkage, d->bank_selection, d->bank_use, d->bank_life, static_cast<unsigned char *>(d->key_image)};
  append_keys_kernel<<<(d->m + 127) / 128, 128, 0, st>>>(a);
  for (int gi = 0; gi < d->n_groups; ++gi) {
    const vosmem_append_group &g = d->group[gi];
    if (g.value == nullptr) continue;     // (a group that receives nothing in this call)
    VOSMEM_CHECK_ARG(g.ref && g.shadow && g.rows >= 1 && g.shadow_ld >= g.rows, "vosmem_store_append: bad value group %d", gi);
    dim3 grid((unsigned)ceil_div64(d->m, 32), (unsigned)((g.rows + 31) / 32)), block(32, 8);
    if (d->value_dtype == VOSMEM_F32)
      append_values_kernel<float><<<grid, block, 0, st>>>(g.value, g.value_ld, g.rows, d->m, g.ref, g.ref_ld,
                                                          static_cast<float *>(g.shadow), g.shadow_ld, g.n);
    else
      append_values_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(g.value, g.value_ld, g.rows, d->m, g.ref, g.ref_ld,
                                                                  static_cast<__nv_bfloat16 *>(g.shadow), g.shadow_ld, g.n);
  }
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_age(float *life_count, int64_t n, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(life_count || n == 0, "vosmem_age: null pointer");
  if (n <= 0) return VOSMEM_OK;
  age_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(life_count, n);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_debug_pack_query(const float *query_key, const float *query_selection, int ck, int hw,
                                       void *image, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(ck == CK_TC && query_key && image, "vosmem_debug_pack_query: CK must be 64, pointers non-null");
  // image buffer layout: [query image tiles | one row of published thresholds (hw_pad) | qvec]; see vosmem_query_image_bytes
  Workspace ws{};
  int64_t n_qtiles = ceil_div64(hw, TQ), hw_pad = n_qtiles * TQ;
  unsigned char *p = static_cast<unsigned char *>(image);
  ws.query_image = p;
  ws.qvec = reinterpret_cast<float *>(p + n_qtiles * QUERY_TILE_BYTES + hw_pad * 4);
  return launch_pack_query(query_key, query_selection, ck, hw, ws, (cudaStream_t)stream);
}

extern "C" int64_t vosmem_query_image_bytes(int ck, int hw) {
  int64_t n_qtiles = ceil_div64(hw, TQ), hw_pad = n_qtiles * TQ;
  return n_qtiles * QUERY_TILE_BYTES + hw_pad * 4 + hw_pad * (2 * (int64_t)ck + 1) * 4;
}
