// Fused similarity + per-query candidate selection on tcgen05 / TMEM (sm_100a, CK == 64).
//
// Replaces get_similarity + torch.topk of the reference (tracker/model/memory_util.py:7-39,46) for one
// object group: the N x HW similarity matrix only ever exists as 128 x 64 fp32 tiles in tensor memory.
//
//   grid  = (query tiles of 128, N-splits)        one CTA per SM (~219 KB shared memory)
//   warp 4  producer : cp.async.bulk (TMA) of the resident query image and the streamed key tiles,
//                      mbarrier full/empty ring of STAGES stages
//   warp 5  MMA      : per key tile 25 tcgen05.mma (M128 N64 K16, bf16 hi/lo split -> fp32 in TMEM)
//                      into one of ACC_BUFS accumulator buffers, tcgen05.commit -> mbarriers
//   warps 0-3 epilogue: tcgen05.ld the tile (thread = query row), keep every score above the
//                      thread's running threshold in a private shared-memory candidate list; when a
//                      list fills the warp prunes it cooperatively to its best 32 and raises the
//                      threshold, which is also published (atomicMax) for the other N-splits.
// Each CTA leaves <= 64 candidates per query in the exchange buffer; merge_splits_kernel finishes.
//
// Invariant that makes the result exact: a threshold is only ever the 32nd-best score of a set of
// real candidates of that query, so it never exceeds the true k-th (k <= 32) best score.
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vosmem {
namespace {

constexpr int STAGES = 3;
constexpr int ACC_BUFS = 4;
constexpr int TMEM_COLS = ACC_BUFS * TK;  // 256
constexpr int EPI_WARPS = 4;
constexpr int TC_THREADS = 192;
constexpr int CSLOTS = 64;            // candidate slots per query in shared memory
constexpr int CS_F = TQ + 1;          // floats per slot row (+1: conflict-free both slot-wise and query-wise)
constexpr int CS_H = TQ + 2;          // u16 per slot row
constexpr int PRUNE_ABOVE = CSLOTS - 8;

// shared memory map (bytes)
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + QUERY_TILE_BYTES;
constexpr int SM_CS = SM_K + STAGES * KEY_TILE_BYTES;
constexpr int SM_CI = SM_CS + CSLOTS * CS_F * 4;
constexpr int SM_BAR = (SM_CI + CSLOTS * CS_H * 2 + 15) / 16 * 16;
constexpr int N_BARS = 2 * STAGES + 2 * ACC_BUFS + 1;
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;
constexpr int SM_TOTAL = SM_TMEM + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget exceeded");

constexpr uint32_t IDESC = ptx::umma_idesc_bf16(TQ, TK);
constexpr uint32_t Q_LBO = (TQ / 8) * 128, K_LBO = (TK / 8) * 128, SBO = 128;

struct TcSeg {
  const unsigned char *image;
  int64_t begin, end;   // candidate keys of the bank
  int64_t tile0;        // first image tile touched
  int64_t tiles;        // number of image tiles touched
};

struct TcArgs {
  TcSeg seg[2];
  int64_t len0;
  int64_t tiles_total;
  int hw, hw_pad, splits;
  const unsigned char *query_image;
  unsigned *tau;
  float *cand_score;
  int *cand_index;
  int *cand_count;
};

// 25 MMAs of one 128 x 64 tile: hi*hi + lo*hi + hi*lo over the 128 packed channels, then the rank-1 tail.
__device__ __forceinline__ void issue_tile(uint32_t q_base, uint32_t k_base, uint32_t tmem_d) {
  uint32_t acc = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    ptx::umma_bf16(tmem_d, ptx::umma_desc(q_base + (2 * s) * Q_LBO, Q_LBO, SBO),
                   ptx::umma_desc(k_base + (2 * s) * K_LBO, K_LBO, SBO), IDESC, acc);
    acc = 1;
  }
#pragma unroll
  for (int s = 0; s < 8; ++s)
    ptx::umma_bf16(tmem_d, ptx::umma_desc(q_base + (16 + 2 * s) * Q_LBO, Q_LBO, SBO),
                   ptx::umma_desc(k_base + (2 * s) * K_LBO, K_LBO, SBO), IDESC, 1);
#pragma unroll
  for (int s = 0; s < 8; ++s)
    ptx::umma_bf16(tmem_d, ptx::umma_desc(q_base + (2 * s) * Q_LBO, Q_LBO, SBO),
                   ptx::umma_desc(k_base + (16 + 2 * s) * K_LBO, K_LBO, SBO), IDESC, 1);
  ptx::umma_bf16(tmem_d, ptx::umma_desc(q_base + 32 * Q_LBO, Q_LBO, SBO), ptx::umma_desc(k_base + 32 * K_LBO, K_LBO, SBO),
                 IDESC, 1);
}

struct TileRef {
  int seg;
  int64_t key0;  // bank index of column 0
};
__device__ __forceinline__ TileRef tile_ref(const TcArgs &a, int64_t g) {
  TileRef t;
  t.seg = g >= a.seg[0].tiles;
  t.key0 = (t.seg ? a.seg[1].tile0 + (g - a.seg[0].tiles) : a.seg[0].tile0 + g) * TK;
  return t;
}

__global__ void __launch_bounds__(TC_THREADS, 1) select_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + SM_BAR);
  uint64_t *bar_empty = bar_full + STAGES;
  uint64_t *bar_tfull = bar_empty + STAGES;
  uint64_t *bar_tempty = bar_tfull + ACC_BUFS;
  uint64_t *bar_q = bar_tempty + ACC_BUFS;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
  float *cs = reinterpret_cast<float *>(smem + SM_CS);
  unsigned short *ci = reinterpret_cast<unsigned short *>(smem + SM_CI);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x;
  const int64_t g_lo = a.tiles_total * blockIdx.y / a.splits;
  const int64_t g_hi = a.tiles_total * (blockIdx.y + 1) / a.splits;
  const int n_tiles = (int)(g_hi - g_lo);

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(bar_full + i, 1); ptx::mbar_init(bar_empty + i, 1); }
    for (int i = 0; i < ACC_BUFS; ++i) { ptx::mbar_init(bar_tfull + i, 1); ptx::mbar_init(bar_tempty + i, EPI_WARPS); }
    ptx::mbar_init(bar_q, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===== producer =====
    if (lane == 0 && n_tiles > 0) {
      ptx::mbar_arrive_expect_tx(bar_q, QUERY_TILE_BYTES);
      const unsigned char *qsrc = a.query_image + (int64_t)qtile * QUERY_TILE_BYTES;
      ptx::bulk_g2s(smem + SM_Q, qsrc, QUERY_TILE_BYTES / 2, bar_q);
      ptx::bulk_g2s(smem + SM_Q + QUERY_TILE_BYTES / 2, qsrc + QUERY_TILE_BYTES / 2, QUERY_TILE_BYTES / 2, bar_q);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % STAGES;
        ptx::mbar_wait(bar_empty + st, ((i / STAGES) & 1) ^ 1);
        const int64_t g = g_lo + i;
        const int sg = g >= a.seg[0].tiles;
        const int64_t tile = sg ? a.seg[1].tile0 + (g - a.seg[0].tiles) : a.seg[0].tile0 + g;
        ptx::mbar_arrive_expect_tx(bar_full + st, KEY_TILE_BYTES);
        ptx::bulk_g2s(smem + SM_K + st * KEY_TILE_BYTES, a.seg[sg].image + tile * KEY_TILE_BYTES, KEY_TILE_BYTES,
                      bar_full + st);
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0 && n_tiles > 0) {
      ptx::mbar_wait(bar_q, 0);
      const uint32_t q_base = ptx::smem_u32(smem + SM_Q);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % STAGES, buf = i % ACC_BUFS;
        ptx::mbar_wait(bar_tempty + buf, ((i / ACC_BUFS) & 1) ^ 1);
        ptx::mbar_wait(bar_full + st, (i / STAGES) & 1);
        ptx::tc_fence_after();
        issue_tile(q_base, ptx::smem_u32(smem + SM_K + st * KEY_TILE_BYTES), tmem_base + buf * TK);
        ptx::umma_commit(bar_empty + st);   // key stage reusable once these MMAs have read it
        ptx::umma_commit(bar_tfull + buf);  // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: warps 0-3 own TMEM lanes 32*warp .. 32*warp+31 =====
    const int row = warp * 32 + lane;              // query row inside the tile
    const int q = qtile * TQ + row;
    float tau = -INFINITY;
    int cnt = 0;
    const volatile unsigned *tau_g = a.tau + q;
    for (int i = 0; i < n_tiles; ++i) {
      const int buf = i % ACC_BUFS;
      const unsigned tg = *tau_g;  // thresholds published by the other N-splits (latency hidden by the wait)
      ptx::mbar_wait(bar_tfull + buf, (i / ACC_BUFS) & 1);
      ptx::tc_fence_after();
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * TK;
      ptx::tmem_ld_32x32(taddr, v0);
      ptx::tmem_ld_32x32(taddr + 32, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + buf);  // accumulator buffer free again
      tau = fmaxf(tau, ord2f(tg));

      const TileRef tr = tile_ref(a, g_lo + i);
      const int64_t kb = a.seg[tr.seg].begin, ke = a.seg[tr.seg].end;
      const bool partial = tr.key0 < kb || tr.key0 + TK > ke;
      const int li0 = i * TK;
#pragma unroll
      for (int grp = 0; grp < 8; ++grp) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = grp * 8 + jj;
          float s = __uint_as_float(j < 32 ? v0[j & 31] : v1[j & 31]);
          if (partial && (tr.key0 + j < kb || tr.key0 + j >= ke)) s = -INFINITY;
          if (s > tau) {
            cs[cnt * CS_F + row] = s;
            ci[cnt * CS_H + row] = (unsigned short)(li0 + j);
            ++cnt;
          }
        }
        // lists that could overflow in the next 8 columns are pruned to their best 32, one query at a time
        unsigned full = __ballot_sync(FULL, cnt > PRUNE_ABOVE);
        while (full) {
          const int src = __ffs(full) - 1;
          full &= full - 1;
          __syncwarp();
          const int n = __shfl_sync(FULL, cnt, src);
          const int srow = warp * 32 + src;
          float sa = -INFINITY, sb = -INFINITY;
          int ia = 0x7fffffff, ib = 0x7fffffff;
          if (lane < n) { sa = cs[lane * CS_F + srow]; ia = ci[lane * CS_H + srow]; }
          if (lane + 32 < n) { sb = cs[(lane + 32) * CS_F + srow]; ib = ci[(lane + 32) * CS_H + srow]; }
          warp_sort_desc(sa, ia, lane);
          warp_sort_desc(sb, ib, lane);
          const float rs = __shfl_sync(FULL, sb, 31 - lane);
          const int ri = __shfl_sync(FULL, ib, 31 - lane);
          if (better(rs, ri, sa, ia)) { sa = rs; ia = ri; }
          const float floor32 = warp_min(sa);
          __syncwarp();
          cs[lane * CS_F + srow] = sa;
          ci[lane * CS_H + srow] = (unsigned short)ia;
          if (lane == src) { cnt = 32; tau = fmaxf(tau, floor32); }
          if (lane == 0) atomicMax(a.tau + qtile * TQ + srow, f2ord(floor32));
          __syncwarp();
        }
      }
    }
    // ---- hand the surviving candidates to the merge kernel ----
    __syncwarp();
    for (int src = 0; src < 32; ++src) {
      const int qs = qtile * TQ + warp * 32 + src;
      if (qs >= a.hw) break;
      const int n = __shfl_sync(FULL, cnt, src);
      const int srow = warp * 32 + src;
      const int64_t slot = ((int64_t)blockIdx.y * a.hw_pad + qs) * CAND_SLOTS;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = lane + 32 * h;
        if (e < n) {
          const int li = ci[e * CS_H + srow];
          const TileRef tr = tile_ref(a, g_lo + (li >> 6));
          const int64_t key = tr.key0 + (li & 63);
          const int64_t cand = tr.seg ? a.len0 + (key - a.seg[1].begin) : key - a.seg[0].begin;
          a.cand_score[slot + e] = cs[e * CS_F + srow];
          a.cand_index[slot + e] = (int)cand;
        }
      }
      if (lane == 0) a.cand_count[(int64_t)blockIdx.y * a.hw_pad + qs] = n;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// Self-test: raw accumulator tile of query tile 0 against key tile 0.
__global__ void __launch_bounds__(TC_THREADS, 1) umma_tile_kernel(const unsigned char *query_image,
                                                                 const unsigned char *key_image, float *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bar_ld = reinterpret_cast<uint64_t *>(smem + SM_BAR);
  uint64_t *bar_mma = bar_ld + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_ld, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 4 && lane == 0) {
    ptx::mbar_arrive_expect_tx(bar_ld, QUERY_TILE_BYTES + KEY_TILE_BYTES);
    ptx::bulk_g2s(smem + SM_Q, query_image, QUERY_TILE_BYTES / 2, bar_ld);
    ptx::bulk_g2s(smem + SM_Q + QUERY_TILE_BYTES / 2, query_image + QUERY_TILE_BYTES / 2, QUERY_TILE_BYTES / 2, bar_ld);
    ptx::bulk_g2s(smem + SM_K, key_image, KEY_TILE_BYTES, bar_ld);
  } else if (warp == 5 && lane == 0) {
    ptx::mbar_wait(bar_ld, 0);
    ptx::tc_fence_after();
    issue_tile(ptx::smem_u32(smem + SM_Q), ptx::smem_u32(smem + SM_K), tmem_base);
    ptx::umma_commit(bar_mma);
  } else if (warp < 4) {
    ptx::mbar_wait(bar_mma, 0);
    ptx::tc_fence_after();
    uint32_t v0[32], v1[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    ptx::tmem_ld_32x32(taddr, v0);
    ptx::tmem_ld_32x32(taddr + 32, v1);
    ptx::tmem_ld_wait();
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      out[row * TK + j] = __uint_as_float(v0[j]);
      out[row * TK + 32 + j] = __uint_as_float(v1[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_select_tc(const vosmem_select_desc &d, const Workspace &ws, int splits, cudaStream_t st) {
  VOSMEM_CHECK_ARG(d.ck == CK_TC, "select(tcgen05): CK must be 64 (got %d)", d.ck);
  TcArgs a{};
  int64_t tiles = 0;
  for (int s = 0; s < 2; ++s) {
    if (s < d.n_segments) {
      const vosmem_segment &g = d.seg[s];
      VOSMEM_CHECK_ARG(g.key_image != nullptr || g.end == g.begin, "select(tcgen05): segment %d has no key image", s);
      TcSeg &t = a.seg[s];
      t.image = static_cast<const unsigned char *>(g.key_image);
      t.begin = g.begin;
      t.end = g.end;
      t.tile0 = g.begin / TK;
      t.tiles = g.end > g.begin ? ceil_div64(g.end, TK) - t.tile0 : 0;
      if (s == 0) a.len0 = g.end - g.begin;
      tiles += t.tiles;
    }
  }
  a.tiles_total = tiles;
  VOSMEM_CHECK_ARG(ceil_div64(tiles, splits) * TK <= 65536, "select(tcgen05): %lld key tiles over %d splits overflow the 16-bit candidate index",
                   (long long)tiles, splits);
  a.hw = d.hw;
  a.hw_pad = (int)round_up64(d.hw, TQ);
  a.splits = splits;
  a.query_image = ws.query_image;
  a.tau = ws.tau;
  a.cand_score = ws.cand_score;
  a.cand_index = ws.cand_index;
  a.cand_count = ws.cand_count;
  static bool attr_set = false;
  if (!attr_set) {
    VOSMEM_CUDA(cudaFuncSetAttribute(select_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div64(d.hw, TQ), splits);
  select_tc_kernel<<<grid, TC_THREADS, SM_TOTAL, st>>>(a);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_debug_umma_tile(const void *query_image, const void *key_image, float *out,
                                      vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(query_image && key_image && out, "vosmem_debug_umma_tile: null pointer");
  VOSMEM_CUDA(cudaFuncSetAttribute(umma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
  umma_tile_kernel<<<1, TC_THREADS, SM_TOTAL, (cudaStream_t)stream>>>(static_cast<const unsigned char *>(query_image),
                                                                      static_cast<const unsigned char *>(key_image), out);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}
