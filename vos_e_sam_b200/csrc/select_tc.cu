// Fused similarity + per-query candidate selection on tcgen05 / TMEM (sm_100a, CK == 64).
//
// Replaces get_similarity + torch.topk of the reference (tracker/model/memory_util.py:7-39,46) for one
// object group: the N x HW similarity matrix only ever exists as 128 x 64 fp32 tiles in tensor memory.
//
//   grid  = (query tiles of 128, N-splits)        one CTA per SM (~220 KB shared memory)
//   warp 4  producer : cp.async.bulk (TMA) of the resident query image and the streamed key tiles,
//                      mbarrier full/empty ring of STAGES stages
//   warp 5  MMA      : per key tile 25 tcgen05.mma (M128 N64 K16, bf16 hi/lo split -> fp32 in TMEM)
//                      into one of ACC_BUFS accumulator buffers, tcgen05.commit -> mbarriers
//   warp 6  refresher: keeps the per-query shared threshold (below) fresh in shared memory
//   warps 0-3 epilogue: tcgen05.ld the tile (thread = query row), append every score above the
//                      thread's threshold to a private shared-memory candidate list with predicated
//                      stores (no branches).  Groups of 8 columns in which no lane has a survivor are
//                      skipped with one warp-wide OR.  When a list fills: drop what fell below the
//                      (risen) threshold, thread-privately; only if that is not enough the warp cuts
//                      lists to their best 32 with a bitonic network over packed 32-bit keys.
// Each CTA leaves <= 64 candidates per query in the exchange buffer; merge_splits_kernel finishes.
//
// Thresholds.  A query's threshold is only ever a LOWER bound of its true 32nd-best score, so no true
// top-k (k <= 32) member is dropped:
//   local : after a cooperative cut, the (truncated) 32nd-best score of this CTA's own candidates;
//   shared: every thread tracks, in registers, lower bounds of the three best scores of G disjoint subsets
//           of its CTA's keys (subset = tile index mod G; updated from the maxima of the 8-column groups,
//           which are scores of distinct keys), G = ceil(11 / splits).  After every tile the CTA publishes
//           the minimum over its subsets of the 3rd best; the minimum over all splits then has at least
//           3 * G * splits >= 33 keys at or above it.  With S splits running concurrently this tracks the
//           quality of a single pass over all keys, which makes list overflows (and sorting) rare.
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vosmem {
namespace {

constexpr int STAGES = 3;
constexpr int ACC_BUFS = 4;
constexpr int TMEM_COLS = ACC_BUFS * TK;  // 256
constexpr int EPI_WARPS = 4;
constexpr int TC_THREADS = 224;        // 4 epilogue warps, producer, MMA issuer, threshold refresher
constexpr int CSLOTS = 64;             // candidate slots per query in shared memory
constexpr int CS_F = TQ + 1;           // floats per slot row (+1: conflict-free both slot-wise and query-wise)
constexpr int CS_H = TQ + 2;           // u16 per slot row
constexpr int PRUNE_ABOVE = CSLOTS - 8;
constexpr uint32_t SS = CS_F * 4, SI = CS_H * 2;  // byte strides between consecutive slots of one list
constexpr uint32_t KEY_SLOT_MASK = 63u;           // low bits of a sort key hold the slot id
constexpr int FIRST_WAIT_CYCLES = 20000;          // bounded wait for the other splits' first publication

// shared memory map (bytes)
constexpr int SM_Q = 0;
constexpr int SM_K = SM_Q + QUERY_TILE_BYTES;
constexpr int SM_CS = SM_K + STAGES * KEY_TILE_BYTES;
constexpr int SM_CI = SM_CS + CSLOTS * CS_F * 4;
constexpr int SM_BAR = (SM_CI + CSLOTS * CS_H * 2 + 15) / 16 * 16;
constexpr int N_BARS = 2 * STAGES + 2 * ACC_BUFS + 1;
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;   // [0] TMEM base address, [1] epilogue warps finished
constexpr int SM_TAU = SM_TMEM + 16;           // TQ floats: shared threshold per query row, kept fresh by warp 6
constexpr int SM_TOTAL = SM_TAU + TQ * 4;
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
  float *pub;          // [splits][hw_pad] published lower bound per (split, query): see "Thresholds"
  float *cand_score;
  int *cand_index;
  int *cand_count;
  long long *dbg;      // optional per-CTA cycle counters (16 per CTA), NULL in production
};

// The operand descriptors of one tile differ only in their start-address field; the MMA thread keeps the
// query-side ones in registers for the whole kernel and derives the key-side ones with one add each.
struct TileDescs {
  uint64_t q_hi[8], q_lo[8], q_tail;
};
__device__ __forceinline__ TileDescs make_query_descs(uint32_t q_base) {
  TileDescs d;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    d.q_hi[s] = ptx::umma_desc(q_base + (2 * s) * Q_LBO, Q_LBO, SBO);
    d.q_lo[s] = ptx::umma_desc(q_base + (16 + 2 * s) * Q_LBO, Q_LBO, SBO);
  }
  d.q_tail = ptx::umma_desc(q_base + 32 * Q_LBO, Q_LBO, SBO);
  return d;
}
// 25 MMAs of one 128 x 64 tile: hi*hi + lo*hi + hi*lo over the 128 packed channels, then the rank-1 tail.
__device__ __forceinline__ void issue_tile(const TileDescs &d, uint32_t k_base, uint32_t tmem_d) {
  const uint64_t k0 = ptx::umma_desc(k_base, K_LBO, SBO);
  constexpr uint64_t STEP = (2 * K_LBO) >> 4;   // two 8-element chunks per K=16 step, in 16-byte units
#pragma unroll
  for (int s = 0; s < 8; ++s) ptx::umma_bf16(tmem_d, d.q_hi[s], k0 + s * STEP, IDESC, s > 0);
#pragma unroll
  for (int s = 0; s < 8; ++s) ptx::umma_bf16(tmem_d, d.q_lo[s], k0 + s * STEP, IDESC, 1);
#pragma unroll
  for (int s = 0; s < 8; ++s) ptx::umma_bf16(tmem_d, d.q_hi[s], k0 + (8 + s) * STEP, IDESC, 1);
  ptx::umma_bf16(tmem_d, d.q_tail, k0 + 16 * STEP, IDESC, 1);
}

struct ListState {
  uint32_t base_s, base_i;  // shared-memory byte address of slot 0 of this thread's score / index list
  uint32_t off_s, off_i;    // next free slot
  float tau;                // current threshold (lower bound of this query's true 32nd-best score)
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned short lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
// if (score CMP tau) { *off_s = score; *off_i = index; advance both }   -- predicated, branch-free
#define VOSMEM_APPEND(CMP, off_s, off_i, score, tau, index)                                         \
  asm volatile(                                                                                     \
      "{\n\t.reg .pred p;\n\t"                                                                      \
      "setp." CMP ".f32 p, %2, %3;\n\t"                                                             \
      "@p st.shared.b32 [%0], %2;\n\t"                                                              \
      "@p st.shared.u16 [%1], %4;\n\t"                                                              \
      "@p add.u32 %0, %0, %5;\n\t"                                                                  \
      "@p add.u32 %1, %1, %6;\n\t}"                                                                 \
      : "+r"(off_s), "+r"(off_i)                                                                    \
      : "f"(score), "f"(tau), "h"((unsigned short)(index)), "n"(CS_F * 4), "n"(CS_H * 2)            \
      : "memory")

// Shared threshold of one query: min over the splits of their published r-th best score.
__device__ __forceinline__ float shared_threshold(const float *pub_q, int splits, int hw_pad) {
  float m = INFINITY;
#pragma unroll 4
  for (int y = 0; y < splits; ++y) m = fminf(m, __ldcg(pub_q + (int64_t)y * hw_pad));
  return m;
}

// Called when some list of the warp could overflow during the next 8 columns.  Kept out of line (eight call
// sites in the unrolled epilogue).  Three steps, cheapest first:
//   1. pick up the latest shared threshold (warp 6 keeps it fresh in shared memory);
//   2. every thread drops the entries of ITS list that fell below its threshold (thread-private, in place);
//   3. lists that are still too long are cut to their best 32 by the whole warp, four queries per round
//      (bitonic network over packed 32-bit keys), which also yields a new local threshold.
// (state goes in and comes back by value so that it stays in registers across the call)
__device__ __noinline__ ListState relieve_lists(ListState st, float *cs, unsigned short *ci, const volatile float *tau_row,
                                                int warp, int lane) {
  st.tau = fmaxf(st.tau, *tau_row);
  // ---- 2. thread-private compaction ----
  {
    const int cnt = (int)((st.off_s - st.base_s) / SS);
    const int nmax = __reduce_max_sync(FULL, cnt);
    uint32_t rd_s = st.base_s, rd_i = st.base_i, wr_s = st.base_s, wr_i = st.base_i;
    for (int e = 0; e < nmax; ++e) {
      const bool in = rd_s < st.off_s;
      const float sc = in ? lds_f32(rd_s) : __int_as_float(0x7fc00000);  // NaN never passes the compare
      const unsigned short ix = lds_u16(in ? rd_i : st.base_i);
      VOSMEM_APPEND("ge", wr_s, wr_i, sc, st.tau, ix);
      rd_s += SS;
      rd_i += SI;
    }
    st.off_s = wr_s;
    st.off_i = wr_i;
  }
  // ---- 3. cooperative cut to the best 32 ----
  unsigned full = __ballot_sync(FULL, st.off_s > st.base_s + PRUNE_ABOVE * SS);
  while (full) {
    int src[4];
    bool on[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      on[u] = full != 0;
      src[u] = on[u] ? __ffs(full) - 1 : src[0];
      full &= full - 1;
    }
    __syncwarp();
    int row[4];
    uint32_t ka[4], kb[4], kc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      row[u] = warp * 32 + src[u];
      const int n = (int)((__shfl_sync(FULL, st.off_s, src[u]) - __shfl_sync(FULL, st.base_s, src[u])) / SS);
      // packed key: order-preserving score bits, low 6 bits = slot id (distinct keys; the truncation can only
      // lower a threshold derived from a key, which keeps it a valid lower bound)
      ka[u] = lane < n ? ((f2ord(cs[lane * CS_F + row[u]]) & ~KEY_SLOT_MASK) | lane) : 0u;
      kb[u] = lane + 32 < n ? ((f2ord(cs[(lane + 32) * CS_F + row[u]]) & ~KEY_SLOT_MASK) | (lane + 32)) : 0u;
    }
    // eight independent 32-key bitonic sorts (descending), interleaved for ILP
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int jx = k >> 1; jx > 0; jx >>= 1) {
        const bool desc = (k == 32) || ((lane & k) == 0);
        const bool keep_max = ((lane & jx) == 0) == desc;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t oa = __shfl_xor_sync(FULL, ka[u], jx), ob = __shfl_xor_sync(FULL, kb[u], jx);
          ka[u] = keep_max ? max(ka[u], oa) : min(ka[u], oa);
          kb[u] = keep_max ? max(kb[u], ob) : min(kb[u], ob);
        }
      }
    }
    // best 32 of the 64: a (descending) against b reversed (ascending), then a bitonic merge -> sorted
#pragma unroll
    for (int u = 0; u < 4; ++u) kc[u] = max(ka[u], __shfl_sync(FULL, kb[u], 31 - lane));
#pragma unroll
    for (int jx = 16; jx > 0; jx >>= 1) {
      const bool keep_max = (lane & jx) == 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t o = __shfl_xor_sync(FULL, kc[u], jx);
        kc[u] = keep_max ? max(kc[u], o) : min(kc[u], o);
      }
    }
    // gather the survivors (lane l = l-th best) and rewrite them as slots 0..31
    float sv[4];
    unsigned short iv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int sl = kc[u] & KEY_SLOT_MASK;
      sv[u] = cs[sl * CS_F + row[u]];
      iv[u] = ci[sl * CS_H + row[u]];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (on[u]) {
        cs[lane * CS_F + row[u]] = sv[u];
        ci[lane * CS_H + row[u]] = iv[u];
        const float floor32 = ord2f(__shfl_sync(FULL, kc[u], 31) & ~KEY_SLOT_MASK);
        if (lane == src[u]) {
          st.off_s = st.base_s + 32 * SS;
          st.off_i = st.base_i + 32 * SI;
          st.tau = fmaxf(st.tau, floor32);
        }
      }
    }
    __syncwarp();
  }
  return st;
}

template <int G>   // G = subsets of the CTA's keys whose three best scores are tracked (see "Thresholds")
__global__ void __launch_bounds__(TC_THREADS, 1) select_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + SM_BAR);
  uint64_t *bar_empty = bar_full + STAGES;
  uint64_t *bar_tfull = bar_empty + STAGES;
  uint64_t *bar_tempty = bar_tfull + ACC_BUFS;
  uint64_t *bar_q = bar_tempty + ACC_BUFS;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
  volatile uint32_t *epi_done = reinterpret_cast<volatile uint32_t *>(smem + SM_TMEM + 4);
  volatile float *tau_sh = reinterpret_cast<volatile float *>(smem + SM_TAU);
  float *cs = reinterpret_cast<float *>(smem + SM_CS);
  unsigned short *ci = reinterpret_cast<unsigned short *>(smem + SM_CI);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x;
  const int64_t g_lo = a.tiles_total * blockIdx.y / a.splits;
  const int64_t g_hi = a.tiles_total * (blockIdx.y + 1) / a.splits;
  const int n_tiles = (int)(g_hi - g_lo);
  long long *dbg = a.dbg ? a.dbg + (blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(bar_full + i, 1); ptx::mbar_init(bar_empty + i, 1); }
    for (int i = 0; i < ACC_BUFS; ++i) { ptx::mbar_init(bar_tfull + i, 1); ptx::mbar_init(bar_tempty + i, EPI_WARPS); }
    ptx::mbar_init(bar_q, 1);
    ptx::fence_barrier_init();
    *epi_done = 0;
  }
  if (threadIdx.x < TQ) tau_sh[threadIdx.x] = -INFINITY;
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 6) {
    // ===== threshold refresher: tau_sh[row] = min over the splits of their published r-th best =====
    // (a stale value is still a valid lower bound, so no ordering with the epilogue is needed)
    const float *pub_row = a.pub + qtile * TQ + lane;
    for (int it = 0; *epi_done < EPI_WARPS; ++it) {
      float m[TQ / 32];
#pragma unroll
      for (int h = 0; h < TQ / 32; ++h) m[h] = INFINITY;
#pragma unroll
      for (int y = 0; y < 16; ++y) {      // unrolled: up to 64 independent loads in flight, one L2 round trip
        if (y < a.splits) {
#pragma unroll
          for (int h = 0; h < TQ / 32; ++h) m[h] = fminf(m[h], __ldcg(pub_row + (int64_t)y * a.hw_pad + 32 * h));
        }
      }
      for (int y = 16; y < a.splits; ++y) {
#pragma unroll
        for (int h = 0; h < TQ / 32; ++h) m[h] = fminf(m[h], __ldcg(pub_row + (int64_t)y * a.hw_pad + 32 * h));
      }
#pragma unroll
      for (int h = 0; h < TQ / 32; ++h) tau_sh[lane + 32 * h] = m[h];
      if (it >= 16) __nanosleep(256);   // thresholds move fastest during the first tiles
    }
  } else if (warp == 4) {
    // ===== producer =====
    if (lane == 0 && n_tiles > 0) {
      ptx::mbar_arrive_expect_tx(bar_q, QUERY_TILE_BYTES);
      const unsigned char *qsrc = a.query_image + (int64_t)qtile * QUERY_TILE_BYTES;
      ptx::bulk_g2s(smem + SM_Q, qsrc, QUERY_TILE_BYTES / 2, bar_q);
      ptx::bulk_g2s(smem + SM_Q + QUERY_TILE_BYTES / 2, qsrc + QUERY_TILE_BYTES / 2, QUERY_TILE_BYTES / 2, bar_q);
      long long t_wait = 0;
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % STAGES;
        const long long t0 = clock64();
        ptx::mbar_wait_backoff(bar_empty + st, ((i / STAGES) & 1) ^ 1, 128);
        t_wait += clock64() - t0;
        const int64_t g = g_lo + i;
        const int sg = g >= a.seg[0].tiles;
        const int64_t tile = sg ? a.seg[1].tile0 + (g - a.seg[0].tiles) : a.seg[0].tile0 + g;
        ptx::mbar_arrive_expect_tx(bar_full + st, KEY_TILE_BYTES);
        ptx::bulk_g2s(smem + SM_K + st * KEY_TILE_BYTES, a.seg[sg].image + tile * KEY_TILE_BYTES, KEY_TILE_BYTES,
                      bar_full + st);
      }
      if (dbg) dbg[0] = t_wait;
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0 && n_tiles > 0) {
      ptx::mbar_wait(bar_q, 0);
      const TileDescs descs = make_query_descs(ptx::smem_u32(smem + SM_Q));
      long long t_acc = 0, t_ld = 0, t_issue = 0;
      const long long t_begin = clock64();
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % STAGES, buf = i % ACC_BUFS;
        const long long t0 = clock64();
        ptx::mbar_wait_backoff(bar_tempty + buf, ((i / ACC_BUFS) & 1) ^ 1, 64);
        const long long t1 = clock64();
        ptx::mbar_wait_backoff(bar_full + st, (i / STAGES) & 1, 32);
        const long long t2 = clock64();
        t_acc += t1 - t0;
        t_ld += t2 - t1;
        ptx::tc_fence_after();
        issue_tile(descs, ptx::smem_u32(smem + SM_K + st * KEY_TILE_BYTES), tmem_base + buf * TK);
        ptx::umma_commit(bar_empty + st);   // key stage reusable once these MMAs have read it
        ptx::umma_commit(bar_tfull + buf);  // accumulator ready for the epilogue
        t_issue += clock64() - t2;
      }
      if (dbg) { dbg[1] = t_acc; dbg[2] = t_ld; dbg[3] = t_issue; dbg[4] = clock64() - t_begin; }
    }
  } else {
    // ===== epilogue: warps 0-3 own TMEM lanes 32*warp .. 32*warp+31 =====
    const int row = warp * 32 + lane;              // query row inside the tile
    ListState st;
    st.base_s = ptx::smem_u32(cs) + row * 4;
    st.base_i = ptx::smem_u32(ci) + row * 2;
    st.off_s = st.base_s;
    st.off_i = st.base_i;
    st.tau = -INFINITY;
    float b1[G], b2[G], b3[G];   // lower bounds of the best / 2nd / 3rd score of each key subset; slot 0 = current tile's
#pragma unroll
    for (int u = 0; u < G; ++u) b1[u] = b2[u] = b3[u] = -INFINITY;
    float *pub_mine = a.pub + (int64_t)blockIdx.y * a.hw_pad + qtile * TQ + warp * 32;
    long long t_wait = 0, t_relieve = 0;
    long long t_first = 0;
    const long long t_begin = clock64();
    for (int i = 0; i < n_tiles; ++i) {
      const int buf = i % ACC_BUFS;
      const long long tw0 = clock64();
      ptx::mbar_wait(bar_tfull + buf, (i / ACC_BUFS) & 1);
      t_wait += clock64() - tw0;
      ptx::tc_fence_after();
      uint32_t v[64];
      {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * TK;
        ptx::tmem_ld_32x32(taddr, v0);
        ptx::tmem_ld_32x32(taddr + 32, v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = v0[j]; v[32 + j] = v1[j]; }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + buf);  // accumulator buffer free again
      st.tau = fmaxf(st.tau, tau_sh[row]);

      // first / last tile of a candidate range: columns outside it never qualify
      {
        const int64_t g = g_lo + i;
        const int sg = g >= a.seg[0].tiles;
        const int64_t key0 = (sg ? a.seg[1].tile0 + (g - a.seg[0].tiles) : a.seg[0].tile0 + g) * TK;
        const int64_t kb = a.seg[sg].begin, ke = a.seg[sg].end;
        if (key0 < kb || key0 + TK > ke) {
          const int jlo = (int)max((int64_t)0, kb - key0), jhi = (int)min((int64_t)TK, ke - key0);
#pragma unroll
          for (int j = 0; j < TK; ++j)
            if (j < jlo || j >= jhi) v[j] = 0xff800000u;  // -inf
        }
      }
      // maxima of the eight 8-column groups: (a) which groups hold a survivor for any lane, (b) running
      // lower bounds of this split's three best scores (group maxima are scores of distinct keys)
      float gm[8];
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        float m = __uint_as_float(v[g8 * 8]);
#pragma unroll
        for (int jj = 1; jj < 8; ++jj) m = fmaxf(m, __uint_as_float(v[g8 * 8 + jj]));
        gm[g8] = m;
        const float t = fminf(b1[0], m);
        b1[0] = fmaxf(b1[0], m);
        const float u = fminf(b2[0], t);
        b2[0] = fmaxf(b2[0], t);
        b3[0] = fmaxf(b3[0], u);
      }
      {
        float pubv = b3[0];
#pragma unroll
        for (int u = 1; u < G; ++u) pubv = fminf(pubv, b3[u]);
        pub_mine[lane] = pubv;
        // rotate the subsets: the next tile updates the next one
        const float r1 = b1[0], r2 = b2[0], r3 = b3[0];
#pragma unroll
        for (int u = 0; u + 1 < G; ++u) { b1[u] = b1[u + 1]; b2[u] = b2[u + 1]; b3[u] = b3[u + 1]; }
        b1[G - 1] = r1; b2[G - 1] = r2; b3[G - 1] = r3;
        if (G == 1 && i == 0) {
          // First tile (single tracked subset, i.e. >= 11 splits): nothing is known yet and all 64 scores would be kept, overflowing every list.  The
          // tile sits in registers, so give the other splits a bounded moment to publish their first values
          // (the MMA warp keeps filling the other accumulator buffers meanwhile) and filter with the shared
          // threshold.  On a timeout (e.g. a grid of several waves) the lists overflow and get sorted instead.
          const long long t0 = clock64();
          float shared = tau_sh[row];
          while (__any_sync(FULL, shared == -INFINITY) && clock64() - t0 < FIRST_WAIT_CYCLES) {
            __nanosleep(100);
            shared = tau_sh[row];
          }
          st.tau = fmaxf(st.tau, shared);
          t_first = clock64() - t0;
        }
      }
      unsigned mine = 0;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) mine |= (gm[g8] > st.tau ? 1u : 0u) << g8;
      const unsigned active = __reduce_or_sync(FULL, mine);

      const uint32_t li0 = (uint32_t)i * TK;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        if (active & (1u << g8)) {   // warp-uniform; in steady state about half of the groups are skipped
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = g8 * 8 + jj;
            VOSMEM_APPEND("gt", st.off_s, st.off_i, __uint_as_float(v[j]), st.tau, li0 + j);
          }
          // lists that could overflow during the next 8 columns
          if (__any_sync(FULL, st.off_s > st.base_s + PRUNE_ABOVE * SS)) {
            const long long tr0 = clock64();
            st = relieve_lists(st, cs, ci, tau_sh + row, warp, lane);
            t_relieve += clock64() - tr0;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) atomicAdd(const_cast<uint32_t *>(epi_done), 1u);   // lets the refresher warp retire
    const long long t_loop = clock64() - t_begin;

    // ---- hand the surviving candidates to the merge kernel: one query at a time, lanes over its entries ----
    {
      __syncwarp();
      const int my_n = (int)((st.off_s - st.base_s) / SS);
      // candidate index of local index li:  64 * (g_lo + li / 64) + li % 64 + (offset of the segment)
      const int tiles0 = (int)a.seg[0].tiles, glo = (int)g_lo;
      const int off0 = (int)(a.seg[0].tile0 * TK - a.seg[0].begin);
      const int off1 = (int)(a.len0 + (a.seg[1].tile0 - a.seg[0].tiles) * TK - a.seg[1].begin);
      const int64_t row0 = (int64_t)blockIdx.y * a.hw_pad + qtile * TQ + warp * 32;
      if (qtile * TQ + row < a.hw) a.cand_count[row0 + lane] = my_n;
#pragma unroll 4
      for (int src = 0; src < 32; ++src) {
        if (qtile * TQ + warp * 32 + src >= a.hw) break;
        const int n = __shfl_sync(FULL, my_n, src);
        const int srow = warp * 32 + src;
        for (int e = lane; e < n; e += 32) {
          const int li = ci[e * CS_H + srow];
          const int g = glo + (li >> 6);
          a.cand_score[(row0 + src) * CAND_SLOTS + e] = cs[e * CS_F + srow];
          a.cand_index[(row0 + src) * CAND_SLOTS + e] = g * TK + (li & 63) + (g >= tiles0 ? off1 : off0);
        }
      }
    }
    if (dbg && lane == 0) {
      dbg[5 + warp * 2] = t_wait;
      dbg[6 + warp * 2] = t_relieve;
      if (warp == 0) { dbg[13] = t_loop; dbg[14] = clock64() - t_begin; }
      if (warp == 1) dbg[15] = t_first;
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
    issue_tile(make_query_descs(ptx::smem_u32(smem + SM_Q)), ptx::smem_u32(smem + SM_K), tmem_base);
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

long long *g_tc_debug = nullptr;  // set through vosmem_debug_set_timing_buffer

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
  VOSMEM_CHECK_ARG(ceil_div64(tiles, splits) * TK <= 65536,
                   "select(tcgen05): %lld key tiles over %d splits overflow the 16-bit candidate index", (long long)tiles,
                   splits);
  a.hw = d.hw;
  a.hw_pad = (int)round_up64(d.hw, TQ);
  a.splits = splits;
  a.query_image = ws.query_image;
  a.pub = ws.pub;
  a.cand_score = ws.cand_score;
  a.cand_index = ws.cand_index;
  a.cand_count = ws.cand_count;
  a.dbg = g_tc_debug;
  dim3 grid((unsigned)ceil_div64(d.hw, TQ), splits);
  // 3 * G * splits >= 33 keys must stand behind a shared threshold
  const int g = (11 + splits - 1) / splits;
#define VOSMEM_LAUNCH_TC(GG)                                                                                     \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      VOSMEM_CUDA(cudaFuncSetAttribute(select_tc_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)); \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    select_tc_kernel<GG><<<grid, TC_THREADS, SM_TOTAL, st>>>(a);                                                 \
  } while (0)
  if (g <= 1) VOSMEM_LAUNCH_TC(1);
  else if (g == 2) VOSMEM_LAUNCH_TC(2);
  else if (g == 3) VOSMEM_LAUNCH_TC(3);
  else if (g == 4) VOSMEM_LAUNCH_TC(4);
  else if (g <= 6) VOSMEM_LAUNCH_TC(6);
  else VOSMEM_LAUNCH_TC(11);
#undef VOSMEM_LAUNCH_TC
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

extern "C" int vosmem_debug_set_timing_buffer(void *device_buffer) {
  vosmem::g_tc_debug = static_cast<long long *>(device_buffer);  // 16 int64 per CTA of the next tcgen05 launches; NULL = off
  return VOSMEM_OK;
}
