// Fused similarity + per-query candidate selection on tcgen05 / TMEM (sm_100a, CK == 64).
//
// Replaces get_similarity + torch.topk of the reference (tracker/model/memory_util.py:7-39,46) for one
// object group: the N x HW similarity matrix only ever exists as 128 x 64 fp32 tiles in tensor memory.
//
//   grid  = (query tiles of 128, N-splits, problems)   one CTA per SM (~229 KB shared memory), 12 warps;
//           blockIdx.z selects one of up to MAX_BATCH independent problems (vosmem_match_batch)
//   prologue         : warps 0-7 pack the CTA's 128-query tile ([-e | 2 q e | -sum e q^2] as bf16 hi / lo, the
//                      shared-memory layout of a K-major no-swizzle UMMA operand) straight from the fp32 query key /
//                      selection -- no separate packing kernel, no query image in global memory
//   warp 8  producer : cp.async.bulk (TMA) of the streamed key tiles, mbarrier full/empty ring of STAGES stages
//   warp 9  MMA      : copies the query operand into tensor memory once (tcgen05.cp), then per key tile
//                      25 tcgen05.mma (M128 N64 K16, A from TMEM, B from shared memory; bf16 hi/lo split
//                      -> fp32) into one of ACC_BUFS accumulator buffers, tcgen05.commit -> mbarriers.  The whole
//                      warp runs the loop convergently with uniform operands (see the kernel body)
//   warps 10-11 refreshers: keep the per-query shared thresholds (below) fresh in shared memory
//   warps 0-7 epilogue: two warps per TMEM lane quarter, each with its own candidate lists and thresholds ("virtual
//                      splits"); the tiles of a quarter are handed out dynamically to its two warps.
//                      tcgen05.ld the tile (thread = query row), append every score above the
//                      thread's threshold to a private shared-memory candidate list with predicated
//                      stores (no branches).  Groups of 8 columns in which no lane has a survivor are
//                      skipped with one warp-wide OR.  When a list fills: drop what fell below the
//                      (risen) threshold, thread-privately; only if that is not enough the warp cuts
//                      lists to their best 32 with a bitonic network over packed 32-bit keys.
// Each CTA leaves <= 120 candidates per query in the exchange buffer; the merge (merge.cuh) finishes.
//
// GROUPS mode (opt-in: VOSMEM_TC_CANDIDATES=groups, see api.cu; measured against single keys in DESIGN.md 3.1).  Appending single scores costs the epilogue ~6.5 instructions per score of
// every 8-column group in which ANY lane of the warp has a survivor -- with ~1 survivor per query row and tile (short
// key streams: DAVIS, a rank's shard of a long bank) that is every group, and the epilogue, not the tensor pipe, sets
// the pace (~1 400 against 800 cycles per tile).  In GROUPS mode the candidates are the 8-column groups themselves,
// scored by their maximum: one compare and one predicated append per group -- the entry {maximum, position} into the
// shared-memory list, the group's 8 scores into a per-list log in global memory (see "GROUPS mode" below).
// Everything said about thresholds below holds with "group" for "key": the best 32 keys of a query lie in its best 32
// groups, a group maximum is the score of a real key, and maxima of distinct groups are scores of distinct keys.
// The consumer (merge.cuh) merges the lists into the best 32 groups and expands them into the best k keys from the
// 8 recorded scores -- the keys are never read again.
//
// Thresholds.  A query's threshold is only ever a LOWER bound of its true 32nd-best score, so no true
// top-k (k <= 32) member is dropped:
//   local : after a cooperative cut, the (truncated) 32nd-best score of this warp's own candidates;
//   shared: every thread tracks, in registers, lower bounds of the R best scores of the keys its warp has seen
//           (a sorted insertion of every score of the first tile when R >= 4, of the maxima of the 8-column groups
//           during the first tiles, of the tile maximum afterwards: maxima of disjoint column sets are scores of
//           distinct keys), R = ceil(33 / virtual splits), and publishes the R-th whenever it rises; every
//           cooperative cut also publishes the exact R-th best of the list it sorted.  The minimum over all virtual
//           splits then has at least 33 keys at or above it.  Published values carry the launch epoch (PubEntry),
//           so nothing has to be reset between launches.  With S splits running concurrently this tracks the
//           quality of a single pass over all keys, which makes list overflows (and sorting) rare.
//           With R == 2 (17-32 virtual splits: the DAVIS shape) the splits are combined four at a time instead, which
//           needs only 6 of a group's 8 published keys: see "Grouped thresholds" at refresh_grouped.
//
// Thresholds across ranks (N-sharded bank, vosmem_select_push with rank_pub set).  Sharding the key axis over G ranks
// would make every rank collect ITS OWN best 33 per query -- G times the appends of a single pass, which is what bounds
// this kernel on short key streams.  Instead the ranks share thresholds while their kernels run.  Every virtual split
// publishes a second bound, its R2-th best with R2 = ceil(33 / (G x virtual splits)) <= R (the tracker holds it anyway);
// the CTA of split 0 of every query tile keeps pushing the rank's summary (min over its virtual splits of those) into row
// `rank` of every rank's peer-mapped summary array (8-byte st.global over NVLink, a few KB per microsecond), and every
// CTA's refresher takes the minimum over the G rows of its LOCAL array: a bound backed by 33 keys of the whole bank.
// The threshold used is the larger of that and the rank's own bound (above), so a rank never does worse than alone --
// whatever the skew between the ranks' kernels -- and towards ~33 / G kept candidates per query once the rows flow.
// Rows carry the launch epoch, which all ranks advance in step (same call sequence on the engine's workspaces); a
// row that is missing or stale reads as -inf, i.e. no cross-rank bound yet.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vosmem {
namespace {

constexpr int STAGES = 3;
constexpr int ACC_BUFS = 5;
constexpr int TMEM_COLS = 512;            // 5 accumulator buffers (320 columns) + the query operand (136 columns)
constexpr int TMEM_A = ACC_BUFS * TK;     // first column of the query operand: [hi 64 | lo 64 | tail 8]
constexpr int HALVES = 2;                 // epilogue warps per TMEM lane quarter = virtual splits (candidate list sets) per CTA
constexpr int EPI_WARPS = 4 * HALVES;
constexpr int W_PRODUCER = EPI_WARPS, W_MMA = EPI_WARPS + 1, W_REFRESH = EPI_WARPS + 2;
constexpr int N_REFRESH = 2;                         // refresher warps: each keeps TQ / N_REFRESH rows of tau_sh fresh
constexpr int RH = TQ / 32 / N_REFRESH;              // rows per lane of a refresher warp
constexpr int TC_THREADS = (EPI_WARPS + 2 + N_REFRESH) * 32;   // 384
constexpr int CSLOTS = 60;             // candidate slots per (virtual split, query) in shared memory
constexpr int CS_E = TQ + 1;           // 8-byte {score, local index} entries per slot row (+1: bank spread)
constexpr int PRUNE_ABOVE = CSLOTS - 8;     // a list this long may overflow during the next 8 columns: relieve the warp
constexpr int SORT_ABOVE = CSLOTS - 16;     // lists still longer than this after the compaction are cut to their best 32
constexpr uint32_t SS = CS_E * 8;      // byte stride between consecutive slots of one list
constexpr uint32_t KEY_SLOT_MASK = 63u;           // low bits of a sort key hold the slot id
constexpr int FIRST_WAIT_CYCLES = 20000;          // bounded wait for the other virtual splits' first publication
constexpr int RB = 22;                            // published rows a refresher warp reads per batch (DAVIS: 22 rows = 1 batch)
constexpr int TRACK_TILES = 16;                   // tiles per warp whose group maxima (not just the tile maximum) feed the tracker

// shared memory map (bytes)
constexpr int SM_K = 0;                                 // key stages; the first two also stage the query image once
constexpr int SM_Q = SM_K;
static_assert(CSLOTS <= 64 && 2 * CSLOTS <= CAND_SLOTS, "list length vs sort width / exchange row");
static_assert(2 * KEY_TILE_BYTES == QUERY_TILE_BYTES && STAGES >= 2, "query image must fit the first two stages");
constexpr int LIST_BYTES = CSLOTS * CS_E * 8;
constexpr int SM_CS = SM_K + STAGES * KEY_TILE_BYTES;   // HALVES candidate lists
constexpr int SM_BAR = (SM_CS + HALVES * LIST_BYTES + 15) / 16 * 16;
constexpr int N_BARS = 2 * STAGES + 2 * ACC_BUFS + 2;
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;   // [0] TMEM base address, [1] epilogue warps finished
constexpr int SM_TAU = SM_TMEM + 16;           // TQ floats: shared threshold per query row, kept fresh by the refresher
constexpr int SM_NA = SM_TAU + TQ * 4;         // TQ ints: list length of warp set 0 per query row (for the hand-off)
constexpr int SM_CTR = SM_NA + TQ * 4;          // 4 ints: next tile of each TMEM lane quarter (dynamic hand-out to its two warps)
constexpr int SM_TOTAL = SM_CTR + 16;
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
  const float *qk, *qe;   // query key / selection, CK x hw fp32 (qe may be NULL: isotropic)
  PubEntry *pub;          // [virtual splits][hw_pad] published lower bound per (virtual split, query): see "Thresholds"
  PubEntry *pub2;         // the same for rank R2 ("Thresholds across ranks"; only written / read when world > 1)
  WsControl *ctl;         // device-side launch epoch (tags this launch's published values), departure counter, error flags
  CandEntry *cand;
  int *cand_count;
  float *gscore;       // GROUPS mode: 8 scores per entry of `cand` (then laid out [virtual split][hw_pad][GSLOTS])
  float *glog;         // GROUPS mode: score logs, VOSMEM_GROUP_LOG records of 8 scores per (virtual split, query)
  int exp;             // measurement switches (VOSMEM_TC_EXP, 0 in production): 1 = GROUPS mode appends nothing (floor of the
                       // epilogue without its stores), 4 = every CTA returns at entry (scripts/launch_cost.py)
  long long *dbg;      // optional per-CTA cycle counters (32 per CTA), NULL in production
};

struct TcBatch {   // kernel parameter: one TcArgs per problem (blockIdx.z)
  TcArgs p[MAX_BATCH];
  // thresholds across ranks (problem 0 only; world == 1: off)
  int world, rank, r2;
  PubEntry *rank_pub[VOSMEM_MAX_RANKS];   // [d]: rank d's [world][hw_pad] summary array; [rank] is the local one
};
static_assert(sizeof(TcBatch) <= 4000, "kernel parameter space");

// The query operand lives in tensor memory: 8 columns per K=16 step, [hi steps 0-7 | lo steps 0-7 | tail].
// One tcgen05.cp (128 rows x 256 bits) per step from the packed shared-memory image.
__device__ __forceinline__ void stage_query_in_tmem(uint32_t q_base, uint32_t tmem_a) {
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    ptx::tmem_cp_128x256b_elect(tmem_a + 8 * s, ptx::umma_desc(q_base + (2 * s) * Q_LBO, Q_LBO, SBO));
    ptx::tmem_cp_128x256b_elect(tmem_a + 64 + 8 * s, ptx::umma_desc(q_base + (16 + 2 * s) * Q_LBO, Q_LBO, SBO));
  }
  ptx::tmem_cp_128x256b_elect(tmem_a + 128, ptx::umma_desc(q_base + 32 * Q_LBO, Q_LBO, SBO));
}
// 25 MMAs of one 128 x 64 tile: hi*hi + lo*hi + hi*lo over the 128 packed channels, then the rank-1 tail.
// Only the key operand is fetched from shared memory (2 KB per MMA), so the tensor pipe is not operand-bound.
__device__ __forceinline__ void issue_tile(uint32_t tmem_a, uint32_t k_base, uint32_t tmem_d) {
  const uint64_t k0 = ptx::umma_desc(k_base, K_LBO, SBO);
  constexpr uint64_t STEP = (2 * K_LBO) >> 4;   // two 8-element chunks per K=16 step, in 16-byte units
#pragma unroll
  for (int s = 0; s < 8; ++s) ptx::umma_bf16_ts_elect(tmem_d, tmem_a + 8 * s, k0 + s * STEP, IDESC, s > 0);
#pragma unroll
  for (int s = 0; s < 8; ++s) ptx::umma_bf16_ts_elect(tmem_d, tmem_a + 64 + 8 * s, k0 + s * STEP, IDESC, 1);
#pragma unroll
  for (int s = 0; s < 8; ++s) ptx::umma_bf16_ts_elect(tmem_d, tmem_a + 8 * s, k0 + (8 + s) * STEP, IDESC, 1);
  ptx::umma_bf16_ts_elect(tmem_d, tmem_a + 128, k0 + 16 * STEP, IDESC, 1);
}

// cycle counter for the optional per-role timing: only read when a timing buffer is attached (the reads and the
// bookkeeping around them are not free in the MMA issue loop)
#define TICK() (dbg ? clock64() : 0ll)

struct ListState {
  uint32_t base;   // shared-memory byte address of slot 0 of this thread's list
  uint32_t off;    // next free slot
  float tau;       // current threshold (lower bound of this query's true 32nd-best score)
  float pub;       // lower bound of the `rank`-th best score of this list's keys that this thread has published
};
struct Entry {
  float score;
  uint32_t index;  // key position inside this CTA's key stream: 64 * tile + column
};

__device__ __forceinline__ Entry lds_entry(uint32_t addr) {
  Entry e;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(e.score), "=r"(e.index) : "r"(addr));
  return e;
}
// if (score CMP tau) { *off = {score, index}; off += stride }   -- predicated, branch-free
#define VOSMEM_APPEND(CMP, off, score, tau, index)                                                  \
  asm volatile(                                                                                     \
      "{\n\t.reg .pred p;\n\t"                                                                      \
      "setp." CMP ".f32 p, %1, %2;\n\t"                                                             \
      "@p st.shared.v2.b32 [%0], {%1, %3};\n\t"                                                     \
      "@p add.u32 %0, %0, %4;\n\t}"                                                                 \
      : "+r"(off)                                                                                   \
      : "f"(score), "f"(tau), "r"(index), "n"(CS_E * 8)                                             \
      : "memory")

// Called when some list of the warp could overflow during the next 8 columns.  Kept out of line (eight call
// sites in the unrolled epilogue).  Three steps, cheapest first:
//   1. pick up the latest shared threshold (warp 6 keeps it fresh in shared memory);
//   2. every thread drops the entries of ITS list that fell below its threshold (thread-private, in place);
//   3. lists that are still too long are cut to their best 32 by the whole warp, four queries per round
//      (bitonic network over packed 32-bit keys), which also yields a new local threshold.
// Every thread drops the entries of its own list that are below its threshold (in place, order kept).
// The keep test is one float below the append threshold: after a cooperative cut the append threshold is "strictly
// above the list's 32nd best" (see relieve_lists), and the entries that tie with that 32nd best must stay.
__device__ __forceinline__ ListState compact_list(ListState st) {
  const int cnt = (int)((st.off - st.base) / SS);
  const int nmax = __reduce_max_sync(FULL, cnt);
  const float keep = nextafterf(st.tau, -INFINITY);
  uint32_t rd = st.base, wr = st.base;
  for (int e = 0; e < nmax; ++e) {
    const bool in = rd < st.off;
    Entry en = lds_entry(in ? rd : st.base);
    if (!in) en.score = __int_as_float(0x7fc00000);  // NaN never passes the compare
    VOSMEM_APPEND("ge", wr, en.score, keep, en.index);
    rd += SS;
  }
  st.off = wr;
  return st;
}

// (state goes in and comes back by value so that it stays in registers across the call)
__device__ __noinline__ ListState relieve_lists(ListState st, Entry *cs, const volatile float *tau_row, int quarter,
                                                int lane, int rank) {
  st.tau = fmaxf(st.tau, *tau_row);
  // ---- 2. thread-private compaction ----
  st = compact_list(st);
  // ---- 3. cooperative cut to the best 32 ----
  unsigned full = __ballot_sync(FULL, st.off > st.base + SORT_ABOVE * SS);
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
      row[u] = quarter * 32 + src[u];
      const int n = (int)((__shfl_sync(FULL, st.off, src[u]) - __shfl_sync(FULL, st.base, src[u])) / SS);
      // packed key: order-preserving score bits, low 6 bits = slot id (distinct keys; the truncation can only
      // lower a threshold derived from a key, which keeps it a valid lower bound)
      ka[u] = lane < n ? ((f2ord(cs[lane * CS_E + row[u]].score) & ~KEY_SLOT_MASK) | lane) : 0u;
      kb[u] = lane + 32 < n ? ((f2ord(cs[(lane + 32) * CS_E + row[u]].score) & ~KEY_SLOT_MASK) | (lane + 32)) : 0u;
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
    // best 32 of the (up to) 48: a (descending) against b reversed (ascending), then a bitonic merge -> sorted
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
    Entry sv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sv[u] = cs[(kc[u] & KEY_SLOT_MASK) * CS_E + row[u]];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (on[u]) {
        cs[lane * CS_E + row[u]] = sv[u];
        // This list now holds 32 keys scoring >= their exact minimum, so a later key only matters if it beats that
        // minimum STRICTLY (a tie is interchangeable with the ones held, as in the reference's torch.topk): without
        // this, exact ties (duplicated frames, uniform regions) would refill and re-sort the list for ever.
        const float exact32 = warp_min(sv[u].score);
        const float at_rank = ord2f(__shfl_sync(FULL, kc[u], rank - 1) & ~KEY_SLOT_MASK);   // exact (truncated) rank-th best
        if (lane == src[u]) {
          st.off = st.base + 32 * SS;
          st.tau = fmaxf(st.tau, nextafterf(exact32, INFINITY));
          st.pub = fmaxf(st.pub, at_rank);
        }
      }
    }
    __syncwarp();
  }
  return st;
}

// ---- GROUPS mode ---------------------------------------------------------------------------------------------------
// The list entries {group maximum, log position << 20 | group index in the CTA's stream} live in the same shared-memory
// lists as the single-key entries (compact_list / relieve_lists treat the index as an opaque payload); the groups' 8
// scores go to a per-list LOG in global memory: one full 32-byte sector per append (a partial-sector store would make
// L2 fetch the rest of the sector first -- measured: three 8 / 16-byte stores per append cost ~160 cycles), written at a
// position that only ever advances, so records never move when the list is compacted or cut.  The log has two halves
// of LOG_HALF records; when the active half is full the (<= CSLOTS) live records are copied to the other half
// (flip_log: rare -- a list sees ~10 appends at the DAVIS shape, ~80 over a 100 000-key stream).
constexpr int LOG_HALF = VOSMEM_GROUP_LOG / 2;
constexpr uint32_t GIDX_BITS = 20, GIDX_MASK = (1u << GIDX_BITS) - 1;   // group index in the CTA's stream: 8 x tile + group
static_assert(LOG_HALF >= CSLOTS + TK / 8, "a flipped log must have room for one more tile");

struct Rec8 { uint32_t w[8]; };
__device__ __forceinline__ Rec8 ldg_rec(const float *p) {
  Rec8 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void stg_rec(float *p, const Rec8 &r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]), "r"(r.w[3]),
               "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7]) : "memory");
}
// if (gmax >= thr) { list[off] = {gmax, lp << 20 | gidx}; log[lp] = 8 scores; off += stride; ++lp }   -- predicated
#define VOSMEM_APPEND_GROUP(off, lp, logp, thr, gmax, gidx, v, j0)                                                   \
  asm volatile(                                                                                                     \
      "{\n\t.reg .pred p;\n\t.reg .b64 ar;\n\t.reg .b32 ix;\n\t"                                                \
      "setp.ge.f32 p, %3, %4;\n\t"                                                                                  \
      "mad.wide.u32 ar, %1, 32, %2;\n\t"                                                                            \
      "shl.b32 ix, %1, 20;\n\t"                                                                                     \
      "or.b32 ix, ix, %5;\n\t"                                                                                      \
      "@p st.shared.v2.b32 [%0], {%6, ix};\n\t"                                                                     \
      "@p st.global.v8.b32 [ar], {%7, %8, %9, %10, %11, %12, %13, %14};\n\t"                                        \
      "@p add.u32 %0, %0, %15;\n\t"                                                                                 \
      "@p add.u32 %1, %1, 1;\n\t}"                                                                                  \
      : "+r"(off), "+r"(lp)                                                                                         \
      : "l"(logp), "f"(gmax), "f"(thr), "r"(gidx), "r"(__float_as_uint(gmax)), "r"((v)[(j0)]), "r"((v)[(j0) + 1]),   \
        "r"((v)[(j0) + 2]), "r"((v)[(j0) + 3]), "r"((v)[(j0) + 4]), "r"((v)[(j0) + 5]), "r"((v)[(j0) + 6]),           \
        "r"((v)[(j0) + 7]), "n"(CS_E * 8)                                                                            \
      : "memory")

// The active half of this thread's log is (nearly) full: copy the records its list still refers to into the other half,
// in list order, and re-point the entries.  Thread-private (every lane of the warp runs it; lanes with short lists
// just move few records).  Returns the new log position.
// (inlined: cicc 12.9 crashes on this function when it is kept out of line)
__device__ __forceinline__ uint32_t flip_log(ListState st, float *logp, uint32_t lp) {
  const uint32_t other = lp >= (uint32_t)LOG_HALF ? 0u : (uint32_t)LOG_HALF;   // first record of the other half
  const int n = (int)((st.off - st.base) / SS);
  const int nmax = __reduce_max_sync(FULL, n);
  for (int e0 = 0; e0 < nmax; e0 += 4) {
    Rec8 r[4];
    uint32_t ix[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool in = e0 + u < n;
      ix[u] = in ? lds_entry(st.base + (e0 + u) * SS).index : 0u;
      r[u] = ldg_rec(logp + (size_t)(ix[u] >> GIDX_BITS) * GROUP_KEYS);   // (lanes past their list read record 0)
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (e0 + u < n) {
        stg_rec(logp + (size_t)(other + e0 + u) * GROUP_KEYS, r[u]);
        const uint32_t addr = st.base + (e0 + u) * SS + 4, val = ((other + e0 + u) << GIDX_BITS) | (ix[u] & GIDX_MASK);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
      }
    }
  }
  return other + n;
}

// One pass of the refresher warp over the published rows: NB x 4 independent 8-byte loads in flight per lane.  Rows
// past the end re-read the last row (harmless for a minimum), so there is no branch between the loads and they are
// all issued before the first use.
template <int NB>
__device__ __forceinline__ void refresh_pass(const PubEntry *pub_row, int vsplits, int hw_pad, uint32_t epoch,
                                             float (&m)[RH]) {
  for (int y0 = 0; y0 < vsplits; y0 += NB) {
    uint2 raw[NB][RH];
#pragma unroll
    for (int y = 0; y < NB; ++y) {
      const int yy = min(y0 + y, vsplits - 1);
#pragma unroll
      for (int h = 0; h < RH; ++h)
        raw[y][h] = __ldcg(reinterpret_cast<const uint2 *>(pub_row + (int64_t)yy * hw_pad + 32 * h));
    }
#pragma unroll
    for (int y = 0; y < NB; ++y)
#pragma unroll
      for (int h = 0; h < RH; ++h)
        m[h] = fminf(m[h], raw[y][h].y == epoch ? __uint_as_float(raw[y][h].x) : -INFINITY);
  }
}

// "Grouped thresholds" (R == 2, i.e. 17-32 virtual splits, the DAVIS shape): every virtual split publishes its best
// score too (in `pub2`, free when no thresholds cross ranks), and the refreshers combine the splits four at a time.
// A group of four has 8 published scores of 8 distinct keys, so its j-th largest published value has j keys of the group
// at or above it; groups hold disjoint key sets, so the minimum over the groups of their want_g-th largest values is a
// valid bound as soon as the want_g sum to >= 33.  With 22 virtual splits: five groups give their 6th largest, the
// remaining pair its 3rd (of 4) -- the bound then sits near overall rank 66 instead of ~120 for the plain minimum of the
// 22 second-best scores (every split must hold 2 of the best keys for that one), i.e. ~1.8x fewer appends and candidates.
// The three smallest of a pair's four values, ascending (b1 >= b2 per split after the clamp):
struct Low3 { float x1, x2, x3; };
__device__ __forceinline__ Low3 pair_low3(float u1, float u2, float v1, float v2) {
  u1 = fmaxf(u1, u2);   // (a cut may have published an exact 2nd best above the tracker's best; an entry not yet
  v1 = fmaxf(v1, v2);   //  written reads -inf)
  const float s = fminf(u2, v2), S = fmaxf(u2, v2), p = fminf(u1, v1);
  return Low3{s, fminf(S, p), fmaxf(S, p)};
}
// the (9 - want)-th smallest of a group of four = want-th largest of its 8 values, want in {6, 7, 8}
__device__ __forceinline__ float quad_bound(const Low3 &x, const Low3 &y, int want) {
  if (want >= 8) return fminf(x.x1, y.x1);
  if (want == 7) return fminf(fminf(x.x2, y.x2), fmaxf(x.x1, y.x1));
  return fminf(fminf(x.x3, y.x3), fminf(fmaxf(x.x1, y.x2), fmaxf(x.x2, y.x1)));
}
constexpr int GB = 12;   // virtual splits per batch of loads (three groups of four)
__device__ __forceinline__ void refresh_grouped(const PubEntry *pub_row, const PubEntry *pub2_row, int vsplits, int hw_pad,
                                                uint32_t epoch, float (&m)[RH]) {
  // how many keys each group has to vouch for: quads 6 (+ up to 2), the trailing pair 3 (+ 1), 33 in total
  const int nq = vsplits >> 2, np = (vsplits >> 1) & 1;
  int deficit = 33 - (6 * nq + 3 * np);
  deficit = deficit > 0 ? deficit : 0;
  const int pair_want = 3 + (np && deficit > 0 ? 1 : 0);
  deficit -= pair_want - 3;
  const int q_add = nq ? deficit / nq : 0, q_rem = nq ? deficit % nq : 0;   // (vsplits >= 17: q_add + 1 <= 2)
  for (int y0 = 0; y0 < vsplits; y0 += GB) {
    uint2 r2[GB][RH], r1[GB][RH];   // pub: 2nd best, pub2: best
#pragma unroll
    for (int y = 0; y < GB; ++y) {
      const int yy = min(y0 + y, vsplits - 1);
#pragma unroll
      for (int h = 0; h < RH; ++h) {
        r2[y][h] = __ldcg(reinterpret_cast<const uint2 *>(pub_row + (int64_t)yy * hw_pad + 32 * h));
        r1[y][h] = __ldcg(reinterpret_cast<const uint2 *>(pub2_row + (int64_t)yy * hw_pad + 32 * h));
      }
    }
#pragma unroll
    for (int g = 0; g < GB / 4; ++g) {
      const int left = vsplits - (y0 + 4 * g);   // virtual splits from this group's first on (warp-uniform)
      if (left <= 0) break;
      const int qi = (y0 >> 2) + g;
      const int want = 6 + q_add + (qi < q_rem ? 1 : 0);
#pragma unroll
      for (int h = 0; h < RH; ++h) {
        float b1[4], b2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          b2[u] = r2[4 * g + u][h].y == epoch ? __uint_as_float(r2[4 * g + u][h].x) : -INFINITY;
          b1[u] = r1[4 * g + u][h].y == epoch ? __uint_as_float(r1[4 * g + u][h].x) : -INFINITY;
        }
        const Low3 x = pair_low3(b1[0], b2[0], b1[1], b2[1]);
        float bound;
        if (left >= 4) bound = quad_bound(x, pair_low3(b1[2], b2[2], b1[3], b2[3]), want);
        else bound = pair_want >= 4 ? x.x1 : x.x2;
        m[h] = fminf(m[h], bound);
      }
    }
  }
}

// Warps 0-7 (256 threads): the query operand of this CTA's 128 queries, written in place as the shared-memory image
// the tcgen05.cp copies expect.  Row y[q] = [-e | 2 q e | -sum_c e q^2] (memory_util.py:20-27; e = 1 and no last term
// when there is no selection, :28-32), every fp32 entry as a bf16 (hi, lo) pair; 16-byte chunk order per row:
// [y1_hi 0-7 | y2_hi 8-15 | y1_lo 16-23 | y2_lo 24-31 | tail 32 | 0 33], tail = (y3_hi, y3_lo, y3_hi, 0...).
// Thread = (query row, half of the channels); loads are coalesced over the rows, stores are conflict-free.
__device__ __forceinline__ void pack_query_tile(const TcArgs &a, int qtile, unsigned char *tile, float *red) {
  const int r = threadIdx.x & (TQ - 1), hsel = threadIdx.x >> 7;
  const int q = qtile * TQ + r;
  const bool live = q < a.hw;
  // all 64 loads are issued before the first use (no branch in between): rows past the end read the last query
  // and are zeroed afterwards
  float kk[32], ee[32];
  const int64_t col = live ? q : a.hw - 1;
#pragma unroll
  for (int j = 0; j < 32; ++j) kk[j] = __ldg(a.qk + (int64_t)(hsel * 32 + j) * a.hw + col);
  if (a.qe) {
#pragma unroll
    for (int j = 0; j < 32; ++j) ee[j] = __ldg(a.qe + (int64_t)(hsel * 32 + j) * a.hw + col);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) ee[j] = 1.f;
  }
  if (!live) {
#pragma unroll
    for (int j = 0; j < 32; ++j) kk[j] = ee[j] = 0.f;
  }
  float y3 = 0.f;
#pragma unroll
  for (int gg = 0; gg < 4; ++gg) {
    const int g = hsel * 4 + gg;
    uint32_t h1[4], l1[4], h2[4], l2[4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      __nv_bfloat16 hi[2][2], lo[2][2];   // [y1 | y2][element of the pair]
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float k = kk[gg * 8 + jp * 2 + u], e = ee[gg * 8 + jp * 2 + u];
        split_bf16(-e, hi[0][u], lo[0][u]);
        split_bf16(2.0f * (k * e), hi[1][u], lo[1][u]);
        if (a.qe) y3 -= e * (k * k);
      }
      auto pack2 = [](__nv_bfloat16 x, __nv_bfloat16 y) {
        return (uint32_t)__bfloat16_as_ushort(x) | ((uint32_t)__bfloat16_as_ushort(y) << 16);
      };
      h1[jp] = pack2(hi[0][0], hi[0][1]); l1[jp] = pack2(lo[0][0], lo[0][1]);
      h2[jp] = pack2(hi[1][0], hi[1][1]); l2[jp] = pack2(lo[1][0], lo[1][1]);
    }
    *reinterpret_cast<uint4 *>(tile + image_offset<TQ>(r, g)) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
    *reinterpret_cast<uint4 *>(tile + image_offset<TQ>(r, 8 + g)) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
    *reinterpret_cast<uint4 *>(tile + image_offset<TQ>(r, 16 + g)) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
    *reinterpret_cast<uint4 *>(tile + image_offset<TQ>(r, 24 + g)) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
  }
  if (hsel == 1) red[r] = y3;
  asm volatile("bar.sync 6, 256;" ::: "memory");   // the eight packing warps
  if (hsel == 0) {
    y3 += red[r];
    __nv_bfloat16 hi, lo;
    split_bf16(y3, hi, lo);
    const uint32_t h = __bfloat16_as_ushort(hi), l = __bfloat16_as_ushort(lo);
    *reinterpret_cast<uint4 *>(tile + image_offset<TQ>(r, 32)) = make_uint4(h | (l << 16), h, 0u, 0u);
    *reinterpret_cast<uint4 *>(tile + image_offset<TQ>(r, 33)) = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async();   // the tcgen05.cp copies read the image through the async proxy
}

// R = tracked / published rank per virtual split (see "Thresholds"); SHARED = thresholds across ranks (a separate
// instantiation: the extra uniform state of that path costs the MMA warp its register -> uniform-register-free issue
// loop, 39.4 against 37.4 us at the DAVIS shape, so single-GPU launches do not carry it)
// GROUPS = candidates are 8-column groups appended to global lists (header comment)
template <int R, bool SHARED, bool GROUPS>
__global__ void __launch_bounds__(TC_THREADS, 1) select_tc_kernel(const __grid_constant__ TcBatch batch) {
  const TcArgs &a = batch.p[blockIdx.z];
  constexpr bool PAIRED = !SHARED && R == 2;   // "Grouped thresholds" (refresh_grouped): the best score is published too
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + SM_BAR);
  uint64_t *bar_empty = bar_full + STAGES;
  uint64_t *bar_tfull = bar_empty + STAGES;
  uint64_t *bar_tempty = bar_tfull + ACC_BUFS;
  uint64_t *bar_q = bar_tempty + ACC_BUFS;       // query image landed in shared memory
  uint64_t *bar_qdone = bar_q + 1;               // query operand copied to tensor memory: stages are free
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
  volatile uint32_t *epi_done = reinterpret_cast<volatile uint32_t *>(smem + SM_TMEM + 4);
  volatile float *tau_sh = reinterpret_cast<volatile float *>(smem + SM_TAU);

  // warp index through a shuffle: provably warp-uniform, so the role branches below are known to be convergent and
  // the MMA warp's address arithmetic can live in uniform registers
  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int qtile = blockIdx.x;
  // launch epoch: a device-side word, so that a replayed CUDA graph runs under a fresh tag every time.  Nobody
  // writes it before the last CTA of this launch leaves (bottom of the kernel).
  const uint32_t epoch = *reinterpret_cast<volatile const uint32_t *>(&a.ctl->epoch);
  const int vsplits = a.splits * HALVES;         // virtual splits: rows of pub / the exchange buffers
  const int64_t g_lo = a.tiles_total * blockIdx.y / a.splits;
  const int64_t g_hi = a.tiles_total * (blockIdx.y + 1) / a.splits;
  const int n_tiles = (int)(g_hi - g_lo);
  long long *dbg = a.dbg ? a.dbg + ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 : nullptr;
  if (a.exp & 4) return;   // experiment: launch + scheduling cost of this grid alone
  const long long t_entry = TICK();
  unsigned long long g_entry = 0;
  if (dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(bar_full + i, 1); ptx::mbar_init(bar_empty + i, 1); }
    for (int i = 0; i < ACC_BUFS; ++i) { ptx::mbar_init(bar_tfull + i, 1); ptx::mbar_init(bar_tempty + i, 4); }
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_qdone, 1);
    ptx::fence_barrier_init();
    *epi_done = 0;
    for (int i = 0; i < 4; ++i) reinterpret_cast<int *>(smem + SM_CTR)[i] = 0;
  }
  if (threadIdx.x < TQ) tau_sh[threadIdx.x] = -INFINITY;
  if (warp == W_MMA) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (warp < EPI_WARPS) pack_query_tile(a, qtile, smem + SM_Q, reinterpret_cast<float *>(smem + SM_NA));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // The CTA owns the SM's whole tensor memory (512 columns, one CTA per SM), so the allocation can only start at
  // column 0, lane 0.  Treating the base as the constant 0 makes every tcgen05 address a compile-time function of
  // uniform loop counters: the MMA warp then needs no register -> uniform-register moves per instruction and is no
  // longer bound by its own issue rate.  Anything else is a configuration this kernel was not built for.
  // Not a trap (which would take the whole context down): the CTA flags the workspace, skips its work and leaves
  // through the common exit; vosmem_workspace_status reports it.
  const bool tmem_ok = __shfl_sync(FULL, *tmem_slot, 0) == 0u;   // (shuffle: provably warp-uniform, like `warp`)
  if (!tmem_ok && threadIdx.x == 0) atomicOr(&a.ctl->error, WS_ERR_TMEM_BASE);
  constexpr uint32_t tmem_base = 0u;
  if (dbg && threadIdx.x == 0) dbg[21] = TICK() - t_entry;   // prologue

  if (!tmem_ok) {
    // (no role runs; the candidate counts of this CTA's queries stay whatever they were: the caller sees the error flag)
  } else if (warp >= W_REFRESH) {
    // ===== threshold refreshers: tau_sh[row] = min over the virtual splits of their published lower bound; each of the
    //       N_REFRESH warps owns TQ / N_REFRESH rows, so a pass over 22 virtual splits is one batch of loads =====
    // (a stale value is still a valid lower bound, so no ordering with the epilogue is needed)
    const int row0 = (warp - W_REFRESH) * (TQ / N_REFRESH);
    const PubEntry *pub_row = a.pub + qtile * TQ + row0 + lane;
    const int world = SHARED ? batch.world : 1, my_rank = SHARED ? batch.rank : 0;
    const bool pusher = SHARED && blockIdx.y == 0;         // one CTA per query tile publishes the rank's summary
    float pushed[RH], across[RH];   // last summary pushed / last bound obtained across the ranks
#pragma unroll
    for (int h = 0; h < RH; ++h) pushed[h] = across[h] = -INFINITY;
    for (int it = 0; *epi_done < EPI_WARPS; ++it) {
      float m[RH];
#pragma unroll
      for (int h = 0; h < RH; ++h) m[h] = INFINITY;
      // rows per batch sized to the number of virtual splits (2: one CTA per query tile, e.g. batched sequences;
      // 4: two splits, LVOS-size query counts; else 22 at a time), so that no pass re-reads rows for nothing
      if constexpr (PAIRED) refresh_grouped(pub_row, a.pub2 + qtile * TQ + row0 + lane, vsplits, a.hw_pad, epoch, m);
      else if (vsplits <= 2) refresh_pass<2>(pub_row, vsplits, a.hw_pad, epoch, m);
      else if (vsplits <= 4) refresh_pass<4>(pub_row, vsplits, a.hw_pad, epoch, m);
      else refresh_pass<RB>(pub_row, vsplits, a.hw_pad, epoch, m);
      if (SHARED && (it < 8 || (it & 3) == 0)) {
        // ---- thresholds across ranks: g = this rank's summary (min over its virtual splits of their R2-th best);
        //      push it when it rose, fold in the other ranks' rows, use the larger of the local and the global bound.
        //      Every fourth pass after the start: these bounds move slowly, the local ones must stay fresh ----
        const int64_t col = qtile * TQ + row0 + lane;
        const PubEntry *pub2_row = a.pub2 + qtile * TQ + row0 + lane;
        float g[RH];
#pragma unroll
        for (int h = 0; h < RH; ++h) g[h] = INFINITY;
        if (vsplits <= 2) refresh_pass<2>(pub2_row, vsplits, a.hw_pad, epoch, g);
        else if (vsplits <= 4) refresh_pass<4>(pub2_row, vsplits, a.hw_pad, epoch, g);
        else refresh_pass<RB>(pub2_row, vsplits, a.hw_pad, epoch, g);
        if (pusher) {
#pragma unroll
          for (int h = 0; h < RH; ++h) {
            if (g[h] > pushed[h]) {
              pushed[h] = g[h];
              for (int d = 0; d < world; ++d)
                if (d != my_rank) pub_store(batch.rank_pub[d] + (int64_t)my_rank * a.hw_pad + col + 32 * h, g[h], epoch);
            }
          }
        }
        const PubEntry *mine_rows = batch.rank_pub[my_rank];
        for (int s0 = 0; s0 < world; s0 += 8) {      // eight ranks' rows in flight at a time
          uint2 raw[8][RH];
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const int ss = s0 + s < world ? s0 + s : my_rank;   // (own row is never used: the fresh local value stands in)
#pragma unroll
            for (int h = 0; h < RH; ++h)
              raw[s][h] = __ldcg(reinterpret_cast<const uint2 *>(mine_rows + (int64_t)ss * a.hw_pad + col + 32 * h));
          }
#pragma unroll
          for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int h = 0; h < RH; ++h)
              if (s0 + s < world && s0 + s != my_rank)
                g[h] = fminf(g[h], raw[s][h].y == epoch ? __uint_as_float(raw[s][h].x) : -INFINITY);
        }
#pragma unroll
        for (int h = 0; h < RH; ++h) across[h] = fmaxf(across[h], g[h]);
      }
      if (SHARED) {
#pragma unroll
        for (int h = 0; h < RH; ++h) m[h] = fmaxf(m[h], across[h]);
      }
#pragma unroll
      for (int h = 0; h < RH; ++h) tau_sh[row0 + lane + 32 * h] = m[h];
      if (it >= 16) __nanosleep(256);   // thresholds move fastest during the first tiles
    }
  } else if (warp == W_PRODUCER) {
    // ===== producer =====
    if (lane == 0 && n_tiles > 0) {
      // stages 0-1 hold the query image until the MMA warp has copied it to tensor memory: key tile i goes to stage
      // (i + 2) % STAGES, so the first key tile streams in meanwhile
      long long t_wait = 0;
      // both segments' tile streams as plain scalars (the descriptor sits in parameter space behind blockIdx.z)
      const int64_t tiles0 = a.seg[0].tiles;
      const unsigned char *src0 = a.seg[0].image + (a.seg[0].tile0 + g_lo) * KEY_TILE_BYTES;
      const unsigned char *src1 = a.seg[1].image + (a.seg[1].tile0 + g_lo - tiles0) * KEY_TILE_BYTES;
      for (int i = 0; i < n_tiles; ++i) {
        const int st = (i + 2) % STAGES;
        if (i == 1) ptx::mbar_wait_backoff(bar_qdone, 0, 64);
        const long long t0 = TICK();
        ptx::mbar_wait_backoff(bar_empty + st, ((i / STAGES) & 1) ^ 1, 20);
        t_wait += TICK() - t0;
        const unsigned char *src = (g_lo + i >= tiles0 ? src1 : src0) + (int64_t)i * KEY_TILE_BYTES;
        ptx::mbar_arrive_expect_tx(bar_full + st, KEY_TILE_BYTES);
        ptx::bulk_g2s(smem + SM_K + st * KEY_TILE_BYTES, src, KEY_TILE_BYTES, bar_full + st);
      }
      if (dbg) dbg[0] = t_wait;
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer: the whole warp runs the loop in lock step (every operand is warp-uniform by construction, so
    //       the uniform-register operands of UTCHMMA need no per-instruction waterfall loop); one elected lane issues
    if (n_tiles > 0) {
      const uint32_t tmem_a = tmem_base + TMEM_A;
      stage_query_in_tmem(ptx::smem_u32(smem + SM_Q), tmem_a);   // the image was packed before the CTA-wide barrier
      ptx::umma_commit_elect(bar_qdone);   // arrives once the copies have read shared memory
      long long t_acc = 0, t_ld = 0, t_issue = 0;
      const long long t_begin = TICK();
      for (int i = 0; i < n_tiles; ++i) {
        const int st = (i + 2) % STAGES, buf = i % ACC_BUFS;
        const long long t0 = TICK();
        ptx::mbar_wait_backoff(bar_tempty + buf, ((i / ACC_BUFS) & 1) ^ 1, 20);
        const long long t1 = TICK();
        ptx::mbar_wait_backoff(bar_full + st, (i / STAGES) & 1, 20);
        const long long t2 = TICK();
        t_acc += t1 - t0;
        t_ld += t2 - t1;
        ptx::tc_fence_after();
        issue_tile(tmem_a, ptx::smem_u32(smem + SM_K + st * KEY_TILE_BYTES), tmem_base + buf * TK);
        ptx::umma_commit_elect(bar_empty + st);   // key stage reusable once these MMAs have read it
        ptx::umma_commit_elect(bar_tfull + buf);  // accumulator ready for the epilogue
        t_issue += TICK() - t2;
      }
      if (dbg && lane == 0) { dbg[1] = t_acc; dbg[2] = t_ld; dbg[3] = t_issue; dbg[4] = TICK() - t_begin; }
    }
  } else {
    // ===== epilogue: warps 0-7; warp w owns TMEM lanes 32*(w%4).. and drains the tiles with i % 2 == w/4 =====
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;           // query row inside the tile
    Entry *cs = reinterpret_cast<Entry *>(smem + SM_CS + half * LIST_BYTES);
    volatile int *n_set0 = reinterpret_cast<volatile int *>(smem + SM_NA);
    const int vsplit = blockIdx.y * HALVES + half;
    ListState st;
    st.base = ptx::smem_u32(cs) + row * 8;
    st.off = st.base;
    // GROUPS mode: this thread's score log and its next free record (see flip_log)
    float *const logp = a.glog + ((int64_t)vsplit * a.hw_pad + qtile * TQ + row) * (VOSMEM_GROUP_LOG * GROUP_KEYS);
    uint32_t lp = 0;
    int log_bound = 0;   // warp-uniform upper bound of the records any lane holds in the active half of its log
    // Candidates are kept when score >= threshold: the shared threshold is the score of a real key, and with exact
    // ties at it (duplicated memory frames, uniform regions) a strict compare would drop keys the reference's
    // torch.topk returns.  The initial threshold is the lowest FINITE float, not -inf, so that masked columns
    // (-inf) never qualify.
    // Rows past the last query (the tail of the last query tile) never keep anything: their operand is all zero,
    // i.e. every key ties at score 0.
    st.tau = qtile * TQ + row < a.hw ? -FLT_MAX : INFINITY;
    st.pub = -INFINITY;
    float best[R];   // lower bounds of the R best scores of this warp set's keys, descending
#pragma unroll
    for (int u = 0; u < R; ++u) best[u] = -INFINITY;
    PubEntry *pub_mine = a.pub + (int64_t)vsplit * a.hw_pad + qtile * TQ + quarter * 32;
    PubEntry *pub2_mine = a.pub2 + (int64_t)vsplit * a.hw_pad + qtile * TQ + quarter * 32;
    float pub2 = -INFINITY;   // what this thread has published for the cross-rank bound
    long long t_wait = 0, t_relieve = 0, t_first = 0, t_ld = 0, t_max = 0, t_app = 0, t_x0 = 0, t_x1 = 0;
    int n_active = 0, n_relieve = 0;
    int len_bound = 0;   // warp-uniform upper bound of the longest list of this warp
    // Tiles are handed out dynamically to the two warps of a lane quarter (shared-memory counter), so a warp that is
    // busy cutting its lists does not hold up the accumulator ring: its partner drains the tiles meanwhile.  Lists,
    // thresholds and the tracker are per warp set and do not care which tiles they see.
    int *tile_ctr = reinterpret_cast<int *>(smem + SM_CTR) + quarter;
    int n_done = 0;      // tiles this warp has drained
    int grabbed = 0;     // lane 0: the next tile of this warp (fetched one iteration ahead to hide the atomic)
    if (lane == 0) grabbed = atomicAdd(tile_ctr, 1);
    // tiles that may hold columns outside the candidate range (first / last tile of a segment), as stream positions
    const int edge0 = -(int)g_lo, edge1 = (int)(a.seg[0].tiles - 1 - g_lo), edge2 = edge1 + 1,
              edge3 = (int)(a.tiles_total - 1 - g_lo);
    const int64_t g_tiles0 = a.seg[0].tiles, g_tile0 = a.seg[0].tile0, g_tile1 = a.seg[1].tile0;   // (GROUPS: locators)
    const long long t_begin = TICK();
    for (int i = __shfl_sync(FULL, grabbed, 0); i < n_tiles; i = __shfl_sync(FULL, grabbed, 0), ++n_done) {
      const int buf = i % ACC_BUFS;
      const long long tw0 = TICK();
      ptx::mbar_wait(bar_tfull + buf, (i / ACC_BUFS) & 1);
      t_wait += TICK() - tw0;
      ptx::tc_fence_after();
      uint32_t v[TK];
      {
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TK;
        ptx::tmem_ld_32x32(taddr, v0);
        ptx::tmem_ld_32x32(taddr + 32, v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = v0[j]; v[32 + j] = v1[j]; }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar_tempty + buf);  // accumulator buffer free again
        grabbed = atomicAdd(tile_ctr, 1);
      }
      const long long tp1 = TICK();
      t_ld += tp1 - tw0;
      st.tau = fmaxf(st.tau, tau_sh[row]);

      // first / last tile of a candidate range: columns outside it never qualify
      if (i == edge0 || i == edge1 || i == edge2 || i == edge3) {
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
      // maxima of the 8-column groups: which groups hold a survivor for any lane
      float gm[TK / 8];
#pragma unroll
      for (int g8 = 0; g8 < TK / 8; ++g8) {
        float m = __uint_as_float(v[g8 * 8]);
#pragma unroll
        for (int jj = 1; jj < 8; ++jj) m = fmaxf(m, __uint_as_float(v[g8 * 8 + jj]));
        gm[g8] = m;
      }
      auto track = [&](float m) {   // sorted insertion, 2 R - 1 min / max
#pragma unroll
        for (int u = 0; u < R; ++u) {
          const float lo = fminf(best[u], m);
          best[u] = fmaxf(best[u], m);
          m = lo;
        }
      };
      if (R >= 4 && n_done == 0) {   // first tile of this warp, few virtual splits: every score, so that the R-th best is
                                     // exact and can be published at once (8 group maxima cannot fill R >= 9 slots)
#pragma unroll
        for (int j = 0; j < TK; ++j) track(__uint_as_float(v[j]));
      } else if (n_done < TRACK_TILES) {   // then the group maxima, tile maxima after the first tiles
#pragma unroll
        for (int g8 = 0; g8 < TK / 8; ++g8) track(gm[g8]);
      } else {
        float tm = gm[0];
#pragma unroll
        for (int g8 = 1; g8 < TK / 8; ++g8) tm = fmaxf(tm, gm[g8]);
        track(tm);
      }
      if (best[R - 1] > st.pub) {
        st.pub = best[R - 1];
        pub_store(pub_mine + lane, st.pub, epoch);
      }
      if (SHARED || PAIRED) {   // SHARED: the R2-th best of the same tracker, for the bound shared across ranks;
                                // PAIRED: the best, combined with the 2nd best by the refreshers (refresh_grouped)
        float b2 = best[0];
        if (SHARED) {
#pragma unroll
          for (int u = 1; u < R; ++u) b2 = (u < batch.r2) ? best[u] : b2;
        }
        if (b2 > pub2) {
          pub2 = b2;
          pub_store(pub2_mine + lane, pub2, epoch);
        }
      }
      if (n_done == 0) {
        // First tile of this warp: nothing is known yet and every score would be kept (and the lists cut by sorting
        // several times before the thresholds bite).  The tile sits in registers, so give the other virtual splits a
        // bounded moment to publish the exact R-th best of their first tiles (the MMA warp keeps filling the other
        // accumulator buffers meanwhile) and filter with the shared threshold.  On a timeout (e.g. a grid of
        // several waves) the lists overflow and get sorted instead.
        const long long t0 = clock64();
        float shared = tau_sh[row];
        while (__any_sync(FULL, shared == -INFINITY) && clock64() - t0 < FIRST_WAIT_CYCLES) {
          __nanosleep(100);
          shared = tau_sh[row];
        }
        st.tau = fmaxf(st.tau, shared);
        t_first = TICK() - t0;
      }
      unsigned mine = 0;
#pragma unroll
      for (int g8 = 0; g8 < TK / 8; ++g8) mine |= (gm[g8] >= st.tau ? 1u : 0u) << g8;
      const unsigned active = __reduce_or_sync(FULL, mine);
      const long long tp2 = TICK();
      t_max += tp2 - tp1;

      if constexpr (GROUPS) {
        if (active) {   // warp-uniform
          // at most 8 appends per list and tile: make room first if some list (or the active half of some log) could
          // overflow.  The warp-uniform bounds make the real (voted) checks rare.
          if (len_bound + TK / 8 > CSLOTS) {
            len_bound = __reduce_max_sync(FULL, (int)((st.off - st.base) / SS));
            if (len_bound + TK / 8 > CSLOTS) {
              const long long tr0 = TICK();
              const float pub_before = st.pub;
              st = relieve_lists(st, cs, tau_sh + row, quarter, lane, R);
              if (st.pub > pub_before) pub_store(pub_mine + lane, st.pub, epoch);
              t_relieve += TICK() - tr0;
              ++n_relieve;
              len_bound = __reduce_max_sync(FULL, (int)((st.off - st.base) / SS));
            }
          }
          if (log_bound + TK / 8 > LOG_HALF) {
            log_bound = __reduce_max_sync(FULL, (int)(lp >= (uint32_t)LOG_HALF ? lp - LOG_HALF : lp));
            if (log_bound + TK / 8 > LOG_HALF) {
              lp = flip_log(st, logp, lp);
              log_bound = len_bound;   // a flipped half holds exactly the list's records
            }
          }
          const uint32_t gi0 = (uint32_t)i * (TK / 8);   // group index in this CTA's stream
          const float thr = (a.exp & 1) ? INFINITY : st.tau;   // experiment: nothing is appended
          const long long tx1 = TICK();
#pragma unroll
          for (int g8 = 0; g8 < TK / 8; ++g8) {
            if (active & (1u << g8)) {   // warp-uniform; in steady state most groups are skipped
              ++n_active;
              VOSMEM_APPEND_GROUP(st.off, lp, logp, thr, gm[g8], gi0 + g8, v, g8 * 8);
            }
          }
          t_x1 += TICK() - tx1;
          const int added = __popc(active);
          len_bound += added;
          log_bound += added;
        }
      } else {
        const uint32_t li0 = (uint32_t)i * TK;
#pragma unroll
        for (int g8 = 0; g8 < TK / 8; ++g8) {
          if (active & (1u << g8)) {   // warp-uniform; in steady state most groups are skipped
            ++n_active;
            // slot addresses first (one select + add per column, each in a fresh register), then the predicated
            // stores: no store waits for the previous store to release its address register
            uint32_t slot[9];
            slot[0] = st.off;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              slot[jj + 1] = slot[jj] + (__uint_as_float(v[g8 * 8 + jj]) >= st.tau ? SS : 0u);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int j = g8 * 8 + jj;
              if (__uint_as_float(v[j]) >= st.tau)
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(slot[jj]), "r"(v[j]), "r"(li0 + j) : "memory");
            }
            st.off = slot[8];
            // Lists that could overflow during the next 8 columns.  The warp-uniform bound makes the real (voted) check
            // rare: lists are short once the shared thresholds work.
            len_bound += 8;
            if (len_bound > PRUNE_ABOVE) {
              len_bound = __reduce_max_sync(FULL, (int)((st.off - st.base) / SS));
              if (len_bound > PRUNE_ABOVE) {
                const long long tr0 = TICK();
                const float pub_before = st.pub;
                st = relieve_lists(st, cs, tau_sh + row, quarter, lane, R);
                if (st.pub > pub_before) pub_store(pub_mine + lane, st.pub, epoch);
                t_relieve += TICK() - tr0;
                ++n_relieve;
                len_bound = __reduce_max_sync(FULL, (int)((st.off - st.base) / SS));
              }
            }
          }
        }
      }
      t_app += TICK() - tp2;
    }
    __syncwarp();
    if (lane == 0) atomicAdd(const_cast<uint32_t *>(epi_done), 1u);   // lets the refresher warp retire
    const long long t_loop = TICK() - t_begin;

    if constexpr (GROUPS) {
      // ---- hand-off: every thread copies its own list -- entries with their final locator (segment bit | group index
      //      inside the bank) and the 8 scores of each from the log -- into row (virtual split, query) of the exchange
      //      arrays, four entries per step (one 32-byte store of entries, four of scores) ----
      st.tau = fmaxf(st.tau, tau_sh[row]);
      st = compact_list(st);
      const int my_n = (int)((st.off - st.base) / SS);
      const int q = qtile * TQ + row;
      if (q < a.hw) a.cand_count[(int64_t)vsplit * a.hw_pad + q] = my_n;
      const int64_t list = ((int64_t)vsplit * a.hw_pad + q) * GSLOTS;
      float *ent_out = reinterpret_cast<float *>(a.cand + list);
      float *rec_out = a.gscore + list * GROUP_KEYS;
      const int nmax = __reduce_max_sync(FULL, my_n);
      for (int e0 = 0; e0 < nmax; e0 += 4) {
        Rec8 r[4], hd;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool in = e0 + u < my_n;
          const Entry en = lds_entry(st.base + (in ? e0 + u : 0) * SS);
          const uint32_t gix = en.index & GIDX_MASK;
          const int64_t g = g_lo + (gix >> 3);
          const uint32_t loc = g >= g_tiles0 ? (GROUP_SEG_BIT | (uint32_t)((g_tile1 + (g - g_tiles0)) * (TK / 8) + (gix & 7)))
                                             : (uint32_t)((g_tile0 + g) * (TK / 8) + (gix & 7));
          hd.w[2 * u] = __float_as_uint(en.score);
          hd.w[2 * u + 1] = loc;
          r[u] = ldg_rec(logp + (size_t)(in ? en.index >> GIDX_BITS : 0u) * GROUP_KEYS);
        }
        if (e0 < my_n) {
          stg_rec(ent_out + e0 * 2, hd);   // (entries past my_n are never read: cand_count)
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (e0 + u < my_n) stg_rec(rec_out + (size_t)(e0 + u) * GROUP_KEYS, r[u]);
        }
      }
    } else
    // ---- hand the surviving candidates to the merge: both sets of a query share one exchange row, set 0 first.
    //      Lanes run over the entries of a row (coalesced 8-byte stores), four rows per step for ILP. ----
    {
      // last look at the shared threshold: what the other splits found meanwhile shortens the hand-off (and the merge)
      st.tau = fmaxf(st.tau, tau_sh[row]);
      st = compact_list(st);
      const int my_n = (int)((st.off - st.base) / SS);
      // only the second warp of a lane quarter depends on the first (its entries follow the first one's in the
      // exchange row): the first arrives on the named barrier 1 + quarter without waiting, the second waits for it
      if (half == 0) {
        n_set0[row] = my_n;
        asm volatile("bar.arrive %0, 64;" ::"r"(1 + quarter) : "memory");
      } else {
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      }
      const int my_first = half == 0 ? 0 : n_set0[row];
      const int q = qtile * TQ + row;
      if (half == 1 && q < a.hw) a.cand_count[(int64_t)blockIdx.y * a.hw_pad + q] = my_first + my_n;
      // candidate index of local index li:  64 * (g_lo + li / 64) + li % 64 + (offset of the segment)
      const int tiles0 = (int)a.seg[0].tiles, glo = (int)g_lo;
      const int off0 = (int)(a.seg[0].tile0 * TK - a.seg[0].begin);
      const int off1 = (int)(a.len0 + (a.seg[1].tile0 - a.seg[0].tiles) * TK - a.seg[1].begin);
      const int q_base = qtile * TQ + quarter * 32;
      CandEntry *xrow0 = a.cand + ((int64_t)blockIdx.y * a.hw_pad + q_base) * CAND_SLOTS;
      // lane = slot, fully unrolled over the 32 rows of the warp (independent load -> store chains); a second pass
      // covers slots 32.. only if some list of the warp is that long
      const int nmax = __reduce_max_sync(FULL, my_n);
      for (int e0 = 0; e0 < nmax; e0 += 32) {
        const int e = e0 + lane;
#pragma unroll 8
        for (int src = 0; src < 32; ++src) {
          const int n = __shfl_sync(FULL, my_n, src);
          const int first = __shfl_sync(FULL, my_first, src);
          if (e < n && q_base + src < a.hw) {
            const Entry en = cs[e * CS_E + quarter * 32 + src];
            const int li = (int)en.index;
            const int g = glo + (li >> 6);
            xrow0[(int64_t)src * CAND_SLOTS + first + e] =
                CandEntry{en.score, g * TK + (li & 63) + (g >= tiles0 ? off1 : off0)};
          }
        }
      }
    }
    if (dbg && lane == 0 && half == 0) {
      dbg[5 + quarter * 2] = t_wait;
      dbg[6 + quarter * 2] = t_relieve;
      if (quarter == 0) { dbg[13] = t_loop; dbg[14] = TICK() - t_begin; }
      if (quarter == 1) dbg[15] = t_first;
      if (quarter == 0) { dbg[16] = t_ld; dbg[17] = t_max; dbg[18] = t_app; dbg[19] = n_active; dbg[20] = n_relieve; }
      if (quarter == 0) { dbg[25] = t_x0; dbg[26] = t_x1; dbg[27] = n_done; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) ptx::tmem_dealloc(*tmem_slot, TMEM_COLS);
  // departure: the last CTA of this problem's grid advances the epoch for the next launch on this workspace (every
  // CTA has read it by then, even in a grid of several waves) and resets the counter
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&a.ctl->departed, 1u) == gridDim.x * gridDim.y - 1) {
      a.ctl->departed = 0u;
      a.ctl->last = epoch;
      a.ctl->epoch = epoch + 1u == 0u ? 1u : epoch + 1u;   // 0 is what never-written entries carry
    }
  }
  if (dbg && threadIdx.x == 0) {
    unsigned long long g_exit;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_exit));
    dbg[22] = TICK() - t_entry;   // whole CTA
    dbg[23] = (long long)g_entry;    // ns, global timer: CTA start / end (spread of the starts, tail of the grid)
    dbg[24] = (long long)g_exit;
  }
}

// Self-test: raw accumulator tile of query tile 0 against key tile 0 through the same staging + MMA sequence.
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
  if (warp == W_MMA) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == W_PRODUCER && lane == 0) {
    // query image in stages 0-1, key tile in stage 2
    ptx::mbar_arrive_expect_tx(bar_ld, QUERY_TILE_BYTES + KEY_TILE_BYTES);
    ptx::bulk_g2s(smem + SM_Q, query_image, QUERY_TILE_BYTES / 2, bar_ld);
    ptx::bulk_g2s(smem + SM_Q + QUERY_TILE_BYTES / 2, query_image + QUERY_TILE_BYTES / 2, QUERY_TILE_BYTES / 2, bar_ld);
    ptx::bulk_g2s(smem + SM_K + 2 * KEY_TILE_BYTES, key_image, KEY_TILE_BYTES, bar_ld);
  } else if (warp == W_MMA) {
    ptx::mbar_wait(bar_ld, 0);
    ptx::tc_fence_after();
    stage_query_in_tmem(ptx::smem_u32(smem + SM_Q), tmem_base + TMEM_A);
    issue_tile(tmem_base + TMEM_A, ptx::smem_u32(smem + SM_K + 2 * KEY_TILE_BYTES), tmem_base);
    ptx::umma_commit_elect(bar_mma);
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
  if (warp == W_MMA) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

thread_local long long *g_tc_debug = nullptr;  // set through vosmem_debug_set_timing_buffer (per host thread)

int launch_select_tc(const vosmem_select_desc *descs, const Workspace *wss, int n, int splits, bool groups, cudaStream_t st,
                     const PeerThresholds *peers) {
  VOSMEM_CHECK_ARG(n >= 1 && n <= MAX_BATCH, "select(tcgen05): batch of %d problems outside [1, %d]", n, MAX_BATCH);
  TcBatch batch{};
  batch.world = 1;
  if (peers != nullptr && peers->world > 1) {
    VOSMEM_CHECK_ARG(n == 1 && peers->world <= VOSMEM_MAX_RANKS && peers->rank >= 0 && peers->rank < peers->world,
                     "select(tcgen05): thresholds across ranks: rank %d of %d, %d problems", peers->rank, peers->world, n);
    batch.world = peers->world;
    batch.rank = peers->rank;
    for (int r = 0; r < peers->world; ++r) {
      VOSMEM_CHECK_ARG(peers->rank_pub[r] != nullptr, "select(tcgen05): no summary array for rank %d", r);
      batch.rank_pub[r] = peers->rank_pub[r];
    }
  }
  for (int b = 0; b < n; ++b) {
    const vosmem_select_desc &d = descs[b];
    const Workspace &ws = wss[b];
    VOSMEM_CHECK_ARG(d.ck == CK_TC, "select(tcgen05): CK must be 64 (got %d)", d.ck);
    VOSMEM_CHECK_ARG(d.hw == descs[0].hw, "select(tcgen05): problems of one batch must share HW (%d vs %d)", d.hw, descs[0].hw);
    TcArgs &a = batch.p[b];
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
    a.hw = d.hw;
    a.hw_pad = (int)round_up64(d.hw, TQ);
    a.splits = splits;
    a.qk = d.query_key;
    a.qe = d.query_selection;
    a.pub = ws.pub;
    a.pub2 = ws.pub2;
    a.ctl = ws.ctl;
    a.cand = ws.cand;
    a.cand_count = ws.cand_count;
    a.gscore = ws.gscore;
    a.glog = ws.glog;
    a.dbg = g_tc_debug;
    { const char *e = getenv("VOSMEM_TC_EXP"); a.exp = e ? atoi(e) : 0; }
    if (groups)
      for (int s = 0; s < d.n_segments; ++s)
        VOSMEM_CHECK_ARG(d.seg[s].end < ((int64_t)1 << 31), "select(tcgen05): segment %d ends at key %lld (group locators hold 31 bits)",
                         s, (long long)d.seg[s].end);
  }
  const vosmem_select_desc &d = descs[0];
  dim3 grid((unsigned)ceil_div64(d.hw, TQ), splits, n);
  // R * (virtual splits) >= 33 keys must stand behind a shared threshold; R2 likewise over the lists of all ranks
  const int r = (33 + HALVES * splits - 1) / (HALVES * splits);
  batch.r2 = (33 + HALVES * splits * batch.world - 1) / (HALVES * splits * batch.world);
  // the shared-memory opt-in is a per-device function attribute: set it on every launch (a host-side table lookup)
  // (thresholds across ranks are only instantiated for single-key candidates: api.cu never asks for both)
#define VOSMEM_LAUNCH_TC_AS(RR, SH, GR)                                                                          \
  do {                                                                                                           \
    VOSMEM_CUDA(cudaFuncSetAttribute(select_tc_kernel<RR, SH, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL)); \
    select_tc_kernel<RR, SH, GR><<<grid, TC_THREADS, SM_TOTAL, st>>>(batch);                                     \
  } while (0)
#define VOSMEM_LAUNCH_TC(RR)                                                                                     \
  do {                                                                                                           \
    if (batch.world > 1) VOSMEM_LAUNCH_TC_AS(RR, true, false);                                                   \
    else if (groups) VOSMEM_LAUNCH_TC_AS(RR, false, true);                                                       \
    else VOSMEM_LAUNCH_TC_AS(RR, false, false);                                                                  \
  } while (0)
  VOSMEM_CHECK_ARG(!groups || batch.world == 1, "select(tcgen05): thresholds across ranks and group candidates together");
  if (r <= 1) VOSMEM_LAUNCH_TC(1);
  else if (r == 2) VOSMEM_LAUNCH_TC(2);
  else if (r == 3) VOSMEM_LAUNCH_TC(3);
  else if (r == 4) VOSMEM_LAUNCH_TC(4);
  else if (r <= 6) VOSMEM_LAUNCH_TC(6);
  else if (r <= 9) VOSMEM_LAUNCH_TC(9);
  else VOSMEM_LAUNCH_TC(17);
#undef VOSMEM_LAUNCH_TC_AS
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
