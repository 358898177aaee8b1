// N-sharded long-term bank: the exchange step between the ranks' local selections and the readout.
//
// The reference is single-GPU (tools/runner.py:32); SURVEY.md section 8e shards the long-term keys along N.  Global top-k
// is a subset of the union of the ranks' local top-k's, so each rank only has to ship HW x k (score, index) pairs --
// and because the readout is sharded over the QUERY axis, a query's list is only needed by the one rank that owns the
// query.  merge_push_kernel therefore writes every list straight into its owner's memory (peer-mapped pointers: plain
// st.global over NVLink, HW / world * k * 8 bytes per (source, owner) pair instead of an all-gather of HW * k * 8 per
// rank), then raises one flag per owner with a system-scope release.  The owner's readout kernel acquires the flags
// (readout.cu, exchange front end).  No collective, no global barrier.
#include "common.cuh"
#include "merge.cuh"

namespace vosmem {

namespace {

constexpr int EXCH_K = VOSMEM_EXCH_K;

struct PushArgs {
  SplitLists lists;
  int hw, top_k, per, world, rank;
  int64_t index_base, index_base1, seg0_len;
  uint2 *dst[VOSMEM_MAX_RANKS];
  uint32_t *flag[VOSMEM_MAX_RANKS];
  uint32_t seq;
  uint32_t *ticket;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All CTAs of a grid have finished their (peer) stores -> the last one to arrive raises the flags.
// Every thread fences its own stores at system scope, the CTA's arrival is a device-scope atomic, and the last CTA
// fences again before the release stores: the flag write is ordered after every data store of the grid.
__device__ __forceinline__ void signal_when_grid_done(uint32_t *ticket, uint32_t *const *flag, int n_flags, uint32_t seq) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    last = atomicAdd(ticket, 1u) == total - 1;
    if (last) *ticket = 0u;   // ready for the next launch (stream-ordered)
  }
  __syncthreads();
  if (last && (int)threadIdx.x < n_flags && flag[threadIdx.x] != nullptr) {
    __threadfence_system();
    st_release_sys(flag[threadIdx.x], seq);
  }
}

// One warp per query: merge the split lists of the local selection into the local top-k (merge.cuh), write the
// VOSMEM_EXCH_K-entry exchange list (global key indices; unused slots {-inf, -1}) into the owner's buffer.
// MB: split lists merged per batch (registers: 2 MB per lane; with few splits the small instantiation doubles the
// resident warps of this latency-bound kernel).
template <int MB>
__global__ void __launch_bounds__(256, MB <= 4 ? 4 : 2) merge_push_kernel(const __grid_constant__ PushArgs a) {
  __shared__ float buf_s[8][MERGE_BUF];
  __shared__ int buf_i[8][MERGE_BUF];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  if (q < a.hw) {
    const WarpTop32 top = merge_lists_of_query<MB>(a.lists, q, a.top_k, buf_s[warp], buf_i[warp], lane);
    const bool have = lane < a.top_k && top.i != 0x7fffffff;
    const int owner = q / a.per, local = q - owner * a.per;
    const int global = top.i < a.seg0_len ? top.i + (int)a.index_base : top.i - (int)a.seg0_len + (int)a.index_base1;
    const uint2 entry = make_uint2(__float_as_uint(have ? top.s : -INFINITY), have ? (uint32_t)global : 0xffffffffu);
    a.dst[owner][(int64_t)local * EXCH_K + lane] = entry;   // 256 contiguous bytes per query
  }
  signal_when_grid_done(a.ticket, a.flag, a.world, a.seq);
}

// A rank without keys launches no selection kernel; its workspace still has to step through the launch epochs in
// step with the other ranks' (thresholds shared across ranks are validated by epoch: select_tc.cu).
__global__ void bump_epoch_kernel(WsControl *ctl) {
  const uint32_t epoch = ctl->epoch;
  ctl->last = epoch;
  ctl->epoch = epoch + 1u == 0u ? 1u : epoch + 1u;
}

struct SliceArgs {
  const float *src;
  int64_t src_ld, dst_ld;
  int rows, cols, n_dst;
  float *dst[VOSMEM_MAX_RANKS];
  uint32_t *flag[VOSMEM_MAX_RANKS];
  uint32_t seq;
  uint32_t *ticket;
};

// grid = (row blocks, destinations): each CTA copies 8 rows of the slice to one destination, 16 bytes per thread and
// step when the alignment allows (cols * 4 bytes per row: 4 KB runs at LVOS / 8 ranks -> full NVLink packets).
__global__ void __launch_bounds__(256) push_slice_kernel(const __grid_constant__ SliceArgs a) {
  float *dst = a.dst[blockIdx.y];
  const bool vec = (a.cols % 4 == 0) && (a.src_ld % 4 == 0) && (a.dst_ld % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(a.src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < a.rows; r += gridDim.x * 8) {
    const float *s = a.src + (int64_t)r * a.src_ld;
    float *d = dst + (int64_t)r * a.dst_ld;
    if (vec) {
      for (int c = (threadIdx.x & 31) * 4; c < a.cols; c += 128)
        *reinterpret_cast<float4 *>(d + c) = *reinterpret_cast<const float4 *>(s + c);
    } else {
      for (int c = threadIdx.x & 31; c < a.cols; c += 32) d[c] = s[c];
    }
  }
  signal_when_grid_done(a.ticket, a.flag, a.n_dst, a.seq);
}

__global__ void wait_flags_kernel(const uint32_t *flags, uint32_t mask, uint32_t seq, uint32_t *status) {
  if (mask >> threadIdx.x & 1u) {
    uint32_t v;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
      if (v == seq) break;
      __nanosleep(100);
    } while (++spins < (1ll << 24));
    if (v != seq && status) *status = 1u;
  }
}

}  // namespace

int run_selection_for_push(const vosmem_select_desc *d, cudaStream_t st, SplitLists &lists, const PeerThresholds *peers);

}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_select_push(const vosmem_select_desc *select, const vosmem_push_desc *push, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(select && push, "vosmem_select_push: null descriptor");
  VOSMEM_CHECK_ARG(push->world >= 1 && push->world <= VOSMEM_MAX_RANKS && push->rank >= 0 && push->rank < push->world,
                   "vosmem_select_push: rank %d of %d", push->rank, push->world);
  VOSMEM_CHECK_ARG(push->per >= 1 && (int64_t)push->per * push->world >= select->hw,
                   "vosmem_select_push: %d queries per rank x %d ranks do not cover HW=%d", push->per, push->world, select->hw);
  VOSMEM_CHECK_ARG(push->ticket != nullptr, "vosmem_select_push: null ticket");
  VOSMEM_CHECK_ARG(select->top_k <= VOSMEM_EXCH_K, "vosmem_select_push: top_k=%d exceeds the exchange list length", select->top_k);
  for (int r = 0; r * (int64_t)push->per < select->hw; ++r)
    VOSMEM_CHECK_ARG(push->dst[r] != nullptr, "vosmem_select_push: no destination for owner rank %d", r);
  cudaStream_t st = (cudaStream_t)stream;
  PeerThresholds peers;
  peers.world = push->world;
  peers.rank = push->rank;
  bool shared = push->world > 1;
  for (int r = 0; r < push->world; ++r) {
    peers.rank_pub[r] = static_cast<PubEntry *>(push->rank_pub[r]);
    shared = shared && push->rank_pub[r] != nullptr;
  }
  PushArgs a{};
  int64_t n_keys = 0;
  for (int s = 0; s < select->n_segments && s < VOSMEM_MAX_SEGMENTS; ++s) n_keys += select->seg[s].end - select->seg[s].begin;
  if (n_keys > 0) {
    int rc = run_selection_for_push(select, st, a.lists, shared ? &peers : nullptr);
    if (rc != VOSMEM_OK) return rc;
  } else {   // no lists (splits == 0): every query's exchange list is pushed empty, the flags are raised as usual
    VOSMEM_CHECK_ARG(select->workspace != nullptr, "vosmem_select_push: null workspace");
    bump_epoch_kernel<<<1, 1, 0, st>>>(static_cast<WsControl *>(select->workspace));
  }
  const int n_lists = a.lists.splits;
  a.hw = select->hw;
  a.top_k = select->top_k;
  a.per = push->per;
  a.world = push->world;
  a.rank = push->rank;
  a.index_base = push->index_base;
  a.seg0_len = push->seg0_len < 0 ? ((int64_t)1 << 40) : push->seg0_len;
  a.index_base1 = push->index_base1;
  for (int r = 0; r < push->world; ++r) {
    a.dst[r] = static_cast<uint2 *>(push->dst[r]);
    a.flag[r] = push->flag[r];
  }
  a.seq = push->seq;
  a.ticket = push->ticket;
  if (n_lists <= 4) merge_push_kernel<4><<<(select->hw + 7) / 8, 256, 0, st>>>(a);
  else merge_push_kernel<MERGE_MAX_SPLITS><<<(select->hw + 7) / 8, 256, 0, st>>>(a);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_push_slice(const float *src, int64_t src_ld, int rows, int cols, float *const *dst, int64_t dst_ld,
                                 uint32_t *const *flag, int n_dst, uint32_t seq, uint32_t *ticket, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(src && dst && ticket, "vosmem_push_slice: null pointer");
  VOSMEM_CHECK_ARG(rows >= 1 && cols >= 1 && n_dst >= 1 && n_dst <= VOSMEM_MAX_RANKS, "vosmem_push_slice: rows=%d cols=%d n_dst=%d",
                   rows, cols, n_dst);
  SliceArgs a{};
  a.src = src;
  a.src_ld = src_ld;
  a.dst_ld = dst_ld;
  a.rows = rows;
  a.cols = cols;
  a.n_dst = n_dst;
  for (int i = 0; i < n_dst; ++i) {
    VOSMEM_CHECK_ARG(dst[i] != nullptr, "vosmem_push_slice: destination %d is null", i);
    a.dst[i] = dst[i];
    a.flag[i] = flag ? flag[i] : nullptr;
  }
  a.seq = seq;
  a.ticket = ticket;
  const int row_blocks = (rows + 7) / 8 < 64 ? (rows + 7) / 8 : 64;
  push_slice_kernel<<<dim3(row_blocks, n_dst), 256, 0, (cudaStream_t)stream>>>(a);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int vosmem_wait_flags(const uint32_t *flags, uint32_t mask, uint32_t seq, uint32_t *status, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(flags && mask != 0u && mask < (1u << VOSMEM_MAX_RANKS), "vosmem_wait_flags: mask=%#x", mask);
  wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, mask, seq, status);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}
