// extern "C" entry points that tie the kernels into the calls declared in include/vosmem.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "common.cuh"
#include "merge.cuh"

namespace vosmem {

static thread_local char g_error[512] = "";
// debug hook state is per host thread (like the error message): begin / after pack / select / merge|readout
static thread_local cudaEvent_t g_stage_events[4] = {nullptr, nullptr, nullptr, nullptr};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// N-splits so that (query tiles x splits) fills the 148 SMs once (tcgen05: one CTA per SM) or gives the
// SIMT kernel ~32 warps per SM; never fewer than ~4 key tiles per split.
int choose_splits(int path, int hw, int64_t n_total, int batch) {
  const int cap = splits_cap(hw);
  int64_t s;
  if (path == VOSMEM_PATH_TCGEN05) {
    int64_t n_qtiles = ceil_div64(hw, TQ);
    s = 148 / (n_qtiles * batch);
    int64_t by_tiles = ceil_div64(n_total, TK) / 4;
    if (s > by_tiles) s = by_tiles;
  } else {
    s = ceil_div64(4736, hw);
    int64_t by_keys = n_total / 256;
    if (s > by_keys) s = by_keys;
  }
  if (s < 1) s = 1;
  if (s > cap) s = cap;
  return (int)s;
}

// Candidate granularity of the tcgen05 selection (select_tc.cu): single keys (default) or 8-key groups
// (VOSMEM_TC_CANDIDATES=groups).  Both are parity-tested; the measurements that keep single keys the default are in
// DESIGN.md section 3.1.  Read once per process.
static bool tc_group_candidates() {
  static const bool groups = [] {
    const char *e = getenv("VOSMEM_TC_CANDIDATES");
    return e != nullptr && strcmp(e, "groups") == 0;
  }();
  return groups;
}

static int resolve_path(const vosmem_select_desc &d) {
  if (d.path == VOSMEM_PATH_SIMT || d.path == VOSMEM_PATH_TCGEN05) return d.path;
  if (d.ck != CK_TC) return VOSMEM_PATH_SIMT;
  for (int s = 0; s < d.n_segments; ++s)
    if (d.seg[s].end > d.seg[s].begin && d.seg[s].key_image == nullptr) return VOSMEM_PATH_SIMT;
  return VOSMEM_PATH_TCGEN05;
}

static int validate_select(const vosmem_select_desc *d) {
  VOSMEM_CHECK_ARG(d != nullptr, "select: null descriptor");
  VOSMEM_CHECK_ARG(d->ck >= 1 && d->ck <= 256, "select: CK=%d outside [1, 256]", d->ck);
  VOSMEM_CHECK_ARG(d->hw >= 1, "select: HW=%d", d->hw);
  VOSMEM_CHECK_ARG(d->top_k >= 1 && d->top_k <= VOSMEM_MAX_TOPK, "select: top_k=%d outside [1, %d]", d->top_k,
                   VOSMEM_MAX_TOPK);
  VOSMEM_CHECK_ARG(d->query_key != nullptr, "select: null query_key");
  VOSMEM_CHECK_ARG(d->n_segments >= 1 && d->n_segments <= VOSMEM_MAX_SEGMENTS, "select: n_segments=%d", d->n_segments);
  int64_t total = 0;
  for (int s = 0; s < d->n_segments; ++s) {
    VOSMEM_CHECK_ARG(d->seg[s].begin >= 0 && d->seg[s].end >= d->seg[s].begin, "select: segment %d range [%lld, %lld)",
                     s, (long long)d->seg[s].begin, (long long)d->seg[s].end);
    total += d->seg[s].end - d->seg[s].begin;
  }
  VOSMEM_CHECK_ARG(total >= 1, "select: no memory elements");
  VOSMEM_CHECK_ARG(total < (int64_t)0x7fffffff, "select: %lld memory elements exceed the 31-bit candidate index",
                   (long long)total);
  VOSMEM_CHECK_ARG(d->workspace != nullptr, "select: null workspace");
  if (d->workspace_bytes < vosmem_workspace_bytes(d->ck, d->hw, total)) {
    set_error("select: workspace of %lld bytes, need %lld", (long long)d->workspace_bytes,
              (long long)vosmem_workspace_bytes(d->ck, d->hw, total));
    return VOSMEM_ENOSPC;
  }
  return VOSMEM_OK;
}

}  // namespace vosmem

using namespace vosmem;

extern "C" int vosmem_abi_version(void) { return VOSMEM_ABI_VERSION; }
extern "C" const char *vosmem_last_error(void) { return g_error; }
extern "C" const char *vosmem_status_string(int status) {
  switch (status) {
    case VOSMEM_OK: return "ok";
    case VOSMEM_EINVAL: return "invalid argument";
    case VOSMEM_ENOTSUP: return "not supported";
    case VOSMEM_ENOSPC: return "workspace too small";
    case VOSMEM_ECUDA: return "CUDA error";
    default: return "unknown status";
  }
}

extern "C" int64_t vosmem_key_image_bytes(int ck, int64_t capacity) {
  if (ck != CK_TC || capacity < 0) return 0;
  return ceil_div64(capacity, TK) * (int64_t)KEY_TILE_BYTES;
}

extern "C" int64_t vosmem_workspace_bytes(int ck, int hw, int64_t n_keys) {
  (void)n_keys;
  if (ck < 1 || hw < 1) return 0;
  return carve_workspace(nullptr, ck, hw, tc_group_candidates()).bytes;
}

// One-time preparation of a workspace: everything zero (published-threshold entries then carry epoch 0, which no
// launch ever uses), launch epoch 0x01010101, departure counter and error flags 0.
extern "C" int vosmem_workspace_init(void *workspace, int64_t workspace_bytes, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(workspace != nullptr && workspace_bytes >= 256, "vosmem_workspace_init: workspace of %lld bytes",
                   (long long)workspace_bytes);
  VOSMEM_CUDA(cudaMemsetAsync(workspace, 0, (size_t)workspace_bytes, (cudaStream_t)stream));
  VOSMEM_CUDA(cudaMemsetAsync(workspace, 1, sizeof(uint32_t), (cudaStream_t)stream));   // WsControl::epoch
  return VOSMEM_OK;
}

// Sticky device-side error flags of a workspace (synchronises the stream; not for the per-frame path).
extern "C" int vosmem_workspace_status(const void *workspace, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(workspace != nullptr, "vosmem_workspace_status: null workspace");
  WsControl host{};
  VOSMEM_CUDA(cudaMemcpyAsync(&host, workspace, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  VOSMEM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (host.error & WS_ERR_TMEM_BASE) {
    set_error("select(tcgen05): tensor memory allocation did not start at column 0 (another TMEM user shares the SM)");
    return VOSMEM_ENOTSUP;
  }
  return VOSMEM_OK;
}

// run the selection kernel(s) of `n` problems: leaves the candidate lists in each problem's workspace and describes
// them in lists[0..n)
static int run_selection(const vosmem_select_desc *d, int n, cudaStream_t st, SplitLists *lists,
                         const PeerThresholds *peers = nullptr) {
  Workspace ws[MAX_BATCH];
  int path = 0, splits = MAX_SPLITS;
  for (int b = 0; b < n; ++b) {
    int rc = validate_select(d + b);
    if (rc != VOSMEM_OK) return rc;
    ws[b] = carve_workspace(d[b].workspace, d[b].ck, d[b].hw, tc_group_candidates());
    int64_t total = 0;
    for (int s = 0; s < d[b].n_segments; ++s) total += d[b].seg[s].end - d[b].seg[s].begin;
    const int p = resolve_path(d[b]);
    VOSMEM_CHECK_ARG(b == 0 || p == path, "match_batch: problems resolve to different kernel paths");
    VOSMEM_CHECK_ARG(n == 1 || p == VOSMEM_PATH_TCGEN05, "match_batch: batches need the tcgen05 path (CK == 64, key images)");
    VOSMEM_CHECK_ARG(d[b].hw == d[0].hw && d[b].top_k == d[0].top_k && d[b].ck == d[0].ck,
                     "match_batch: problems must share CK, HW and top_k");
    for (int o = 0; o < b; ++o)
      VOSMEM_CHECK_ARG(d[o].workspace != d[b].workspace, "match_batch: problems %d and %d share a workspace", o, b);
    path = p;
    const int sp = choose_splits(p, d[b].hw, total, n);
    splits = sp < splits ? sp : splits;
  }
  const bool groups = path == VOSMEM_PATH_TCGEN05 && tc_group_candidates() && !(peers != nullptr && peers->world > 1);
  const int n_lists = groups ? splits * LISTS_PER_SPLIT : splits;                 // candidate lists left per query
  const int n_pub = path == VOSMEM_PATH_TCGEN05 ? splits * LISTS_PER_SPLIT : 0;   // published threshold rows (SIMT: none)
  for (int b = 0; b < n; ++b) lists[b] = split_lists_of(ws[b], n_lists, n_pub, d[b].hw, groups, d + b);
  if (g_stage_events[0]) cudaEventRecord(g_stage_events[0], st);
  int rc;
  if (path != VOSMEM_PATH_TCGEN05) {   // the tcgen05 kernel packs its query tile itself
    rc = launch_pack_query(d->query_key, d->query_selection, d->ck, d->hw, ws[0], st);
    if (rc != VOSMEM_OK) return rc;
  }
  if (g_stage_events[1]) cudaEventRecord(g_stage_events[1], st);
  if (peers != nullptr && peers->world > 1)
    VOSMEM_CHECK_ARG(path == VOSMEM_PATH_TCGEN05 && n == 1, "select: thresholds across ranks need the tcgen05 path, one problem");
  rc = path == VOSMEM_PATH_TCGEN05 ? launch_select_tc(d, ws, n, splits, groups, st, peers)
                                   : launch_select_simt(*d, ws[0], splits, st);
  if (rc != VOSMEM_OK) return rc;
  if (g_stage_events[2]) cudaEventRecord(g_stage_events[2], st);
  return VOSMEM_OK;
}

namespace vosmem {
// the selection stage alone, for the sharded path (exchange.cu): candidate lists stay in the workspace
int run_selection_for_push(const vosmem_select_desc *d, cudaStream_t st, SplitLists &lists, const PeerThresholds *peers) {
  return run_selection(d, 1, st, &lists, peers);
}
}  // namespace vosmem

extern "C" int vosmem_select_topk(const vosmem_select_desc *d, float *out_score, int64_t *out_index,
                                  vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(out_score && out_index, "select: null output");
  cudaStream_t st = (cudaStream_t)stream;
  SplitLists lists;
  int rc = run_selection(d, 1, st, &lists);
  if (rc != VOSMEM_OK) return rc;
  rc = launch_merge_splits(lists, d->hw, d->top_k, d->index_base, out_score, out_index, st);
  if (g_stage_events[3]) cudaEventRecord(g_stage_events[3], st);
  return rc;
}

namespace vosmem {
int launch_fused_readout(const vosmem_readout_desc *d, const SplitLists *lists, int n, cudaStream_t st);
}

// One object group of match_memory: pack -> select -> (merge + softmax + usage + readout in one kernel).
// scratch_score / scratch_index are not needed by this fused path and may be NULL.
extern "C" int vosmem_match(const vosmem_select_desc *select, const vosmem_readout_desc *readout, float *scratch_score,
                            int64_t *scratch_index, vosmem_stream_t stream) {
  (void)scratch_score;
  (void)scratch_index;
  VOSMEM_CHECK_ARG(select && readout, "vosmem_match: null descriptor");
  VOSMEM_CHECK_ARG(select->hw == readout->hw && select->top_k == readout->top_k,
                   "vosmem_match: select (hw=%d, k=%d) and readout (hw=%d, k=%d) disagree", select->hw, select->top_k,
                   readout->hw, readout->top_k);
  VOSMEM_CHECK_ARG(select->index_base == 0, "vosmem_match: index_base must be 0 (use the staged calls for sharded banks)");
  cudaStream_t st = (cudaStream_t)stream;
  SplitLists lists;
  int rc = run_selection(select, 1, st, &lists);
  if (rc != VOSMEM_OK) return rc;
  rc = launch_fused_readout(readout, &lists, 1, st);
  if (g_stage_events[3]) cudaEventRecord(g_stage_events[3], st);
  return rc;
}

// n independent problems (sequences) in one selection launch + one readout launch (blockIdx.z = problem)
extern "C" int vosmem_match_batch(const vosmem_select_desc *select, const vosmem_readout_desc *readout, int n,
                                  vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(select && readout, "vosmem_match_batch: null descriptor array");
  VOSMEM_CHECK_ARG(n >= 1 && n <= MAX_BATCH, "vosmem_match_batch: n=%d outside [1, %d]", n, MAX_BATCH);
  for (int b = 0; b < n; ++b) {
    VOSMEM_CHECK_ARG(select[b].hw == readout[b].hw && select[b].top_k == readout[b].top_k,
                     "vosmem_match_batch: problem %d: select (hw=%d, k=%d) and readout (hw=%d, k=%d) disagree", b,
                     select[b].hw, select[b].top_k, readout[b].hw, readout[b].top_k);
    VOSMEM_CHECK_ARG(select[b].index_base == 0, "vosmem_match_batch: index_base must be 0");
  }
  cudaStream_t st = (cudaStream_t)stream;
  SplitLists lists[MAX_BATCH];
  int rc = run_selection(select, n, st, lists);
  if (rc != VOSMEM_OK) return rc;
  rc = launch_fused_readout(readout, lists, n, st);
  if (g_stage_events[3]) cudaEventRecord(g_stage_events[3], st);
  return rc;
}

extern "C" int vosmem_debug_set_stage_events(void *begin, void *after_pack, void *after_select, void *after_last) {
  g_stage_events[0] = (cudaEvent_t)begin;
  g_stage_events[1] = (cudaEvent_t)after_pack;
  g_stage_events[2] = (cudaEvent_t)after_select;
  g_stage_events[3] = (cudaEvent_t)after_last;
  return VOSMEM_OK;
}
