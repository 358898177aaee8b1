/*
 * vosmem.h -- C ABI of the B200-native space-time memory readout (libvosmem.so).
 *
 * The reference (jpitagc/VOS-E-SAM, XMem tracker) has no FFI layer: its hot path is Python
 * calling torch ops.  This header is the boundary a maintainer binds instead -- plain device
 * pointers, sizes and a cudaStream_t (passed as void*), no torch types.  Every entry point
 * names the reference lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; fp32 unless stated otherwise
 *   - functions only enqueue work on `stream`; they never synchronise, never allocate
 *   - return 0 on success, a negative VOSMEM_E* code otherwise; vosmem_last_error() gives
 *     a thread-local message.  No C++ exception crosses this boundary.
 *   - "key axis" = memory elements N (reference: last dim of k / v), "query axis" = HW
 *   - matrices named like the reference: key is CK x N with N contiguous (row pitch key_ld),
 *     queries are CK x HW with HW contiguous, values are rows x N with N contiguous.
 */
#ifndef VOSMEM_H_
#define VOSMEM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOSMEM_ABI_VERSION 3

typedef void *vosmem_stream_t; /* cudaStream_t */

enum vosmem_status {
  VOSMEM_OK = 0,
  VOSMEM_EINVAL = -22,  /* bad argument (shape, null pointer, unsupported top_k ...) */
  VOSMEM_ENOTSUP = -95, /* configuration not supported by the requested kernel path */
  VOSMEM_ENOSPC = -28,  /* workspace too small */
  VOSMEM_ECUDA = -5     /* a CUDA runtime call failed (message has the CUDA error string) */
};

enum vosmem_dtype { VOSMEM_F32 = 0, VOSMEM_BF16 = 1 };

/* kernel path for the similarity + selection stage */
enum vosmem_path {
  VOSMEM_PATH_AUTO = 0,   /* tcgen05 when CK == 64, else SIMT */
  VOSMEM_PATH_SIMT = 1,   /* exact fp32 CUDA-core kernel, any CK <= 256 */
  VOSMEM_PATH_TCGEN05 = 2 /* TMA(bulk) + tcgen05/TMEM kernel, CK == 64 */
};

#define VOSMEM_MAX_TOPK 32     /* per-query survivors held one per lane of a warp */
#define VOSMEM_MAX_LISTS 16   /* candidate lists one merge call takes (ranks of a sharded bank) */
#define VOSMEM_MAX_BATCH 12   /* problems per vosmem_match_batch call */
#define VOSMEM_KEY_TILE 64     /* keys per packed image tile */
#define VOSMEM_QUERY_TILE 128  /* queries per packed image tile (= TMEM lanes) */
#define VOSMEM_MAX_SEGMENTS 2  /* [long-term | working] */

int vosmem_abi_version(void);
const char *vosmem_last_error(void);
const char *vosmem_status_string(int status);

/* ---------------------------------------------------------------------------------------------
 * Layout helpers (host-only arithmetic; usable without a GPU).
 * ------------------------------------------------------------------------------------------- */

/* bytes of the packed key image holding `capacity` keys (rounded up to VOSMEM_KEY_TILE) */
int64_t vosmem_key_image_bytes(int ck, int64_t capacity);
/* scratch bytes vosmem_select_topk / vosmem_match need for `hw` queries over `n_keys` keys */
int64_t vosmem_workspace_bytes(int ck, int hw, int64_t n_keys);
/* Prepare a freshly allocated workspace (zero fill; launch epoch, CTA departure counter and error flags at its
 * head).  Call ONCE per allocation, before the first select / match call that uses it.  Afterwards the library
 * maintains the contents: calls that share a workspace must be ordered on one stream (or otherwise serialised), and a
 * captured CUDA graph of such calls may be replayed with new query / memory contents -- the launch epoch that
 * validates the published thresholds lives in the workspace and advances on the device with every launch. */
int vosmem_workspace_init(void *workspace, int64_t workspace_bytes, vosmem_stream_t stream);
/* Sticky device-side error flags of a workspace; synchronises `stream`.  VOSMEM_OK, or VOSMEM_ENOTSUP when a tcgen05
 * launch found tensor memory shared with another kernel (its results are then undefined).  Not for the per-frame path. */
int vosmem_workspace_status(const void *workspace, vosmem_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Store-side packing.  Replaces the torch.cat growth of KeyValueMemoryStore.add
 * (tracker/inference/kv_memory_store.py:36-90): the store owns preallocated buffers and packs
 * only the appended range.
 * ------------------------------------------------------------------------------------------- */

/* Pack keys [begin, end) of a bank into the tensor-core operand image.
 * Row n of the operand is  s[n]/sqrt(CK) * [ k[:,n]^2 , k[:,n] , 1 ]  split into bf16 hi/lo halves
 * (memory_util.py:24-25,35: the shrinkage scale is folded into the memory side).
 * shrinkage may be NULL (memory_util.py:36-37).  Only CK == 64 has an image format. */
int vosmem_pack_keys(const float *key, int64_t key_ld, const float *shrinkage, int ck, int64_t begin,
                     int64_t end, void *image, int64_t capacity, vosmem_stream_t stream);

/* Transpose-append values into the gather-friendly shadow:
 *   shadow[(dst_begin + i) * shadow_ld + r] = value[r * value_ld + src_begin + i],  r < rows, i < n
 * (the reference keeps values rows x N and concatenates on N: kv_memory_store.py:70,88). */
int vosmem_pack_values(const float *value, int64_t value_ld, int rows, int64_t src_begin, int64_t n,
                       void *shadow, int64_t shadow_ld, int64_t dst_begin, int shadow_dtype,
                       vosmem_stream_t stream);

/* Append `m` memory elements to a bank in ONE call (KeyValueMemoryStore.add, kv_memory_store.py:36-90, without its
 * torch.cat re-allocation): keys / shrinkage / selection copied behind the `n` elements the preallocated buffers hold,
 * use_count = 0 and life_count = 1e-7 for the new elements (:37-38), the tensor-core key image packed for them, and the
 * values of every object group written both in the reference layout (rows x capacity) and, transposed, into the
 * gather-friendly shadow.  Two launches (keys, values).  The caller has made sure every buffer holds n + m elements. */
#define VOSMEM_MAX_GROUPS 8
typedef struct vosmem_append_group {
  const float *value;   /* rows x m new values, row pitch value_ld                               */
  int64_t value_ld;
  int rows;             /* n_g * CV                                                              */
  float *ref;           /* rows x capacity fp32, row pitch ref_ld (reference layout)             */
  int64_t ref_ld;
  void *shadow;         /* capacity x rows (vosmem_pack_values layout), row pitch shadow_ld      */
  int64_t shadow_ld;
  int64_t n;            /* elements this group already holds                                     */
} vosmem_append_group;

typedef struct vosmem_append_desc {
  int ck, m;
  int64_t n;                         /* elements the bank already holds                           */
  const float *key;                  /* CK x m, row pitch key_ld                                  */
  int64_t key_ld;
  const float *shrinkage;            /* m, or NULL                                                */
  const float *selection;            /* CK x m (row pitch selection_ld), or NULL                  */
  int64_t selection_ld;
  float *bank_key;                   /* CK x capacity, row pitch bank_ld                          */
  int64_t bank_ld;
  float *bank_shrinkage;             /* capacity, or NULL                                         */
  float *bank_selection;             /* CK x capacity (row pitch bank_ld), or NULL                */
  float *bank_use, *bank_life;       /* capacity each, or NULL                                    */
  void *key_image;                   /* packed image of the bank (CK == 64), or NULL              */
  int64_t capacity;
  int value_dtype;                   /* enum vosmem_dtype of the shadows                          */
  int n_groups;
  vosmem_append_group group[VOSMEM_MAX_GROUPS];
} vosmem_append_desc;

int vosmem_store_append(const vosmem_append_desc *desc, vosmem_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The hot path: MemoryManager.match_memory for one object group
 * (tracker/inference/memory_manager.py:57-150 driving memory_util.py:7-65).
 * ------------------------------------------------------------------------------------------- */

/* One contiguous candidate range of one bank ([long-term] or [working]). */
typedef struct vosmem_segment {
  const float *key;       /* CK x N fp32, row pitch key_ld              (SIMT path)            */
  int64_t key_ld;
  const float *shrinkage; /* N fp32 or NULL                              (SIMT path)            */
  const void *key_image;  /* packed image from vosmem_pack_keys or NULL  (tcgen05 path)         */
  int64_t begin, end;     /* candidate keys [begin, end) of the bank; a group that entered the
                             video late only sees a suffix (memory_manager.py:88-98,137-140)   */
} vosmem_segment;

/* Similarity + per-query top-k.  Candidate index = position on the concatenated candidate axis
 * [seg0 | seg1] plus index_base (a rank's offset when the bank is sharded along N). */
typedef struct vosmem_select_desc {
  int ck, hw, top_k;
  const float *query_key;       /* CK x HW                                                     */
  const float *query_selection; /* CK x HW or NULL (memory_util.py:28-32)                      */
  int n_segments;
  vosmem_segment seg[VOSMEM_MAX_SEGMENTS];
  int64_t index_base;
  int path;                     /* enum vosmem_path                                            */
  void *workspace;        /* device scratch of >= vosmem_workspace_bytes(), prepared once by
                           * vosmem_workspace_init(); not shared by concurrent calls                       */
  int64_t workspace_bytes;
} vosmem_select_desc;

/* out_score / out_index: HW x top_k, sorted by descending score; missing candidates (fewer than
 * top_k keys) are (-inf, -1).  Replaces get_similarity + torch.topk (memory_util.py:7-39,46)
 * without materialising the N x HW matrix. */
int vosmem_select_topk(const vosmem_select_desc *desc, float *out_score, int64_t *out_index,
                       vosmem_stream_t stream);

/* Merge `n_lists` candidate lists (each HW x top_k, e.g. one per rank after an all-gather) into the
 * global top_k per query.  lists are laid out [list][HW][top_k]. */
int vosmem_merge_topk(const float *scores, const int64_t *indices, int n_lists, int hw, int top_k,
                      float *out_score, int64_t *out_index, vosmem_stream_t stream);

/* The same with one pointer pair per list (host arrays of `n_lists` device pointers, each list HW x top_k).  The lists
 * may sit in peer-mapped memory of other GPUs: the kernel then loads them over NVLink itself, i.e. the all-gather of
 * the sharded long-term readout and the merge are one kernel (vos_e_sam_b200/sharded.py, exchange='peer'). */
int vosmem_merge_topk_ptrs(const float *const *scores, const int64_t *const *indices, int n_lists, int hw, int top_k,
                           float *out_score, int64_t *out_index, vosmem_stream_t stream);

/* Values of one object group on the candidate axis. */
typedef struct vosmem_value_segment {
  const void *shadow;   /* rows of n_g*CV values per memory element (vosmem_pack_values layout) */
  int64_t shadow_ld;    /* elements between consecutive memory elements                         */
  int64_t first;        /* global candidate index of shadow row 0                               */
  int64_t count;        /* candidates [first, first+count) live in this shadow                  */
  float *use_count;     /* += sum_q affinity[n, q]   or NULL (kv_memory_store.py:92-99)         */
  float *life_count;    /* [0, count) += 1 in the same launch, or NULL (kv_memory_store.py:99;   */
                        /* otherwise call vosmem_age)                                            */
} vosmem_value_segment;

typedef struct vosmem_readout_desc {
  int hw, top_k;
  int rows;             /* n_g * CV                                                            */
  int value_dtype;      /* enum vosmem_dtype of the shadows                                    */
  int n_segments;
  vosmem_value_segment seg[VOSMEM_MAX_SEGMENTS];
  float *out;           /* rows x HW, row pitch out_ld (the reference's n_g x CV x h x w)      */
  int64_t out_ld;
  float *out_weight;    /* optional HW x top_k softmax weights, or NULL                        */
  /* Optional grouped output rows: row r lands at out[(r / out_group_rows) * out_group_stride +         */
  /* (r % out_group_rows) * out_ld + q].  With out_group_rows = CV and out_group_stride = (CV + CH) *  */
  /* HW the readout is written straight into the channels [0, CV) of the decoder's concatenated       */
  /* num_objects x (CV + CH) x h x w input (the torch.cat of model/modules.py:232-233).  0 = plain.    */
  int out_group_rows;
  int64_t out_group_stride;
} vosmem_readout_desc;

/* softmax over the top_k survivors (max-subtracted; memory_util.py:48-49), usage scatter-add
 * (memory_util.py:62-63) and the affinity x value product restricted to the survivors
 * (memory_manager.py:53-55,145-148).  Candidates whose index falls outside every value segment
 * contribute to the softmax normalisation but not to `out` (values held by another rank). */
int vosmem_softmax_readout(const vosmem_readout_desc *desc, const float *score, const int64_t *index,
                           vosmem_stream_t stream);

/* select + softmax + readout for one group in one call (what match_memory does per group). */
int vosmem_match(const vosmem_select_desc *select, const vosmem_readout_desc *readout,
                 float *scratch_score, int64_t *scratch_index, vosmem_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N-sharded long-term bank (BASELINE configs[3]; no reference counterpart: tools/runner.py:32 is single-GPU).
 * Rank r owns keys [lo_r, hi_r) and HW / world of the query rows.  Per frame:
 *   vosmem_select_push      fused similarity + LOCAL top-k over the rank's keys, then every query's (score, global
 *                           index) list is written straight into the exchange buffer of the rank that OWNS that query
 *                           (peer-mapped memory: st.global over NVLink), followed by one flag per owner;
 *   vosmem_exchange_readout waits for the flags of all source ranks, merges the `world` lists of each of its own
 *                           queries, softmax, usage, sparse readout of its query slice;
 *   vosmem_push_slice / vosmem_wait_flags   optional: replicate the readout slices on every rank.
 * No collective and no global barrier: ranks only wait for the data they consume.  Exchange entries are 8 bytes
 * {fp32 score, int32 global key index}, VOSMEM_EXCH_K per (source rank, query).
 * ------------------------------------------------------------------------------------------- */
#define VOSMEM_EXCH_K 32       /* entries per (source rank, query) in an exchange buffer (top_k <= 32, padded)  */
#define VOSMEM_MAX_RANKS 16

typedef struct vosmem_push_desc {
  int world, rank;
  int per;                               /* queries owned by each rank: owner(q) = q / per                         */
  int64_t index_base;                    /* global index of this rank's first key                                  */
  void *dst[VOSMEM_MAX_RANKS];           /* per owner: [per][VOSMEM_EXCH_K] x 8 B, this rank's lists for the owner's
                                            queries (a peer-mapped address for owner != rank)                      */
  uint32_t *flag[VOSMEM_MAX_RANKS];      /* per owner: word set to `seq` (system-scope release) once the lists have
                                            landed, or NULL                                                        */
  uint32_t seq;
  uint32_t *ticket;                      /* zero-initialised device word of this rank (last-CTA detection)         */
  /* Optional (all NULL = off): thresholds shared across the ranks while the selection kernels run.  rank_pub[d] is a
   * zero-initialised [world][round_up(HW, 128)] x 8-byte array in rank d's memory (peer-mapped for d != rank).  Every
   * rank keeps writing, per query, a lower bound that world x (its candidate lists) x R of ITS keys reach into row
   * `rank` of every array and reads the other rows of its own array, so that a query's threshold is backed by 33
   * keys of the WHOLE bank: each rank then keeps ~33 / world candidates per query instead of 33, which takes the
   * selection out of its append-bound regime.  Requires that all ranks issue the same sequence of calls on the
   * workspaces they pass (their launch epochs must agree). */
  void *rank_pub[VOSMEM_MAX_RANKS];
  /* Two candidate segments ([long-term shard | working-memory shard] of a sharded MemoryManager): local candidate i
   * becomes global index  i + index_base  for i < seg0_len  and  i - seg0_len + index_base1  otherwise.
   * seg0_len < 0: one base (index_base) for every candidate. */
  int64_t seg0_len;
  int64_t index_base1;
} vosmem_push_desc;

/* select (desc->index_base is ignored: push->index_base is applied) + merge of the split lists + push.  A rank
 * whose segments are all empty (fewer 64-key tiles than ranks) pushes empty lists and still raises its flags. */
int vosmem_select_push(const vosmem_select_desc *select, const vosmem_push_desc *push, vosmem_stream_t stream);

typedef struct vosmem_exchange_desc {
  const void *lists;        /* [n_lists][list_stride entries] exchange entries of this rank's queries (local memory)  */
  int n_lists;              /* source ranks                                                                           */
  int64_t list_stride;      /* entries between consecutive source ranks' lists                                        */
  int64_t first_entry;      /* entry of query 0 of the readout inside a list (0, or q_lo * VOSMEM_EXCH_K when the lists
                               cover all queries, e.g. after an NCCL all-gather)                                      */
  const uint32_t *flags;    /* n_lists words; the kernel waits until each equals `seq` (NULL: no wait)                */
  uint32_t seq;
  uint32_t *status;         /* device word, set to 1 if the wait timed out (bounded spin), or NULL                    */
} vosmem_exchange_desc;

/* readout->hw = number of queries of this rank's slice, readout->out = first column of the slice */
int vosmem_exchange_readout(const vosmem_readout_desc *readout, const vosmem_exchange_desc *exchange,
                            vosmem_stream_t stream);

/* Copy the rows x cols block at `src` (row pitch src_ld) to dst[i] (row pitch dst_ld) for every i < n_dst -- the
 * readout slice of this rank into every rank's full readout -- then set flag[i] = seq (system-scope release). */
int vosmem_push_slice(const float *src, int64_t src_ld, int rows, int cols, float *const *dst, int64_t dst_ld,
                      uint32_t *const *flag, int n_dst, uint32_t seq, uint32_t *ticket, vosmem_stream_t stream);
/* Stream-ordered wait until flags[i] == seq for every bit i of `mask` (bounded spin; *status = 1 on timeout). */
int vosmem_wait_flags(const uint32_t *flags, uint32_t mask, uint32_t seq, uint32_t *status, vosmem_stream_t stream);

/* life_count[0:n] += 1  (kv_memory_store.py:99) */
int vosmem_age(float *life_count, int64_t n, vosmem_stream_t stream);

/* `n` (1..VOSMEM_MAX_BATCH) independent match problems -- e.g. one frame of each of n sequences, every one with its
 * own MemoryManager (tools/runner.py:61-63 builds one tracker per sequence) -- in ONE selection launch and ONE
 * readout launch (blockIdx.z = problem).  All problems share ck == 64, hw and top_k; memory sizes, segments, value
 * rows and workspaces are per problem (the workspaces must be distinct).  With n x query tiles >= 148 every CTA
 * streams a problem's whole key axis, which is where the tcgen05 path is most efficient.  Equivalent to n
 * vosmem_match calls. */
int vosmem_match_batch(const vosmem_select_desc *select, const vosmem_readout_desc *readout, int n,
                       vosmem_stream_t stream);


/* ---------------------------------------------------------------------------------------------
 * Dense twins of tracker/model/memory_util.py, for API parity (training-time read_memory,
 * memory consolidation).  Batch 1; callers loop over B.
 * ------------------------------------------------------------------------------------------- */

/* get_similarity (memory_util.py:7-39): out is N x HW, row pitch hw. shrinkage / selection may be NULL */
int vosmem_similarity_dense(const float *key, int64_t key_ld, const float *shrinkage, const float *query_key,
                            const float *query_selection, int ck, int64_t n, int hw, float *out,
                            vosmem_stream_t stream);

/* do_softmax (memory_util.py:41-65) over the key axis of an N x HW matrix (row pitch sim_ld).
 * top_k <= 0 means the max-subtracted dense softmax; 1 <= top_k <= 512 here (the fused per-frame calls above keep
 * top_k <= VOSMEM_MAX_TOPK); usage (N) may be NULL.  affinity may alias similarity (the reference's inplace=True). */
int vosmem_softmax_dense(const float *similarity, int64_t sim_ld, int64_t n, int hw, int top_k, float *affinity,
                         int64_t aff_ld, float *usage, vosmem_stream_t stream);

/* The same with an optional scratch of vosmem_softmax_dense_scratch_bytes(n, hw) bytes: the dense (top_k <= 0) softmax
 * then runs parallel over the key axis as well (two launches; a consolidation's 8 100 x 128 matrix: 630 -> ~15 us). */
int64_t vosmem_softmax_dense_scratch_bytes(int64_t n, int hw);
int vosmem_softmax_dense_ws(const float *similarity, int64_t sim_ld, int64_t n, int hw, int top_k, float *affinity,
                            int64_t aff_ld, float *usage, void *scratch, int64_t scratch_bytes, vosmem_stream_t stream);

/* readout / MemoryManager._readout (memory_util.py:73-80, memory_manager.py:53-55):
 * out[rows x hw] = value[rows x n] @ affinity[n x hw] */
int vosmem_readout_dense(const float *value, int64_t value_ld, const float *affinity, int64_t aff_ld, int rows,
                         int64_t n, int hw, float *out, int64_t out_ld, vosmem_stream_t stream);

/* The dense readout on tcgen05 (bf16 hi / lo split of both operands, split-K, fp32 accumulation in tensor memory), for
 * the shapes of a consolidation (memory_manager.py:280-284: 2 560 x 8 100 @ 8 100 x 128).  hw <= 256; `workspace` of
 * vosmem_readout_dense_tc_workspace_bytes(rows, n, hw) bytes (0: shape not supported, use vosmem_readout_dense). */
int64_t vosmem_readout_dense_tc_workspace_bytes(int rows, int64_t n, int hw);
int vosmem_readout_dense_tc(const float *value, int64_t value_ld, const float *affinity, int64_t aff_ld, int rows,
                            int64_t n, int hw, float *out, int64_t out_ld, void *workspace, int64_t workspace_bytes,
                            vosmem_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * KeyProjection (tracker/model/modules.py:194-211): key_proj, d_proj and e_proj -- three 3x3 convolutions (padding 1)
 * of the 1/16-scale feature map, the producers of the query key / shrinkage / selection that the readout consumes
 * (tracker/model/network.py:55) -- as ONE implicit GEMM with 2 * key_dim + 1 output channels on tcgen05 (bf16 hi / lo
 * split, ~fp32 accuracy).  key_dim == 64, in_dim a multiple of 32 (XMem: 1024), batch 1, w <= 240.
 * ------------------------------------------------------------------------------------------- */

/* bytes of the packed weights for vosmem_keyproj_pack_weights (0: unsupported dimensions) */
int64_t vosmem_keyproj_weight_bytes(int in_dim, int key_dim);
/* key_w: key_dim x in_dim x 3 x 3, d_w: 1 x in_dim x 3 x 3, e_w: key_dim x in_dim x 3 x 3 (nn.Conv2d layout).  Once per model. */
int vosmem_keyproj_pack_weights(const float *key_w, const float *d_w, const float *e_w, int in_dim, int key_dim,
                                void *packed, vosmem_stream_t stream);
int64_t vosmem_keyproj_workspace_bytes(int in_dim, int key_dim, int h, int w);
/* x: in_dim x h x w.  key: key_dim x h x w = key_proj(x); shrinkage: h x w = d_proj(x)^2 + 1 or NULL (need_s false);
 * selection: key_dim x h x w = sigmoid(e_proj(x)) or NULL (need_e false)  (modules.py:206-211). */
int vosmem_keyproj_forward(const float *x, int in_dim, int key_dim, int h, int w, const void *packed_weights,
                           const float *key_bias, const float *d_bias, const float *e_bias, float *key,
                           float *shrinkage, float *selection, void *workspace, int64_t workspace_bytes,
                           vosmem_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Self-test hook: D[128 x 64] = A[128 x 272] . B[64 x 272]^T through the same descriptors the
 * tcgen05 path uses, on packed images.  Returns the raw accumulator tile (fp32, 128 x 64).
 * ------------------------------------------------------------------------------------------- */
int vosmem_debug_umma_tile(const void *query_image, const void *key_image, float *out, vosmem_stream_t stream);
int vosmem_debug_pack_query(const float *query_key, const float *query_selection, int ck, int hw, void *image,
                            vosmem_stream_t stream);
int64_t vosmem_query_image_bytes(int ck, int hw);
/* per-role cycle counters of the tcgen05 kernel (16 int64 per CTA) into `device_buffer`; NULL switches it off */
int vosmem_debug_set_timing_buffer(void *device_buffer);
/* cudaEvent_t handles recorded on the stream before the pack kernel and after the pack / select / last (merge, or fused
 * merge+readout) kernel of the next vosmem_select_topk / vosmem_match calls; NULLs switch it off */
int vosmem_debug_set_stage_events(void *begin, void *after_pack, void *after_select, void *after_last);

#ifdef __cplusplus
}
#endif
#endif /* VOSMEM_H_ */
