// KeyProjection on tcgen05: the step right before the memory readout.
//
// Reference: tracker/model/modules.py:194-211 -- three 3x3 convolutions (padding 1) of the 1/16-scale feature map
// f16 (B x 1024 x h x w): key_proj (64 channels), d_proj (1 channel; shrinkage = d^2 + 1) and e_proj (64 channels;
// selection = sigmoid(e)).  They share their input, so here they are ONE implicit GEMM with 129 output channels:
//
//   D[p, n] = sum over taps (ky, kx) and channels c of  Xpad[p + (ky - 1) * Wp + (kx - 1), c] * W[n, c, ky, kx]
//
// over the positions p of the zero-padded (h + 2) x (w + 2) grid (row-major, Wp = w + 2): a tap is a CONSTANT shift of
// p, so the activations of a 128-position tile plus a halo of Wp + 1 rows on either side are loaded into shared memory
// once per channel block and all nine taps read them through nine different operand start addresses.  Border positions
// compute junk that the finalize kernel drops (10 % of the work at DAVIS-480p, 5 % at 1080p).
//
//   pack_x_kernel      fp32 NCHW -> bf16 (hi, lo) pairs, K-major operand rows [hl][c / 8][padded position][8 channels]
//                      (16 bytes per row and chunk; halo and border rows written as zeros every call)
//   keyproj_mma_kernel grid = (position tiles of 128, K splits); per 32-channel block: one halo tile of activations
//                      (cp.async.bulk, two stages) and, per tap, an 18 KB block of packed weights (four stages);
//                      6 tcgen05.mma (M128 N144 K16: hi*hi + lo*hi + hi*lo, ~fp32 accuracy like the selection kernel)
//                      per tap into one 128 x 144 fp32 accumulator in tensor memory; partial sums -> global
//   finalize_kernel    sum of the K splits + bias, d^2 + 1, sigmoid, transposed into the reference's NCHW outputs
//
// Weights are packed once per model (vosmem_keyproj_pack_weights): [32-channel block][tap][hl][c / 8][n = 144][8].
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vosmem {
namespace {

constexpr int KP_M = 128;            // positions per tile (TMEM lanes)
constexpr int KP_N = 144;            // 64 key + 1 shrinkage + 64 selection channels, padded to a multiple of 16
constexpr int KP_KC = 32;            // channels per K block
constexpr int KP_K8 = KP_KC / 8;     // 16-byte chunks per operand row and K block
constexpr int KP_TAPS = 9;
constexpr int KP_A_STAGES = 2, KP_B_STAGES = 4;
constexpr int KP_B_BYTES = 2 * KP_K8 * KP_N * 16;     // one (K block, tap) of packed weights: hi + lo = 18 432 B
template <int NN>
constexpr int b_bytes() { return 2 * KP_K8 * NN * 16; }
constexpr int KP_THREADS = 192;      // warps 0-3 epilogue (TMEM lane quarters), warp 4 producer, warp 5 MMA
constexpr int KP_TMEM_COLS = 256;

struct KpGeom {
  int in_dim, h, w, wp, p_pad, q_rows, halo, splits;   // q_rows = p_pad + 2 * (wp + 1)
};

__host__ __device__ inline KpGeom kp_geom(int in_dim, int h, int w) {
  KpGeom g;
  g.in_dim = in_dim;
  g.h = h;
  g.w = w;
  g.wp = w + 2;
  g.p_pad = (int)round_up64((int64_t)(h + 2) * (w + 2), KP_M);
  g.q_rows = g.p_pad + 2 * (g.wp + 1);
  g.halo = KP_M + 2 * (g.wp + 1);
  // K splits: fill the 148 SMs once, at least one channel block per CTA
  const int tiles = g.p_pad / KP_M, blocks = in_dim / KP_KC;
  int s = 148 / tiles;
  if (s < 1) s = 1;
  while (s > 1 && blocks % s != 0) --s;
  g.splits = s;
  return g;
}

// ---- activations: fp32 NCHW -> [hl][c8][q][8] bf16, q = padded position + wp + 1 -----------------------------------
__global__ void __launch_bounds__(256) pack_x_kernel(const float *__restrict__ x, KpGeom g, uint4 *__restrict__ xp) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  const int c8 = blockIdx.y;
  if (q >= g.q_rows) return;
  const int p = q - (g.wp + 1);
  const int yp = p >= 0 ? p / g.wp : -1, xq = p >= 0 ? p % g.wp : -1;
  const bool inside = p >= 0 && yp >= 1 && yp <= g.h && xq >= 1 && xq <= g.w;
  uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
  if (inside) {
    const float *src = x + ((int64_t)c8 * 8) * g.h * g.w + (int64_t)(yp - 1) * g.w + (xq - 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(src[(int64_t)(2 * j) * g.h * g.w], h0, l0);
      split_bf16(src[(int64_t)(2 * j + 1) * g.h * g.w], h1, l1);
      hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
  const int64_t c8_total = g.in_dim / 8;
  xp[((int64_t)c8) * g.q_rows + q] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  xp[(c8_total + c8) * g.q_rows + q] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- weights: [n][c][ky][kx] fp32 (three tensors) -> [K block][tap][hl][k8][n = 144][8] bf16 -------------------------
__global__ void __launch_bounds__(KP_N) pack_w_kernel(const float *__restrict__ key_w, const float *__restrict__ d_w,
                                                      const float *__restrict__ e_w, int in_dim, int key_dim,
                                                      uint4 *__restrict__ out) {
  const int n = threadIdx.x, k8 = blockIdx.x % KP_K8, tap = (blockIdx.x / KP_K8) % KP_TAPS, kb = blockIdx.x / (KP_K8 * KP_TAPS);
  const float *src = nullptr;
  if (n < key_dim) src = key_w + (int64_t)n * in_dim * KP_TAPS;
  else if (n == key_dim) src = d_w;
  else if (n < 2 * key_dim + 1) src = e_w + (int64_t)(n - key_dim - 1) * in_dim * KP_TAPS;
  uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
  if (src != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      const int c = kb * KP_KC + k8 * 8 + 2 * j;
      split_bf16(src[(int64_t)c * KP_TAPS + tap], h0, l0);
      split_bf16(src[(int64_t)(c + 1) * KP_TAPS + tap], h1, l1);
      hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
  uint4 *blk = out + (int64_t)(kb * KP_TAPS + tap) * (KP_B_BYTES / 16);
  blk[k8 * KP_N + n] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  blk[(KP_K8 + k8) * KP_N + n] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---- the implicit GEMM -----------------------------------------------------------------------------------------------
struct KpArgs {
  const unsigned char *xp;   // packed activations
  const unsigned char *wq;   // packed weights
  float *partial;            // [splits][p_pad][NN]
  KpGeom g;
};

// NN output columns (a multiple of 16, <= 256), TAPS row-shifted reads of the activation tile per K block.  With
// TAPS == 1 and a halo of 128 rows this is a plain split-K GEMM  D[128 x NN] += A[128 x K] . B[NN x K]^T  over packed
// operands, which the dense readout of the consolidation uses (tc_readout_dense below).
template <int NN, int TAPS>
__global__ void __launch_bounds__(KP_THREADS, 1) keyproj_mma_kernel(const __grid_constant__ KpArgs a) {
  constexpr int KP_N = NN, KP_TAPS = TAPS, KP_B_BYTES = b_bytes<NN>();
  constexpr uint32_t KP_IDESC = ptx::umma_idesc_bf16(KP_M, NN);
  extern __shared__ __align__(128) unsigned char smem[];
  const KpGeom &g = a.g;
  const uint32_t a_stage_bytes = 2u * KP_K8 * g.halo * 16u;   // hi + lo halo tiles of one K block
  unsigned char *sm_a = smem;
  unsigned char *sm_b = smem + KP_A_STAGES * a_stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm_b + KP_B_STAGES * KP_B_BYTES);
  uint64_t *a_full = bars, *a_empty = bars + KP_A_STAGES, *b_full = a_empty + KP_A_STAGES, *b_empty = b_full + KP_B_STAGES,
           *d_full = b_empty + KP_B_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_full + 1);

  const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * KP_M;
  const int blocks_total = g.in_dim / KP_KC, blocks_mine = blocks_total / g.splits, kb0 = blockIdx.y * blocks_mine;

  if (threadIdx.x == 0) {
    for (int i = 0; i < KP_A_STAGES; ++i) { ptx::mbar_init(a_full + i, 1); ptx::mbar_init(a_empty + i, 1); }
    for (int i = 0; i < KP_B_STAGES; ++i) { ptx::mbar_init(b_full + i, 1); ptx::mbar_init(b_empty + i, 1); }
    ptx::mbar_init(d_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, KP_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 4) {
    // ===== producer: one halo tile of activations per K block, one weight block per (K block, tap) =====
    if (lane == 0) {
      const int64_t c8_total = g.in_dim / 8;
      int bi = 0;   // weight blocks issued
      for (int kb = 0; kb < blocks_mine; ++kb) {
        const int st = kb % KP_A_STAGES;
        ptx::mbar_wait_backoff(a_empty + st, ((kb / KP_A_STAGES) & 1) ^ 1, 20);
        ptx::mbar_arrive_expect_tx(a_full + st, a_stage_bytes);
        for (int hl = 0; hl < 2; ++hl)
          for (int k8 = 0; k8 < KP_K8; ++k8) {
            const int64_t c8 = (int64_t)(kb0 + kb) * KP_K8 + k8;
            const unsigned char *src = a.xp + ((hl * c8_total + c8) * g.q_rows + p0) * 16;
            ptx::bulk_g2s(sm_a + st * a_stage_bytes + (hl * KP_K8 + k8) * g.halo * 16, src, g.halo * 16, a_full + st);
          }
        for (int tap = 0; tap < KP_TAPS; ++tap, ++bi) {
          const int sb = bi % KP_B_STAGES;
          ptx::mbar_wait_backoff(b_empty + sb, ((bi / KP_B_STAGES) & 1) ^ 1, 20);
          ptx::mbar_arrive_expect_tx(b_full + sb, KP_B_BYTES);
          ptx::bulk_g2s(sm_b + sb * KP_B_BYTES, a.wq + ((int64_t)(kb0 + kb) * KP_TAPS + tap) * KP_B_BYTES, KP_B_BYTES, b_full + sb);
        }
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer (whole warp in lock step, one elected lane issues: see select_tc.cu) =====
    const uint32_t a_lbo = (uint32_t)g.halo * 16u, b_lbo = KP_N * 16u;
    int bi = 0;
    for (int kb = 0; kb < blocks_mine; ++kb) {
      const int st = kb % KP_A_STAGES;
      ptx::mbar_wait_backoff(a_full + st, (kb / KP_A_STAGES) & 1, 20);
      const uint32_t a_base = ptx::smem_u32(sm_a + st * a_stage_bytes);
      for (int tap = 0; tap < KP_TAPS; ++tap, ++bi) {
        const int sb = bi % KP_B_STAGES;
        ptx::mbar_wait_backoff(b_full + sb, (bi / KP_B_STAGES) & 1, 20);
        ptx::tc_fence_after();
        const uint32_t b_base = ptx::smem_u32(sm_b + sb * KP_B_BYTES);
        const uint32_t shift = (uint32_t)((tap / 3) * g.wp + (tap % 3)) * 16u;   // the tap's constant row shift
#pragma unroll
        for (int s = 0; s < KP_KC / 16; ++s) {   // K = 16 steps: two 8-channel chunks each
          const uint64_t a_hi = ptx::umma_desc(a_base + shift + 2 * s * a_lbo, a_lbo, 128);
          const uint64_t a_lo = ptx::umma_desc(a_base + KP_K8 * a_lbo + shift + 2 * s * a_lbo, a_lbo, 128);
          const uint64_t b_hi = ptx::umma_desc(b_base + 2 * s * b_lbo, b_lbo, 128);
          const uint64_t b_lo = ptx::umma_desc(b_base + KP_K8 * b_lbo + 2 * s * b_lbo, b_lbo, 128);
          ptx::umma_bf16_ss_elect(tmem_d, a_hi, b_hi, KP_IDESC, (kb | tap | s) != 0);
          ptx::umma_bf16_ss_elect(tmem_d, a_lo, b_hi, KP_IDESC, 1);
          ptx::umma_bf16_ss_elect(tmem_d, a_hi, b_lo, KP_IDESC, 1);
        }
        ptx::umma_commit_elect(b_empty + sb);
      }
      ptx::umma_commit_elect(a_empty + st);
    }
    ptx::umma_commit_elect(d_full);
  } else {
    // ===== epilogue: warp w reads TMEM lanes 32 w .. (positions p0 + 32 w + lane), 144 columns -> partial sums =====
    ptx::mbar_wait(d_full, 0);
    ptx::tc_fence_after();
    float *dst = a.partial + ((int64_t)blockIdx.y * g.p_pad + p0 + warp * 32 + lane) * KP_N;
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 + 32 <= NN; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(taddr + c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4 *>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if constexpr (NN % 32 == 16) {
      constexpr int C0 = NN - 16;
      uint32_t t0[8], t1[8];
      ptx::tmem_ld_32x8(taddr + C0, t0);
      ptx::tmem_ld_32x8(taddr + C0 + 8, t1);
      ptx::tmem_ld_wait();
      *reinterpret_cast<uint4 *>(dst + C0) = make_uint4(t0[0], t0[1], t0[2], t0[3]);
      *reinterpret_cast<uint4 *>(dst + C0 + 4) = make_uint4(t0[4], t0[5], t0[6], t0[7]);
      *reinterpret_cast<uint4 *>(dst + C0 + 8) = make_uint4(t1[0], t1[1], t1[2], t1[3]);
      *reinterpret_cast<uint4 *>(dst + C0 + 12) = make_uint4(t1[4], t1[5], t1[6], t1[7]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_d, KP_TMEM_COLS);
}

// ---- finalize: K-split sum + bias + activation, positions x channels -> NCHW ------------------------------------------
// block = 128 threads, 32 consecutive padded positions; the 32 x 144 block of every split is contiguous in memory
struct KpOut {
  const float *partial, *key_b, *d_b, *e_b;
  float *key, *shrinkage, *selection;
  KpGeom g;
  int key_dim;
};

constexpr int KP_FIN_CH = 36;   // channels per finalize CTA (grid.y = KP_N / 36 = 4: enough CTAs to hide the load latency)
__global__ void __launch_bounds__(128) keyproj_finalize_kernel(const __grid_constant__ KpOut o) {
  __shared__ float tile[32][KP_FIN_CH + 1];
  const KpGeom &g = o.g;
  const int pb = blockIdx.x * 32, n0 = blockIdx.y * KP_FIN_CH;
  // 16-byte loads, the K splits of an element in flight together (the sum order is fixed: split 0, 1, ...)
  for (int e = threadIdx.x; e < 32 * (KP_FIN_CH / 4); e += 128) {
    const int r = e / (KP_FIN_CH / 4), c = (e % (KP_FIN_CH / 4)) * 4;
    const float4 *src = reinterpret_cast<const float4 *>(o.partial + ((int64_t)pb + r) * KP_N + n0 + c);
    const int64_t split_stride = (int64_t)g.p_pad * KP_N / 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < g.splits; s0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = s0 + u < g.splits ? src[(s0 + u) * split_stride] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    tile[r][c] = acc.x; tile[r][c + 1] = acc.y; tile[r][c + 2] = acc.z; tile[r][c + 3] = acc.w;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = pb + lane;
  const int yp = p / g.wp, xq = p % g.wp;
  const bool inside = yp >= 1 && yp <= g.h && xq >= 1 && xq <= g.w;
  const int64_t at = (int64_t)(yp - 1) * g.w + (xq - 1), hw = (int64_t)g.h * g.w;
  for (int c = warp; c < KP_FIN_CH; c += 4) {
    const int n = n0 + c;
    if (!inside || n >= 2 * o.key_dim + 1) continue;
    const float v = tile[lane][c];
    if (n < o.key_dim) {
      o.key[n * hw + at] = v + o.key_b[n];
    } else if (n == o.key_dim) {
      const float d = v + o.d_b[0];
      if (o.shrinkage) o.shrinkage[at] = d * d + 1.0f;
    } else if (o.selection) {
      const int ce = n - o.key_dim - 1;
      o.selection[ce * hw + at] = 1.0f / (1.0f + expf(-(v + o.e_b[ce])));
    }
  }
}

struct KpWorkspace {
  unsigned char *xp;
  float *partial;
  int64_t bytes;
};
KpWorkspace kp_carve(void *base, const KpGeom &g) {
  KpWorkspace w;
  unsigned char *p = static_cast<unsigned char *>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    unsigned char *r = p ? p + off : nullptr;
    off += round_up64(bytes, 256);
    return r;
  };
  w.xp = take((int64_t)2 * (g.in_dim / 8) * g.q_rows * 16);
  w.partial = reinterpret_cast<float *>(take((int64_t)g.splits * g.p_pad * KP_N * 4));
  w.bytes = off;
  return w;
}

int kp_check(int in_dim, int key_dim, int h, int w) {
  VOSMEM_CHECK_ARG(key_dim == 64, "keyproj: key_dim=%d (the tcgen05 kernel is built for XMem's 64)", key_dim);
  VOSMEM_CHECK_ARG(in_dim >= KP_KC && in_dim % KP_KC == 0, "keyproj: in_dim=%d must be a multiple of %d", in_dim, KP_KC);
  VOSMEM_CHECK_ARG(h >= 1 && w >= 1 && w <= 240, "keyproj: feature map %d x %d (w <= 240: the halo tile must fit shared memory)", h, w);
  return VOSMEM_OK;
}

// ======================================================================================================================
// Dense readout on tcgen05:  out[rows x hw] = value[rows x n] @ affinity[n x hw]   (MemoryManager._readout,
// memory_manager.py:53-55, as consolidation calls it: memory_manager.py:280-284 -- 2 560 x 8 100 @ 8 100 x 128 at the
// DAVIS shape, 5.3 GFLOP that the 64 x 64 SIMT twin needs 1.6 ms for).  Both operands are packed into bf16 (hi, lo)
// K-major rows (K = memory elements, zero-padded to a multiple of 32) and go through the split-K GEMM above.
// ======================================================================================================================
// value rows x n fp32 (n contiguous) -> [hl][k8][row][8]
__global__ void __launch_bounds__(256) pack_rows_kernel(const float *__restrict__ v, int64_t v_ld, int rows, int64_t n,
                                                        int rows_pad, int k8_total, uint4 *__restrict__ out) {
  const int row = blockIdx.x * 256 + threadIdx.x, k8 = blockIdx.y;
  if (row >= rows_pad) return;
  uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
  if (row < rows) {
    const float *src = v + (int64_t)row * v_ld + (int64_t)k8 * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = (int64_t)k8 * 8 + 2 * j;
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(k < n ? src[2 * j] : 0.f, h0, l0);
      split_bf16(k + 1 < n ? src[2 * j + 1] : 0.f, h1, l1);
      hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
  out[(int64_t)k8 * rows_pad + row] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  out[((int64_t)k8_total + k8) * rows_pad + row] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// affinity n x hw fp32 (hw contiguous) -> [K block][hl][k8][col = NN][8]
template <int NN>
__global__ void __launch_bounds__(NN) pack_cols_kernel(const float *__restrict__ aff, int64_t aff_ld, int64_t n, int hw,
                                                       uint4 *__restrict__ out) {
  const int col = threadIdx.x, k8 = blockIdx.x % KP_K8, kb = blockIdx.x / KP_K8;
  uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
  if (col < hw) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = (int64_t)kb * KP_KC + k8 * 8 + 2 * j;
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(k < n ? aff[k * aff_ld + col] : 0.f, h0, l0);
      split_bf16(k + 1 < n ? aff[(k + 1) * aff_ld + col] : 0.f, h1, l1);
      hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
  uint4 *blk = out + (int64_t)kb * (b_bytes<NN>() / 16);
  blk[k8 * NN + col] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  blk[(KP_K8 + k8) * NN + col] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// out[row][col] = sum over the K splits of partial[s][row][col]
__global__ void __launch_bounds__(256) sum_partials_kernel(const float *__restrict__ partial, int splits, int rows_pad, int nn,
                                                           int rows, int hw, float *__restrict__ out, int64_t out_ld) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int row = (int)(e / nn), col = (int)(e % nn);
  if (row >= rows || col >= hw) return;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[((int64_t)s * rows_pad + row) * nn + col];
  out[(int64_t)row * out_ld + col] = acc;
}

struct DenseGeom {
  KpGeom g;
  int nn, rows_pad;
  int64_t k_pad, a_bytes, b_bytes_total, partial_bytes, total;
};
DenseGeom dense_geom(int rows, int64_t n, int hw) {
  DenseGeom d{};
  d.nn = hw <= 64 ? 64 : hw <= 128 ? 128 : 256;
  d.rows_pad = (int)round_up64(rows, KP_M);
  KpGeom &g = d.g;
  {   // K splits: fill the SMs once; the K axis is zero-padded to a whole number of 32-element blocks per split
    const int tiles = d.rows_pad / KP_M;
    const int64_t blocks = ceil_div64(n, KP_KC);
    int64_t s = 148 / tiles;
    if (s < 1) s = 1;
    if (s > blocks) s = blocks;
    g.splits = (int)s;
    d.k_pad = round_up64(blocks, s) * KP_KC;
  }
  g.in_dim = (int)d.k_pad;
  g.h = g.w = 0;
  g.wp = -1;                 // no halo: every "tap" shift is zero
  g.p_pad = d.rows_pad;
  g.q_rows = d.rows_pad;
  g.halo = KP_M;
  d.a_bytes = round_up64((int64_t)2 * (d.k_pad / 8) * d.rows_pad * 16, 256);
  d.b_bytes_total = round_up64((d.k_pad / KP_KC) * (int64_t)(2 * KP_K8 * d.nn * 16), 256);
  d.partial_bytes = round_up64((int64_t)g.splits * d.rows_pad * d.nn * 4, 256);
  d.total = d.a_bytes + d.b_bytes_total + d.partial_bytes;
  return d;
}

template <int NN>
int run_dense_tc(const float *value, int64_t value_ld, const float *aff, int64_t aff_ld, int rows, int64_t n, int hw,
                 float *out, int64_t out_ld, const DenseGeom &d, unsigned char *ws, cudaStream_t st) {
  const KpGeom &g = d.g;
  unsigned char *ap = ws, *bp = ws + d.a_bytes;
  float *partial = reinterpret_cast<float *>(ws + d.a_bytes + d.b_bytes_total);
  const int k8_total = (int)(d.k_pad / 8);
  pack_rows_kernel<<<dim3((d.rows_pad + 255) / 256, k8_total), 256, 0, st>>>(value, value_ld, rows, n, d.rows_pad, k8_total,
                                                                            reinterpret_cast<uint4 *>(ap));
  pack_cols_kernel<NN><<<(unsigned)(d.k_pad / KP_KC) * KP_K8, NN, 0, st>>>(aff, aff_ld, n, hw, reinterpret_cast<uint4 *>(bp));
  KpArgs a{ap, bp, partial, g};
  const size_t smem = (size_t)KP_A_STAGES * 2 * KP_K8 * g.halo * 16 + (size_t)KP_B_STAGES * b_bytes<NN>() +
                      (2 * KP_A_STAGES + 2 * KP_B_STAGES + 1) * 8 + 16;
  VOSMEM_CUDA(cudaFuncSetAttribute(keyproj_mma_kernel<NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  keyproj_mma_kernel<NN, 1><<<dim3(d.rows_pad / KP_M, g.splits), KP_THREADS, smem, st>>>(a);
  sum_partials_kernel<<<(unsigned)ceil_div64((int64_t)rows * NN, 256), 256, 0, st>>>(partial, g.splits, d.rows_pad, NN, rows, hw,
                                                                                      out, out_ld);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

}  // namespace
}  // namespace vosmem

using namespace vosmem;

extern "C" int64_t vosmem_readout_dense_tc_workspace_bytes(int rows, int64_t n, int hw) {
  if (rows < 1 || n < 1 || hw < 1 || hw > 256 || n > ((int64_t)1 << 24)) return 0;
  return dense_geom(rows, n, hw).total;
}

extern "C" int vosmem_readout_dense_tc(const float *value, int64_t value_ld, const float *affinity, int64_t aff_ld, int rows,
                                       int64_t n, int hw, float *out, int64_t out_ld, void *workspace, int64_t workspace_bytes,
                                       vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(value && affinity && out && workspace, "vosmem_readout_dense_tc: null pointer");
  VOSMEM_CHECK_ARG(rows >= 1 && n >= 1 && hw >= 1 && hw <= 256, "vosmem_readout_dense_tc: rows=%d n=%lld hw=%d (hw <= 256)", rows,
                   (long long)n, hw);
  const DenseGeom d = dense_geom(rows, n, hw);
  if (workspace_bytes < d.total) {
    set_error("readout_dense_tc: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)d.total);
    return VOSMEM_ENOSPC;
  }
  unsigned char *ws = static_cast<unsigned char *>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  if (d.nn == 64) return run_dense_tc<64>(value, value_ld, affinity, aff_ld, rows, n, hw, out, out_ld, d, ws, st);
  if (d.nn == 128) return run_dense_tc<128>(value, value_ld, affinity, aff_ld, rows, n, hw, out, out_ld, d, ws, st);
  return run_dense_tc<256>(value, value_ld, affinity, aff_ld, rows, n, hw, out, out_ld, d, ws, st);
}

extern "C" int64_t vosmem_keyproj_weight_bytes(int in_dim, int key_dim) {
  if (key_dim != 64 || in_dim < KP_KC || in_dim % KP_KC != 0) return 0;
  return (int64_t)(in_dim / KP_KC) * KP_TAPS * KP_B_BYTES;
}

extern "C" int vosmem_keyproj_pack_weights(const float *key_w, const float *d_w, const float *e_w, int in_dim, int key_dim,
                                           void *packed, vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(key_w && d_w && e_w && packed, "vosmem_keyproj_pack_weights: null pointer");
  int rc = kp_check(in_dim, key_dim, 1, 1);
  if (rc != VOSMEM_OK) return rc;
  pack_w_kernel<<<(in_dim / KP_KC) * KP_TAPS * KP_K8, KP_N, 0, (cudaStream_t)stream>>>(key_w, d_w, e_w, in_dim, key_dim,
                                                                                     static_cast<uint4 *>(packed));
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}

extern "C" int64_t vosmem_keyproj_workspace_bytes(int in_dim, int key_dim, int h, int w) {
  if (kp_check(in_dim, key_dim, h, w) != VOSMEM_OK) return 0;
  return kp_carve(nullptr, kp_geom(in_dim, h, w)).bytes;
}

extern "C" int vosmem_keyproj_forward(const float *x, int in_dim, int key_dim, int h, int w, const void *packed_weights,
                                      const float *key_bias, const float *d_bias, const float *e_bias, float *key,
                                      float *shrinkage, float *selection, void *workspace, int64_t workspace_bytes,
                                      vosmem_stream_t stream) {
  VOSMEM_CHECK_ARG(x && packed_weights && key_bias && d_bias && e_bias && key && workspace, "vosmem_keyproj_forward: null pointer");
  int rc = kp_check(in_dim, key_dim, h, w);
  if (rc != VOSMEM_OK) return rc;
  const KpGeom g = kp_geom(in_dim, h, w);
  const KpWorkspace ws = kp_carve(workspace, g);
  if (workspace_bytes < ws.bytes) {
    set_error("keyproj: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)ws.bytes);
    return VOSMEM_ENOSPC;
  }
  cudaStream_t st = (cudaStream_t)stream;
  pack_x_kernel<<<dim3((g.q_rows + 255) / 256, in_dim / 8), 256, 0, st>>>(x, g, reinterpret_cast<uint4 *>(ws.xp));
  KpArgs a{ws.xp, static_cast<const unsigned char *>(packed_weights), ws.partial, g};
  const size_t smem = (size_t)KP_A_STAGES * 2 * KP_K8 * g.halo * 16 + (size_t)KP_B_STAGES * KP_B_BYTES +
                      (2 * KP_A_STAGES + 2 * KP_B_STAGES + 1) * 8 + 16;
  VOSMEM_CHECK_ARG(smem <= 232448, "keyproj: %zu bytes of shared memory for w=%d", smem, w);
  VOSMEM_CUDA(cudaFuncSetAttribute(keyproj_mma_kernel<KP_N, KP_TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  keyproj_mma_kernel<KP_N, KP_TAPS><<<dim3(g.p_pad / KP_M, g.splits), KP_THREADS, smem, st>>>(a);
  KpOut o{ws.partial, key_bias, d_bias, e_bias, key, shrinkage, selection, g, key_dim};
  keyproj_finalize_kernel<<<dim3(g.p_pad / 32, KP_N / KP_FIN_CH), 128, 0, st>>>(o);
  VOSMEM_CUDA(cudaGetLastError());
  return VOSMEM_OK;
}
