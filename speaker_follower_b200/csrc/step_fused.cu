// step_fused.cu — the first half of a decode step as ONE launch: 36-view attention gather + LSTM cell.
//
//   feature_b = sum_v softmax_v(V_b[v] . q_b) V_b[v]                       (VisualSoftDotAttention, model.py:310-326)
//   gates     = W_ih [u_prev ; feature] + W_hh h_0 (+ biases) -> (h_1, c_1) (nn.LSTMCell, model.py:393)
//
// One resident wave of tiles x S CTAs (16 x 9 = 144 at H = 512: one per SM), every CTA carrying two roles:
//   * GEMM role (all CTAs): CTA (tile, rank) owns 128 gate rows (4 gates x 32 hidden units, interleaved at pack time)
//     and a 1/S slice of K.  Warp 9 feeds GEMM stages with cp.async.bulk (weights + packed activations), warp 8
//     issues tcgen05.mma (bf16 x 3, fp32 accumulation in TMEM).  The K blocks that multiply u_prev and h_0 ("pre" part)
//     depend on nothing computed in this step: they are consumed WHILE the gather streams.  The blocks that multiply
//     the attention output ("post" part) have their weights prefetched and start the moment the gather CTAs have
//     signalled a device-wide counter.
//   * gather role (CTAs 0..B-1): one CTA owns ONE batch element, so there is no cross-CTA softmax merge at all.
//     Warp 10 streams the element's slab (36 contiguous rows of the feature table) as a ring of multi-row bulk copies
//     (one cp.async.bulk per RB rows; the orientation rows arrive once as one extra copy), warps 0-7 consume RB rows
//     per block barrier (block-wide dot products, online softmax, running weighted sum) and finally write the
//     normalised feature as fp32 and as bf16 (hi, lo) straight into the packed activation operand of the post part.
// Shared memory is time-multiplexed: while the gather runs, the GEMM role owns one stage and the ring owns the rest;
// when the CTA's gather is done the ring becomes two more GEMM stages for the post part.
// Afterwards all CTAs run the split-K reduction (partials through L2, one sense-reversing barrier per tile) and the
// LSTM cell update, which also emits h_1 in packed form for the NEXT step's pre part.
#include "step_fused_a.cuh"

namespace sfb {

// grid = (tiles, S), FNT threads.  Dynamic smem: [GEMM stage 0 | ring (later GEMM stages 1, 2) | orientation rows | barriers | scratch]
__global__ void __launch_bounds__(FNT, 1) vis_lstm_fused_kernel(const FusedVisLstmParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint32_t tmem_unused;
  vis_lstm_body<false>(q, smem, (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, (int)gridDim.y, NoStepHooks{}, nullptr, tmem_unused);
}

// ------------------------------------------------------------------ host side

FusedPlan vis_lstm_fused_plan(int B, int H, int nkb, int R, int D, int lenA, int lenB, int num_sms) {
  FusedPlan pl{};
  pl.ok = false;
  if (H < 32 || (H % 32) != 0 || B < 1 || B > 128 || D != F_D || R < 1 || R > 256) return pl;
  if (lenA + lenB != D || !((lenA == 2048 && lenB == 128) || (lenA == D && lenB == 0))) return pl;   // the reference's slab layouts
  pl.tiles = H / 32;
  pl.S = num_sms / pl.tiles;
  if (pl.S > 16) pl.S = 16;
  if (pl.S > nkb) pl.S = nkb;
  if (pl.S < 1 || pl.tiles * pl.S < B) return pl;           // every batch element needs its own CTA
  pl.NB = (B + 15) & ~15;
  const size_t stage = 2 * (size_t)FA_HALF + 2 * (size_t)(pl.NB / 8) * FSBO;
  const size_t chunk = (size_t)F_RB * lenA * 4;
  const size_t fixed = stage + (size_t)R * lenB * 4 + (2 * FGS + 1 + 2 * F_MAXCH + 2) * sizeof(uint64_t) +
                       (2 * 8 * 8 + ((R + 3) & ~3)) * sizeof(float) + 96;
  const size_t budget = 227 * 1024 - 2048;                  // static shared + alignment slack
  if (fixed + 2 * chunk > budget) return pl;
  int nch = (int)((budget - fixed) / chunk);
  const int nchunks = (R + F_RB - 1) / F_RB;
  if (nch > nchunks) nch = nchunks;
  if (nch > F_MAXCH) nch = F_MAXCH;
  pl.nch = nch;
  pl.chunk_rows = F_RB;
  pl.gstages = 1 + (int)(((size_t)nch * chunk) / stage);
  if (pl.gstages > FGS) pl.gstages = FGS;
  pl.smem = fixed + (size_t)nch * chunk;
  pl.data_bytes = stage + (size_t)nch * chunk + (size_t)R * lenB * 4;
  pl.sem_bytes = 4096;
  pl.bytes = pl.sem_bytes + (((size_t)pl.tiles * pl.S * FBM * pl.NB * sizeof(float) + 255) & ~size_t(255));
  pl.ok = (size_t)(2 * pl.tiles + 16 + 4) * sizeof(unsigned int) <= pl.sem_bytes;
  return pl;
}

int32_t launch_vis_lstm_fused(const FusedVisLstmParams& q_in, cudaStream_t stream, void* ws, size_t ws_bytes) {
  FusedVisLstmParams q = q_in;
  GemmParams& p = q.g;
  p.trace = next_trace_slot();
  const int H = p.lstm.H;
  const FusedPlan pl = vis_lstm_fused_plan(q.B, H, q.nkb, q.R, q.D, q.lenA, q.lenB, device_num_sms());
  SFB_CHECK_ARG(pl.ok, "fused gather + LSTM step: unsupported shape");
  SFB_CHECK_ARG(q.a_pk && q.b_pk && (reinterpret_cast<uintptr_t>(q.a_pk) & 127u) == 0 && (reinterpret_cast<uintptr_t>(q.b_pk) & 127u) == 0,
                "fused step: packed operands missing / misaligned");
  SFB_CHECK_ARG(q.post_kb0 >= 0 && q.post_kb1 <= q.nkb && q.post_kb0 <= q.feat_kb0 && q.feat_kb0 + (q.D + FBK - 1) / FBK <= q.post_kb1,
                "fused step: the post K range must cover the feature blocks");
  SFB_CHECK_ARG(q.q && q.segA && (q.lenB == 0 || q.segB) && (reinterpret_cast<uintptr_t>(q.segA) & 15u) == 0 &&
                    (q.strideA_b % 4) == 0 && (q.lenB == 0 || ((reinterpret_cast<uintptr_t>(q.segB) & 15u) == 0 && (q.strideB_b % 4) == 0)),
                "fused step: visual source missing / misaligned");
  SFB_CHECK_ARG(ws && ws_bytes >= pl.bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "fused step: workspace");
  // barrier words: [0, 2*tiles) split-K {count, generation} pairs; the gather counter block sits behind them
  q.sem = static_cast<unsigned int*>(ws);
  q.sync = q.sem + 2 * pl.tiles + 16;
  q.partial = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.sem_bytes);
  q.NB = pl.NB;
  q.nch = pl.nch;
  q.chunk_rows = pl.chunk_rows;
  p.M = q.B;
  static SmemMarks marks;
  SFB_CHECK_CUDA(ensure_dynamic_smem(vis_lstm_fused_kernel, pl.smem, marks));
  SFB_CHECK_CUDA(launch_ex(vis_lstm_fused_kernel, dim3(pl.tiles, pl.S, 1), dim3(FNT, 1, 1), pl.smem, stream, dim3(1, 1, 1), q));
  count_launch();
  return 0;
}

}  // namespace sfb
