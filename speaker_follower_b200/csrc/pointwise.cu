// pointwise.cu — the small fused kernels around the GEMMs: action scoring (EltwiseProdScoring, model.py:342-352, in its re-associated form) and
// the per-step tail of the follower rollout (follower.py:476-505).
#include <cuda_bf16.h>

#include "kernels.h"
#include "tail.cuh"

namespace sfb {

// ---------------------------------------------------------------- action scoring
// w_out . ((W_h ht + b_h) (.) (W_a u + b_a)) + b_out  ==  u . g + c   with  tp = w_out (.) (W_h ht + b_h),
// g = W_a^T tp  and  c = sum_d b_a[d] tp[d] + b_out   (SURVEY.md §7 hard part 1).
// The A candidate rows of a batch element are step inputs: with PDL they are pulled into shared memory (bulk
// async copies) while the kernels producing g / t' are still running.
__global__ void __launch_bounds__(256) action_scoring_kernel(const ScoringParams p, const int stage_rows) {
  extern __shared__ __align__(128) unsigned char sraw[];
  float* gs = reinterpret_cast<float*>(sraw);                       // [E]
  float* us = gs + p.E;                                             // [A][E] when stage_rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(us + (stage_rows ? (size_t)p.A * p.E : 0));
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = p.E >> 2;
  trace_mark(p.trace, 0);
  pdl_launch_dependents();
  const bool gather = p.cand_table != nullptr;   // candidates = rows of the device-resident feature table + 4 angles
  if (stage_rows && tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    const uint64_t pol = policy_evict_first();
    if (!gather) {
      mbar_expect_tx(bar, (uint32_t)((size_t)p.A * p.E * 4));
      for (int a = 0; a < p.A; ++a)
        bulk_g2s_hint(us + (size_t)a * p.E, p.all_u_t + ((size_t)b * p.A + a) * p.E, (uint32_t)p.E * 4u, bar, pol);
    } else {
      int n = 0;
      for (int a = 0; a < p.A; ++a) n += p.cand_view[(size_t)b * p.A + a] >= 0;
      mbar_expect_tx(bar, (uint32_t)((size_t)n * p.img_dim * 4));
      const float* slab = p.cand_table + (size_t)p.vp_idx[b] * p.cand_V * p.img_dim;
      for (int a = 0; a < p.A; ++a) {
        const int v = p.cand_view[(size_t)b * p.A + a];
        if (v >= 0) bulk_g2s_hint(us + (size_t)a * p.E, slab + (size_t)v * p.img_dim, (uint32_t)p.img_dim * 4u, bar, pol);
      }
    }
  }
  if (gather) {
    // env.py:60-75: row a = [feature[absViewIndex_a, :img_dim], sin(rh) x n, cos(rh) x n, sin(re) x n, cos(re) x n],
    // n = (E - img_dim) / 4; rows without a view (the stop action, padding) are zero
    const int loc = p.E - p.img_dim, grp = loc >> 2;
    for (int i = tid; i < p.A * loc; i += 256) {
      const int a = i / loc, j = i - a * loc;
      const bool ok = p.cand_view[(size_t)b * p.A + a] >= 0;
      us[(size_t)a * p.E + p.img_dim + j] = ok ? p.cand_trig[((size_t)b * p.A + a) * 4 + j / grp] : 0.f;
    }
    for (int a = 0; a < p.A; ++a)
      if (p.cand_view[(size_t)b * p.A + a] < 0)
        for (int i = tid; i < p.img_dim; i += 256) us[(size_t)a * p.E + i] = 0.f;
  }
  pdl_wait();
  trace_mark(p.trace, 1);
  const float4* g4 = reinterpret_cast<const float4*>(p.g + (size_t)b * p.ldg);
  for (int j = tid; j < nvec; j += 256) reinterpret_cast<float4*>(gs)[j] = g4[j];
  float c = 0.f;
  if (p.tp)
    for (int d = tid; d < p.D; d += 256) c = fmaf(__ldg(p.b_a + d), p.tp[(size_t)b * p.D + d], c);
  c = warp_sum(c);
  if (lane == 0) red[warp] = c;
  __syncthreads();
  float cst = p.tp ? __ldg(p.b_out) : p.g[(size_t)b * p.ldg + p.E];   // folded weights carry the constant as g[E]
#pragma unroll
  for (int w = 0; w < 8; ++w) cst += red[w];
  if (stage_rows) mbar_wait(bar, 0);
  for (int a = warp; a < p.A; a += 8) {
    const float4* u4 = stage_rows ? reinterpret_cast<const float4*>(us + (size_t)a * p.E)
                                  : reinterpret_cast<const float4*>(p.all_u_t + ((size_t)b * p.A + a) * p.E);
    float acc = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      const float4 u = u4[j];
      const float4 g = reinterpret_cast<const float4*>(gs)[j];
      acc = fmaf(u.x, g.x, acc);
      acc = fmaf(u.y, g.y, acc);
      acc = fmaf(u.z, g.z, acc);
      acc = fmaf(u.w, g.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) p.logit[(size_t)b * p.A + a] = acc + cst;
  }
  __syncthreads();
  if (p.has_tail && warp == 0)   // logits of this row were written by this CTA
    tail_row(p.tail, b, lane, stage_rows ? us : p.all_u_t + (size_t)b * p.A * p.E);
  trace_mark(p.trace, 2);
}

int32_t launch_action_scoring(const ScoringParams& p_in, cudaStream_t stream) {
  ScoringParams p = p_in;
  p.trace = next_trace_slot();
  SFB_CHECK_ARG((p.E % 4) == 0, "scoring: E % 4");
  if (p.ldg == 0) p.ldg = p.E;
  SFB_CHECK_ARG((p.ldg % 4) == 0, "scoring: ldg % 4");
  const size_t staged = ((size_t)p.A + 1) * p.E * sizeof(float) + 16;
  const int stage_rows = staged <= 160 * 1024 ? 1 : 0;
  if (p.cand_table) {
    SFB_CHECK_ARG(stage_rows, "scoring: too many action candidates for the gather source (A * E * 4 must fit 160 KB)");
    SFB_CHECK_ARG(p.vp_idx && p.cand_view && p.cand_trig && p.img_dim > 0 && p.img_dim < p.E && (p.img_dim % 4) == 0 &&
                      ((p.E - p.img_dim) % 4) == 0 && p.cand_V > 0, "scoring: bad gather source");
  } else {
    SFB_CHECK_ARG(p.all_u_t, "scoring: all_u_t is NULL");
  }
  const size_t smem = stage_rows ? staged : (size_t)p.E * sizeof(float) + 16;
  SFB_CHECK_ARG(smem <= 200 * 1024, "scoring: E too large");
  static SmemMarks marks;
  SFB_CHECK_CUDA(ensure_dynamic_smem(action_scoring_kernel, smem, marks));
  SFB_CHECK_CUDA(launch_ex(action_scoring_kernel, dim3(p.B, 1, 1), dim3(256, 1, 1), smem, stream, dim3(1, 1, 1), p, stage_rows));
  count_launch();
  return 0;
}

// ---------------------------------------------------------------- table-driven navigation environment
// One thread per batch row: apply the previous action to the row's discretised world state (viewpoint, heading bin)
// through the precomputed transition table, then emit everything the next decode step needs from the observation
// tables — what R2RBatch.step + R2RBatch.observe (env.py:628-641, 763-804) produce per row, as pure table look-ups.
__global__ void nav_step_kernel(const NavStepParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  int s = p.state[b];
  int ended = p.ended[b];
  if (p.a_prev) {
    const int a = p.a_prev[b];
    if (p.actions_log) p.actions_log[b] = ended ? -1 : a;
    if (!ended) {
      if (a > 0 && a < p.A) s = p.next[(size_t)s * p.A + a];
      if (a == 0) ended = 1;        // follower.py:531-533: the row ends after its stop action has been recorded
    }
    p.state[b] = s;
    p.ended[b] = ended;
  }
  p.vp_idx[b] = p.vp[s];
  p.view_idx[b] = p.view[s];
  const int nv = p.nvalid[s];
  for (int a = 0; a < p.A; ++a) {
    p.cand_view[(size_t)b * p.A + a] = p.cv[(size_t)s * p.A + a];
    p.is_valid[(size_t)b * p.A + a] = a < nv ? 1.f : 0.f;
    const float4 t = reinterpret_cast<const float4*>(p.trig)[(size_t)s * p.A + a];
    reinterpret_cast<float4*>(p.cand_trig)[(size_t)b * p.A + a] = t;
  }
  if (p.target) p.target[b] = (ended || !p.teach) ? -1 : p.teach[(size_t)s * p.G + p.goal[b]];
}

int32_t launch_nav_step(const NavStepParams& p, cudaStream_t stream) {
  SFB_CHECK_CUDA(launch_ex(nav_step_kernel, dim3((p.B + 127) / 128, 1, 1), dim3(128, 1, 1), 0, stream, dim3(1, 1, 1), p));
  count_launch();
  return 0;
}

// ---------------------------------------------------------------- state-factored search bookkeeping
// follower.py:886-924 for successor_size = 1 over the table-driven environment, one CTA per instance.  The world-state
// key (scan, viewpoint, heading, elevation) is the discretised state id, so the reference's per-instance dicts are dense
// arrays over the S states: `cache` (best open inference state per world state), `holding` (best finished one),
// `completed`.  Inference states live in a per-instance node pool (parent, state, action, count, score, and the slot of
// the expansion whose (h, c, alpha) they carry).  Per iteration: the successors of the state expanded in this iteration
// are inserted where they strictly improve their table entry, then the best not-yet-expanded entry is selected: an open
// one becomes the next iteration's state, a finished one moves to `completed`.
__global__ void __launch_bounds__(128) sf_search_update_kernel(const SfSearchParams p) {
  const int i = blockIdx.x, tid = threadIdx.x;
  if (p.flags[0]) return;                                   // the search has ended (no instance had a state to expand)
  // iteration index: given by the host, or (iter < 0: the launch is replayed from a CUDA graph) the device's own count
  const int iter = p.iter >= 0 ? p.iter : p.flags[3];
  if (iter >= p.max_iter) { if (tid == 0) p.flags[2] = 1; return; }
  const size_t so = (size_t)i * p.S, no = (size_t)i * p.M;
  __shared__ float s_best[128];
  __shared__ int s_arg[128];
  const int n = p.beam_node[i];
  const bool full = p.n_done[i] >= p.completion_size;      // follower.py:889-891: nothing is inserted or selected any more
  if (tid == 0 && n >= 0 && !full) {
    const int s = p.node_state[no + n];
    const int nv = p.nav_nvalid[s];
    const float base = p.node_score[no + n];
    const int cnt = p.node_count[no + n] + 1;
    for (int a = 0; a < nv && a < p.A; ++a) {
      const float sc = base + p.lp[(size_t)i * p.A + a];   // np.float32(st.score + lp) (follower.py:851)
      const int ns = a == 0 ? s : p.nav_next[(size_t)s * p.A + a];
      const bool fin = a == 0 || cnt == p.episode_len;     // follower.py:895
      float* tsc = fin ? p.h_score : p.c_score;
      if (tsc[so + ns] < sc) {                             // absent (-inf) or strictly better (follower.py:896-900)
        int m = p.n_nodes[i];
        if (m >= p.M) { p.flags[2] = 1; break; }           // node pool exhausted: reported, the host falls back
        p.n_nodes[i] = m + 1;
        p.node_parent[no + m] = n; p.node_state[no + m] = ns; p.node_action[no + m] = a; p.node_count[no + m] = cnt;
        p.node_score[no + m] = sc; p.node_slot[no + m] = iter + 1;     // this iteration's outputs sit in slot iter + 1
        tsc[so + ns] = sc;
        (fin ? p.h_node : p.c_node)[so + ns] = m;
        (fin ? p.h_exp : p.c_exp)[so + ns] = 0;
      }
    }
  }
  __syncthreads();
  // best not-yet-expanded entry over cache and holding (heapq.nlargest(1, ...), follower.py:903-908); ties: open before
  // finished, then the lower state id
  float best = -INFINITY;
  int arg = -1;                                            // state id, + S for a holding entry
  if (!full) {
    for (int s = tid; s < p.S; s += 128) {
      const float c = p.c_exp[so + s] ? -INFINITY : p.c_score[so + s];
      if (c > best) { best = c; arg = s; }
    }
    for (int s = tid; s < p.S; s += 128) {
      const float h = p.h_exp[so + s] ? -INFINITY : p.h_score[so + s];
      if (h > best) { best = h; arg = p.S + s; }
    }
  }
  s_best[tid] = best; s_arg[tid] = arg;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) {
      const float ob = s_best[tid + o];
      const int oa = s_arg[tid + o];
      if (oa >= 0 && (s_arg[tid] < 0 || ob > s_best[tid] || (ob == s_best[tid] && oa < s_arg[tid]))) { s_best[tid] = ob; s_arg[tid] = oa; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    int beam = -1;
    const int a = s_arg[0];
    if (!full && a >= 0) {
      if (a >= p.S) {                                      // a finished state: mark expanded, move to completed (912-916)
        const int s = a - p.S;
        p.h_exp[so + s] = 1;
        if (p.d_score[so + s] < s_best[0]) {
          if (p.d_score[so + s] == -INFINITY) p.n_done[i] += 1;
          p.d_score[so + s] = s_best[0];
          p.d_node[so + s] = p.h_node[so + s];
        }
      } else {                                             // an open state: it is expanded in the next iteration
        p.c_exp[so + a] = 1;
        beam = p.c_node[so + a];
      }
    }
    if (p.n_done[i] >= p.completion_size) beam = -1;       // follower.py:921
    p.beam_node[i] = beam;
    p.trav[(size_t)i * p.max_iter + iter] = beam;
    if (beam >= 0) atomicAdd(&p.flags[1], 1);              // flags[1]: instances with a state to expand (reset by the flags kernel)
  }
}

// follower.py:925-926 `if not any(beams): break`, evaluated on the device so that the host need not look every iteration
__global__ void sf_search_flags_kernel(int* flags) {
  if (threadIdx.x == 0) {
    if (!flags[0]) {
      flags[3] += 1;                                        // iterations the search really ran
      if (flags[1] == 0) flags[0] = 1;
    }
    flags[1] = 0;
  }
}

int32_t launch_sf_search_update(const SfSearchParams& p, cudaStream_t stream) {
  SFB_CHECK_ARG(p.B >= 1 && p.A >= 1 && p.S >= 1 && p.M >= 1 && p.iter < p.max_iter, "sf_search_update: bad sizes");
  sf_search_update_kernel<<<p.B, 128, 0, stream>>>(p);
  SFB_CHECK_LAUNCH();
  sf_search_flags_kernel<<<1, 32, 0, stream>>>(p.flags);
  SFB_CHECK_LAUNCH();
  count_launch();
  count_launch();
  return 0;
}

// ---------------------------------------------------------------- follower rollout tail (one warp per row)
__global__ void __launch_bounds__(128) follower_tail_kernel(const TailParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  trace_mark(p.trace, 0);
  pdl_launch_dependents();
  pdl_wait();
  trace_mark(p.trace, 1);
  if (b >= p.B) return;
  tail_row(p, b, lane, p.all_u_t + (size_t)b * p.A * p.E);
  __syncwarp();
  trace_mark(p.trace, 2);
}

int32_t launch_follower_tail(const TailParams& p_in, cudaStream_t stream) {
  TailParams p = p_in;
  p.trace = next_trace_slot();
  SFB_CHECK_CUDA(launch_ex(follower_tail_kernel, dim3((p.B + 3) / 4, 1, 1), dim3(128, 1, 1), 0, stream, dim3(1, 1, 1), p));
  count_launch();
  return 0;
}

}  // namespace sfb
