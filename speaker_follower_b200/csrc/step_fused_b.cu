// step_fused_b.cu — the second half of a follower decode step as ONE launch (model.py:394-396 + follower.py:476-505):
//
//   hh   = W_out_h h1d                     q' = M_q h_1 + b_q   (next step's visual query, off the critical path)
//   alpha = softmax_l(ctx_k[b,l] . h1d[b]) (masked),  h~ = tanh(sum_l alpha_l ctx_o[b,l] + hh[b])     (SoftDotAttention
//           with the per-episode key / value projections of ctx, see sfb_follower_project_ctx)
//   g    = M_g h~ + b_g                    logit[b,a] = u_{b,a} . g[b] + g[b][E]    (EltwiseProdScoring, folded)
//   rollout tail of the row (mask, log-softmax, teacher / argmax / sample, next-u gather, score / CE terms)
//
// One resident wave of CTAs (one per SM) in clusters of two, with two roles:
//   * projection pairs (the first 2P CTAs): a pair owns one 128-row weight tile per job and splits K between its two
//     CTAs.  Warp 9 pulls the job's packed weights (cp.async.bulk) as soon as the buffer is free — for the g projection
//     that is while the attention is still running — and streams the packed activation blocks through a 2-stage ring;
//     warp 8 issues tcgen05.mma (bf16 x 3, fp32 accumulation in TMEM); the two partial accumulators are exchanged
//     through DISTRIBUTED SHARED MEMORY (each CTA parks the columns its peer owns, one cluster barrier, pull) — no
//     global partial tiles, no device-wide split-K barrier.
//   * row CTAs (one per batch element): warp 9 streams the element's un-masked key / value rows into a ring before
//     the dependency wait (they are per-episode constants); each of the 8 compute warps owns whole rows (warp-shuffle
//     dot product, online softmax, no block barrier while streaming); the 8 partial results merge through shared
//     memory; h~ is formed once hh has arrived and is published both as fp32 and as the packed operand of the g
//     projection.  The same CTA then stages the element's action-candidate rows (from the feature table or the dense
//     tensor) in the freed ring and, when g has arrived, forms the logits and runs the rollout tail.
// The three hand-offs (hh -> rows, h~ -> g pairs, g -> rows) are device-wide arrival counters polled by one thread.
//
// step_kernel (bottom of this file) runs BOTH halves of the step — the gather + gate GEMM + LSTM cell of step_fused.cu
// and the above — as ONE launch over one resident wave of CTAs: the second half starts in every CTA as soon as the
// CTA's own share of the first half is finished, its input-independent reads (key / value rows, projection weights)
// are issued by an otherwise idle warp while the first half's epilogue is still running, and the hand-over between
// the halves is one more device-wide arrival counter instead of a kernel boundary.
#include "step_fused_a.cuh"
#include "tail.cuh"

namespace sfb {

namespace {
constexpr int TBM = 128, TBK = 64;
constexpr int TNT = 320;                         // 8 compute warps + MMA issuer + producer
constexpr uint32_t TCORE = 128;
constexpr uint32_t TSBO = (TBK / 8) * TCORE;
constexpr uint32_t TLBO = TCORE;
constexpr uint32_t TA_HALF = (TBM / 8) * TSBO;   // 16 KB: one (hi | lo) weight tile block
constexpr int T_RW = 8;                          // rows per ring chunk = compute warps
constexpr int T_MAXCH = 8;                       // ring depth in chunks
constexpr int T_MAXKH = 4;                       // K blocks per CTA of a pair (K <= 512)
constexpr int T_BST = 3;                         // activation stages of a projection CTA

__device__ __forceinline__ bool t_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    __nanosleep(32);   // a hot try_wait loop competes with the compute warps for the shared-memory pipe
  }
  return false;
}
__device__ __forceinline__ uint64_t t_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(TLBO >> 4) << 16;
  d |= (uint64_t)(TSBO >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void t_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void t_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void t_bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned int t_ld_acq(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool t_spin_ge(const unsigned int* p, unsigned int target) {
  for (uint32_t i = 0; i < (1u << 22); ++i) {   // one poller per CTA, backing off: the arrivals must not be starved
    if (t_ld_acq(p) >= target) return true;
    __nanosleep(40);
  }
  return false;
}
}  // namespace

// One CTA of the kernel.  Stand-alone (MERGED = false): STAGE = T_ALL does everything.  Inside the one-launch step
// kernel the same code is entered four times: T_INIT (all threads, before the first half: barriers), T_LISTS (producer
// warp, early: the un-masked positions), T_PREFETCH (producer warp, when the first half has released the data region:
// the first key / value rows or the first job's weights), T_MAIN (all threads, after the first half).
// `top` = barriers + position lists: in the merged kernel a region the first half never touches.
enum { T_ALL = 0, T_INIT = 1, T_LISTS = 2, T_PREFETCH = 3, T_MAIN = 4 };
struct TShared {   // static shared state of a CTA (declared by the kernel: one copy whatever the number of stages)
  int fail, pre, nvalid[2], at;
  uint32_t tmem;
};
template <bool MERGED, int STAGE>
__device__ __forceinline__ void text_score_body(const FusedTextScoreParams& q, unsigned char* smem, unsigned char* top_in,
                                                TShared& sh, const uint32_t tmem_in, const unsigned int* phase) {
  int& s_fail = sh.fail;
  uint32_t& s_tmem = sh.tmem;
  int* s_nvalid = sh.nvalid;
  int& s_at = sh.at;
  int& s_pre = sh.pre;                    // merged: key / value chunks (row CTA) or weight sets (pair CTA) issued by T_PREFETCH
  constexpr int PW = MERGED ? 10 : 9;     // producer warp (the merged kernel's warp 9 belongs to the first half)
  constexpr bool DO_INIT = STAGE == T_ALL || STAGE == T_INIT;
  constexpr bool DO_LISTS = STAGE == T_ALL || STAGE == T_LISTS;
  constexpr bool DO_MAIN = STAGE == T_ALL || STAGE == T_MAIN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cid = blockIdx.x;
  const int P = q.P, B = q.B, H = q.H, NB = q.NB;
  const bool is_pair = cid < 2 * P;
  unsigned int* cnt_hh = q.sync;         // arrivals: 2 per hh tile
  unsigned int* cnt_ht = q.sync + 1;     // arrivals: 1 per batch element (h~ published)
  unsigned int* cnt_g = q.sync + 2;      // arrivals: 2 per g tile
  unsigned int* cnt_exit = q.sync + 3;
  unsigned int* status = q.sync + 4;
  // merged: "the first half is complete device-wide" replaces the dependency on the preceding kernel
  auto wait_inputs = [&]() -> bool {
    if (!MERGED) {
      pdl_wait();
      return true;
    }
    return t_spin_ge(phase, gridDim.x);
  };

  if (DO_INIT) {
    if (tid == 0) {
      s_fail = 0;
      s_pre = 0;
    }
  }
  if (STAGE == T_ALL) {
    trace_mark(q.trace, 0);
    pdl_launch_dependents();
  }
  if (STAGE == T_MAIN) trace_mark(q.trace, 0);

  if (is_pair) {
    // =====================================================================================================
    // projection pair
    // =====================================================================================================
    const int pair = cid >> 1, rank = cid & 1;
    const uint32_t b_half = (uint32_t)(NB / 8) * TSBO;
    const uint32_t bstage = 2 * b_half;
    unsigned char* wbuf = smem;                                             // [T_MAXKH][32 KB]; later the parked partial
    unsigned char* bst = smem + (size_t)T_MAXKH * 2 * TA_HALF;              // [T_BST][bstage]
    uint64_t* wfull = reinterpret_cast<uint64_t*>(MERGED ? top_in : bst + (size_t)T_BST * bstage);
    uint64_t* bfull = wfull + 1;      // [T_BST]
    uint64_t* bempty = bfull + T_BST; // [T_BST]
    uint64_t* done = bempty + T_BST;
    uint64_t* pdone = done + 1;       // the PEER's MMAs of the job have retired (remote arrival): its weight buffer may be written
    const uint32_t tmem_cols = NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256;
    if (DO_INIT && warp == (MERGED ? 1 : 0)) {
      if (lane == 0) {
        mbar_init(wfull, 1);
        for (int i = 0; i < T_BST; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        mbar_init(done, 1);
        mbar_init(pdone, 1);
      }
      mbar_fence_init();
      __syncwarp();
      if (!MERGED) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
      }
    }
    if (STAGE == T_ALL) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t tmem_d = MERGED ? tmem_in : s_tmem;

    // jobs of this pair, in order: phase 1 = hh tiles then q' tiles (inputs ready at kernel start), phase 2 = g tiles
    const int n1 = q.hh_tiles + q.q_tiles, n2 = q.g_tiles;
    const int nkb = q.nkb, kh0 = rank == 0 ? 0 : (nkb + 1) / 2, kh1 = rank == 0 ? (nkb + 1) / 2 : nkb, nkh = kh1 - kh0;
    // the weights of a job: all K blocks of this CTA's half into the weight buffer, one barrier phase
    auto issue_weights = [&](const unsigned char* a_base, int tile) {
      const uint64_t pol = policy_evict_last();
      mbar_expect_tx(wfull, (uint32_t)nkh * 2 * TA_HALF);
      for (int k = 0; k < nkh; ++k)
        bulk_g2s_hint(wbuf + (size_t)k * 2 * TA_HALF, a_base + ((size_t)tile * nkb + kh0 + k) * (2 * TA_HALF), 2 * TA_HALF, wfull, pol);
    };
    if (STAGE == T_PREFETCH) {   // the first job's weights (a pair always has a phase-1 job when it has any)
      if (lane == 0 && pair < n1) {
        if (pair < q.hh_tiles) issue_weights(q.a_hh, pair);
        else issue_weights(q.a_q, pair - q.hh_tiles);
        s_pre = 1;
      }
    }
    if (DO_MAIN) {
    // batch columns whose epilogue this CTA runs: rank 0 the first groups of 16, rank 1 the rest
    const int ngrp = NB / 16, g_split = (ngrp + 1) / 2;
    const int my_g0 = rank == 0 ? 0 : g_split, my_g1 = rank == 0 ? g_split : ngrp;
    const int peer_g0 = rank == 0 ? g_split : 0, peer_g1 = rank == 0 ? ngrp : g_split;
    int njobs_done = 0, bcount = 0;   // running counters -> mbarrier parities
    bool waited_pdl = false, fail = false;

    for (int jp = 1; jp <= 2; ++jp) {   // job phase
      const int njobs = jp == 1 ? n1 : n2;
      for (int job = pair; job < njobs; job += P) {
        // ---- job description
        const unsigned char* a_base; const unsigned char* b_base; float* out; int ldo, ncols, tile; const float* bias; unsigned int* cnt;
        if (jp == 1 && job < q.hh_tiles) {
          tile = job; a_base = q.a_hh; b_base = q.hdpk; out = q.hh; ldo = q.ldhh; ncols = H; bias = nullptr; cnt = cnt_hh;
        } else if (jp == 1) {
          tile = job - q.hh_tiles; a_base = q.a_q; b_base = q.hpk; out = q.q_next; ldo = q.ldq; ncols = q.q_cols; bias = q.b_q; cnt = nullptr;
        } else {
          tile = job; a_base = q.a_g; b_base = q.htpk; out = q.g; ldo = q.ldg; ncols = q.g_cols; bias = q.b_g; cnt = cnt_g;
        }
        const uint32_t jpar = (uint32_t)njobs_done & 1u;
        if (warp == PW) {
          if (lane == 0) {
            // weights of the job: the buffer is free (the end-of-job cluster barrier of the previous job was passed)
            if (!(MERGED && njobs_done == 0 && jp == 1 && s_pre)) issue_weights(a_base, tile);
            if (!waited_pdl) {
              fail = !wait_inputs() || fail;   // packed h1d / h_1 come from the first half of the step
              if (MERGED) asm volatile("fence.proxy.async;" ::: "memory");   // ... written with generic-proxy stores in this launch
              waited_pdl = true;
              if (q.trace && cid == 0) q.trace[11] = globaltimer_ns();
            }
            if (jp == 2) {   // the packed h~ operand is complete when every row CTA has published its element
              fail = !t_spin_ge(cnt_ht, (unsigned int)B) || fail;
              asm volatile("fence.proxy.async;" ::: "memory");
            }
            for (int k = 0; k < nkh; ++k, ++bcount) {
              const int s = bcount % T_BST;
              if (bcount >= T_BST) fail = !t_wait(&bempty[s], (uint32_t)(bcount / T_BST - 1) & 1u) || fail;
              mbar_expect_tx(&bfull[s], bstage);
              bulk_g2s(bst + (size_t)s * bstage, b_base + (size_t)(kh0 + k) * bstage, bstage, &bfull[s]);
            }
          } else {
            bcount += nkh;
          }
        } else if (warp == 8) {
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
          fail = !t_wait(wfull, jpar) || fail;
          for (int k = 0; k < nkh; ++k, ++bcount) {
            const int s = bcount % T_BST;
            fail = !t_wait(&bfull[s], (uint32_t)(bcount / T_BST) & 1u) || fail;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
              const uint32_t a_hi = smem_u32(wbuf + (size_t)k * 2 * TA_HALF), a_lo = a_hi + TA_HALF;
              const uint32_t b_hi = smem_u32(bst + (size_t)s * bstage), b_lo = b_hi + b_half;
#pragma unroll
              for (int j = 0; j < TBK / 16; ++j) {
                const uint32_t ko = (uint32_t)j * 2u * TCORE;
                const uint64_t dah = t_desc(a_hi + ko), dal = t_desc(a_lo + ko);
                const uint64_t dbh = t_desc(b_hi + ko), dbl = t_desc(b_lo + ko);
                t_umma(tmem_d, dal, dbh, idesc, (k > 0 || j > 0) ? 1u : 0u);
                t_umma(tmem_d, dah, dbl, idesc, 1u);
                t_umma(tmem_d, dah, dbh, idesc, 1u);
              }
              t_commit(&bempty[s]);
              if (k + 1 == nkh) t_commit(done);
            }
            __syncwarp();
          }
        } else if (warp >= 8) {
          bcount += nkh;   // merged: warps of the first half without a role here
        } else {
          bcount += nkh;
          // ---- epilogue part 1: PUSH the columns the peer owns into the peer's shared memory (TMEM -> registers ->
          // st.shared::cluster, [col][row] over the peer's weight buffer, which is free once the peer's MMAs have retired)
          fail = !t_wait(done, jpar) || fail;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (q.trace && cid == 0 && tid == 0 && njobs_done == 0) q.trace[12] = globaltimer_ns();
          if (tid == 0) mbar_arrive_remote(dsmem_addr(pdone, (uint32_t)(rank ^ 1)));
          fail = !t_wait(pdone, jpar) || fail;
          const int lq = warp & 3, hf = warp >> 2, row = lq * 32 + lane;
          const uint32_t peer_park = dsmem_addr(wbuf, (uint32_t)(rank ^ 1));
          for (int gq = peer_g0 + hf; gq < peer_g1; gq += 2) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(gq * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j)
              dsmem_st_f32(peer_park + (uint32_t)(((gq - peer_g0) * 16 + j) * TBM + row) * 4u, __uint_as_float(v[j]));
          }
        }
        cluster_sync_all();   // both partials delivered (release / acquire at cluster scope)
        if (q.trace && cid == 0 && tid == 0 && njobs_done == 0) q.trace[13] = globaltimer_ns();
        if (warp < 8) {
          // ---- epilogue part 2: own columns = own accumulator + the peer's parked values, bias, store
          const int lq = warp & 3, hf = warp >> 2, row = lq * 32 + lane;
          const float* park = reinterpret_cast<const float*>(wbuf);   // what the peer pushed: this CTA's columns of ITS partial
          const int n = tile * TBM + row;
          const float bv = (bias && n < ncols) ? __ldg(bias + n) : 0.f;
          for (int gq = my_g0 + hf; gq < my_g1; gq += 2) {
            if (gq * 16 >= B) break;
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(gq * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            float pv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pv[j] = park[(size_t)((gq - my_g0) * 16 + j) * TBM + row];
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n < ncols) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int m = gq * 16 + j;
                // rank 0 holds the low K half: fixed summation order (low + high) on both sides
                const float lo = rank == 0 ? __uint_as_float(v[j]) : pv[j], hi = rank == 0 ? pv[j] : __uint_as_float(v[j]);
                if (m < B) out[(size_t)m * ldo + n] = lo + hi + bv;
              }
            }
          }
          if (q.trace && cid == 0 && tid == 0 && njobs_done == 0) q.trace[14] = globaltimer_ns();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cluster_sync_all();   // outputs stored, parked data consumed: the weight buffer is free again
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && cnt) {   // the CTA's stores are ordered before this thread by the barrier: one device-scope fence, then arrive
          __threadfence();
          atomicAdd(cnt, 1u);
        }
        if (q.trace && cid == 0 && tid == 0 && njobs_done == 0) q.trace[15] = globaltimer_ns();
        ++njobs_done;
      }
    }
    if (!MERGED && warp == PW && lane == 0 && !waited_pdl) pdl_wait();
    if (fail) s_fail = 1;
    __syncthreads();
    if (!MERGED && warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }   // DO_MAIN
  } else if (((cid - 2 * P) >> 1) < (B + 1) / 2) {
    // =====================================================================================================
    // row CTAs: a cluster of two serves TWO batch elements — a long one and a short one (element c and B-1-c: the
    // batch is sorted by instruction length, so the pair streams about the same number of rows as every other pair).
    // CTA `rank` streams every other un-masked row of BOTH elements; each CTA then finalises one element (rank 0 the
    // first, rank 1 the second): the partial softmax state of the other element is parked in shared memory and pulled
    // by its owner through distributed shared memory.  The owner goes on to h~, action scoring and the rollout tail.
    // =====================================================================================================
    const int rc = (cid - 2 * P) >> 1, rank = cid & 1;
    const int L = q.L, A = q.A, E = q.E, NCH = q.nch;
    const int eA = rc < B ? rc : -1, eB = (B - 1 - rc > rc) ? B - 1 - rc : -1;
    const int own = rank == 0 ? eA : eB, other = rank == 0 ? eB : eA;
    const int els[2] = {other, own};
    auto rmark = [&](int which) {   // bring-up timeline of the first row CTA
      if (q.trace && tid == 0 && cid == 2 * P) q.trace[which] = globaltimer_ns();
    };
    const int nv = H >> 2;                              // float4 per row; a lane owns up to 4 of them (H <= 512)
    const int Lr = (L + 3) & ~3;
    const uint32_t row_bytes = (uint32_t)H * 4u, chunk_bytes = (uint32_t)T_RW * 2u * row_bytes;
    float* ring = reinterpret_cast<float*>(smem);      // [NCH][T_RW][2][H]  (key row, value row); later the candidate rows
    const size_t ring_bytes = (size_t)NCH * chunk_bytes;
    float* wacc = reinterpret_cast<float*>(smem + ring_bytes);               // [8][H] per-warp partial sums
    float* xacc = wacc + 8 * H;                                              // [H] this CTA's partial of the OTHER element (pulled by the peer)
    float* xstat = xacc + H;                                                 // (m, Z) of: other, own  (read by the peer)
    float* wst = xstat + 4;                                                  // [8][2] per-warp (max, sum)
    float* gs = wst + 16;                                                    // [E + 4]
    float* sc = gs + ((E + 4 + 3) & ~3);                                     // [2][Lr] raw scores by list position
    float* slog = sc + 2 * Lr;                                               // [Ar] logits of the own element
    float* sval = slog + ((A + 3) & ~3);                                     // [Ar] validity flags
    unsigned char* top = MERGED ? top_in : reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sval + ((A + 3) & ~3)) + 15) & ~uintptr_t(15));
    uint64_t* rfull = reinterpret_cast<uint64_t*>(top);
    uint64_t* rempty = rfull + T_MAXCH;
    uint64_t* candfull = rempty + T_MAXCH;
    uint64_t* ring_free = candfull + 1;
    int* list = reinterpret_cast<int*>(ring_free + 1);                       // [2][Lr] un-masked positions
    if (DO_INIT && warp == (MERGED ? 1 : 0)) {
      if (lane < T_MAXCH) {
        mbar_init(&rfull[lane], 1);
        mbar_init(&rempty[lane], T_RW);
      }
      if (lane == 0) {
        mbar_init(candfull, 1);
        mbar_init(ring_free, 1);
      }
      mbar_fence_init();
    }
    if (DO_LISTS && warp == (MERGED ? PW : 0)) {
      // compact the un-masked positions of both elements (padding masks are suffixes in practice; any mask works)
      for (int k = 0; k < 2; ++k) {
        const int e = els[k];
        int n = 0;
        if (e >= 0) {
          const uint8_t* mrow = q.mask ? q.mask + (size_t)e * q.ldmask : nullptr;
          for (int base = 0; base < L; base += 32) {
            const int l = base + lane;
            const bool ok = l < L && !(mrow && mrow[l]);
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) list[k * Lr + n + __popc(bal & ((1u << lane) - 1u))] = l;
            n += __popc(bal);
          }
        }
        if (lane == 0) s_nvalid[k] = n;
      }
      __syncwarp();
    }
    if (STAGE == T_ALL) __syncthreads();
    if (STAGE == T_ALL || STAGE == T_PREFETCH || STAGE == T_MAIN) {
    // rows of element k this CTA streams: list positions rank, rank + 2, ...
    int nmine[2], nchk[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      nmine[k] = s_nvalid[k] > rank ? (s_nvalid[k] - rank + 1) / 2 : 0;
      nchk[k] = (nmine[k] + T_RW - 1) / T_RW;
    }
    if (warp == PW) {
      // ---- producer: key / value rows (per-episode constants: no dependency wait), later the candidate rows.
      // T_PREFETCH issues the chunks that fit the empty ring and leaves; T_MAIN resumes behind them.
      const uint64_t pol = policy_evict_normal();
      bool ok = true;
      int cg = 0;
      const int resume = STAGE == T_MAIN ? s_pre : 0;
      for (int k = 0; k < 2; ++k) {
        if (els[k] < 0) continue;
        const float* kb_ = q.ctx_k + (size_t)els[k] * L * H;
        const float* vb_ = q.ctx_o + (size_t)els[k] * L * H;
        for (int c = 0; c < nchk[k]; ++c, ++cg) {
          if (cg < resume) continue;
          if (STAGE == T_PREFETCH && cg >= NCH) break;
          const int slot = cg % NCH;
          if (cg >= NCH) ok = t_wait(&rempty[slot], (uint32_t)(cg / NCH - 1) & 1u) && ok;
          const int rows = min(T_RW, nmine[k] - c * T_RW);
          if (lane == 0) mbar_expect_tx(&rfull[slot], (uint32_t)rows * 2u * row_bytes);
          __syncwarp();
          if (lane < 2 * rows) {   // lanes 0..rows-1: key rows, rows..2rows-1: value rows
            const int r = lane < rows ? lane : lane - rows;
            const int l = list[k * Lr + rank + 2 * (c * T_RW + r)];
            float* dst = ring + ((size_t)slot * T_RW + r) * 2 * H + (lane < rows ? 0 : H);
            bulk_g2s_hint(dst, (lane < rows ? kb_ : vb_) + (size_t)l * H, row_bytes, &rfull[slot], pol);
          }
        }
      }
      if (STAGE == T_PREFETCH) {
        if (lane == 0) s_pre = cg < NCH ? cg : NCH;
      }
      if (DO_MAIN) {
      cluster_sync_all();   // X1: partials parked
      if (own >= 0) {
        // candidate rows of the own element into the freed ring (step inputs)
        ok = t_wait(ring_free, 0) && ok;
        if (lane == 0) {
          const uint64_t pol2 = policy_evict_first();
          float* us = ring;
          if (q.cand_table == nullptr) {
            mbar_expect_tx(candfull, (uint32_t)((size_t)A * E * 4));
            for (int a = 0; a < A; ++a)
              bulk_g2s_hint(us + (size_t)a * E, q.all_u_t + ((size_t)own * A + a) * E, (uint32_t)E * 4u, candfull, pol2);
          } else {
            int n = 0;
            for (int a = 0; a < A; ++a) n += q.cand_view[(size_t)own * A + a] >= 0;
            mbar_expect_tx(candfull, (uint32_t)((size_t)n * q.img_dim * 4));
            const float* slab = q.cand_table + (size_t)q.vp_idx[own] * q.cand_V * q.img_dim;
            for (int a = 0; a < A; ++a) {
              const int v = q.cand_view[(size_t)own * A + a];
              if (v >= 0) bulk_g2s_hint(us + (size_t)a * E, slab + (size_t)v * q.img_dim, (uint32_t)q.img_dim * 4u, candfull, pol2);
            }
          }
        }
      }
      if (!MERGED) pdl_wait();
      }   // DO_MAIN
      if (!ok) s_fail = 1;
    } else if (!DO_MAIN) {
    } else if (warp >= 8) {
      cluster_sync_all();   // X1
      if (!MERGED) pdl_wait();
    } else {
      // ---- compute warps: each warp owns whole rows (row r of a chunk -> warp r)
      if (!MERGED) {
        pdl_wait();   // h1d is produced by the kernel before this one
      } else {        // ... by the first half of this launch, in every CTA of the grid
        if (tid == 0 && !wait_inputs()) s_fail = 1;
        t_bar256();
      }
      rmark(1);
      // inputs of the tail, fetched while everything else is still on its way
      if (q.has_tail && own >= 0)
        for (int a = tid; a < A; a += 256) sval[a] = q.tail.is_valid[(size_t)own * A + a];
      float4 o_own = make_float4(0.f, 0.f, 0.f, 0.f);   // threads < nv: the CTA-level partial of the own element
      int cg = 0;
      for (int k = 0; k < 2; ++k) {
        const int e = els[k];
        float m = -INFINITY, Z = 0.f;
        float4 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e >= 0) {
          float4 qv[4];
          const float4* q4 = reinterpret_cast<const float4*>(q.h1d + (size_t)e * q.ldh);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int idx = lane + 32 * j;
            qv[j] = idx < nv ? __ldcg(q4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          for (int c = 0; c < nchk[k]; ++c, ++cg) {
            const int slot = cg % NCH, i = c * T_RW + warp;
            if (!t_wait(&rfull[slot], (uint32_t)(cg / NCH) & 1u)) s_fail = 1;
            if (i < nmine[k]) {
              const float4* key4 = reinterpret_cast<const float4*>(ring + ((size_t)slot * T_RW + warp) * 2 * H);
              const float4* val4 = key4 + nv;
              float part = 0.f;
              float4 v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int idx = lane + 32 * j;
                if (idx < nv) {
                  const float4 k4 = key4[idx];
                  v[j] = val4[idx];
                  part = fmaf(k4.x, qv[j].x, part); part = fmaf(k4.y, qv[j].y, part);
                  part = fmaf(k4.z, qv[j].z, part); part = fmaf(k4.w, qv[j].w, part);
                } else {
                  v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
              part = warp_sum(part);
              if (lane == 0) sc[k * Lr + rank + 2 * i] = part;
              const float mn = fmaxf(m, part), corr = __expf(m - mn), ex = __expf(part - mn);
              Z = fmaf(Z, corr, ex);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc[j].x = fmaf(ex, v[j].x, acc[j].x * corr); acc[j].y = fmaf(ex, v[j].y, acc[j].y * corr);
                acc[j].z = fmaf(ex, v[j].z, acc[j].z * corr); acc[j].w = fmaf(ex, v[j].w, acc[j].w * corr);
              }
              m = mn;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rempty[slot]);
          }
        }
        // the 8 per-warp partials of this element -> one CTA partial
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = lane + 32 * j;
          if (idx < nv) reinterpret_cast<float4*>(wacc + (size_t)warp * H)[idx] = acc[j];
        }
        if (lane == 0) { wst[2 * warp] = m; wst[2 * warp + 1] = Z; }
        t_bar256();
        float Mc = -INFINITY, Zc = 0.f, wk[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) Mc = fmaxf(Mc, wst[2 * w]);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          wk[w] = wst[2 * w] != -INFINITY ? __expf(wst[2 * w] - Mc) : 0.f;
          Zc = fmaf(wst[2 * w + 1], wk[w], Zc);
        }
        if (tid < nv) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float4 pw = reinterpret_cast<const float4*>(wacc + (size_t)w * H)[tid];
            o.x = fmaf(wk[w], pw.x, o.x); o.y = fmaf(wk[w], pw.y, o.y); o.z = fmaf(wk[w], pw.z, o.z); o.w = fmaf(wk[w], pw.w, o.w);
          }
          if (k == 0) reinterpret_cast<float4*>(xacc)[tid] = o;
          else o_own = o;
        }
        if (tid == 0) { xstat[2 * k] = Mc; xstat[2 * k + 1] = Zc; }
        t_bar256();   // wacc / wst are free again
      }
      rmark(4);
      cluster_sync_all();   // X1: both CTAs have parked the partial of the element they do not own + both stats
      // merged softmax state of both elements (this CTA's partial + the peer's)
      const uint32_t pstat = dsmem_addr(xstat, (uint32_t)(rank ^ 1)), pacc = dsmem_addr(xacc, (uint32_t)(rank ^ 1));
      float Mm[2], inv[2], wmine[2], wpeer[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {   // my slot k pairs with the peer's slot 1 - k (its "own" is my "other")
        const float m0 = xstat[2 * k], z0 = xstat[2 * k + 1];
        const float m1 = dsmem_ld_f32(pstat + (uint32_t)(2 * (1 - k)) * 4u), z1 = dsmem_ld_f32(pstat + (uint32_t)(2 * (1 - k) + 1) * 4u);
        Mm[k] = fmaxf(m0, m1);
        wmine[k] = m0 != -INFINITY ? __expf(m0 - Mm[k]) : 0.f;
        wpeer[k] = m1 != -INFINITY ? __expf(m1 - Mm[k]) : 0.f;
        const float zt = z0 * wmine[k] + z1 * wpeer[k];
        inv[k] = zt > 0.f ? 1.0f / zt : 0.f;            // every position masked -> zeros (the reference would give NaN)
      }
      if (own >= 0 && tid < nv) {
        const float4 pp = dsmem_ld_f32x4(pacc + (uint32_t)tid * 16u);
        // fixed summation order (rank 0's partial first) whichever CTA owns the element
        const float wa = (rank == 0 ? wmine[1] : wpeer[1]) * inv[1], wb = (rank == 0 ? wpeer[1] : wmine[1]) * inv[1];
        const float4 pa = rank == 0 ? o_own : pp, pb = rank == 0 ? pp : o_own;
        o_own.x = pa.x * wa + pb.x * wb; o_own.y = pa.y * wa + pb.y * wb;
        o_own.z = pa.z * wa + pb.z * wb; o_own.w = pa.w * wa + pb.w * wb;
      }
      auto write_alpha = [&]() {   // attention weights of the rows this CTA streamed (both elements) — off the critical path
        if (!q.alpha) return;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (els[k] < 0) continue;
          float* arow = q.alpha + (size_t)els[k] * q.ldalpha;
          for (int i = tid; i < nmine[k]; i += 256) {
            const int li = rank + 2 * i;
            arow[list[k * Lr + li]] = __expf(sc[k * Lr + li] - Mm[k]) * inv[k];
          }
        }
        if (own >= 0 && q.mask) {   // masked positions of the own element: zero
          const uint8_t* mrow = q.mask + (size_t)own * q.ldmask;
          for (int l = tid; l < L; l += 256)
            if (mrow[l]) q.alpha[(size_t)own * q.ldalpha + l] = 0.f;
        }
      };
      if (own < 0) write_alpha();
      if (own >= 0) {
        const int b = own;
        // h~ = tanh(sum_l alpha_l ctx_o[l] + hh) once hh (W_out_h h1d, from the projection pairs) has arrived
        if (tid == 0 && !t_spin_ge(cnt_hh, 2u * (unsigned int)q.hh_tiles)) s_fail = 1;
        t_bar256();
        const size_t half = (size_t)NB * 128;
        if (tid < nv) {
          const int col = tid;
          float4 o = o_own;
          const float4 hh4 = __ldcg(reinterpret_cast<const float4*>(q.hh + (size_t)b * q.ldhh + col * 4));
          o.x = tanhf(o.x + hh4.x); o.y = tanhf(o.y + hh4.y); o.z = tanhf(o.z + hh4.z); o.w = tanhf(o.w + hh4.w);
          if (q.h_tilde) *reinterpret_cast<float4*>(q.h_tilde + (size_t)b * H + col * 4) = o;
          const int k = col * 4;
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
          unsigned char* dst = q.htpk + (size_t)(k >> 6) * (2 * half) + (size_t)(b >> 3) * 1024 + (size_t)((k & 63) >> 3) * 128 +
                               (size_t)(b & 7) * 16 + (size_t)(k & 7) * 2;
          *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          *reinterpret_cast<uint2*>(dst + half) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
        t_bar256();   // every thread's stores are ordered before thread 0's fences + arrival (cumulativity); the ring can take the candidate rows
        if (tid == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence();
          atomicAdd(cnt_ht, 1u);   // h~ of this element is published
          mbar_arrive(ring_free);
        }
        rmark(5);
        write_alpha();
        // ---- action scoring + rollout tail (model.py:396, follower.py:476-505)
        float* us = ring;   // [A][E]
        if (q.cand_table) {
          // env.py:60-75: [feature[absViewIndex, :img_dim], sin(rh) x n, cos(rh) x n, sin(re) x n, cos(re) x n]; rows
          // without a view (stop, padding) are zero.  The image part arrives by bulk copy (disjoint bytes); this fills
          // the rest.
          const int loc = E - q.img_dim, grp = loc >> 2;
          for (int i = tid; i < A * loc; i += 256) {
            const int a = i / loc, j = i - a * loc;
            const bool okv = q.cand_view[(size_t)b * A + a] >= 0;
            us[(size_t)a * E + q.img_dim + j] = okv ? q.cand_trig[((size_t)b * A + a) * 4 + j / grp] : 0.f;
          }
          for (int a = 0; a < A; ++a)
            if (q.cand_view[(size_t)b * A + a] < 0)
              for (int i = tid; i < q.img_dim; i += 256) us[(size_t)a * E + i] = 0.f;
        }
        // the tail's per-row scalars, requested before the wait for g
        int tgt_pre = -1;
        float u_pre = 0.f;
        if (q.has_tail && warp == 0) {
          tgt_pre = q.tail.target ? q.tail.target[b] : -1;
          u_pre = (q.tail.feedback == 2 && q.tail.sample_u) ? q.tail.sample_u[b] : 0.f;
        }
        if (tid == 0 && !t_spin_ge(cnt_g, 2u * (unsigned int)q.g_tiles)) s_fail = 1;
        t_bar256();
        rmark(6);
        // logit[a] = u_a . g + g[E]: warp w owns a 1/8 slice of the E columns for ALL candidates — its slice of g comes
        // straight from L2 into registers (<= 3 float4 per lane), the candidate rows are in the ring; up to 8 candidates
        // are reduced together (transposed butterfly), the 8 per-warp partials meet in shared memory
        {
          const int nE4 = E >> 2, per_w = (nE4 + 7) >> 3, j0 = warp * per_w;
          const float4* g4 = reinterpret_cast<const float4*>(q.g + (size_t)b * q.ldg);
          float4 gv[3];
          int jj[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int jl = lane + 32 * i, j = j0 + jl;
            const bool okj = jl < per_w && j < nE4;
            jj[i] = okj ? j : 0;
            gv[i] = okj ? __ldcg(g4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const float cst = (warp == 0) ? __ldcg(q.g + (size_t)b * q.ldg + E) : 0.f;
          if (!t_wait(candfull, 0)) s_fail = 1;
          rmark(8);
          float* spart = wacc;   // [8 warps][Ar] (the attention's scratch is free)
          const int Ar = (A + 3) & ~3;
          for (int a0 = 0; a0 < A; a0 += 8) {
            float part[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const int a = a0 + r < A ? a0 + r : A - 1;   // rows past the end repeat the last one (discarded below)
              const float4* u4 = reinterpret_cast<const float4*>(us + (size_t)a * E);
              float acc0 = 0.f;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float4 u = u4[jj[i]];
                acc0 = fmaf(u.x, gv[i].x, acc0); acc0 = fmaf(u.y, gv[i].y, acc0);
                acc0 = fmaf(u.z, gv[i].z, acc0); acc0 = fmaf(u.w, gv[i].w, acc0);
              }
              part[r] = acc0;
            }
            const bool u4b = (lane & 16) != 0;
            float a4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a4[i] = (u4b ? part[4 + i] : part[i]) + __shfl_xor_sync(0xffffffffu, u4b ? part[i] : part[4 + i], 16);
            const bool u3 = (lane & 8) != 0;
            float a2[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) a2[i] = (u3 ? a4[2 + i] : a4[i]) + __shfl_xor_sync(0xffffffffu, u3 ? a4[i] : a4[2 + i], 8);
            const bool u2 = (lane & 4) != 0;
            float t = (u2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, u2 ? a2[0] : a2[1], 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            if ((lane & 3) == 0 && a0 + (lane >> 2) < A) spart[warp * Ar + a0 + (lane >> 2)] = t;
          }
          t_bar256();
          rmark(9);
          if (warp == 0) {
            for (int a = lane; a < A; a += 32) {
              float sum = 0.f;
#pragma unroll
              for (int w = 0; w < 8; ++w) sum += spart[w * Ar + a];   // fixed order
              slog[a] = sum + cst;
            }
            __syncwarp();
          }
        }
        if (q.has_tail) {
          if (warp == 0) {
            const int a_t = tail_row(q.tail, b, lane, us, slog, sval, false, tgt_pre, q.tail.feedback == 2 && q.tail.sample_u ? u_pre : -1.f);
            if (lane == 0) s_at = a_t;
          }
          t_bar256();
          rmark(10);
          tail_copy_u(q.tail, b, s_at, us, tid, 256);
        } else {
          if (warp == 0)
            for (int a = lane; a < A; a += 32) q.logit[(size_t)b * A + a] = slog[a];
        }
        rmark(7);
      }
    }
    if (DO_MAIN) {
      cluster_sync_all();   // X2: the peer has pulled what it needed from this CTA's shared memory (it may go away now)
      __syncthreads();
    }
    }   // T_ALL / T_PREFETCH / T_MAIN
  }
  if (!DO_MAIN) return;
  // ---- exit bookkeeping: the last CTA resets the hand-off counters (no launch of this kernel overlaps another)
  if (tid == 0) {
    if (s_fail) atomicExch(status, 1u);
    __threadfence();
    const unsigned int seen = atomicAdd(cnt_exit, 1u);
    if (seen == gridDim.x - 1u) {
      atomicExch(cnt_hh, 0u);
      atomicExch(cnt_ht, 0u);
      atomicExch(cnt_g, 0u);
      atomicExch(cnt_exit, 0u);
      if (MERGED) atomicExch(const_cast<unsigned int*>(phase), 0u);   // every CTA has passed its wait on it: it arrived here
    }
  }
  trace_mark(q.trace, 2);
}

// grid = 2P + (row CTAs), cluster (2,1,1), TNT threads
__global__ void __launch_bounds__(TNT, 1) text_score_fused_kernel(const FusedTextScoreParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ TShared sh;
  text_score_body<false, T_ALL>(q, smem, nullptr, sh, 0u, nullptr);
}

// ------------------------------------------------------------------ the whole step as one launch
namespace {
struct StepHooks {
  const FusedTextScoreParams& qb;
  unsigned char* smem;
  unsigned char* top;
  TShared& sh;
  __device__ __forceinline__ void early(int) const { text_score_body<true, T_LISTS>(qb, smem, top, sh, 0u, nullptr); }
  int mode;   // bring-up: 0 = no early reads, 1 = projection weights only, 2 = weights + key / value rows
  __device__ __forceinline__ void smem_free(int) const {
    if (mode == 0 || (mode == 1 && (int)blockIdx.x >= 2 * qb.P)) return;
    text_score_body<true, T_PREFETCH>(qb, smem, top, sh, 0u, nullptr);
  }
};
}  // namespace

// grid = tiles * S CTAs (one per SM, the first half's arrangement: CTA c = tile c % tiles, rank c / tiles), cluster
// (2,1,1) for the second half's pairs, FNT threads.  Dynamic smem = the first half's layout + the second half's `top`
// region (barriers + position lists) behind it; the second half's data region aliases the first half's.
__global__ void __launch_bounds__(FNT, 1) step_kernel(const FusedStepParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ TShared sh;
  unsigned char* top = smem + q.top_off;
  text_score_body<true, T_INIT>(q.b, smem, top, sh, 0u, nullptr);   // barriers of the second half (made visible by the first half's first block barrier)
  uint32_t tmem_d = 0;
  const int cid = blockIdx.x;
  vis_lstm_body<true>(q.a, smem, cid % q.tiles, q.tiles, cid / q.tiles, q.S, StepHooks{q.b, smem, top, sh, q.prefetch}, q.phase, tmem_d);
  text_score_body<true, T_MAIN>(q.b, smem, top, sh, tmem_d, q.phase);
  if (threadIdx.x < 32) {
    const uint32_t tmem_cols = q.a.NB <= 32 ? 32 : q.a.NB <= 64 ? 64 : q.a.NB <= 128 ? 128 : 256;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side

FusedTextPlan text_score_fused_plan(int B, int L, int A, int H, int E, int F, bool with_q, int num_sms) {
  FusedTextPlan pl{};
  pl.ok = false;
  if (H < 128 || H > 512 || (H % 128) != 0 || (E % 4) != 0 || E > 3072 || B < 1 || L < 1 || A < 1) return pl;
  const int nkb = H / TBK;
  if ((nkb + 1) / 2 > T_MAXKH) return pl;
  pl.NB = (B + 15) & ~15;
  if (pl.NB > 256 || B > 128) return pl;
  const int grid_max = (num_sms / 2) * 2;
  int rows = (B + 1) & ~1;                          // row CTAs (padded to whole clusters)
  pl.P = (grid_max - rows) / 2;
  const int n1 = H / 128 + (with_q ? (F + 127) / 128 : 0), n2 = (E + 1 + 127) / 128;
  const int want = n1 > n2 ? n1 : n2;
  if (pl.P > want) pl.P = want;
  if (pl.P < 4) return pl;                          // too few projection pairs: the unfused chain is faster
  pl.grid = 2 * pl.P + rows;
  // shared memory: the larger of the two roles
  const size_t b_half = (size_t)(pl.NB / 8) * TSBO;
  const size_t pair_bytes = (size_t)T_MAXKH * 2 * TA_HALF + (size_t)T_BST * 2 * b_half + (2 * T_BST + 3) * sizeof(uint64_t) + 64;
  const size_t chunk = (size_t)T_RW * 2 * H * 4;
  const size_t Lr = (size_t)(L + 3) & ~size_t(3), Ar = (size_t)(A + 3) & ~size_t(3);
  const size_t row_fixed = ((size_t)8 * H + H + 4 + 16) * 4 + ((size_t)(E + 4 + 3) & ~size_t(3)) * 4 + 4 * Lr * 4 + 2 * Ar * 4 + 32 +
                           (2 * T_MAXCH + 2) * sizeof(uint64_t) + 64;
  const size_t budget = 227 * 1024 - 2048;
  const size_t cand = (size_t)A * E * 4;
  size_t ring_min = cand;
  if (ring_min < 2 * chunk) ring_min = 2 * chunk;
  if (row_fixed + ring_min > budget) return pl;
  int nch = (int)((budget - row_fixed) / chunk);
  if (nch > T_MAXCH) nch = T_MAXCH;
  const int nchunks = (L + T_RW - 1) / T_RW;
  if (nch > nchunks && (size_t)nchunks * chunk >= ring_min) nch = nchunks;
  while ((size_t)nch * chunk < ring_min) ++nch;
  if (nch > T_MAXCH || row_fixed + (size_t)nch * chunk > budget) return pl;
  pl.nch = nch;
  const size_t row_bytes = row_fixed + (size_t)nch * chunk;
  pl.smem = row_bytes > pair_bytes ? row_bytes : pair_bytes;
  pl.sync_bytes = 256;
  pl.ok = true;
  return pl;
}

int32_t launch_text_score_fused(const FusedTextScoreParams& q_in, cudaStream_t stream, void* sync_ws, size_t sync_bytes) {
  FusedTextScoreParams q = q_in;
  const FusedTextPlan pl = text_score_fused_plan(q.B, q.L, q.A, q.H, q.E, q.q_cols, q.q_tiles > 0, device_num_sms());
  SFB_CHECK_ARG(pl.ok, "fused text attention + scoring step: unsupported shape");
  SFB_CHECK_ARG(sync_ws && sync_bytes >= pl.sync_bytes && (reinterpret_cast<uintptr_t>(sync_ws) & 15u) == 0, "fused text step: counter words");
  SFB_CHECK_ARG(q.h1d && q.ctx_k && q.ctx_o && q.hh && q.g && q.logit && q.a_hh && q.a_g && q.hdpk && q.htpk, "fused text step: NULL argument");
  SFB_CHECK_ARG(q.q_tiles == 0 || (q.a_q && q.hpk && q.q_next), "fused text step: next-query operands");
  SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(q.ctx_k) & 15u) == 0 && (reinterpret_cast<uintptr_t>(q.ctx_o) & 15u) == 0 &&
                    (reinterpret_cast<uintptr_t>(q.h1d) & 15u) == 0 && (q.ldh % 4) == 0 && (q.ldhh % 4) == 0 && (q.ldg % 4) == 0,
                "fused text step: alignment");
  q.trace = next_trace_slot();
  q.sync = static_cast<unsigned int*>(sync_ws);
  q.NB = pl.NB;
  q.P = pl.P;
  q.nch = pl.nch;
  q.nkb = q.H / TBK;
  static SmemMarks marks;
  SFB_CHECK_CUDA(ensure_dynamic_smem(text_score_fused_kernel, pl.smem, marks));
  SFB_CHECK_CUDA(launch_ex(text_score_fused_kernel, dim3(pl.grid, 1, 1), dim3(TNT, 1, 1), pl.smem, stream, dim3(2, 1, 1), q));
  count_launch();
  return 0;
}

// ------------------------------------------------------------------ one-launch step: host side

FusedStepPlan step_fused_plan(int B, int L, int A, int H, int E, int F, bool with_q, int nkb, int R, int D, int lenA, int lenB, int num_sms) {
  FusedStepPlan pl{};
  pl.ok = false;
  pl.a = vis_lstm_fused_plan(B, H, nkb, R, D, lenA, lenB, num_sms);
  pl.b = text_score_fused_plan(B, L, A, H, E, F, with_q, num_sms);
  if (!pl.a.ok || !pl.b.ok || pl.a.NB != pl.b.NB) return pl;
  const int grid = pl.a.tiles * pl.a.S;
  const int rows = (B + 1) & ~1;
  if ((grid & 1) || grid < rows + 8) return pl;
  // second half inside the first half's geometry: projection pairs from the CTAs that are not row CTAs, the data
  // region (ring / weight buffer / scratch) inside the first half's data region, barriers + lists behind everything
  int P = (grid - rows) / 2;
  const int n1 = H / 128 + (with_q ? (F + 127) / 128 : 0), n2 = (E + 1 + 127) / 128;
  const int want = n1 > n2 ? n1 : n2;
  if (P > want) P = want;
  if (P < 4) return pl;
  const size_t b_half = (size_t)(pl.b.NB / 8) * TSBO;
  const size_t pair_data = (size_t)T_MAXKH * 2 * TA_HALF + (size_t)T_BST * 2 * b_half;
  const size_t chunk = (size_t)T_RW * 2 * H * 4;
  const size_t Lr = (size_t)(L + 3) & ~size_t(3), Ar = (size_t)(A + 3) & ~size_t(3);
  const size_t row_arrays = ((size_t)8 * H + H + 4 + 16) * 4 + ((size_t)(E + 4 + 3) & ~size_t(3)) * 4 + 2 * Lr * 4 + 2 * Ar * 4 + 16;
  const size_t limit = pl.a.data_bytes;
  if (pair_data > limit || row_arrays + 2 * chunk > limit) return pl;
  int nch = (int)((limit - row_arrays) / chunk);
  if (nch > T_MAXCH) nch = T_MAXCH;
  size_t ring_min = (size_t)A * E * 4;   // the candidate rows are staged in the ring
  if (ring_min < 2 * chunk) ring_min = 2 * chunk;
  if ((size_t)nch * chunk < ring_min) return pl;
  const int nchunks = (L + T_RW - 1) / T_RW;
  if (nch > nchunks && (size_t)nchunks * chunk >= ring_min) nch = nchunks;
  pl.b.P = P;
  pl.b.nch = nch;
  pl.b.grid = grid;
  const size_t top_rows = (2 * T_MAXCH + 2) * sizeof(uint64_t) + 2 * Lr * 4, top_pairs = (2 * T_BST + 3) * sizeof(uint64_t);
  pl.top_off = (pl.a.smem + 15) & ~size_t(15);
  pl.smem = pl.top_off + (top_rows > top_pairs ? top_rows : top_pairs) + 16;
  if (pl.smem > 227 * 1024 - 1024 - 64) return pl;   // static shared memory of the kernel
  pl.ok = true;
  return pl;
}

int32_t launch_step_fused(const FusedVisLstmParams& qa, const FusedTextScoreParams& qb, cudaStream_t stream, void* ws, size_t ws_bytes,
                          void* sync_ws, size_t sync_bytes) {
  FusedStepParams q{};
  q.a = qa;
  q.b = qb;
  GemmParams& p = q.a.g;
  const FusedStepPlan pl = step_fused_plan(qb.B, qb.L, qb.A, qb.H, qb.E, qb.q_cols, qb.q_tiles > 0, qa.nkb, qa.R, qa.D, qa.lenA, qa.lenB,
                                           device_num_sms());
  SFB_CHECK_ARG(pl.ok && qa.B == qb.B, "one-launch step: unsupported shape");
  // the argument checks of the two stand-alone launchers
  SFB_CHECK_ARG(qa.a_pk && qa.b_pk && (reinterpret_cast<uintptr_t>(qa.a_pk) & 127u) == 0 && (reinterpret_cast<uintptr_t>(qa.b_pk) & 127u) == 0,
                "one-launch step: packed operands missing / misaligned");
  SFB_CHECK_ARG(qa.post_kb0 >= 0 && qa.post_kb1 <= qa.nkb && qa.post_kb0 <= qa.feat_kb0 && qa.feat_kb0 + (qa.D + FBK - 1) / FBK <= qa.post_kb1,
                "one-launch step: the post K range must cover the feature blocks");
  SFB_CHECK_ARG(qa.q && qa.segA && (qa.lenB == 0 || qa.segB) && (reinterpret_cast<uintptr_t>(qa.segA) & 15u) == 0 &&
                    (qa.strideA_b % 4) == 0 && (qa.lenB == 0 || ((reinterpret_cast<uintptr_t>(qa.segB) & 15u) == 0 && (qa.strideB_b % 4) == 0)),
                "one-launch step: visual source missing / misaligned");
  SFB_CHECK_ARG(ws && ws_bytes >= pl.a.bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "one-launch step: workspace");
  SFB_CHECK_ARG(sync_ws && sync_bytes >= 64 && (reinterpret_cast<uintptr_t>(sync_ws) & 15u) == 0, "one-launch step: counter words");
  SFB_CHECK_ARG(qb.h1d && qb.ctx_k && qb.ctx_o && qb.hh && qb.g && qb.logit && qb.a_hh && qb.a_g && qb.hdpk && qb.htpk, "one-launch step: NULL argument");
  SFB_CHECK_ARG(qb.q_tiles == 0 || (qb.a_q && qb.hpk && qb.q_next), "one-launch step: next-query operands");
  SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(qb.ctx_k) & 15u) == 0 && (reinterpret_cast<uintptr_t>(qb.ctx_o) & 15u) == 0 &&
                    (reinterpret_cast<uintptr_t>(qb.h1d) & 15u) == 0 && (qb.ldh % 4) == 0 && (qb.ldhh % 4) == 0 && (qb.ldg % 4) == 0,
                "one-launch step: alignment");
  p.trace = next_trace_slot();
  q.b.trace = next_trace_slot();
  q.a.sem = static_cast<unsigned int*>(ws);
  q.a.sync = q.a.sem + 2 * pl.a.tiles + 16;
  q.a.partial = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.a.sem_bytes);
  q.a.NB = pl.a.NB;
  q.a.nch = pl.a.nch;
  q.a.chunk_rows = pl.a.chunk_rows;
  p.M = qa.B;
  q.b.sync = static_cast<unsigned int*>(sync_ws);
  q.b.NB = pl.b.NB;
  q.b.P = pl.b.P;
  q.b.nch = pl.b.nch;
  q.b.nkb = qb.H / TBK;
  q.phase = q.b.sync + 8;
  q.top_off = (int)pl.top_off;
  q.prefetch = g_merged_prefetch;
  q.tiles = pl.a.tiles;
  q.S = pl.a.S;
  static SmemMarks marks;
  SFB_CHECK_CUDA(ensure_dynamic_smem(step_kernel, pl.smem, marks));
  SFB_CHECK_CUDA(launch_ex(step_kernel, dim3(pl.a.tiles * pl.a.S, 1, 1), dim3(FNT, 1, 1), pl.smem, stream, dim3(2, 1, 1), q));
  count_launch();
  return 0;
}

}  // namespace sfb
