// step_fused_a.cuh — device code of the gather + gate GEMM + LSTM cell part of a decode step (see step_fused.cu),
// shared by the stand-alone kernel and the one-launch step kernel (step_fused_b.cu).
#pragma once
#include <cuda_bf16.h>

#include "epilogue.cuh"
#include "kernels.h"
#include "pack.cuh"

namespace sfb {

namespace {   // one copy per translation unit
constexpr int FBM = 128, FBK = 64;
constexpr int FNT = 384;                         // 8 compute warps + MMA issuer + GEMM producer + gather producer + 1 epilogue helper
constexpr int FGS = 3;                           // GEMM stages at most (1 dedicated + 2 carved from the ring)
constexpr uint32_t FCORE = 128;
constexpr uint32_t FSBO = (FBK / 8) * FCORE;
constexpr uint32_t FLBO = FCORE;
constexpr uint32_t FA_HALF = (FBM / 8) * FSBO;   // 16 KB
constexpr int F_RB = 6;                          // slab rows per ring chunk (one bulk copy); <= 8 so that a warp owns at most one row of a chunk
constexpr int F_D = 2176;                        // row length this kernel is built for (2048 image + 128 orientation floats)
constexpr int F_MAXCH = 8;                       // ring depth in chunks

__device__ __forceinline__ bool f_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    __nanosleep(32);   // a hot try_wait loop competes with the compute warps for the shared-memory pipe
  }
  return false;   // never hang the device: the launch is flagged as failed instead
}
__device__ __forceinline__ uint64_t f_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(FLBO >> 4) << 16;
  d |= (uint64_t)(FSBO >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void f_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void f_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The stage an iteration uses and how often that stage was used before — the same sequence in the producer and the
// MMA issuer.  While the gather owns the ring (pre part of a gather CTA) only stage 0 exists.
struct StageSeq {
  int ng, rr, use[FGS];
  bool pinned;   // pre iterations of a gather CTA stay on stage 0
  __device__ StageSeq(int ng_, bool pinned_) : ng(ng_), rr(0), pinned(pinned_) {
#pragma unroll
    for (int i = 0; i < FGS; ++i) use[i] = 0;
  }
  __device__ void next(bool is_pre, int& s, int& u) {
    s = (is_pre && pinned) ? 0 : (rr++ % ng);
    u = use[s]++;
  }
};
}  // namespace

// Hooks of the one-launch step kernel: what the gather producer warp (idle once the slab copies are issued) does for
// the SECOND half of the step inside this CTA.  The stand-alone kernel uses the empty ones.
struct NoStepHooks {
  __device__ __forceinline__ void early(int) const {}        // right after the slab copies are issued
  __device__ __forceinline__ void smem_free(int) const {}    // the data region of shared memory is no longer used by this half
};

// One CTA of the kernel: (tile, rank) of a (tiles x S) arrangement.  MERGED: the tensor-memory allocation is handed to
// the caller (tmem_out) instead of being released, and `phase` receives one arrival when every output of this CTA
// is visible device-wide.
template <bool MERGED, class Hooks>
__device__ __forceinline__ void vis_lstm_body(const FusedVisLstmParams& q, unsigned char* smem, const int tile, const int tiles,
                                              const int rank, const int S, const Hooks& hooks, unsigned int* phase,
                                              uint32_t& tmem_out) {
  const GemmParams& p = q.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cid = rank * tiles + tile;          // gather element of this CTA (ranks spread the elements over all tiles)
  const int NB = q.NB, B = q.B;
  const bool has_gather = cid < B;
  const int D = q.D, nvec = D >> 2, NCH = q.nch, R = q.R;
  const int lenA = q.lenA, lenB = q.lenB;

  const uint32_t b_half = (uint32_t)(NB / 8) * FSBO;
  const uint32_t stage_bytes = 2 * FA_HALF + 2 * b_half;
  const uint32_t chunk_floats = (uint32_t)F_RB * (uint32_t)lenA;
  const uint32_t ring_bytes = (uint32_t)NCH * chunk_floats * 4u;
  unsigned char* ring_raw = smem + stage_bytes;
  float* ring = reinterpret_cast<float*>(ring_raw);                                     // [NCH][F_RB][lenA]
  float* locb = reinterpret_cast<float*>(ring_raw + ring_bytes);                        // [R][lenB]
  uint64_t* gfull = reinterpret_cast<uint64_t*>(locb + (size_t)R * lenB);
  uint64_t* gempty = gfull + FGS;
  uint64_t* done = gempty + FGS;
  uint64_t* rfull = done + 1;                                                           // [F_MAXCH]
  uint64_t* rempty = rfull + F_MAXCH;                                                   // [F_MAXCH]
  uint64_t* locfull = rempty + F_MAXCH;
  uint64_t* ring_free = locfull + 1;
  float* red = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ring_free + 1) + 15) & ~uintptr_t(15));   // [2][8 warps][8 rows]
  float* sc = red + 2 * 8 * 8;                                                          // [R] raw scores
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc + ((R + 3) & ~3));
  __shared__ int s_fail;

  const int NG = min(FGS, 1 + (int)(ring_bytes / stage_bytes));   // GEMM stages once the ring is released
  auto stage_ptr = [&](int s) -> unsigned char* { return s == 0 ? smem : ring_raw + (size_t)(s - 1) * stage_bytes; };

  const uint32_t tmem_cols = NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < FGS; ++s) {
        mbar_init(&gfull[s], 1);
        mbar_init(&gempty[s], 1);
      }
      mbar_init(done, 1);
      mbar_init(locfull, 1);
      mbar_init(ring_free, 1);
      s_fail = 0;
    }
    if (lane < F_MAXCH) {
      mbar_init(&rfull[lane], 1);
      mbar_init(&rempty[lane], 1);
    }
    mbar_fence_init();
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // K blocks of this CTA: a slice of the "pre" blocks (everything outside [post_kb0, post_kb1)) first, then a slice of
  // the "post" blocks (the attention output)
  const int npost_all = q.post_kb1 - q.post_kb0, npre_all = q.nkb - npost_all;
  const int per_post = (npost_all + S - 1) / S;
  // pre blocks: CTAs that also gather take a small share (their shared-memory bandwidth belongs to the slab stream),
  // the others proportionally more: rank r of this tile gathers iff r * tiles + tile < B
  const int ng = B > tile ? min(S, (B - tile + tiles - 1) / tiles) : 0;
  const int wf = q.pre_weight_free > 0 ? q.pre_weight_free : 1;
  const int wsum = ng + (S - ng) * wf;
  auto cumw = [&](int r) { return r <= ng ? r : ng + (r - ng) * wf; };
  const int pre0 = (int)((long long)npre_all * cumw(rank) / wsum), pre1 = (int)((long long)npre_all * cumw(rank + 1) / wsum);
  const int post0 = min(npost_all, rank * per_post), post1 = min(npost_all, post0 + per_post);
  const int n_pre = pre1 - pre0, n_post = post1 - post0, nit = n_pre + n_post;
  auto kblock = [&](int it) {   // iteration -> K block index in the packed operands
    if (it < n_pre) {
      const int i = pre0 + it;
      return i < q.post_kb0 ? i : i + npost_all;
    }
    return q.post_kb0 + post0 + (it - n_pre);
  };

  trace_mark(p.trace, 0);
  pdl_launch_dependents();

  if (warp == 10) {
    // =============================== gather producer: the slab of this CTA's batch element ===============================
    if (q.idx_dependent) pdl_wait();   // the slab indices come from the kernel before this one (table-driven environment)
    if (has_gather && lane == 0) {
      const int b = cid;
      const long long ia = q.idxA ? (long long)q.idxA[b] : (long long)b;
      const long long ib = q.idxB ? (long long)q.idxB[b] : (long long)b;
      const float* srcA = q.segA + (size_t)ia * q.strideA_b;
      const uint64_t pol = policy_evict_first();
      if (lenB > 0) {   // the R orientation rows of this element: one contiguous block
        mbar_expect_tx(locfull, (uint32_t)R * (uint32_t)lenB * 4u);
        bulk_g2s_hint(locb, q.segB + (size_t)ib * q.strideB_b, (uint32_t)R * (uint32_t)lenB * 4u, locfull, pol);
      }
      const int nchunks = (R + F_RB - 1) / F_RB;
      bool ok = true;
      for (int c = 0; c < nchunks; ++c) {
        const int slot = c % NCH;
        if (c >= NCH) ok = f_wait(&rempty[slot], (uint32_t)(c / NCH - 1) & 1u) && ok;
        const int rows = min(F_RB, R - c * F_RB);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)lenA * 4u;
        mbar_expect_tx(&rfull[slot], bytes);
        bulk_g2s_hint(ring + (size_t)slot * chunk_floats, srcA + (size_t)c * F_RB * lenA, bytes, &rfull[slot], pol);
      }
      if (!ok) s_fail = 1;
    }
    if (MERGED) {
      // the warp is idle while the slab streams: set up the second half's reads, and start them the moment this half no
      // longer needs the data region of shared memory (all MMAs retired; the gather released the ring before that)
      hooks.early(lane);
      bool ok = true;
      if (has_gather) ok = f_wait(ring_free, 0) && ok;
      if (nit > 0) ok = f_wait(done, 0) && ok;
      if (!ok) s_fail = 1;
      hooks.smem_free(lane);
    }
    pdl_wait();
  } else if (warp == 9) {
    // =============================== GEMM producer ===============================
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();   // weights are re-read every step: keep them in L2
      const unsigned char* a_base = q.a_pk + (size_t)tile * q.nkb * (2 * FA_HALF);
      const unsigned char* b_base = q.b_pk;
      const uint32_t tx = 2 * FA_HALF + 2 * b_half;
      bool ok = true, ring_waited = !has_gather;
      StageSeq seq(NG, has_gather);
      int st_of[4], n_issued_a = 0, n_issued_b = 0;   // stages of the (<= FGS) iterations whose A load is ahead of their B load
      auto issue_a = [&](int it) {   // claim the iteration's stage (waiting for it to drain) and start its weight copy
        int s, u;
        seq.next(it < n_pre, s, u);
        if (s > 0 && !ring_waited) {
          ok = f_wait(ring_free, 0) && ok;
          ring_waited = true;
        }
        if (u > 0) ok = f_wait(&gempty[s], (uint32_t)(u - 1) & 1u) && ok;
        mbar_expect_tx(&gfull[s], tx);
        bulk_g2s_hint(stage_ptr(s), a_base + (size_t)kblock(it) * (2 * FA_HALF), 2 * FA_HALF, &gfull[s], pol);
        st_of[n_issued_a & 3] = s;
        ++n_issued_a;
      };
      auto issue_b = [&](int it) {
        const int s = st_of[n_issued_b & 3];
        ++n_issued_b;
        bulk_g2s(stage_ptr(s) + 2 * FA_HALF, b_base + (size_t)kblock(it) * (2 * b_half), 2 * b_half, &gfull[s]);
      };
      // weights of the first pre iterations before the dependency wait (never written inside a step)
      const int ahead = has_gather ? 1 : NG;
      const int first = min(n_pre, ahead);
      for (int it = 0; it < first; ++it) issue_a(it);
      pdl_wait();   // the packed u_prev / h_0 blocks come from the previous step's kernels
      for (int it = 0; it < first; ++it) issue_b(it);
      for (int it = first; it < n_pre; ++it) {
        issue_a(it);
        issue_b(it);
      }
      // post part: weights into every stage that frees up, activations once all gather CTAs have signalled
      const int pf = min(n_post, NG);
      for (int j = 0; j < pf; ++j) issue_a(n_pre + j);
      if (n_post > 0) {
        bool arrived = false;
        for (uint32_t i = 0; i < (1u << 22); ++i) {   // one poller per SM: back off so the arrivals are not starved
          if (ld_acquire_u32(q.sync) >= (unsigned int)B) { arrived = true; break; }
          __nanosleep(40);
        }
        ok = ok && arrived;
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[11] = globaltimer_ns();
        asm volatile("fence.proxy.async;" ::: "memory");   // the feature blocks were written with generic-proxy stores
        for (int j = 0; j < pf; ++j) issue_b(n_pre + j);
        for (int j = pf; j < n_post; ++j) {
          issue_a(n_pre + j);
          issue_b(n_pre + j);
        }
      }
      if (!ok) s_fail = 1;
      // the last CTA to get here resets the arrival counter: every waiter has observed it by then, and no two
      // launches of this kernel overlap (the kernel that follows waits for this grid's completion)
      const unsigned int seen = atomicAdd(q.sync + 1, 1u);
      if (seen == (unsigned int)(tiles * S) - 1u) {
        atomicExch(q.sync, 0u);
        atomicExch(q.sync + 1, 0u);
      }
    } else {
      pdl_wait();
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(FBM >> 4) << 24);
    bool ok = true;
    StageSeq seq(NG, has_gather);
    for (int it = 0; it < nit; ++it) {
      int s, u;
      seq.next(it < n_pre, s, u);
      ok = f_wait(&gfull[s], (uint32_t)u & 1u) && ok;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(stage_ptr(s)), a_lo = a_hi + FA_HALF, b_hi = a_hi + 2 * FA_HALF, b_lo = b_hi + b_half;
#pragma unroll
        for (int j = 0; j < FBK / 16; ++j) {
          const uint32_t ko = (uint32_t)j * 2u * FCORE;
          const uint64_t dah = f_desc(a_hi + ko), dal = f_desc(a_lo + ko);
          const uint64_t dbh = f_desc(b_hi + ko), dbl = f_desc(b_lo + ko);
          f_umma(tmem_d, dal, dbh, idesc, (it > 0 || j > 0) ? 1u : 0u);   // small terms first
          f_umma(tmem_d, dah, dbl, idesc, 1u);
          f_umma(tmem_d, dah, dbh, idesc, 1u);
        }
        f_commit(&gempty[s]);
        if (it + 1 == nit) f_commit(done);
      }
      __syncwarp();
    }
    if (!ok && lane == 0) s_fail = 1;
    pdl_wait();
  } else if (warp >= 8) {
    pdl_wait();   // epilogue helper warp
  } else {
    // =============================== warps 0-7: gather consumers ===============================
    pdl_wait();   // q is produced by the previous step
    trace_mark(p.trace, 1);
    if (has_gather) {
      // Column ownership: thread t owns float4 columns t and 256 + t of the feature-table part of a row and (threads
      // 0-31) column t of its tail (orientation block, or the last 128 floats of a dense row).  Few registers per
      // thread, so the F_RB x 3 shared-memory loads of a chunk are all issued before the first use (branch-free:
      // addresses are always valid, lanes without a tail column select zero), F_RB independent dot products, one
      // transposed-butterfly reduction and ONE block barrier per chunk.
      const int b = cid;
      const int nvA = lenA >> 2;
      const bool has_tail = tid < nvec - 512;           // D = 2176: 544 float4 per row = 2 x 256 + 32
      const int tcol = tid & 31;
      float4 qv[3], acc[3];
      {
        const float4* q4 = reinterpret_cast<const float4*>(q.q + (size_t)b * q.ldq);
        qv[0] = q4[tid]; qv[1] = q4[256 + tid];
        qv[2] = has_tail ? q4[512 + tcol] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // tail columns: orientation block [R][lenB/4] (gather source) or columns 512.. of the dense chunk row
      const float4* tail_base = lenB > 0 ? reinterpret_cast<const float4*>(locb) + tcol : nullptr;
      const int tail_stride = lenB > 0 ? (lenB >> 2) : nvA;
      if (lenB > 0 && !f_wait(locfull, 0)) s_fail = 1;
      float m = -INFINITY, Z = 0.f;
      int c = 0;
      long long waited = 0;   // bring-up: cycles thread 0 spent waiting for slab data
      for (int i0 = 0; i0 < R; i0 += F_RB, ++c) {
        const int nb = min(F_RB, R - i0);
        const int slot = c % NCH;
        const long long t0 = p.trace ? clock64() : 0;
        if (!f_wait(&rfull[slot], (uint32_t)(c / NCH) & 1u)) s_fail = 1;
        if (p.trace) waited += clock64() - t0;
        if (c == 0) trace_mark(p.trace, 8);
        if (i0 + F_RB >= R) trace_mark(p.trace, 9);
        const float4* pA = reinterpret_cast<const float4*>(ring + (size_t)slot * chunk_floats) + tid;
        const float4* pT = lenB > 0 ? tail_base + (size_t)i0 * tail_stride : pA - tid + 512 + tcol;
        float4 v[F_RB][3];
        if (q.dbg == 1) {   // bring-up: stream only
          bar_sync_256();
          if (tid == 0) mbar_arrive(&rempty[slot]);
          continue;
        }
#pragma unroll
        for (int r = 0; r < F_RB; ++r) {
          const int rr = r < nb ? r : nb - 1;           // rows past the end re-read the last row (their weight is zero)
          v[r][0] = pA[(size_t)rr * nvA];
          v[r][1] = pA[(size_t)rr * nvA + 256];
          v[r][2] = pT[(size_t)rr * tail_stride];
        }
        float part[8];
#pragma unroll
        for (int r = 0; r < F_RB; ++r) {
          if (!has_tail) v[r][2] = make_float4(0.f, 0.f, 0.f, 0.f);
          const float d0 = fmaf(v[r][0].w, qv[0].w, fmaf(v[r][0].z, qv[0].z, fmaf(v[r][0].y, qv[0].y, v[r][0].x * qv[0].x)));
          const float d1 = fmaf(v[r][1].w, qv[1].w, fmaf(v[r][1].z, qv[1].z, fmaf(v[r][1].y, qv[1].y, v[r][1].x * qv[1].x)));
          const float d2 = fmaf(v[r][2].w, qv[2].w, fmaf(v[r][2].z, qv[2].z, fmaf(v[r][2].y, qv[2].y, v[r][2].x * qv[2].x)));
          part[r] = (d0 + d1) + d2;
        }
#pragma unroll
        for (int r = F_RB; r < 8; ++r) part[r] = 0.f;
        // up to 8 row sums per warp with 9 shuffles (transposed butterfly): three halving exchanges leave every lane with
        // ONE row's partial (row = lane bits 4,3,2), two more steps finish the sum inside each group of 4 lanes
        float* rb = red + (c & 1) * 8 * 8;
        if (q.dbg == 2) {   // bring-up: loads + dot products only
          if (part[0] + part[1] + part[2] + part[3] + part[4] + part[5] == 123.456f) sc[0] = 1.f;
          bar_sync_256();
          if (tid == 0) mbar_arrive(&rempty[slot]);
          continue;
        }
        {
          const bool u4 = (lane & 16) != 0;
          float a4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a4[i] = (u4 ? part[4 + i] : part[i]) + __shfl_xor_sync(0xffffffffu, u4 ? part[i] : part[4 + i], 16);
          const bool u3 = (lane & 8) != 0;
          float a2[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) a2[i] = (u3 ? a4[2 + i] : a4[i]) + __shfl_xor_sync(0xffffffffu, u3 ? a4[i] : a4[2 + i], 8);
          const bool u2 = (lane & 4) != 0;
          float t = (u2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, u2 ? a2[0] : a2[1], 4);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          if ((lane & 3) == 0) rb[warp * 8 + (lane >> 2)] = t;   // lanes 0, 4, ..., 28 hold rows 0..7
        }
        bar_sync_256();   // every thread holds its slices in registers -> the chunk is free
        if (tid == 0) mbar_arrive(&rempty[slot]);
        if (q.dbg == 3) continue;   // bring-up: no softmax / accumulation
        float sr[F_RB], mn = m;
        {
          float4 lo4 = make_float4(0.f, 0.f, 0.f, 0.f), hi4 = lo4;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float4 x0 = *reinterpret_cast<const float4*>(rb + w * 8), x1 = *reinterpret_cast<const float4*>(rb + w * 8 + 4);
            lo4.x += x0.x; lo4.y += x0.y; lo4.z += x0.z; lo4.w += x0.w;
            hi4.x += x1.x; hi4.y += x1.y; hi4.z += x1.z; hi4.w += x1.w;
          }
          const float all[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
          for (int r = 0; r < F_RB; ++r) {
            sr[r] = r < nb ? all[r] : -INFINITY;
            mn = fmaxf(mn, sr[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < F_RB; ++r)
          if (tid == r && r < nb) sc[i0 + r] = sr[r];
        const float corr = __expf(m - mn);
        float e[F_RB], esum = 0.f;
#pragma unroll
        for (int r = 0; r < F_RB; ++r) {
          e[r] = __expf(sr[r] - mn);   // rows past the end: exp(-inf) = 0
          esum += e[r];
        }
        Z = fmaf(Z, corr, esum);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float4 a = acc[j];
          a.x *= corr; a.y *= corr; a.z *= corr; a.w *= corr;
#pragma unroll
          for (int r = 0; r < F_RB; ++r) {
            a.x = fmaf(e[r], v[r][j].x, a.x);
            a.y = fmaf(e[r], v[r][j].y, a.y);
            a.z = fmaf(e[r], v[r][j].z, a.z);
            a.w = fmaf(e[r], v[r][j].w, a.w);
          }
          acc[j] = a;
        }
        m = mn;
      }
      bar_sync_256();   // all scores are in sc[]; nobody reads the ring any more
      if (tid == 0) mbar_arrive(ring_free);   // the ring becomes GEMM stages 1, 2
      trace_mark(p.trace, 4);
      if (p.trace && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) p.trace[10] = p.trace[1] + (unsigned long long)(waited / 2);   // ~ns at ~2 GHz
      // the CTA owns the whole batch element: normalise and emit (fp32 + packed bf16 hi/lo for the post part)
      const float inv = Z > 0.f ? 1.0f / Z : 0.f;
      if (q.alpha)
        for (int i = tid; i < R; i += 256) q.alpha[(size_t)b * q.ldalpha + i] = __expf(sc[i] - m) * inv;
      const size_t half = (size_t)NB * 128;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = j < 2 ? tid + 256 * j : 512 + tcol;
        if (j < 2 || has_tail) {
          float4 o = acc[j];
          o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
          if (q.feat) *reinterpret_cast<float4*>(q.feat + (size_t)b * q.ldfeat + col * 4) = o;
          const int k = col * 4;
          if (q.pk_scale) {
            const float4 s4 = *reinterpret_cast<const float4*>(q.pk_scale + (size_t)b * q.pk_ldscale + k);
            o.x *= s4.x; o.y *= s4.y; o.z *= s4.z; o.w *= s4.w;
          }
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
          unsigned char* dst = q.b_pk + (size_t)(q.feat_kb0 + (k >> 6)) * (2 * half) + (size_t)(b >> 3) * 1024 +
                               (size_t)((k & 63) >> 3) * 128 + (size_t)(b & 7) * 16 + (size_t)(k & 7) * 2;
          *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          *reinterpret_cast<uint2*>(dst + half) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
      }
      bar_sync_256();   // every thread's stores are ordered before thread 0's device-scope fence (cumulativity)
      if (tid == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        atomicAdd(q.sync, 1u);   // this batch element's feature blocks are visible device-wide
      }
      trace_mark(p.trace, 5);
    }
  }

  // ---- epilogue 1: TMEM -> registers -> [col][row] partial tile in L2
  float* mypart = q.partial + ((size_t)tile * S + rank) * (size_t)(FBM * NB);
  if (nit > 0 && warp < 8) {
    if (!f_wait(done, 0)) s_fail = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  trace_mark(p.trace, 6);
  if (warp < 8) {
    const int lq = warp & 3, hf = warp >> 2;
    const int row = lq * 32 + lane;
    for (int c = hf * 16; c < NB; c += 32) {
      if (c >= B) break;
      uint32_t v[16];
      if (nit > 0) {
        const uint32_t taddr = tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)c;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c + j < B) __stcg(mypart + (size_t)(c + j) * FBM + row, __uint_as_float(v[j]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  trace_mark(p.trace, 12);
  // epilogue operands of this thread's first-pass elements: requested now, consumed after the barrier
  LstmPre1 lpre0, lpre1;
  lpre0.ok = false; lpre1.ok = false;
  const int total = B * 32, share = (((total + S - 1) / S) + 31) & ~31;
  const int e_beg = rank * share, e_end = min(total, e_beg + share);
  {
    const int e0 = e_beg + tid, e1 = e0 + FNT;
    if (e0 < e_end) lpre0 = lstm_preload1(p, e0 >> 5, tile * 32 + (e0 & 31));
    if (e1 < e_end) lpre1 = lstm_preload1(p, e1 >> 5, tile * 32 + (e1 & 31));
  }
  __syncthreads();   // the CTA's partial stores are ordered before thread 0's device-scope fence below
  trace_mark(p.trace, 13);
  // split-K barrier of this tile: a monotonic arrival counter (S arrivals per launch, never reset: one atomic per CTA,
  // the release of the last arriver IS its arrival); a launch ends when the count reaches the next multiple of S
  // (64-bit: never wraps)
  unsigned long long* my_sem = reinterpret_cast<unsigned long long*>(q.sem) + tile;
  if (S > 1) {
    if (tid == 0) {
      __threadfence();
      const unsigned long long old = atomicAdd(my_sem, 1ull);
      const unsigned long long target = (old / (unsigned long long)S + 1ull) * (unsigned long long)S;
      bool ok = false;
      for (uint32_t i = 0; i < (1u << 24); ++i) {
        unsigned long long cur;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(my_sem) : "memory");
        if (cur >= target) { ok = true; break; }
      }
      if (!ok) s_fail = 1;
      __threadfence();
    }
    __syncthreads();
  }
  trace_mark(p.trace, 7);
  // ---- epilogue 2: reduce the S partial tiles in split order + LSTM cell, one (batch row, hidden unit) per thread
  {
    const float* tbase = q.partial + (size_t)tile * S * (size_t)(FBM * NB);
    const size_t pstride = (size_t)FBM * NB;
    for (int e = e_beg + tid; e < e_end; e += FNT) {
      const int col = e >> 5, ul = e & 31;
      const float* pk = tbase + (size_t)col * FBM + ul;
      float v[16][4];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (k < S) {
          const float* pp = pk + (size_t)k * pstride;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) v[k][gq] = __ldcg(pp + 32 * gq);
        }
      }
      float g4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < S) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) g4[gq] += v[k][gq];
        }
      trace_mark(p.trace, 14);
      const int which = (e - e_beg - tid) / FNT;
      if (which == 0) lstm_update1_pre(p, col, tile * 32 + ul, g4[0], g4[1], g4[2], g4[3], lpre0);
      else if (which == 1) lstm_update1_pre(p, col, tile * 32 + ul, g4[0], g4[1], g4[2], g4[3], lpre1);
      else lstm_update(p, col, tile * 32 + ul, g4[0], g4[1], g4[2], g4[3]);
    }
  }
  trace_mark(p.trace, 15);
  __syncthreads();
  if (s_fail && tid == 0) atomicExch(q.sync + 2, 1u);   // sticky status word (sfb_debug_status)
  trace_mark(p.trace, 2);
  if (MERGED) {
    tmem_out = tmem_d;
    if (tid == 0) {   // every store of this CTA (ordered by the barrier above) becomes visible before the arrival
      __threadfence();
      atomicAdd(phase, 1u);
    }
  } else if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

}  // namespace sfb
