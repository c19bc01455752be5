// encoder_persist.cu — the recurrent part of EncoderLSTM (model.py:81-104) as ONE launch for all time steps.
//
//   for s in 0..maxlen-1:  gates = W_hh h_{s-1} + (W_ih x_t + biases)   (the input projection is hoisted: api.cu)
//                          (h_s, c_s) = LSTM cell, rows past their length carry their state (packed sequence)
//
// W_hh never leaves the SMs: as bf16 (hi, lo) it is 4 MB at Hd = 512, spread over tiles x 2 CTAs of 128 KB each — CTA
// (tile, rank) keeps the 128 gate rows of 32 hidden units (gate-interleaved at pack time) for one half of K.  The
// batch is cut into independent groups (different batch rows never interact), each group with its own set of
// tiles x 2 CTAs, so that a group's per-step tensor-core work is tiny (N = 16..64 columns) and up to 4 x 32 = 128 SMs
// work at once.  Per step and CTA:
//   producer warp : waits for the group's step counter (every CTA of the group has published h_{s-1}), then pulls its
//                   K half of the packed h_{s-1} (bf16 hi/lo operand blocks, a few KB) with cp.async.bulk
//   MMA warp      : tcgen05.mma (bf16 x 3, fp32 accumulation in TMEM) against the resident weights
//   compute warps : TMEM -> registers; the columns the PEER CTA finalises are pushed into its shared memory
//                   (st.shared::cluster), the own ones staged locally; one cluster barrier; LSTM cell for
//                   (column, unit) pairs with c and h kept in REGISTERS across all steps; h is written to ctx (fp32) and
//                   to the group's packed operand buffer of the other parity; one arrival on the step counter
// The only device-wide traffic per step is the packed h (B x Hd x 4 bytes) and one counter per group.
#include <cuda_bf16.h>

#include "kernels.h"
#include "pack.cuh"

namespace sfb {

namespace {
constexpr int EBM = 128, EBK = 64;
constexpr int ENT = 320;                         // 8 compute warps + MMA issuer + producer
constexpr uint32_t ECORE = 128;
constexpr uint32_t ESBO = (EBK / 8) * ECORE;
constexpr uint32_t ELBO = ECORE;
constexpr uint32_t EA_HALF = (EBM / 8) * ESBO;   // 16 KB
constexpr int E_MAXKH = 4;                       // K blocks per CTA (Hd <= 512)
constexpr int E_MAXN = 64;                       // batch columns per group
constexpr int E_NP = (E_MAXN / 2) * 32 / 256;    // (column, unit) pairs per thread at most

__device__ __forceinline__ bool e_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 24); ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    __nanosleep(20);
  }
  return false;
}
__device__ __forceinline__ uint64_t e_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(ELBO >> 4) << 16;
  d |= (uint64_t)(ESBO >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void e_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void e_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void e_bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float e_sigmoid(float x) { return sigmoidf_fast(x); }
__device__ __forceinline__ float e_tanh(float x) { return tanhf_fast(x); }
}  // namespace

// grid = ndir * NG * tiles * 2, cluster (2,1,1), ENT threads
__global__ void __launch_bounds__(ENT, 1) encoder_persist_kernel(const EncPersistParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ int s_fail;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cid = blockIdx.x, rank = cid & 1, pairid = cid >> 1;
  const int tiles = q.Hd / 32, tile = pairid % tiles, grp = pairid / tiles, dir = grp / q.NG, bg = grp % q.NG;
  const int N = q.N, Hd = q.Hd, nkb = Hd / EBK;
  const int kh0 = rank == 0 ? 0 : nkb / 2, nkh = nkb / 2;
  const uint32_t b_half = (uint32_t)(N / 8) * ESBO, bblock = 2 * b_half;
  // batch columns (of the group's N) this CTA finalises: whole groups of 16, rank 0 the first ones
  const int ngrp = N / 16, g_split = (ngrp + 1) / 2;
  const int my_g0 = rank == 0 ? 0 : g_split, my_g1 = rank == 0 ? g_split : ngrp;
  const int cmax = g_split * 16;                                     // columns of the larger half (buffer stride)
  unsigned char* wbuf = smem;                                        // [nkh][32 KB] resident weights
  unsigned char* bst = smem + (size_t)E_MAXKH * 2 * EA_HALF;         // [nkh][bblock] packed h of the step
  float* mine = reinterpret_cast<float*>(bst + (size_t)E_MAXKH * bblock);   // [4 gates][cmax][32] own partial of the own columns
  float* park = mine + 4 * cmax * 32;                                // [4 gates][cmax][32] the peer's partial of the own columns
  uint64_t* wfull = reinterpret_cast<uint64_t*>(park + 4 * cmax * 32);
  uint64_t* bfull = wfull + 1;
  uint64_t* done = bfull + 1;
  const uint32_t tmem_cols = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 128;   // accumulator columns [0, N) and [N, 2N), see the MMA issuer
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(wfull, 1);
      mbar_init(bfull, 1);
      mbar_init(done, 1);
      s_fail = 0;
    }
    mbar_fence_init();
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = s_tmem;
  pdl_wait();   // packed weights and the hoisted input projection come from the launches before this one
  pdl_launch_dependents();

  const int maxlen = q.maxlen, B = q.B, Bg = q.Bg;
  const int ncta_grp = tiles * 2;
  unsigned long long* step_cnt = q.bar + grp;
  const size_t grp_bytes = (size_t)nkb * bblock;                     // one parity of one group's packed h
  unsigned char* hpk0 = q.hpk + (size_t)grp * 2 * grp_bytes;         // [2 parities][nkb][bblock]
  bool fail = false;
  // bring-up timeline of step 8 of CTA 0 (slots 4..): poll done, operand landed, MMAs retired, exchanged, cell done, arrived
  auto mark = [&](int s, int which) {
    if (q.trace && cid == 0 && s == 8) q.trace[which] = globaltimer_ns();
  };
  if (tid == 0) trace_mark(q.trace, 0);

  if (warp == 9) {
    // =============================== producer ===============================
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();
      mbar_expect_tx(wfull, (uint32_t)nkh * 2 * EA_HALF);
      const unsigned char* a_base = q.whh_pk[dir];
      for (int k = 0; k < nkh; ++k)
        bulk_g2s_hint(wbuf + (size_t)k * 2 * EA_HALF, a_base + ((size_t)tile * nkb + kh0 + k) * (2 * EA_HALF), 2 * EA_HALF, wfull, pol);
    }
    for (int s = 0; s < maxlen; ++s) {
      if (lane == 0) {
        // h_{s-1} of the whole group is published when every CTA of the group has arrived s times
        const unsigned long long target = (unsigned long long)ncta_grp * (unsigned long long)s;
        bool ok = s == 0;
        for (uint32_t i = 0; !ok && i < (1u << 24); ++i) {
          unsigned long long cur;
          asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(step_cnt) : "memory");
          ok = cur >= target;
        }
        fail = fail || !ok;
        mark(s, 4);
        asm volatile("fence.proxy.async;" ::: "memory");
        const unsigned char* src = hpk0 + (size_t)(s & 1) * grp_bytes + (size_t)kh0 * bblock;
        mbar_expect_tx(bfull, (uint32_t)nkh * bblock);
        for (int k = 0; k < nkh; ++k) bulk_g2s(bst + (size_t)k * bblock, src + (size_t)k * bblock, bblock, bfull);
      }
      __syncwarp();
      cluster_sync_all();   // the step's exchange barrier (all threads of both CTAs take part)
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    // bf16 x 3 with TWO instructions per K step instead of three: the packed operand holds the hi rows followed by the lo
    // rows, so A_hi x [B_hi | B_lo] is ONE N = 2N instruction whose accumulator columns [N, 2N) collect the A_hi x B_lo term
    // (added by the epilogue); A_lo x B_hi goes into columns [0, N).  The resident A tile — whose shared-memory read
    // paces these tiny-N instructions — is read twice per K step instead of three times.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(EBM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(EBM >> 4) << 24);
    fail = !e_wait(wfull, 0) || fail;
    // the operands sit at the same shared-memory addresses every step: descriptors once, per MMA only an add
    // (the address field counts 16-byte units: a K step of 16 elements = 2 core matrices = 256 bytes = +16)
    uint64_t dA[E_MAXKH][2], dB[E_MAXKH][2];
#pragma unroll
    for (int k = 0; k < E_MAXKH; ++k) {
      const uint32_t a_hi = smem_u32(wbuf + (size_t)k * 2 * EA_HALF), b_hi = smem_u32(bst + (size_t)k * bblock);
      dA[k][0] = e_desc(a_hi); dA[k][1] = e_desc(a_hi + EA_HALF);
      dB[k][0] = e_desc(b_hi); dB[k][1] = e_desc(b_hi + b_half);
    }
    for (int s = 0; s < maxlen; ++s) {
      fail = !e_wait(bfull, (uint32_t)s & 1u) || fail;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        mark(s, 5);
#pragma unroll
        for (int k = 0; k < E_MAXKH; ++k) {
          if (k < nkh) {
#pragma unroll
            for (int j = 0; j < EBK / 16; ++j) {
              const uint64_t ko = (uint64_t)(j * 2 * (int)ECORE / 16);
              e_umma(tmem_d, dA[k][0] + ko, dB[k][0] + ko, idesc2, (k > 0 || j > 0) ? 1u : 0u);   // A_hi x [B_hi | B_lo]
              e_umma(tmem_d, dA[k][1] + ko, dB[k][0] + ko, idesc, 1u);                              // A_lo x B_hi
            }
          }
        }
        e_commit(done);
      }
      __syncwarp();
      cluster_sync_all();
    }
  } else {
    // =============================== compute warps ===============================
    // TMEM lane r of the tile = gate r / 32 of hidden unit tile * 32 + r % 32 (pack.cu): warp (lq, hf) reads gate lq of
    // all 32 units for the 16-column groups gq with gq % 2 == hf
    const int lq = warp & 3, hf = warp >> 2;
    const uint32_t peer_park = dsmem_addr(park, (uint32_t)(rank ^ 1));
    // (column, unit) pairs of this thread: pair i -> column (tid + 256 i) / 32 of the own columns, unit = lane
    const int ncols = (my_g1 - my_g0) * 16;
    const int unit = tile * 32 + lane;
    float c_reg[E_NP], h_reg[E_NP], bsum[4];
    int len_reg[E_NP], m_reg[E_NP];
#pragma unroll
    for (int g = 0; g < 4; ++g) bsum[g] = __ldg(q.b_ih[dir] + g * Hd + unit) + __ldg(q.b_hh[dir] + g * Hd + unit);
#pragma unroll
    for (int i = 0; i < E_NP; ++i) {
      const int col = (tid + 256 * i) >> 5;                          // own-column index
      const int gcol = my_g0 * 16 + col;                             // column inside the group
      const int m = bg * Bg + gcol;                                  // batch row
      const bool okc = col < ncols && gcol < Bg && m < B;
      m_reg[i] = okc ? m : -1;
      len_reg[i] = okc ? q.lengths[m] : 0;
      c_reg[i] = 0.f;
      h_reg[i] = 0.f;
    }
    const float* xp = q.xproj[dir];
    const size_t state = (size_t)B * Hd;
    // row of the hoisted input projection for (batch row, time): position-indexed [B * maxlen] rows, or — eval mode,
    // no dropout on the embedding — the row of the per-token table T = Emb W_ih^T selected by the word id
    int row_next[E_NP];
    auto row_of = [&](int i, int t) -> int {
      if (m_reg[i] < 0 || t < 0 || t >= maxlen) return 0;
      const int pos = m_reg[i] * maxlen + t;
      return q.seq ? __ldg(q.seq + pos) : pos;
    };
#pragma unroll
    for (int i = 0; i < E_NP; ++i) row_next[i] = row_of(i, dir == 0 ? 0 : maxlen - 1);
    for (int s = 0; s < maxlen; ++s) {
      const int t = dir == 0 ? s : maxlen - 1 - s;
      // hoisted input projection of this step, requested before the recurrent product is awaited
      float add[E_NP][4];
#pragma unroll
      for (int i = 0; i < E_NP; ++i) {
        const bool act = m_reg[i] >= 0 && t < len_reg[i];
        const float* a = xp + (size_t)(act ? row_next[i] : 0) * 4 * Hd + unit;
#pragma unroll
        for (int g = 0; g < 4; ++g) add[i][g] = act ? __ldg(a + g * Hd) : 0.f;
        row_next[i] = row_of(i, dir == 0 ? t + 1 : t - 1);   // the word id of the next step, while this one computes
      }
      fail = !e_wait(done, (uint32_t)s & 1u) || fail;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) mark(s, 6);
      for (int gq = hf; gq < ngrp; gq += 2) {
        uint32_t v[16], v2[16];
        const uint32_t taddr = tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(gq * 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]),
              "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]), "=r"(v2[15])
            : "r"(taddr + (uint32_t)N));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));   // + the A_hi x B_lo term
        const bool own = gq >= my_g0 && gq < my_g1;
        if (own) {
          float* dst = mine + ((size_t)lq * cmax + (size_t)(gq - my_g0) * 16) * 32 + lane;
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[j * 32] = __uint_as_float(v[j]);
        } else {
          const int pg0 = rank == 0 ? g_split : 0;                   // first group of the peer's columns
          const uint32_t dst = peer_park + (uint32_t)(((lq * cmax + (gq - pg0) * 16) * 32 + lane) * 4);
#pragma unroll
          for (int j = 0; j < 16; ++j) dsmem_st_f32(dst + (uint32_t)(j * 32 * 4), __uint_as_float(v[j]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      cluster_sync_all();   // both partials of every column are with the CTA that finalises it
      if (tid == 0) mark(s, 7);
      unsigned char* hdst = hpk0 + (size_t)((s + 1) & 1) * grp_bytes;
#pragma unroll
      for (int i = 0; i < E_NP; ++i) {
        const int col = (tid + 256 * i) >> 5;
        if (col >= ncols) continue;
        const int gcol = my_g0 * 16 + col, m = m_reg[i];
        const bool act = m >= 0 && t < len_reg[i];
        float ig = 0.f, fg = 0.f, gt = 0.f, og = 0.f;
        if (act) {
          float g4[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float a = mine[((size_t)g * cmax + col) * 32 + lane], b = park[((size_t)g * cmax + col) * 32 + lane];
            // fixed summation order: the low K half (rank 0's partial) first, whichever CTA finalises
            g4[g] = (rank == 0 ? a + b : b + a) + (add[i][g] + bsum[g]);
          }
          ig = e_sigmoid(g4[0]); fg = e_sigmoid(g4[1]); gt = e_tanh(g4[2]); og = e_sigmoid(g4[3]);
          c_reg[i] = fg * c_reg[i] + ig * gt;
          h_reg[i] = og * e_tanh(c_reg[i]);
        }
        // packed h for the next step (every column of the group, so that the operand block is fully defined)
        {
          const float h1 = h_reg[i];
          const __nv_bfloat16 hi = __float2bfloat16_rn(h1);
          const __nv_bfloat16 lo = __float2bfloat16_rn(h1 - __bfloat162float(hi));
          unsigned char* d = hdst + (size_t)(unit >> 6) * bblock + (size_t)(gcol >> 3) * 1024 + (size_t)((unit & 63) >> 3) * 128 +
                             (size_t)(gcol & 7) * 16 + (size_t)(unit & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(d) = hi;
          *reinterpret_cast<__nv_bfloat16*>(d + b_half) = lo;
        }
        if (m >= 0) {
          q.ctx[(size_t)m * q.ld_ctx + (size_t)t * q.H + dir * Hd + unit] = act ? h_reg[i] : 0.f;
          if (q.tape_h) {
            q.tape_h[(size_t)(s + 1) * state + (size_t)m * Hd + unit] = h_reg[i];
            q.tape_c[(size_t)(s + 1) * state + (size_t)m * Hd + unit] = c_reg[i];
            if (act) {
              float* ga = q.tape_g + ((size_t)s * B + m) * 4 * Hd + unit;
              ga[0] = ig; ga[Hd] = fg; ga[2 * Hd] = gt; ga[3 * Hd] = og;
            }
          }
        }
      }
      if (s + 1 < maxlen) {
        if (tid == 0) mark(s, 8);
        e_bar256();   // every thread's stores are ordered before thread 0's fences + arrival (cumulativity)
        if (tid == 0) {
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence();
          atomicAdd(step_cnt, 1ull);
          mark(s, 9);
        }
      }
    }
    // final states (model.py:92-99)
#pragma unroll
    for (int i = 0; i < E_NP; ++i)
      if (m_reg[i] >= 0) {
        q.h_fin[(size_t)dir * state + (size_t)m_reg[i] * Hd + unit] = h_reg[i];
        q.c_fin[(size_t)dir * state + (size_t)m_reg[i] * Hd + unit] = c_reg[i];
      }
  }
  if (fail) s_fail = 1;
  cluster_sync_all();   // nobody pushes into a CTA that has gone away
  __syncthreads();
  if (s_fail && tid == 0) atomicExch(q.status, 1u);
  if (tid == 0) trace_mark(q.trace, 2);
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------ host side

EncPersistPlan encoder_persist_plan(int ndir, int Hd, int B, int num_sms) {
  EncPersistPlan pl{};
  pl.ok = false;
  if (Hd < 128 || Hd > 512 || (Hd % 128) != 0 || B < 1 || (ndir != 1 && ndir != 2)) return pl;
  const int tiles = Hd / 32, per_grp = tiles * 2;
  int ng = (num_sms / per_grp) / ndir;               // batch groups per direction that fit one resident wave
  if (ng < 1) return pl;
  if (ng > 4) ng = 4;
  const int need = (B + 15) / 16;                    // never more groups than 16-column blocks
  if (ng > need) ng = need;
  const int Bg = (B + ng - 1) / ng;
  const int N = (Bg + 15) & ~15;
  if (N > E_MAXN) return pl;                         // larger batches: the caller runs the batch in chunks
  pl.NG = ng; pl.Bg = Bg; pl.N = N;
  pl.grid = ndir * ng * per_grp;
  const size_t b_half = (size_t)(N / 8) * ESBO;
  const int cmax = ((N / 16 + 1) / 2) * 16;
  pl.smem = (size_t)E_MAXKH * 2 * EA_HALF + (size_t)E_MAXKH * 2 * b_half + 2 * (size_t)4 * cmax * 32 * sizeof(float) + 3 * sizeof(uint64_t) + 64;
  if (pl.smem > 227 * 1024 - 1024) return pl;
  pl.hpk_bytes = (size_t)ndir * ng * 2 * (Hd / EBK) * 2 * b_half;
  pl.bar_bytes = 256;
  pl.ok = true;
  return pl;
}

int32_t launch_encoder_persist(const EncPersistParams& q_in, cudaStream_t stream) {
  EncPersistParams q = q_in;
  const EncPersistPlan pl = encoder_persist_plan(q.ndir, q.Hd, q.B, device_num_sms());
  SFB_CHECK_ARG(pl.ok, "persistent encoder: unsupported shape");
  SFB_CHECK_ARG(q.hpk && q.bar && (reinterpret_cast<uintptr_t>(q.hpk) & 127u) == 0 && (reinterpret_cast<uintptr_t>(q.bar) & 15u) == 0,
                "persistent encoder: workspace");
  q.NG = pl.NG; q.Bg = pl.Bg; q.N = pl.N;
  q.status = reinterpret_cast<unsigned int*>(q.bar + 16);
  q.trace = next_trace_slot();
  // h_0 = 0 in packed form (parity 0 of every group) and the step counters
  SFB_CHECK_CUDA(cudaMemsetAsync(q.hpk, 0, pl.hpk_bytes, stream));
  SFB_CHECK_CUDA(cudaMemsetAsync(q.bar, 0, pl.bar_bytes, stream));
  static SmemMarks marks;
  SFB_CHECK_CUDA(ensure_dynamic_smem(encoder_persist_kernel, pl.smem, marks));
  SFB_CHECK_CUDA(launch_ex(encoder_persist_kernel, dim3(pl.grid, 1, 1), dim3(ENT, 1, 1), pl.smem, stream, dim3(2, 1, 1), q));
  count_launch();
  return 0;
}

}  // namespace sfb
