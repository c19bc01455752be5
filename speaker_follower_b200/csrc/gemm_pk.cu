// gemm_pk.cu — projections on Blackwell tensor cores from PRE-PACKED weights:  out[m, n] = sum_k X[m,k] W[n,k]
//
// The weight operand is packed once per weight version (pack.cu) into bf16 (hi, lo) pairs laid out exactly as
// tcgen05.mma wants them in shared memory (no-swizzle K-major core matrices), one contiguous 32 KB block per
// (128-row tile, 64-wide K block).  The kernel is then the canonical Blackwell pipeline with nothing to convert on
// the weight side:
//   warp 9  : producer — ONE cp.async.bulk (TMA engine) per stage for the weights (+ one for the activations when
//             they come pre-packed too), completing on the stage's "full" mbarrier; weight copies are issued BEFORE
//             the PDL dependency wait (weights are never written inside a step);
//   warp 8  : one elected thread issues tcgen05.mma kind::f16 (bf16x3: hi·hi + hi·lo + lo·hi, fp32 accumulate in
//             TMEM) and commits each stage to its "empty" mbarrier;
//   warps 0-7: when the activations are fp32 in global memory (the small projections: K-range <= 3 blocks per CTA)
//             they split them into bf16 hi/lo core matrices on the fly; afterwards they are the epilogue warps.
// UMMA M = 128 weight rows (for the LSTM gates: 4 gates x 32 hidden units, interleaved at pack time), UMMA N = the
// batch rounded up to 16, K = 16 per instruction, 64 per stage, 3 stages (~180 KB of loads in flight per SM).
// K is split over S CTAs per tile so that tiles x S ~ #SMs in ONE resident wave; partial accumulators go
// TMEM -> registers -> an L2-resident [col][row] partial buffer (coalesced), the S CTAs meet at a self-resetting
// semaphore and each reduces its share in split order (deterministic) and applies the epilogue (bias / tanh /
// addend / scale, or the whole LSTM cell update).
#include <cuda_bf16.h>

#include "epilogue.cuh"
#include "kernels.h"
#include "pack.cuh"

namespace sfb {

namespace {

constexpr int PBM = 128;   // weight rows per tile
constexpr int PBK = 64;    // K per stage
constexpr int PSTAGES = 3;
constexpr uint32_t PCORE = 128;
constexpr uint32_t PSBO = (PBK / 8) * PCORE;   // 1024: byte stride between 8-row groups
constexpr uint32_t PLBO = PCORE;               // byte stride between K-adjacent core matrices
constexpr uint32_t PA_HALF = (PBM / 8) * PSBO; // 16 KB: one (hi | lo) weight tile

__device__ __forceinline__ void pk_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 27); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();   // never hang the device
}
__device__ __forceinline__ uint64_t pk_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(PLBO >> 4) << 16;
  d |= (uint64_t)(PSBO >> 4) << 32;
  d |= 1ull << 46;   // descriptor version 1 (Blackwell); base_offset 0, SWIZZLE_NONE
  return d;
}
__device__ __forceinline__ void pk_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void pk_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace

// grid = (tiles, S, batch tiles), 320 threads, dynamic smem = PSTAGES stages + barriers
template <bool B_PACKED, bool HAS_XS>
__global__ void __launch_bounds__(320, 1) gemm_pk_kernel(const PkParams q) {
  extern __shared__ __align__(128) unsigned char smem[];
  const GemmParams& p = q.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, S = gridDim.y, rank = blockIdx.y, tiles = gridDim.x, z = blockIdx.z;
  const int NB = q.NB;
  const int m0 = z * q.rows_per_z, m_end = min(p.M, m0 + q.rows_per_z);
  const bool lstm = p.lstm.H > 0;

  const uint32_t b_half = (uint32_t)(NB / 8) * PSBO;          // bytes of one (hi | lo) activation tile
  const uint32_t stage_bytes = 2 * PA_HALF + 2 * b_half;
  const int NST = q.nstages;   // allocated stages (<= PSTAGES)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NST * stage_bytes);
  uint64_t* empty = full + PSTAGES;
  uint64_t* done = empty + PSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const uint32_t tmem_cols = NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < PSTAGES; ++s) {
        mbar_init(&full[s], B_PACKED ? 1 : 1 + 256);
        mbar_init(&empty[s], 1);
      }
      mbar_init(done, 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // K blocks of this CTA
  const int per = (q.nkb + S - 1) / S;
  const int kb_begin = rank * per, kb_end = min(q.nkb, kb_begin + per);
  const int nit = max(0, kb_end - kb_begin);

  trace_mark(p.trace, 0);
  if (!q.late_trigger) pdl_launch_dependents();
  unsigned int sem_gen = 0;

  if (warp == 9) {
    // =============================== producer: bulk copies ===============================
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();   // weights are re-read every step: keep them in L2
      const unsigned char* a_src = q.a_pk + ((size_t)tile * q.nkb + kb_begin) * (2 * PA_HALF);
      const unsigned char* b_src = B_PACKED ? q.b_pk + ((size_t)z * q.nkb + kb_begin) * (2 * b_half) : nullptr;
      const uint32_t tx = 2 * PA_HALF + (B_PACKED ? 2 * b_half : 0u);
      const int pre = min(nit, NST);
      for (int it = 0; it < pre; ++it) {   // weights first: not produced inside the step
        mbar_expect_tx(&full[it], tx);
        bulk_g2s_hint(smem + (size_t)it * stage_bytes, a_src + (size_t)it * (2 * PA_HALF), 2 * PA_HALF, &full[it], pol);
      }
      pdl_wait();
      if (q.late_trigger) pdl_launch_dependents();
      if (B_PACKED)
        for (int it = 0; it < pre; ++it)
          bulk_g2s(smem + (size_t)it * stage_bytes + 2 * PA_HALF, b_src + (size_t)it * (2 * b_half), 2 * b_half, &full[it]);
      for (int it = pre; it < nit; ++it) {
        const int s = it % NST, use = it / NST;
        pk_wait(&empty[s], (uint32_t)(use - 1) & 1u);
        mbar_expect_tx(&full[s], tx);
        bulk_g2s_hint(smem + (size_t)s * stage_bytes, a_src + (size_t)it * (2 * PA_HALF), 2 * PA_HALF, &full[s], pol);
        if (B_PACKED)
          bulk_g2s(smem + (size_t)s * stage_bytes + 2 * PA_HALF, b_src + (size_t)it * (2 * b_half), 2 * b_half, &full[s]);
      }
    } else {
      pdl_wait();
      if (q.late_trigger) pdl_launch_dependents();
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ===============================
    pdl_wait();
    if (q.late_trigger) pdl_launch_dependents();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(PBM >> 4) << 24);
    for (int it = 0; it < nit; ++it) {
      const int s = it % NST, use = it / NST;
      pk_wait(&full[s], (uint32_t)use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + PA_HALF, b_hi = a_hi + 2 * PA_HALF,
                       b_lo = b_hi + b_half;
#pragma unroll
        for (int j = 0; j < PBK / 16; ++j) {
          const uint32_t ko = (uint32_t)j * 2u * PCORE;
          const uint64_t dah = pk_desc(a_hi + ko), dal = pk_desc(a_lo + ko);
          const uint64_t dbh = pk_desc(b_hi + ko), dbl = pk_desc(b_lo + ko);
          pk_umma(tmem_d, dal, dbh, idesc, (it > 0 || j > 0) ? 1u : 0u);   // small terms first
          pk_umma(tmem_d, dah, dbl, idesc, 1u);
          pk_umma(tmem_d, dah, dbh, idesc, 1u);
        }
        pk_commit(&empty[s]);
        if (it + 1 == nit) pk_commit(done);
      }
      __syncwarp();
    }
  } else {
    // =============================== warps 0-7 ===============================
    pdl_wait();
    if (q.late_trigger) pdl_launch_dependents();
    trace_mark(p.trace, 1);
    if (!B_PACKED) {
      // fp32 activations -> bf16 hi/lo core matrices, one K block per stage of the ring
      const int r_in = lane & 7, kc_in = lane >> 3;
      float4 rb[2][4][2], rs[HAS_XS ? 2 : 1][HAS_XS ? 4 : 1][2];
      auto load_block = [&](int blk, int set) {
        int s = 0, cc = blk;
        while (s + 1 < p.nseg) {
          const int n = (p.seg[s].k + PBK - 1) / PBK;
          if (cc < n) break;
          cc -= n;
          ++s;
        }
        const GemmSeg& g = (q.alt_tile0 > 0 && tile >= q.alt_tile0) ? q.alt_seg : p.seg[s];
        const int kofs = cc * PBK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int wu = warp + 8 * i;
          const int rg = wu >> 1, k = kofs + ((wu & 1) * 4 + kc_in) * 8;
          const int m = m0 + rg * 8 + r_in;
          if (rg * 8 < NB && m < m_end && k < g.k) {
            const int xr = g.xrow ? g.xrow[m] : m;
            const float* src = g.x + (size_t)xr * g.ldx + k;
            const bool hi4 = k + 4 < g.k;   // K % 8 == 4: the last group is half full
            rb[set][i][0] = *reinterpret_cast<const float4*>(src);
            rb[set][i][1] = hi4 ? *reinterpret_cast<const float4*>(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.xadd) {
              const float* ap = g.xadd + (size_t)m * g.ldxadd + k;
              const float4 a0 = *reinterpret_cast<const float4*>(ap);
              const float4 a1 = hi4 ? *reinterpret_cast<const float4*>(ap + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              float4& r0 = rb[set][i][0];
              float4& r1 = rb[set][i][1];
              r0.x += a0.x; r0.y += a0.y; r0.z += a0.z; r0.w += a0.w;
              r1.x += a1.x; r1.y += a1.y; r1.z += a1.z; r1.w += a1.w;
              if (g.xtanh) {
                r0.x = tanhf(r0.x); r0.y = tanhf(r0.y); r0.z = tanhf(r0.z); r0.w = tanhf(r0.w);
                r1.x = tanhf(r1.x); r1.y = tanhf(r1.y); r1.z = tanhf(r1.z); r1.w = tanhf(r1.w);
              }
            }
            if (HAS_XS) {
              if (g.xs) {
                const float* sp = g.xs + (size_t)m * g.ldxs + k;
                rs[set][i][0] = __ldg(reinterpret_cast<const float4*>(sp));
                rs[set][i][1] = hi4 ? __ldg(reinterpret_cast<const float4*>(sp + 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
              } else {
                rs[set][i][0] = rs[set][i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
              }
            }
          } else {
            rb[set][i][0] = rb[set][i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_XS) rs[set][i][0] = rs[set][i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
          }
        }
      };
      if (nit > 0) load_block(kb_begin, 0);
#pragma unroll 1
      for (int it0 = 0; it0 < nit; it0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int it = it0 + u;
          if (it >= nit) break;
          if (it + 1 < nit) load_block(kb_begin + it + 1, u ^ 1);
          const int sidx = it % NST;
          if (it >= NST) pk_wait(&empty[sidx], (uint32_t)(it / NST - 1) & 1u);   // ring: the MMAs that read it retired
          unsigned char* st = smem + (size_t)sidx * stage_bytes + 2 * PA_HALF;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int wu = warp + 8 * i;
            if ((wu >> 1) * 8 < NB) {
              const uint32_t off = (uint32_t)(wu >> 1) * PSBO + (uint32_t)((wu & 1) * 4 + kc_in) * PCORE + (uint32_t)r_in * 16u;
              float4 b0 = rb[u][i][0], b1 = rb[u][i][1];
              if (HAS_XS) {
                const float4 s0 = rs[u][i][0], s1 = rs[u][i][1];
                b0.x *= s0.x; b0.y *= s0.y; b0.z *= s0.z; b0.w *= s0.w;
                b1.x *= s1.x; b1.y *= s1.y; b1.z *= s1.z; b1.w *= s1.w;
              }
              uint4 hi, lo;
              bf16_split8(b0, b1, hi, lo);
              *reinterpret_cast<uint4*>(st + off) = hi;
              *reinterpret_cast<uint4*>(st + b_half + off) = lo;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&full[sidx]);
        }
      }
    }
    if (tid == 0 && S > 1)   // generation of this tile's split-K barrier, read long before it is needed
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(sem_gen) : "l"(q.sem + 2 * (z * tiles + tile) + 1) : "memory");
    if (q.has_side) {   // side job while the tensor core works: pack another operand of the step (grid-strided)
      const long long total = (long long)q.side.ntile * q.side.nkb * q.side.R * 8;
      const long long cta = ((long long)z * S + rank) * tiles + tile, nthreads = (long long)tiles * S * gridDim.z * 256;
      for (long long idx = cta * 256 + tid; idx < total; idx += nthreads) pack_item(q.side, idx);
    }
  }
  trace_mark(p.trace, 4);

  // ---- epilogue 1: TMEM -> registers -> (direct output | [col][row] partial tile in L2)
  const bool direct = (S == 1) && !lstm;
  float* mypart = q.partial + ((size_t)(z * tiles + tile) * S + rank) * (size_t)(PBM * NB);
  if (nit > 0 && warp < 8) {
    pk_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp < 8) {
    // warp w may only read TMEM lanes 32*(w%4)..+31; warps w and w+4 split the columns
    const int lq = warp & 3, half = warp >> 2;
    const int row = lq * 32 + lane;
    const int ncols = m_end - m0;
    for (int c = half * 16; c < NB; c += 32) {
      if (c >= ncols) break;
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
      if (direct) {
        const int n = tile * PBM + row;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < ncols && n < p.N) plain_store(p, m0 + c + j, n, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < ncols) __stcg(mypart + (size_t)(c + j) * PBM + row, __uint_as_float(v[j]));   // lanes -> consecutive rows
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  trace_mark(p.trace, 5);
  if (!direct) {
    // epilogue operands of this thread's first-pass elements: requested now, consumed after the barrier
    const int ncols_e = m_end - m0;
    LstmPre1 lpre0, lpre1;
    PlainPre ppre0, ppre1;
    lpre0.ok = false; lpre1.ok = false; ppre0.ok = false; ppre1.ok = false;
    {
      if (lstm) {
        const int total = ncols_e * 32, share = (((total + S - 1) / S) + 31) & ~31;
        const int e_end = min(total, rank * share + share);
        const int e0 = rank * share + tid, e1 = e0 + 320;
        if (e0 < e_end) lpre0 = lstm_preload1(p, m0 + (e0 >> 5), tile * 32 + (e0 & 31));
        if (e1 < e_end) lpre1 = lstm_preload1(p, m0 + (e1 >> 5), tile * 32 + (e1 & 31));
      } else {
        const int total = ncols_e * (PBM / 4), share = (((total + S - 1) / S) + 7) & ~7;
        const int e_end = min(total, rank * share + share);
        const int e0 = rank * share + tid, e1 = e0 + 320;
        if (e0 < e_end) ppre0 = plain_preload(p, m0 + (e0 >> 5), tile * PBM + (e0 & 31) * 4);
        if (e1 < e_end) ppre1 = plain_preload(p, m0 + (e1 >> 5), tile * PBM + (e1 & 31) * 4);
      }
    }
    __threadfence();
    __syncthreads();
    // ---- the S CTAs of this tile meet at a sense-reversing barrier {count, generation} (all co-resident:
    // grid <= #SMs at 1 CTA/SM).  The last arriver resets the count BEFORE it releases the others, so nothing is left
    // to do on the exit path and the barrier is valid for any S on its next use.
    unsigned int* my_sem = q.sem + 2 * (z * tiles + tile);
    if (S > 1) {
      if (tid == 0) {
        const unsigned int old = atomicAdd(my_sem, 1u);
        if (old == (unsigned int)S - 1u) {
          atomicExch(my_sem, 0u);
          __threadfence();
          atomicAdd(my_sem + 1, 1u);
        } else {
          unsigned int gen = sem_gen;
          for (uint32_t i = 0; i < (1u << 26) && gen == sem_gen; ++i)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(my_sem + 1) : "memory");
          if (gen == sem_gen) __trap();
        }
        __threadfence();
      }
      __syncthreads();
    }
    trace_mark(p.trace, 6);
    // ---- epilogue 2: reduce the S partial tiles in split order + epilogue op; consecutive threads -> consecutive rows
    const int ncols = m_end - m0;
    const float* tbase = q.partial + (size_t)(z * tiles + tile) * S * (size_t)(PBM * NB);
    const size_t pstride = (size_t)PBM * NB;
    if (lstm) {
      // one (batch row, hidden unit) per thread: every thread of the CTA is busy with the transcendental-heavy cell
      // update; the 4 gates of a unit sit 32 rows apart in the tile; all S x 4 partial loads of a thread are issued
      // together (predicated full unroll), lanes -> consecutive units (coalesced)
      const int total = ncols * 32, share = (((total + S - 1) / S) + 31) & ~31;
      const int e_beg = rank * share, e_end = min(total, e_beg + share);
      for (int e = e_beg + tid; e < e_end; e += 320) {
        const int col = e >> 5, ul = e & 31;
        const float* pk = tbase + (size_t)col * PBM + ul;
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
        const int which = (e - e_beg - tid) / 320;
        if (which == 0) lstm_update1_pre(p, m0 + col, tile * 32 + ul, g4[0], g4[1], g4[2], g4[3], lpre0);
        else if (which == 1) lstm_update1_pre(p, m0 + col, tile * 32 + ul, g4[0], g4[1], g4[2], g4[3], lpre1);
        else lstm_update(p, m0 + col, tile * 32 + ul, g4[0], g4[1], g4[2], g4[3]);
      }
    } else {
      // float4 = 4 consecutive output features; two groups per thread per pass so all partial loads are in flight
      const int total = ncols * (PBM / 4), share = (((total + S - 1) / S) + 7) & ~7;
      const int e_beg = rank * share, e_end = min(total, e_beg + share);
      for (int e0 = e_beg + tid; e0 < e_end; e0 += 640) {
        const int e1 = e0 + 320;
        const bool has1 = e1 < e_end;
        const float* pa = tbase + (size_t)e0 * 4;
        const float* pb = tbase + (size_t)(has1 ? e1 : e0) * 4;
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
#pragma unroll 8
        for (int k = 0; k < S; ++k) {
          const float4 x = __ldcg(reinterpret_cast<const float4*>(pa + (size_t)k * pstride));
          const float4 y = __ldcg(reinterpret_cast<const float4*>(pb + (size_t)k * pstride));
          va.x += x.x; va.y += x.y; va.z += x.z; va.w += x.w;
          vb.x += y.x; vb.y += y.y; vb.z += y.z; vb.w += y.w;
        }
        if (e0 == e_beg + tid) {
          plain_store4_pre(p, m0 + (e0 >> 5), tile * PBM + (e0 & 31) * 4, va, ppre0);
          if (has1) plain_store4_pre(p, m0 + (e1 >> 5), tile * PBM + (e1 & 31) * 4, vb, ppre1);
        } else {
          plain_store4(p, m0 + (e0 >> 5), tile * PBM + (e0 & 31) * 4, va);
          if (has1) plain_store4(p, m0 + (e1 >> 5), tile * PBM + (e1 & 31) * 4, vb);
        }
      }
    }
    __syncthreads();
  } else {
    __syncthreads();
  }
  trace_mark(p.trace, 2);
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------ host side

int pk_num_kblocks(const int* seg_k, int nseg) {
  int n = 0;
  for (int s = 0; s < nseg; ++s) n += (seg_k[s] + PBK - 1) / PBK;
  return n;
}

PkPlan gemm_pk_plan(int M, int N_rows, int nkb, bool b_packed, int num_sms, bool wide) {
  PkPlan pl{};
  pl.tiles = (N_rows + PBM - 1) / PBM;
  // batch tiles of <= 128 rows (the weights are re-read per tile); very tall batches (per-episode projections over
  // B*L rows) use the widest UMMA N = 256 to halve that re-read
  const int zrows = (wide && b_packed && M > 2048) ? 256 : 128;
  pl.nz = (M + zrows - 1) / zrows;
  pl.rows_per_z = (M + pl.nz - 1) / pl.nz;
  pl.NB = (pl.rows_per_z + 15) & ~15;
  int s = num_sms / (pl.tiles * pl.nz);            // co-residency: tiles*S*nz <= #SMs (spin semaphore)
  if (s > nkb) s = nkb;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  (void)b_packed;
  pl.S = s;
  // fixed-size barrier region: projections of different geometry share one workspace, so the partial tiles must
  // start at the same offset for all of them (split-K happens only when tiles*nz <= #SMs/2, far below 512 pairs)
  pl.sem_bytes = 4096;
  if (pl.tiles * pl.nz > 512) pl.S = 1;
  pl.bytes = pl.sem_bytes + (((size_t)pl.tiles * pl.nz * pl.S * PBM * pl.NB * sizeof(float) + 255) & ~size_t(255));
  return pl;
}

size_t pk_weight_bytes(int N_rows, int nkb) { return (size_t)((N_rows + PBM - 1) / PBM) * nkb * 2 * PA_HALF; }
size_t pk_act_bytes(int M, int nkb, bool wide) {
  const int zrows = (wide && M > 2048) ? 256 : 128;
  const int nz = (M + zrows - 1) / zrows, rpz = (M + nz - 1) / nz, NB = (rpz + 15) & ~15;
  return (size_t)nz * nkb * 2 * (size_t)(NB / 8) * PSBO;
}

// bring-up: how many clusters of `cluster` CTAs of this kernel (320 threads, `smem` bytes) can be resident at once
int pk_max_active_clusters(int cluster, int smem) {
  auto kern = gemm_pk_kernel<true, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  if (cluster > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster * 64, 1, 1);
  cfg.blockDim = dim3(320, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -3; }
  return n;
}

int32_t launch_gemm_pk(const PkParams& q_in, cudaStream_t stream, void* ws, size_t ws_bytes) {
  PkParams q = q_in;
  GemmParams& p = q.g;
  p.trace = next_trace_slot();
  const bool b_packed = q.b_pk != nullptr;
  const bool lstm = p.lstm.H > 0;
  SFB_CHECK_ARG(q.a_pk && (reinterpret_cast<uintptr_t>(q.a_pk) & 127u) == 0, "gemm_pk: packed weights missing / misaligned");
  SFB_CHECK_ARG(p.M >= 1 && q.nkb >= 1, "gemm_pk: bad sizes");
  SFB_CHECK_ARG(!lstm || (p.lstm.H % 32) == 0, "gemm_pk: LSTM epilogue needs H % 32 == 0");
  const int n_rows = lstm ? 4 * p.lstm.H : p.N;
  bool has_xs = false;
  if (!b_packed) {
    SFB_CHECK_ARG(p.nseg >= 1 && p.nseg <= 3, "gemm_pk: 1..3 K segments");
    int kk[3];
    for (int s = 0; s < p.nseg; ++s) {
      const GemmSeg& g = p.seg[s];
      kk[s] = g.k;
      has_xs |= g.xs != nullptr;
      SFB_CHECK_ARG(g.x && (g.k % 4) == 0 && (g.ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(g.x) & 15u) == 0,
                    "gemm_pk: fp32 activations must be 16-byte aligned, K % 4 == 0");
      SFB_CHECK_ARG(!g.xs || ((reinterpret_cast<uintptr_t>(g.xs) & 15u) == 0 && (g.ldxs % 4) == 0), "gemm_pk: scale alignment");
      SFB_CHECK_ARG(!g.xadd || ((reinterpret_cast<uintptr_t>(g.xadd) & 15u) == 0 && (g.ldxadd % 4) == 0), "gemm_pk: addend alignment");
    }
    SFB_CHECK_ARG(pk_num_kblocks(kk, p.nseg) == q.nkb, "gemm_pk: K segments do not match the packed weights");
    SFB_CHECK_ARG(q.alt_tile0 == 0 || (p.nseg == 1 && q.alt_seg.x && q.alt_seg.k == p.seg[0].k && !q.alt_seg.xs && !p.seg[0].xs &&
                                       (q.alt_seg.ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(q.alt_seg.x) & 15u) == 0),
                  "gemm_pk: alternative activation source must match the single K segment");
  } else {
    SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(q.b_pk) & 127u) == 0, "gemm_pk: packed activations misaligned");
  }
  const PkPlan pl = gemm_pk_plan(p.M, n_rows, q.nkb, b_packed, device_num_sms(), q.wide != 0);
  const size_t ws_need = (pl.S == 1 && !lstm) ? pl.sem_bytes : pl.bytes;   // direct epilogue: no partial tiles
  SFB_CHECK_ARG(ws && ws_bytes >= ws_need && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "gemm_pk: workspace");
  if (q.has_side) SFB_PROPAGATE(pack_prepare(q.side));
  q.sem = static_cast<unsigned int*>(ws);
  q.partial = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.sem_bytes);
  q.NB = pl.NB;
  q.rows_per_z = pl.rows_per_z;
  const size_t stage_bytes = 2 * (size_t)PA_HALF + 2 * (size_t)(pl.NB / 8) * PSBO;
  // only as many stages as this launch can fill: a 1-block K range leaves room for the NEXT kernel's CTAs to become
  // resident early (PDL) and prefetch their own operands
  int nst = (q.nkb + pl.S - 1) / pl.S;
  if (nst > PSTAGES) nst = PSTAGES;
  while (nst > 1 && (size_t)nst * stage_bytes > 200 * 1024) --nst;   // wide batch tiles: fewer, larger stages
  q.nstages = nst;
  const size_t smem = (size_t)nst * stage_bytes + 8 * sizeof(uint64_t) + 16;
  const dim3 grid(pl.tiles, pl.S, pl.nz), block(320, 1, 1), cl(1, 1, 1);
#define SFB_PK_LAUNCH(BP, XS)                                                                                       \
  do {                                                                                                              \
    static SmemMarks marks;                                                                                         \
    SFB_CHECK_CUDA(ensure_dynamic_smem(gemm_pk_kernel<BP, XS>, smem, marks));                                       \
    SFB_CHECK_CUDA(launch_ex(gemm_pk_kernel<BP, XS>, grid, block, smem, stream, cl, q));                            \
  } while (0)
  if (b_packed) SFB_PK_LAUNCH(true, false);
  else if (has_xs) SFB_PK_LAUNCH(false, true);
  else SFB_PK_LAUNCH(false, false);
#undef SFB_PK_LAUNCH
  count_launch();
  return 0;
}

}  // namespace sfb
