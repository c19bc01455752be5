// gemm_tc.cu — the LSTM-gate GEMM on Blackwell tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   gates[4H, B] = W_ih[:, seg0|seg1] · [x0 | x1]^T + W_hh · h0^T      (nn.LSTMCell, model.py:393)
//
// fp32 weights and activations are split ON THE FLY into bf16 (hi, lo) pairs and multiplied as
// hi·hi + hi·lo + lo·hi with fp32 accumulation in TMEM ("bf16x3"): relative error per product ~2^-17,
// i.e. ~1e-5 on a gate pre-activation, an order of magnitude inside the 1e-4 contract that single-pass
// TF32/bf16 miss (SURVEY.md §7 hard part 2).  Nothing is pre-converted or cached: the reference's
// nn.Parameter storage is read in place, every step.
//
// Mapping (one CTA = one UMMA tile):
//   UMMA M = 128 weight rows = 4 gates x 32 hidden units (gate-interleaved, so the epilogue owns whole cells),
//   UMMA N = batch rounded up to 16 (<= 256), UMMA K = 16 per instruction, 64 per pipeline stage.
//   K is split over S CTAs per tile so that tiles x S ~ number of SMs (16 x 9 = 144 on a B200), all resident in
//   ONE wave.  Partial accumulators go TMEM -> registers -> an L2-resident partial buffer; the S CTAs of a tile
//   meet at a self-resetting semaphore (they are co-resident: grid <= #SMs, 1 CTA/SM) and each reduces its share
//   in split order (deterministic) and runs the LSTM cell update, so h1 / c1 leave this kernel.
//   Warp roles: warps 0-7 are producers (coalesced 128-bit fp32 loads, next block prefetched into a second
//   register set, bf16 hi/lo split, 8x16B core matrices, no swizzle, K-major, fence.proxy.async, arrive on the
//   stage's "full" mbarrier); warp 8 issues the tcgen05.mma's and commits them to the stage's "empty" mbarrier
//   (3-stage ring).  Producers never wait for the issuer.
#include <cuda_bf16.h>

#include "epilogue.cuh"
#include "kernels.h"

namespace sfb {

namespace {

constexpr int TBM = 128;   // weight rows per CTA
constexpr int TBK = 64;    // K per stage
constexpr uint32_t CORE = 128;                 // bytes of one 8x16B core matrix
constexpr uint32_t SBO = (TBK / 8) * CORE;     // byte stride between 8-row groups (1024)
constexpr uint32_t LBO = CORE;                 // byte stride between K-adjacent core matrices

static int g_tc_debug = 0;   // bit0: swap LBO/SBO, bit1: descriptor version 0, bit2: phase timestamps (bring-up only)
__device__ long long g_tc_ts[256];
#define TS(i)                                                              \
  do {                                                                     \
    if ((dbg & 4) && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (i) < 256) \
      g_tc_ts[(i)] = clock64();                                            \
  } while (0)

__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 27); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();   // never hang the device: a lost commit is a bug, fail loudly
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int dbg) {
  uint64_t lbo = LBO >> 4, sbo = SBO >> 4;
  if (dbg & 1) { uint64_t t = lbo; lbo = sbo; sbo = t; }
  uint64_t d = (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= lbo << 16;
  d |= sbo << 32;
  if (!(dbg & 2)) d |= 1ull << 46;   // descriptor version 1 (Blackwell)
  return d;                          // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// split 8 consecutive fp32 into 8 bf16 "hi" and 8 bf16 "lo" (residual), packed in K order
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace

// grid = (4H/128, S, batch tiles), 288 threads (8 producer warps + 1 MMA-issuer warp), dynamic smem
template <bool HAS_XS>
__global__ void __launch_bounds__(288, 1) gemm_tc_lstm_kernel(const GemmParams p, const int NB, const int rows_per_z,
                                                             float* __restrict__ partial, unsigned int* sem,
                                                             const int dbg) {
  constexpr int STAGES = 3;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, S = gridDim.y, rank = blockIdx.y, tiles = gridDim.x;
  const int H = p.lstm.H;
  const int m0 = blockIdx.z * rows_per_z, m_end = min(p.M, m0 + rows_per_z);   // batch rows of this CTA

  const uint32_t a_bytes = (TBM / 8) * SBO;          // 16 KB per (hi|lo) A tile
  const uint32_t b_bytes = (uint32_t)(NB / 8) * SBO;  // per (hi|lo) B tile
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes);   // [STAGES] producers -> issuer
  uint64_t* empty = full + STAGES;                                                     // [STAGES] tensor core -> producers
  uint64_t* done = empty + STAGES;                                                     // [1] all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  // ---- one-time setup: barriers, TMEM allocation (warp 0), visible to everyone after the sync
  const uint32_t tmem_cols = NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 256);
        mbar_init(&empty[s], 1);
      }
      mbar_init(done, 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  TS(0);

  // ---- K blocks of this CTA
  int nblk = 0;
  for (int s = 0; s < p.nseg; ++s) nblk += (p.seg[s].k + TBK - 1) / TBK;
  const int per = (nblk + S - 1) / S;
  const int b_begin = rank * per, b_end = min(nblk, b_begin + per);
  const int nit = max(0, b_end - b_begin);

  trace_mark(p.trace, 0);
  pdl_launch_dependents();

  if (warp < 8) {
    // =============================== producers ===============================
    const int r_in = lane & 7, kc_in = lane >> 3;
    float4 ra[2][4][2], rb[2][4][2], rs[HAS_XS ? 2 : 1][HAS_XS ? 4 : 1][2];

    auto load_block = [&](int blk, int set, bool do_a, bool do_b) {
      int s = 0, cc = blk;
      while (s + 1 < p.nseg) {
        const int n = (p.seg[s].k + TBK - 1) / TBK;
        if (cc < n) break;
        cc -= n;
        ++s;
      }
      const GemmSeg& g = p.seg[s];
      const int kofs = cc * TBK;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int wu = warp + 8 * i;
        const int rg = wu >> 1, k = kofs + ((wu & 1) * 4 + kc_in) * 8;
        if (do_a) {   // weight rows, gate-interleaved: tile row = gate*32 + unit_local
          const int row = rg * 8 + r_in;
          const int wrow = (row >> 5) * H + tile * 32 + (row & 31);
          if (k < g.k) {
            const float* src = g.w + (size_t)wrow * g.ldw + k;
            ra[set][i][0] = ldg4(src);
            ra[set][i][1] = ldg4(src + 4);
          } else {
            ra[set][i][0] = ra[set][i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (do_b) {   // activations (batch rows), optional row indirection; dropout scale kept separate
          const int m = m0 + rg * 8 + r_in;
          if (rg * 8 < NB && m < m_end && k < g.k) {
            const int xr = g.xrow ? g.xrow[m] : m;
            const float* src = g.x + (size_t)xr * g.ldx + k;
            rb[set][i][0] = *reinterpret_cast<const float4*>(src);   // produced by the previous kernel: coherent loads
            rb[set][i][1] = *reinterpret_cast<const float4*>(src + 4);
            if (HAS_XS) {
              if (g.xs) {
                const float* sp = g.xs + (size_t)m * g.ldxs + k;
                rs[set][i][0] = ldg4(sp);
                rs[set][i][1] = ldg4(sp + 4);
              } else {
                rs[set][i][0] = rs[set][i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
              }
            }
          } else {
            rb[set][i][0] = rb[set][i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HAS_XS) rs[set][i][0] = rs[set][i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
          }
        }
      }
    };

    // PDL: weights are step inputs -> the first weight block is in flight before the activations' producer is done
    if (nit > 0) load_block(b_begin, 0, true, false);
    pdl_wait();
    trace_mark(p.trace, 1);
    TS(1);
    if (nit > 0) load_block(b_begin, 0, false, true);

    for (int it0 = 0; it0 < nit; it0 += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int it = it0 + u;
        if (it >= nit) break;
        const int s = it % STAGES, use = it / STAGES;
        if (it + 1 < nit) load_block(b_begin + it + 1, u ^ 1, true, true);   // next block's loads fly during this convert
        TS(8 + it * 8 + 0);
        if (use >= 1) mbar_wait_bounded(&empty[s], (uint32_t)(use - 1) & 1u);  // MMAs that read this stage retired
        TS(8 + it * 8 + 1);
        unsigned char* st = smem + (size_t)s * stage_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int wu = warp + 8 * i;
          const uint32_t off = (uint32_t)(wu >> 1) * SBO + (uint32_t)((wu & 1) * 4 + kc_in) * CORE + (uint32_t)r_in * 16u;
          uint4 hi, lo;
          split8(ra[u][i][0], ra[u][i][1], hi, lo);
          *reinterpret_cast<uint4*>(st + off) = hi;
          *reinterpret_cast<uint4*>(st + a_bytes + off) = lo;
          if ((wu >> 1) * 8 < NB) {
            float4 b0 = rb[u][i][0], b1 = rb[u][i][1];
            if (HAS_XS) {
              const float4 s0 = rs[u][i][0], s1 = rs[u][i][1];
              b0.x *= s0.x; b0.y *= s0.y; b0.z *= s0.z; b0.w *= s0.w;
              b1.x *= s1.x; b1.y *= s1.y; b1.z *= s1.z; b1.w *= s1.w;
            }
            split8(b0, b1, hi, lo);
            *reinterpret_cast<uint4*>(st + 2 * a_bytes + off) = hi;
            *reinterpret_cast<uint4*>(st + 2 * a_bytes + b_bytes + off) = lo;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(&full[s]);
        TS(8 + it * 8 + 2);
      }
    }
  } else {
    // =============================== MMA issuer (warp 8) ===============================
    pdl_wait();
    // instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=NB
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    for (int it = 0; it < nit; ++it) {
      const int s = it % STAGES, use = it / STAGES;
      mbar_wait_bounded(&full[s], (uint32_t)use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes,
                       b_lo = b_hi + b_bytes;
#pragma unroll
        for (int j = 0; j < TBK / 16; ++j) {
          const uint32_t ko = (uint32_t)j * 2u * CORE;   // 16 K elements = 2 core matrices
          const uint64_t dah = make_desc(a_hi + ko, dbg), dal = make_desc(a_lo + ko, dbg);
          const uint64_t dbh = make_desc(b_hi + ko, dbg), dbl = make_desc(b_lo + ko, dbg);
          umma_bf16(tmem_d, dah, dbh, idesc, (it > 0 || j > 0) ? 1u : 0u);
          umma_bf16(tmem_d, dah, dbl, idesc, 1u);
          umma_bf16(tmem_d, dal, dbh, idesc, 1u);
        }
        umma_commit(&empty[s]);
        if (it + 1 == nit) umma_commit(done);
      }
      __syncwarp();
    }
  }
  TS(2);

  // ---- epilogue 1: TMEM -> registers -> shared [128][NB+4] -> coalesced store of this CTA's partial tile into the
  // (L2-resident) partial buffer.  Stage memory is free: every MMA has retired when `done` flips.
  float* mypart = partial + ((size_t)(blockIdx.z * tiles + tile) * S + rank) * (size_t)(TBM * NB);
  const int NBS = NB + 4;
  float* stg = reinterpret_cast<float*>(smem);
  if (nit > 0) {
    mbar_wait_bounded(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp < 8) {
    // warp w reads TMEM lanes 32*(w%4).. (hardware restriction); warps w and w+4 split the columns
    const int lq = warp & 3, half = warp >> 2;
    float* prow = stg + (size_t)(lq * 32 + lane) * NBS;
    for (int c = half * 16; c < NB; c += 32) {
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
        for (int q = 0; q < 16; ++q) v[q] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(prow + c + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {
    const int nq4 = NB >> 2;
    for (int i = tid; i < TBM * nq4; i += 288) {
      const int row = i / nq4, c4 = i - row * nq4;
      __stcg(reinterpret_cast<float4*>(mypart) + i, *reinterpret_cast<const float4*>(stg + (size_t)row * NBS + c4 * 4));
    }
    __threadfence();
  }
  TS(3);
  __syncthreads();

  // ---- the S CTAs of this tile meet at a semaphore (all co-resident: grid <= #SMs at 1 CTA/SM)
  unsigned int* my_sem = sem + 2 * (blockIdx.z * tiles + tile);
  if (S > 1) {
    if (tid == 0) {
      atomicAdd(my_sem, 1u);
      unsigned int seen = 0;
      for (uint32_t i = 0; i < (1u << 26); ++i) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_sem) : "memory");
        if (seen >= (unsigned int)S) break;
      }
      if (seen < (unsigned int)S) __trap();
      __threadfence();
    }
    __syncthreads();
  }
  TS(4);

  // ---- epilogue 2: reduce the S partial tiles (split order) + LSTM cell update; one thread per (unit, batch row),
  // consecutive threads -> consecutive batch rows (coalesced partial reads)
  {
    const int total = 32 * NB, share = (total + S - 1) / S;
    const int e_beg = rank * share, e_end = min(total, e_beg + share);
    const float* tbase = partial + (size_t)(blockIdx.z * tiles + tile) * S * (size_t)(TBM * NB);
    for (int e = e_beg + tid; e < e_end; e += 288) {
      const int ul = e / NB, bm = e - ul * NB;
      const int m = m0 + bm;
      if (m >= m_end) continue;
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < S; ++k) {
        const float* pk = tbase + (size_t)k * (TBM * NB) + (size_t)ul * NB + bm;
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] += __ldcg(pk + (size_t)q * 32 * NB);
      }
      lstm_update(p, m, tile * 32 + ul, g[0], g[1], g[2], g[3]);
    }
  }
  TS(5);
  __syncthreads();
  if (S > 1 && tid == 0) {   // last CTA to leave re-arms the semaphore for the next launch
    const unsigned int gone = atomicAdd(my_sem + 1, 1u);
    if (gone == (unsigned int)S - 1) {
      atomicExch(my_sem + 1, 0u);
      atomicExch(my_sem, 0u);
    }
  }
  TS(6);
  trace_mark(p.trace, 2);
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------ host side

void gemm_tc_set_debug(int flags) { g_tc_debug = flags; }
int gemm_tc_read_timestamps(long long* out, int n) {
  if (n > 256) n = 256;
  return cudaMemcpyFromSymbol(out, g_tc_ts, n * sizeof(long long)) == cudaSuccess ? 0 : -1;
}

TcPlan gemm_tc_plan(int M, int H, int ktotal, int nseg, int num_sms) {
  TcPlan pl{};
  pl.tiles = (4 * H) / TBM;
  pl.nz = (M + 127) / 128;                              // batch tiles of <= 128 rows (weights re-read per tile)
  pl.rows_per_z = (M + pl.nz - 1) / pl.nz;
  pl.NB = (pl.rows_per_z + 15) & ~15;
  const int nblk_max = (ktotal + TBK - 1) / TBK + nseg;  // upper bound on K blocks
  int s = num_sms / (pl.tiles * pl.nz);                 // co-residency: tiles*S*nz <= #SMs (spin semaphore)
  if (s > nblk_max / 2) s = nblk_max / 2;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  pl.S = s;
  pl.sem_bytes = ((size_t)pl.tiles * pl.nz * 2 * sizeof(unsigned int) + 255) & ~size_t(255);
  pl.bytes = pl.sem_bytes + (size_t)pl.tiles * pl.nz * pl.S * TBM * pl.NB * sizeof(float);
  return pl;
}

bool gemm_tc_supported(const GemmParams& p) {
  if (p.lstm.H <= 0 || (p.lstm.H % 32) != 0 || p.N != 4 * p.lstm.H) return false;
  if (p.M < 1) return false;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& g = p.seg[s];
    if (g.w_kn || (g.k % 8) != 0 || (g.ldx % 4) != 0 || (g.ldw % 4) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.w)) & 15u) return false;
    if (g.xs && ((reinterpret_cast<uintptr_t>(g.xs) & 15u) || (g.ldxs % 4) != 0)) return false;
  }
  return true;
}

int32_t launch_gemm_tc(const GemmParams& p_in, cudaStream_t stream, void* ws, size_t ws_bytes) {
  SFB_CHECK_ARG(gemm_tc_supported(p_in), "gemm_tc: unsupported shape");
  GemmParams p = p_in;
  p.trace = next_trace_slot();
  int ktotal = 0;
  for (int s = 0; s < p.nseg; ++s) ktotal += p.seg[s].k;
  const TcPlan pl = gemm_tc_plan(p.M, p.lstm.H, ktotal, p.nseg, device_num_sms());
  SFB_CHECK_ARG(ws && ws_bytes >= pl.bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "gemm_tc: workspace");
  unsigned int* sem = static_cast<unsigned int*>(ws);
  float* partial = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.sem_bytes);
  const size_t stage_bytes = 2 * (size_t)(TBM / 8) * SBO + 2 * (size_t)(pl.NB / 8) * SBO;
  const size_t smem = 3 * stage_bytes + 8 * sizeof(uint64_t) + 16;
  bool has_xs = false;
  for (int s = 0; s < p.nseg; ++s) has_xs |= p.seg[s].xs != nullptr;
  static SmemMarks marks[2];
  if (has_xs) SFB_CHECK_CUDA(ensure_dynamic_smem(gemm_tc_lstm_kernel<true>, smem, marks[1]));
  else SFB_CHECK_CUDA(ensure_dynamic_smem(gemm_tc_lstm_kernel<false>, smem, marks[0]));
  const dim3 grid(pl.tiles, pl.S, pl.nz), cl(1, 1, 1);
  if (has_xs) SFB_CHECK_CUDA(launch_ex(gemm_tc_lstm_kernel<true>, grid, dim3(288, 1, 1), smem, stream, cl, p, pl.NB, pl.rows_per_z, partial, sem, g_tc_debug));
  else SFB_CHECK_CUDA(launch_ex(gemm_tc_lstm_kernel<false>, grid, dim3(288, 1, 1), smem, stream, cl, p, pl.NB, pl.rows_per_z, partial, sem, g_tc_debug));
  count_launch();
  return 0;
}

}  // namespace sfb
