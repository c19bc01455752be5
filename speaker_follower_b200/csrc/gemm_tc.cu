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
//   K is split over the CTAs of a thread-block cluster (<= 8); partial accumulators go TMEM -> registers ->
//   shared memory and are reduced across the cluster through DSMEM in rank order (deterministic), then the
//   LSTM cell update runs in the same kernel and writes h1 / c1.
//   Operand staging: all 8 warps load fp32 with coalesced 128-bit loads (next stage prefetched in registers),
//   split, and store bf16 core matrices (8 rows x 16 B, no swizzle, K-major); one thread issues the MMAs and
//   commits them to the stage's mbarrier, which is what frees the stage for re-use (3-stage ring).
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

// grid = (4H/128, S, batch tiles), cluster (1, S, 1), 256 threads, dynamic smem
// HAS_XS: some K segment carries a dropout keep-mask (training); kept out of the eval instantiation so the
// prefetch loads have no consumer until the conversion phase (a predicated-off FMUL still waits on its inputs).
template <bool HAS_XS>
__global__ void __launch_bounds__(256, 1) gemm_tc_lstm_kernel(const GemmParams p, const int NB, const int STAGES,
                                                             const int rows_per_z, const int dbg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, S = p.splitk, rank = blockIdx.y;
  const int H = p.lstm.H;
  const int m0 = blockIdx.z * rows_per_z, m_end = min(p.M, m0 + rows_per_z);   // batch rows of this CTA

  const uint32_t a_bytes = (TBM / 8) * SBO;          // 16 KB per (hi|lo) A tile
  const uint32_t b_bytes = (uint32_t)(NB / 8) * SBO;  // per (hi|lo) B tile
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  unsigned char* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes);   // [STAGES] empty + [1] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + STAGES + 1);
  const int NBS = NB + 4;
  float* part = reinterpret_cast<float*>(smem);       // [128][NBS], aliases the stages after the last MMA

  // ---- one-time setup: barriers, TMEM allocation (warp 0), visible to everyone after the sync
  const uint32_t tmem_cols = NB <= 32 ? 32 : NB <= 64 ? 64 : NB <= 128 ? 128 : 256;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s <= STAGES; ++s) mbar_init(&bars[s], 1);
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

  // per-thread staging coordinates: warp-unit wu = warp + 8*i -> (row group, K half); lane -> (row in group, core)
  const int r_in = lane & 7, kc_in = lane >> 3;
  float4 ra[4][2], rb[4][2], rs[HAS_XS ? 4 : 1][2];

  auto load_block = [&](int blk, bool do_a, bool do_b) {
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
      // A operand: weight rows, gate-interleaved: tile row = gate*32 + unit_local
      if (do_a) {
        const int row = rg * 8 + r_in;
        const int wrow = (row >> 5) * H + tile * 32 + (row & 31);
        if (k < g.k) {
          const float* src = g.w + (size_t)wrow * g.ldw + k;
          ra[i][0] = ldg4(src);
          ra[i][1] = ldg4(src + 4);
        } else {
          ra[i][0] = ra[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      // B operand: activations (batch rows), optional row indirection and dropout scale
      if (do_b) {
        const int m = m0 + rg * 8 + r_in;
        if (rg * 8 < NB && m < m_end && k < g.k) {
          const int xr = g.xrow ? g.xrow[m] : m;
          const float* src = g.x + (size_t)xr * g.ldx + k;
          rb[i][0] = *reinterpret_cast<const float4*>(src);       // produced by the previous kernel: coherent loads
          rb[i][1] = *reinterpret_cast<const float4*>(src + 4);
          if (HAS_XS) {
            if (g.xs) {
              const float* sp = g.xs + (size_t)m * g.ldxs + k;
              rs[i][0] = ldg4(sp);
              rs[i][1] = ldg4(sp + 4);
            } else {
              rs[i][0] = rs[i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
            }
          }
        } else {
          rb[i][0] = rb[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (HAS_XS) rs[i][0] = rs[i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
      }
    }
  };

  // instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=NB
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);

  // PDL: weights are step inputs -> first weight block is in flight before the activations' producer has finished
  pdl_launch_dependents();
  if (b_begin < b_end) load_block(b_begin, true, false);
  pdl_wait();
  TS(1);
  if (b_begin < b_end) load_block(b_begin, false, true);
  for (int blk = b_begin; blk < b_end; ++blk) {
    const int it = blk - b_begin, s = it % STAGES, use = it / STAGES;
    TS(8 + it * 8 + 0);
    if (use >= 1) mbar_wait_bounded(&bars[s], (uint32_t)(use - 1) & 1u);   // MMAs that read this stage are done
    TS(8 + it * 8 + 1);
    unsigned char* st = stage_base + (size_t)s * stage_bytes;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int wu = warp + 8 * i;
      const uint32_t off = (uint32_t)(wu >> 1) * SBO + (uint32_t)((wu & 1) * 4 + kc_in) * CORE + (uint32_t)r_in * 16u;
      uint4 hi, lo;
      split8(ra[i][0], ra[i][1], hi, lo);
      *reinterpret_cast<uint4*>(st + off) = hi;
      *reinterpret_cast<uint4*>(st + a_bytes + off) = lo;
      if ((wu >> 1) * 8 < NB) {
        if (HAS_XS) {
          rb[i][0].x *= rs[i][0].x; rb[i][0].y *= rs[i][0].y; rb[i][0].z *= rs[i][0].z; rb[i][0].w *= rs[i][0].w;
          rb[i][1].x *= rs[i][1].x; rb[i][1].y *= rs[i][1].y; rb[i][1].z *= rs[i][1].z; rb[i][1].w *= rs[i][1].w;
        }
        split8(rb[i][0], rb[i][1], hi, lo);
        *reinterpret_cast<uint4*>(st + 2 * a_bytes + off) = hi;
        *reinterpret_cast<uint4*>(st + 2 * a_bytes + b_bytes + off) = lo;
      }
    }
    TS(8 + it * 8 + 2);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    TS(8 + it * 8 + 3);
    if (blk + 1 < b_end) load_block(blk + 1, true, true);         // next stage's global loads fly during the MMAs
    TS(8 + it * 8 + 4);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + a_bytes, b_hi = a_hi + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
      for (int j = 0; j < TBK / 16; ++j) {
        const uint32_t ko = (uint32_t)j * 2u * CORE;   // 16 K elements = 2 core matrices
        const uint64_t dah = make_desc(a_hi + ko, dbg), dal = make_desc(a_lo + ko, dbg);
        const uint64_t dbh = make_desc(b_hi + ko, dbg), dbl = make_desc(b_lo + ko, dbg);
        umma_bf16(tmem_d, dah, dbh, idesc, (it > 0 || j > 0) ? 1u : 0u);
        umma_bf16(tmem_d, dah, dbl, idesc, 1u);
        umma_bf16(tmem_d, dal, dbh, idesc, 1u);
      }
      umma_commit(&bars[s]);
      if (blk + 1 == b_end) umma_commit(&bars[STAGES]);
    }
    TS(8 + it * 8 + 5);
  }
  TS(2);

  // ---- epilogue: TMEM -> registers -> shared partial tile [128][NBS]
  if (b_begin < b_end) {
    mbar_wait_bounded(&bars[STAGES], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp < 4) {
    float* prow = part + (size_t)(warp * 32 + lane) * NBS;
    for (int c = 0; c < NB; c += 16) {
      uint32_t v[16];
      if (b_begin < b_end) {
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
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
  TS(3);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (S > 1) cluster_sync_all(); else __syncthreads();
  TS(4);

  // ---- reduce over the cluster (rank order) + LSTM cell update: CTA `rank` owns 32/S hidden units of the tile
  {
    const int units_per = 32 / S;
    const int nq = NB >> 2;                       // batch quads
    uint32_t peer[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) peer[k] = (S > 1 && k < S) ? dsmem_addr(part, k) : smem_u32(part);
    for (int e = tid; e < units_per * nq; e += 256) {
      const int ul = rank * units_per + e / nq, bq = e % nq;
      float g[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q][0] = g[q][1] = g[q][2] = g[q][3] = 0.f;
      for (int k = 0; k < S; ++k) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = dsmem_ld_f32x4(peer[k] + (uint32_t)((q * 32 + ul) * NBS + bq * 4) * 4u);
          g[q][0] += v.x; g[q][1] += v.y; g[q][2] += v.z; g[q][3] += v.w;
        }
      }
      const int unit = tile * 32 + ul;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + bq * 4 + j;
        if (m < m_end) lstm_update(p, m, unit, g[0][j], g[1][j], g[2][j], g[3][j]);
      }
    }
  }
  TS(5);
  if (S > 1) cluster_sync_all(); else __syncthreads();   // peers may still be reading this CTA's partial tile
  TS(6);
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

// ------------------------------------------------------------------ host side

void gemm_tc_set_debug(int flags) { g_tc_debug = flags; }
int gemm_tc_read_timestamps(long long* out, int n) {
  if (n > 256) n = 256;
  return cudaMemcpyFromSymbol(out, g_tc_ts, n * sizeof(long long)) == cudaSuccess ? 0 : -1;
}

static int tc_pick_splitk(const GemmParams& p) {
  int nblk = 0;
  for (int s = 0; s < p.nseg; ++s) nblk += (p.seg[s].k + TBK - 1) / TBK;
  int s = 1;
  while (s * 2 <= 8 && nblk / (s * 2) >= 2) s *= 2;
  return s;
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

int32_t launch_gemm_tc(const GemmParams& p_in, cudaStream_t stream) {
  SFB_CHECK_ARG(gemm_tc_supported(p_in), "gemm_tc: unsupported shape");
  GemmParams p = p_in;
  p.splitk = tc_pick_splitk(p);
  const int nz = (p.M + 127) / 128;                 // batch tiles of <= 128 rows (weights are re-read per tile)
  const int rows_per_z = (p.M + nz - 1) / nz;
  const int NB = (rows_per_z + 15) & ~15;
  const int stages = 3;
  const size_t stage_bytes = 2 * (size_t)(TBM / 8) * SBO + 2 * (size_t)(NB / 8) * SBO;
  size_t smem = stages * stage_bytes + (stages + 1) * sizeof(uint64_t) + 16;
  const size_t part_bytes = (size_t)TBM * (NB + 4) * sizeof(float);
  if (part_bytes + 64 > smem) smem = part_bytes + 64;
  SFB_CHECK_ARG(part_bytes <= stages * stage_bytes, "gemm_tc: partial tile does not fit the stage memory");
  bool has_xs = false;
  for (int s = 0; s < p.nseg; ++s) has_xs |= p.seg[s].xs != nullptr;
  static size_t configured[2] = {0, 0};
  if (smem > configured[has_xs]) {
    if (has_xs) SFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_lstm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_lstm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[has_xs] = smem;
  }
  const dim3 grid(p.N / TBM, p.splitk, nz), cl(1, p.splitk, 1);
  if (has_xs) SFB_CHECK_CUDA(launch_ex(gemm_tc_lstm_kernel<true>, grid, dim3(256, 1, 1), smem, stream, cl, p, NB, stages, rows_per_z, g_tc_debug));
  else SFB_CHECK_CUDA(launch_ex(gemm_tc_lstm_kernel<false>, grid, dim3(256, 1, 1), smem, stream, cl, p, NB, stages, rows_per_z, g_tc_debug));
  count_launch();
  return 0;
}

}  // namespace sfb
