// gemm_simt.cu — "skinny" GEMM:  out[M,N] = sum_s X_s[M,k_s] · W_s^T  (+epilogue), two arithmetic paths that share
// the tile movement: exact fp32 on the FFMA pipe, and 3xTF32 error-compensated mma.sync (hi·hi + hi·lo + lo·hi,
// fp32 accumulate, per-product error ~2^-21) for the projections on the step's dependency chain, where the FFMA
// pipe (64 FMA/clk/SM) — not memory — set the latency.
//
// M is the batch (<= a few hundred rows), the weights are streamed once.  K may be the concatenation of
// up to three segments with their own activation/weight pointers, which is how cat(u_prev, feature)
// (model.py:391) and the two addmm's of nn.LSTMCell (model.py:393) run as ONE pass without materialising
// the concatenation.  Dropout keep-masks (model.py:392,394) are applied while the A tile is read, an
// embedding lookup (model.py:497) is an optional row indirection.
//
// B200 mapping: these GEMMs are tiny (<= 2.2 MB of weights) and sit on the step's dependency chain, so the
// design goal is LATENCY, i.e. as few dependent memory round trips as possible:
//   * a 128x32 output tile is split along K over the CTAs of a thread-block CLUSTER (<= 8 CTAs);
//   * a CTA requests ALL its K chunks at once with cp.async (weights before the PDL dependency wait,
//     activations after), waits once, and runs the FFMA loop out of shared memory without further syncs;
//   * partial tiles are PUSHED into the owner CTA's shared memory (st.shared::cluster) and reduced there in
//     rank order after a single cluster barrier (deterministic, no atomics, no second kernel);
//   * the epilogue is bias + activation or — with gate-interleaved tiles — the whole LSTM cell update.
#include "epilogue.cuh"
#include "kernels.h"

namespace sfb {

namespace {
constexpr int BN = 32, BK = 32, SA = 36, SR = 33;
// weight-tile row stride: 36 ([N,k] tiles; FFMA and mma B fragments conflict-free), 40 for [k,N] tiles on the mma path
template <bool KN, bool TC> struct SBStride { static constexpr int v = (KN && TC) ? 40 : 36; };
constexpr int MAXC = 4;   // K chunks resident in shared memory per pass

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// x = hi + lo with hi, lo representable in TF32 (round-to-nearest)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace

// grid = (N tiles, splitk, M tiles); cluster = (1, splitk, 1); dynamic smem = MAXC*(As+Bs) [+ scale tiles] + recv
// TC (TM == 4 only): warp w owns rows 16w..16w+15 of the 128-row tile and all 32 columns (4 m16n8k8 tiles).
template <int TM, bool KN, bool HAS_XS, bool TC>
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const GemmParams p) {
  constexpr int BM = 32 * TM;
  constexpr int SB = SBStride<KN, TC>::v;
  static_assert(!TC || TM == 4, "mma path is built for the 128-row tile");
  extern __shared__ __align__(16) float dsm[];
  float* As = dsm;                                   // [MAXC][BM][SA]
  float* Bs = As + MAXC * BM * SA;                   // [MAXC][32][SB]
  float* Xs = Bs + MAXC * 32 * SB;                   // [MAXC][BM][SA] dropout scale tiles (HAS_XS only)
  float* recv = Xs + (HAS_XS ? MAXC * BM * SA : 0);  // [S][rows_per][SR] partial rows pushed by the cluster peers

  const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
  const int m0 = blockIdx.z * BM, n0 = blockIdx.x * BN;
  const int S = p.splitk, rank = blockIdx.y;
  const bool lstm = p.lstm.H > 0;

  int nch = 0;
  for (int s = 0; s < p.nseg; ++s) nch += (p.seg[s].k + BK - 1) / BK;
  const int per = (nch + S - 1) / S;
  const int c_begin = rank * per, c_end = min(nch, c_begin + per);

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // weight row of local tile column `ty` (gate-interleaved for the LSTM epilogue: col = gate*8 + unit)
  const int wrow = lstm ? (ty >> 3) * p.lstm.H + blockIdx.x * 8 + (ty & 7) : n0 + ty;
  const bool wrow_ok = lstm ? (blockIdx.x * 8 + (ty & 7)) < p.lstm.H : wrow < p.N;

  auto locate = [&](int c, int& kofs) -> const GemmSeg& {
    int s = 0, cc = c;
    while (s + 1 < p.nseg) {
      const int n = (p.seg[s].k + BK - 1) / BK;
      if (cc < n) break;
      cc -= n;
      ++s;
    }
    kofs = cc * BK;
    return p.seg[s];
  };
  // request the weight / activation tile of chunk c into slot `slot` (16-byte cp.async, zero-filled out of range)
  auto request_b = [&](int c, int slot) {
    int kofs;
    const GemmSeg& g = locate(c, kofs);
    float* dst = Bs + slot * 32 * SB + ty * SB + tx * 4;
    if (!KN) {
      const int kk = kofs + tx * 4;
      const bool ok = wrow_ok && kk < g.k;
      cp_async16(dst, ok ? g.w + (size_t)wrow * g.ldw + kk : g.w, ok);
    } else {
      const int kk = kofs + ty, nn = n0 + tx * 4;
      const bool ok = kk < g.k && nn < p.N;
      cp_async16(dst, ok ? g.w + (size_t)kk * g.ldw + nn : g.w, ok);
    }
  };
  auto request_a = [&](int c, int slot) {
    int kofs;
    const GemmSeg& g = locate(c, kofs);
    const int kk = kofs + tx * 4;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty + 32 * i;
      const bool ok = m < p.M && kk < g.k;
      const int xr = ok ? (g.xrow ? g.xrow[m] : m) : 0;
      cp_async16(As + slot * BM * SA + (ty + 32 * i) * SA + tx * 4, ok ? g.x + (size_t)xr * g.ldx + kk : g.x, ok);
      if (HAS_XS) {
        float* xd = Xs + slot * BM * SA + (ty + 32 * i) * SA + tx * 4;
        if (g.xs) cp_async16(xd, ok ? g.xs + (size_t)m * g.ldxs + kk : g.xs, ok);
        else *reinterpret_cast<float4*>(xd) = make_float4(1.f, 1.f, 1.f, 1.f);
      }
    }
  };

  trace_mark(p.trace, 0);
  pdl_launch_dependents();
  // PDL: weights are step inputs -> all weight chunks of the first pass are in flight before the dependency wait
  const int first = min(MAXC, c_end - c_begin);
  for (int k = 0; k < first; ++k) request_b(c_begin + k, k);
  pdl_wait();
  trace_mark(p.trace, 1);

  for (int c0 = c_begin; c0 < c_end; c0 += MAXC) {
    const int nc = min(MAXC, c_end - c0);
    if (c0 > c_begin) {
      __syncthreads();   // previous pass fully consumed
      for (int k = 0; k < nc; ++k) request_b(c0 + k, k);
    }
    for (int k = 0; k < nc; ++k) request_a(c0 + k, k);
    cp_async_wait_all();
    __syncthreads();
    trace_mark(p.trace, 4);
    for (int k = 0; k < nc; ++k) {
      const float* Ak = As + k * BM * SA;
      const float* Bk = Bs + k * 32 * SB;
      const float* Xk = Xs + k * BM * SA;
      if constexpr (TC) {
        const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
        if (m0 + 16 * warp < p.M) {   // warps whose 16 rows are all padding do nothing
#pragma unroll
          for (int k8 = 0; k8 < BK / 8; ++k8) {
            const int ra = (16 * warp + g) * SA + k8 * 8 + t;
            float av[4] = {Ak[ra], Ak[ra + 8 * SA], Ak[ra + 4], Ak[ra + 8 * SA + 4]};
            if (HAS_XS) {
              av[0] *= Xk[ra]; av[1] *= Xk[ra + 8 * SA]; av[2] *= Xk[ra + 4]; av[3] *= Xk[ra + 8 * SA + 4];
            }
            uint32_t ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_tf32(av[q], ah[q], al[q]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              float b0, b1;
              if (!KN) {
                b0 = Bk[(nt * 8 + g) * SB + k8 * 8 + t];
                b1 = Bk[(nt * 8 + g) * SB + k8 * 8 + t + 4];
              } else {
                b0 = Bk[(k8 * 8 + t) * SB + nt * 8 + g];
                b1 = Bk[(k8 * 8 + t + 4) * SB + nt * 8 + g];
              }
              uint32_t bh0, bl0, bh1, bl1;
              split_tf32(b0, bh0, bl0);
              split_tf32(b1, bh1, bl1);
              mma_tf32(acc[nt], al, bh0, bh1);   // small terms first
              mma_tf32(acc[nt], ah, bl0, bl1);
              mma_tf32(acc[nt], ah, bh0, bh1);
            }
          }
        }
      } else {
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
          float4 a[TM];
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            a[i] = *reinterpret_cast<const float4*>(&Ak[(ty + 32 * i) * SA + k4 * 4]);
            if (HAS_XS) {
              const float4 sc = *reinterpret_cast<const float4*>(&Xk[(ty + 32 * i) * SA + k4 * 4]);
              a[i].x *= sc.x; a[i].y *= sc.y; a[i].z *= sc.z; a[i].w *= sc.w;
            }
          }
          if (!KN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(&Bk[(tx + 8 * j) * SB + k4 * 4]);
#pragma unroll
              for (int i = 0; i < TM; ++i) {
                acc[i][j] = fmaf(a[i].x, b.x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b.y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b.z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b.w, acc[i][j]);
              }
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const float4 b = *reinterpret_cast<const float4*>(&Bk[(k4 * 4 + kk) * SB + tx * 4]);
#pragma unroll
              for (int i = 0; i < TM; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, b.x, acc[i][0]);
                acc[i][1] = fmaf(av, b.y, acc[i][1]);
                acc[i][2] = fmaf(av, b.z, acc[i][2]);
                acc[i][3] = fmaf(av, b.w, acc[i][3]);
              }
            }
          }
        }
      }
    }
  }
  trace_mark(p.trace, 5);

  // tile-local (row, column) of accumulator acc[i][j] of this thread
  auto row_of = [&](int i, int j) {
    if constexpr (TC) return 16 * (tid >> 5) + ((tid & 31) >> 2) + 8 * (j >> 1);
    else return ty + 32 * i;
  };
  auto col_of = [&](int i, int j) {
    if constexpr (TC) return i * 8 + 2 * (tid & 3) + (j & 1);
    else return KN ? tx * 4 + j : tx + 8 * j;
  };

  if (S == 1) {
    if (lstm) {
      // gate-interleaved tile: column = gate*8 + unit.  FFMA: acc[i][gate] of unit tx, row ty+32i.
      // mma: acc[gate][2*half+e] of unit 2t+e, row 16w+g+8*half.
      if constexpr (TC) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int m = m0 + row_of(0, 2 * hh), unit = blockIdx.x * 8 + 2 * (tid & 3) + e;
            if (m < p.M && unit < p.lstm.H)
              lstm_update(p, m, unit, acc[0][2 * hh + e], acc[1][2 * hh + e], acc[2][2 * hh + e], acc[3][2 * hh + e]);
          }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int m = m0 + ty + 32 * i, unit = blockIdx.x * 8 + tx;
          if (m < p.M && unit < p.lstm.H) lstm_update(p, m, unit, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int m = m0 + row_of(i, j), n = n0 + col_of(i, j);
          if (m < p.M && n < p.N) plain_store(p, m, n, acc[i][j]);
        }
    }
    trace_mark(p.trace, 2);
    return;
  }

  // ---- split-K: push every partial row to the CTA that owns it, one cluster barrier, reduce locally in rank order
  const int rows_per = (BM + S - 1) / S;
  {
    uint32_t peer[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) peer[k] = k < S ? dsmem_addr(recv, k) : 0u;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = row_of(i, j), owner = r / rows_per, lr = r - owner * rows_per;
        if (m0 + r < p.M)   // padding rows are never read back
          st_cluster_f32(peer[owner] + (uint32_t)((rank * rows_per + lr) * SR + col_of(i, j)) * 4u, acc[i][j]);
      }
  }
  trace_mark(p.trace, 6);
  cluster_sync_all();   // release/acquire: the pushed rows are visible to their owner
  trace_mark(p.trace, 7);
  {
    const int rbeg = rank * rows_per, rend = min(BM, rbeg + rows_per);
    if (lstm) {
      // one thread per (row, unit): the 4 gate columns unit, 8+unit, 16+unit, 24+unit
      for (int e = tid; e < (rend - rbeg) * 8; e += 256) {
        const int lr = e >> 3, unit_l = e & 7;
        const int m = m0 + rbeg + lr, unit = blockIdx.x * 8 + unit_l;
        if (m >= p.M || unit >= p.lstm.H) continue;
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < S; ++k)
#pragma unroll
          for (int q = 0; q < 4; ++q) g4[q] += recv[(k * rows_per + lr) * SR + q * 8 + unit_l];
        lstm_update(p, m, unit, g4[0], g4[1], g4[2], g4[3]);
      }
    } else {
      for (int e = tid; e < (rend - rbeg) * BN; e += 256) {
        const int lr = e >> 5, cn = e & 31;
        const int m = m0 + rbeg + lr, n = n0 + cn;
        if (m >= p.M || n >= p.N) continue;
        float v = 0.f;
        for (int k = 0; k < S; ++k) v += recv[(k * rows_per + lr) * SR + cn];
        plain_store(p, m, n, v);
      }
    }
  }
  trace_mark(p.trace, 2);
}

int gemm_pick_splitk(int M, int N, int ktotal, int num_sms) {
  const int bm = M <= 32 ? 32 : 128;
  const int tiles = ((N + BN - 1) / BN) * ((M + bm - 1) / bm);
  const int nch = (ktotal + BK - 1) / BK;
  int want = (num_sms + tiles - 1) / tiles;       // aim at one CTA per SM: the SM's FFMA pipe is the limit
  int s = 1;
  while (s * 2 <= want && s * 2 <= 8 && nch / (s * 2) >= 2) s *= 2;
  return s;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int TM, bool KN, bool HAS_XS, bool TC>
static int32_t launch_t(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int BM = 32 * TM;
  constexpr int SB = SBStride<KN, TC>::v;
  const int rows_per = (BM + p.splitk - 1) / p.splitk;
  const size_t tiles = ((size_t)MAXC * (BM * SA + 32 * SB) + (HAS_XS ? (size_t)MAXC * BM * SA : 0)) * sizeof(float);
  const size_t smem = tiles + (size_t)p.splitk * rows_per * SR * sizeof(float);
  static bool configured = false;   // per instantiation
  if (!configured) {
    const size_t mx = tiles + (size_t)(BM + 8) * SR * sizeof(float);
    SFB_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<TM, KN, HAS_XS, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx));
    configured = true;
  }
  SFB_CHECK_CUDA(launch_ex(gemm_skinny_kernel<TM, KN, HAS_XS, TC>, grid, dim3(256, 1, 1), smem, stream, dim3(1, p.splitk, 1), p));
  count_launch();
  return 0;
}

int32_t launch_gemm(const GemmParams& p_in, cudaStream_t stream) {
  GemmParams p = p_in;
  p.trace = next_trace_slot();
  SFB_CHECK_ARG(p.nseg >= 1 && p.nseg <= 3, "gemm: 1..3 K segments");
  SFB_CHECK_ARG(p.M >= 1 && p.N >= 1, "gemm: bad sizes");
  SFB_CHECK_ARG(p.splitk == 1 || p.splitk == 2 || p.splitk == 4 || p.splitk == 8, "gemm: splitk must be 1, 2, 4 or 8");
  const int kn = p.seg[0].w_kn;
  const bool lstm = p.lstm.H > 0;
  SFB_CHECK_ARG(!lstm || (!kn && p.N == 4 * p.lstm.H && (p.lstm.H % 8) == 0), "gemm: LSTM epilogue needs [4H,K] weights, H % 8 == 0");
  SFB_CHECK_ARG(lstm || p.out, "gemm: output is NULL");
  bool has_xs = false;
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& g = p.seg[s];
    has_xs |= g.xs != nullptr;
    SFB_CHECK_ARG(g.w_kn == kn, "gemm: mixed weight layouts");
    SFB_CHECK_ARG(g.k >= 4 && (g.k % 4) == 0, "gemm: K segment must be a multiple of 4");
    SFB_CHECK_ARG(g.x && g.w && aligned16(g.x) && aligned16(g.w) && (g.ldx % 4) == 0 && (g.ldw % 4) == 0,
                  "gemm: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    SFB_CHECK_ARG(!g.xs || (aligned16(g.xs) && (g.ldxs % 4) == 0), "gemm: scale operand alignment");
    SFB_CHECK_ARG(!kn || (p.N % 4) == 0, "gemm: [K,N] weights need N % 4 == 0");
  }
  const int tm = p.M <= 32 ? 1 : 4;
  dim3 grid((p.N + BN - 1) / BN, p.splitk, (p.M + 32 * tm - 1) / (32 * tm));
  if (tm == 1) {
    if (has_xs) return kn ? launch_t<1, true, true, false>(p, grid, stream) : launch_t<1, false, true, false>(p, grid, stream);
    return kn ? launch_t<1, true, false, false>(p, grid, stream) : launch_t<1, false, false, false>(p, grid, stream);
  }
  if (!g_disable_tc && !p.exact) {   // 3xTF32 mma.sync
    if (has_xs) return kn ? launch_t<4, true, true, true>(p, grid, stream) : launch_t<4, false, true, true>(p, grid, stream);
    return kn ? launch_t<4, true, false, true>(p, grid, stream) : launch_t<4, false, false, true>(p, grid, stream);
  }
  if (has_xs) return kn ? launch_t<4, true, true, false>(p, grid, stream) : launch_t<4, false, true, false>(p, grid, stream);
  return kn ? launch_t<4, true, false, false>(p, grid, stream) : launch_t<4, false, false, false>(p, grid, stream);
}

}  // namespace sfb
