// gemm_simt.cu — exact-fp32 "skinny" GEMM on the FFMA pipe:  out[M,N] = sum_s X_s[M,k_s] · W_s^T  (+epilogue)
//
// M is the batch (<= a few hundred rows), the weights are streamed once.  K may be the concatenation of
// up to three segments with their own activation/weight pointers, which is how cat(u_prev, feature)
// (model.py:391) and the two addmm's of nn.LSTMCell (model.py:393) run as ONE pass without materialising
// the concatenation.  Dropout keep-masks (model.py:392,394) are applied while the A tile is loaded, an
// embedding lookup (model.py:497) is an optional row indirection.
//
// B200 mapping: these GEMMs are tiny (<= 2.2 MB of weights) and sit on the step's dependency chain, so the
// goal is latency: a 128x32 output tile is split along K over the CTAs of a thread-block CLUSTER (up to 8),
// each CTA does a few 32-wide K chunks, and the partial tiles are reduced through distributed shared
// memory in a fixed order (deterministic, no atomics, no second kernel).  The epilogue is either
// bias + activation or — with gate-interleaved tiles — the whole LSTM cell update, so h1/c1 leave the
// GEMM directly.
#include "kernels.h"
#include "epilogue.cuh"

namespace sfb {

namespace {
constexpr int BN = 32, BK = 32, SA = 36, SB = 36, SR = 33;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace

// grid = (N tiles, splitk, M tiles); cluster = (1, splitk, 1)
template <int TM, bool KN>
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const GemmParams p) {
  constexpr int BM = 32 * TM;
  __shared__ __align__(16) float As[BM * SA];   // re-used as the [BM][SR] partial tile for the cluster reduction
  __shared__ __align__(16) float Bs[32 * SB];

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

  float4 pa[TM], pb;
  // weight row of local tile column `ty` (gate-interleaved for the LSTM epilogue: col = gate*8 + unit)
  const int wrow = lstm ? (ty >> 3) * p.lstm.H + blockIdx.x * 8 + (ty & 7) : n0 + ty;
  const bool wrow_ok = lstm ? (blockIdx.x * 8 + (ty & 7)) < p.lstm.H : wrow < p.N;

  auto load_chunk = [&](int c, bool do_a, bool do_b) {
    int s = 0, cc = c;
    while (s + 1 < p.nseg) {
      const int n = (p.seg[s].k + BK - 1) / BK;
      if (cc < n) break;
      cc -= n;
      ++s;
    }
    const GemmSeg& g = p.seg[s];
    const int kofs = cc * BK;
    if (do_a) {
      const int kk = kofs + tx * 4;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty + 32 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M && kk < g.k) {
          const int xr = g.xrow ? g.xrow[m] : m;
          v = *reinterpret_cast<const float4*>(g.x + (size_t)xr * g.ldx + kk);
          if (g.xs) {
            const float4 sc = ldg4(g.xs + (size_t)m * g.ldxs + kk);
            v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
          }
        }
        pa[i] = v;
      }
    }
    if (do_b) {
      pb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!KN) {
        const int kk = kofs + tx * 4;
        if (wrow_ok && kk < g.k) pb = ldg4(g.w + (size_t)wrow * g.ldw + kk);
      } else {
        const int kk = kofs + ty, nn = n0 + tx * 4;
        if (kk < g.k && nn < p.N) pb = ldg4(g.w + (size_t)kk * g.ldw + nn);
      }
    }
  };

  // PDL: weights are step inputs -> fetch the first weight chunk while the producer of the activations still runs
  pdl_launch_dependents();
  if (c_begin < c_end) load_chunk(c_begin, false, true);
  pdl_wait();
  if (c_begin < c_end) load_chunk(c_begin, true, false);
  for (int c = c_begin; c < c_end; ++c) {
#pragma unroll
    for (int i = 0; i < TM; ++i) *reinterpret_cast<float4*>(&As[(ty + 32 * i) * SA + tx * 4]) = pa[i];
    *reinterpret_cast<float4*>(&Bs[ty * SB + tx * 4]) = pb;
    __syncthreads();
    if (c + 1 < c_end) load_chunk(c + 1, true, true);
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[(ty + 32 * i) * SA + k4 * 4]);
      if (!KN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[(tx + 8 * j) * SB + k4 * 4]);
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
          const float4 b = *reinterpret_cast<const float4*>(&Bs[(k4 * 4 + kk) * SB + tx * 4]);
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
    __syncthreads();
  }

  // local tile column of accumulator j of this thread
  auto col_of = [&](int j) { return KN ? tx * 4 + j : tx + 8 * j; };

  if (S == 1) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty + 32 * i;
      if (m >= p.M) continue;
      if (lstm) {
        const int unit = blockIdx.x * 8 + tx;   // NT layout: accumulator j is gate j of unit tx
        if (unit < p.lstm.H) lstm_update(p, m, unit, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + col_of(j);
          if (n < p.N) plain_store(p, m, n, acc[i][j]);
        }
      }
    }
    return;
  }

  // ---- split-K: partial tile -> shared memory, reduce across the cluster through DSMEM in rank order
  float* red = As;   // [BM][SR]
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(ty + 32 * i) * SR + col_of(j)] = acc[i][j];
  cluster_sync_all();
  {
    const int rows_per = (BM + S - 1) / S;
    const int rbeg = rank * rows_per, rend = min(BM, rbeg + rows_per);
    uint32_t peer[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) peer[k] = k < S ? dsmem_addr(red, k) : 0u;
    if (lstm) {
      // one thread per (row, unit): the 4 gate columns unit, 8+unit, 16+unit, 24+unit
      for (int e = tid; e < (rend - rbeg) * 8; e += 256) {
        const int r = rbeg + (e >> 3), unit_l = e & 7;
        const int m = m0 + r, unit = blockIdx.x * 8 + unit_l;
        if (m >= p.M || unit >= p.lstm.H) continue;
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < S; ++k)
#pragma unroll
          for (int q = 0; q < 4; ++q) g4[q] += dsmem_ld_f32(peer[k] + (uint32_t)(r * SR + q * 8 + unit_l) * 4u);
        lstm_update(p, m, unit, g4[0], g4[1], g4[2], g4[3]);
      }
    } else {
      for (int e = tid; e < (rend - rbeg) * BN; e += 256) {
        const int r = rbeg + (e >> 5), cn = e & 31;
        const int m = m0 + r, n = n0 + cn;
        if (m >= p.M || n >= p.N) continue;
        float v = 0.f;
        for (int k = 0; k < S; ++k) v += dsmem_ld_f32(peer[k] + (uint32_t)(r * SR + cn) * 4u);
        plain_store(p, m, n, v);
      }
    }
  }
  cluster_sync_all();   // peers may still be reading this CTA's partial tile
}

int gemm_pick_splitk(int M, int N, int ktotal, int num_sms) {
  const int bm = M <= 32 ? 32 : 128;
  const int tiles = ((N + BN - 1) / BN) * ((M + bm - 1) / bm);
  const int nch = (ktotal + BK - 1) / BK;
  int want = (2 * num_sms + tiles - 1) / tiles;   // aim at ~2 CTAs per SM
  int s = 1;
  while (s * 2 <= want && s * 2 <= 8 && nch / (s * 2) >= 2) s *= 2;
  return s;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int TM, bool KN>
static int32_t launch_t(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  SFB_CHECK_CUDA(launch_ex(gemm_skinny_kernel<TM, KN>, grid, dim3(256, 1, 1), 0, stream, dim3(1, p.splitk, 1), p));
  count_launch();
  return 0;
}

int32_t launch_gemm(const GemmParams& p, cudaStream_t stream) {
  SFB_CHECK_ARG(p.nseg >= 1 && p.nseg <= 3, "gemm: 1..3 K segments");
  SFB_CHECK_ARG(p.M >= 1 && p.N >= 1, "gemm: bad sizes");
  SFB_CHECK_ARG(p.splitk == 1 || p.splitk == 2 || p.splitk == 4 || p.splitk == 8, "gemm: splitk must be 1, 2, 4 or 8");
  const int kn = p.seg[0].w_kn;
  const bool lstm = p.lstm.H > 0;
  SFB_CHECK_ARG(!lstm || (!kn && p.N == 4 * p.lstm.H && (p.lstm.H % 8) == 0), "gemm: LSTM epilogue needs [4H,K] weights, H % 8 == 0");
  SFB_CHECK_ARG(lstm || p.out, "gemm: output is NULL");
  for (int s = 0; s < p.nseg; ++s) {
    const GemmSeg& g = p.seg[s];
    SFB_CHECK_ARG(g.w_kn == kn, "gemm: mixed weight layouts");
    SFB_CHECK_ARG(g.k >= 4 && (g.k % 4) == 0, "gemm: K segment must be a multiple of 4");
    SFB_CHECK_ARG(g.x && g.w && aligned16(g.x) && aligned16(g.w) && (g.ldx % 4) == 0 && (g.ldw % 4) == 0,
                  "gemm: operands must be 16-byte aligned with leading dimensions % 4 == 0");
    SFB_CHECK_ARG(!g.xs || (aligned16(g.xs) && (g.ldxs % 4) == 0), "gemm: scale operand alignment");
    SFB_CHECK_ARG(!kn || (p.N % 4) == 0, "gemm: [K,N] weights need N % 4 == 0");
  }
  const int tm = p.M <= 32 ? 1 : 4;
  dim3 grid((p.N + BN - 1) / BN, p.splitk, (p.M + 32 * tm - 1) / (32 * tm));
  if (tm == 1) return kn ? launch_t<1, true>(p, grid, stream) : launch_t<1, false>(p, grid, stream);
  return kn ? launch_t<4, true>(p, grid, stream) : launch_t<4, false>(p, grid, stream);
}

}  // namespace sfb
