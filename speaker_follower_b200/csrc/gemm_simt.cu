// gemm_simt.cu — exact-fp32 "skinny" GEMM on the FFMA pipe:  out[M,N] = sum_s X_s[M,k_s] · W_s^T  (+bias, act)
//
// M is the batch (<= a few hundred rows), the weights are streamed once.  K may be the concatenation of
// up to three segments with their own activation/weight pointers, which is how cat(u_prev, feature)
// (model.py:391) and the two addmm's of nn.LSTMCell (model.py:393) run as ONE pass without materialising
// the concatenation.  Dropout keep-masks (model.py:392,394) are applied while the A tile is loaded, an
// embedding lookup (model.py:497) is an optional row indirection.  Split-K writes raw partial sums that
// the consumer kernel reduces in a fixed order (deterministic, no atomics).
//
// This is the general, always-available path (any M, N, K % 4 == 0).  The large LSTM-gate GEMM has a
// tensor-core (tcgen05) implementation in gemm_tc.cu that is used when its shape constraints hold.
#include "kernels.h"

namespace sfb {

namespace {
constexpr int BN = 32, BK = 32, SA = 36, SB = 36;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
}  // namespace

template <int TM, bool KN>
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const GemmParams p) {
  constexpr int BM = 32 * TM;
  __shared__ __align__(16) float As[BM * SA];
  __shared__ __align__(16) float Bs[32 * SB];

  const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
  const int m0 = blockIdx.z * BM, n0 = blockIdx.x * BN;

  int nch = 0;
  for (int s = 0; s < p.nseg; ++s) nch += (p.seg[s].k + BK - 1) / BK;
  const int per = (nch + p.splitk - 1) / p.splitk;
  const int c_begin = blockIdx.y * per, c_end = min(nch, c_begin + per);

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 pa[TM], pb;

  auto load_chunk = [&](int c) {
    int s = 0, cc = c;
    while (s + 1 < p.nseg) {
      const int n = (p.seg[s].k + BK - 1) / BK;
      if (cc < n) break;
      cc -= n;
      ++s;
    }
    const GemmSeg& g = p.seg[s];
    const int kofs = cc * BK;
    {
      const int kk = kofs + tx * 4;
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty + 32 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M && kk < g.k) {
          const int xr = g.xrow ? g.xrow[m] : m;
          v = ldg4(g.x + (size_t)xr * g.ldx + kk);
          if (g.xs) {
            const float4 sc = ldg4(g.xs + (size_t)m * g.ldxs + kk);
            v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
          }
        }
        pa[i] = v;
      }
    }
    pb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!KN) {
      const int n = n0 + ty, kk = kofs + tx * 4;
      if (n < p.N && kk < g.k) pb = ldg4(g.w + (size_t)n * g.ldw + kk);
    } else {
      const int kk = kofs + ty, nn = n0 + tx * 4;
      if (kk < g.k && nn < p.N) pb = ldg4(g.w + (size_t)kk * g.ldw + nn);
    }
  };

  if (c_begin < c_end) load_chunk(c_begin);
  for (int c = c_begin; c < c_end; ++c) {
#pragma unroll
    for (int i = 0; i < TM; ++i) *reinterpret_cast<float4*>(&As[(ty + 32 * i) * SA + tx * 4]) = pa[i];
    *reinterpret_cast<float4*>(&Bs[ty * SB + tx * 4]) = pb;
    __syncthreads();
    if (c + 1 < c_end) load_chunk(c + 1);
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

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + 32 * i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = KN ? n0 + tx * 4 + j : n0 + tx + 8 * j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.splitk > 1) {
        p.out[((size_t)blockIdx.y * p.M + m) * p.N + n] = v;
      } else {
        if (p.bias0) v += __ldg(p.bias0 + n);
        if (p.bias1) v += __ldg(p.bias1 + n);
        if (p.act == 1) v = tanhf(v);
        p.out[(size_t)m * p.ldo + n] = v;
      }
    }
  }
}

int gemm_pick_splitk(int M, int N, int ktotal, int num_sms) {
  const int bm = M <= 32 ? 32 : 128;
  const int tiles = ((N + BN - 1) / BN) * ((M + bm - 1) / bm);
  const int nch = (ktotal + BK - 1) / BK;
  int s = (3 * num_sms + tiles / 2) / tiles;
  if (s > nch / 4) s = nch / 4;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  return s;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int32_t launch_gemm(const GemmParams& p, cudaStream_t stream) {
  SFB_CHECK_ARG(p.nseg >= 1 && p.nseg <= 3, "gemm: 1..3 K segments");
  SFB_CHECK_ARG(p.M >= 1 && p.N >= 1 && p.splitk >= 1, "gemm: bad sizes");
  SFB_CHECK_ARG(p.splitk == 1 || (!p.bias0 && !p.bias1 && p.act == 0), "gemm: no epilogue with split-K");
  const int kn = p.seg[0].w_kn;
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
  if (tm == 1) {
    if (kn) gemm_skinny_kernel<1, true><<<grid, 256, 0, stream>>>(p);
    else    gemm_skinny_kernel<1, false><<<grid, 256, 0, stream>>>(p);
  } else {
    if (kn) gemm_skinny_kernel<4, true><<<grid, 256, 0, stream>>>(p);
    else    gemm_skinny_kernel<4, false><<<grid, 256, 0, stream>>>(p);
  }
  SFB_CHECK_LAUNCH();
  return 0;
}

}  // namespace sfb
