// pack.cu — operand packing for gemm_pk.cu and the pack-time weight folding.
//
// pack_rows_kernel: a logical fp32 matrix X[r][k] (K = concatenation of up to three segments, optional elementwise
// scale = dropout keep mask, optional row indirection) -> bf16 (hi, lo) pairs in tcgen05's no-swizzle K-major
// core-matrix layout, one contiguous block per (row tile, 64-wide K block): [hi: R*128 B][lo: R*128 B] with
// element (r, k) at (r/8)*1024 + (k/8)*128 + (r%8)*16 + (k%8)*2.  Used once per weight version for the weights
// (R = 128; LSTM weights gate-interleaved so a tile holds the 4 gates of 32 hidden units) and once per step for the
// activations of the gate GEMM (R = batch rounded up to 16).
//
// fold_kernel (pack time only): out[n][j] = sum_d A[d][n] s[d] Bm[d][j], obias[n] = sum_d A[d][n] s[d] bv[d] —
// collapses the two chained projections of VisualSoftDotAttention (model.py:316-320) and EltwiseProdScoring
// (model.py:348-351) into one matrix each (DESIGN.md "Arithmetic re-association").
#include "pack.cuh"

namespace sfb {

// one thread per (tile, kblock, row, 8-wide k group)
__global__ void __launch_bounds__(256) pack_rows_kernel(const PackParams p) {
  trace_mark(p.trace, 0);
  pdl_launch_dependents();
  pdl_wait();
  trace_mark(p.trace, 1);
  const long long total = (long long)p.ntile * p.nkb * p.R * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    pack_item(p, idx);
  trace_mark(p.trace, 2);
}

int32_t pack_prepare(PackParams& p) {
  SFB_CHECK_ARG(p.out && (reinterpret_cast<uintptr_t>(p.out) & 127u) == 0, "pack: output missing / misaligned");
  SFB_CHECK_ARG(p.nseg >= 1 && p.nseg <= 3 && p.R >= 8 && (p.R % 8) == 0 && p.ntile >= 1, "pack: bad sizes");
  int nkb = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const PackSeg& g = p.seg[s];
    SFB_CHECK_ARG((g.k % 4) == 0 && (g.ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(g.x) & 15u) == 0,
                  "pack: source must be 16-byte aligned with K % 4 == 0");
    SFB_CHECK_ARG(!g.xs || ((reinterpret_cast<uintptr_t>(g.xs) & 15u) == 0 && (g.ldxs % 4) == 0), "pack: scale alignment");
    nkb += (g.k + 63) / 64;
  }
  p.nkb = nkb;
  return 0;
}

int32_t launch_pack_rows(const PackParams& p_in, cudaStream_t stream) {
  PackParams p = p_in;
  p.trace = next_trace_slot();
  SFB_PROPAGATE(pack_prepare(p));
  const long long total = (long long)p.ntile * p.nkb * p.R * 8;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SFB_CHECK_CUDA(launch_ex(pack_rows_kernel, dim3((unsigned)blocks, 1, 1), dim3(256, 1, 1), 0, stream, dim3(1, 1, 1), p));
  count_launch();
  return 0;
}

// out[n][j] = sum_d A[d*lda + n] * s[d] * Bm[d*ldb + j]   (n < NA, j < NJ), fp64 accumulation; pack time only.
// obias[n] = sum_d A[d*lda + n] * s[d] * bv[d] (+ obias_add) when obias != NULL.
__global__ void __launch_bounds__(256) fold_kernel(const float* __restrict__ A, int lda, const float* __restrict__ s,
                                                   const float* __restrict__ Bm, int ldb, const float* __restrict__ bv,
                                                   int D, int NA, int NJ, float* __restrict__ out, int ldo,
                                                   float* __restrict__ obias, const float* __restrict__ obias_add) {
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int n0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  double bacc = 0.0;
  for (int d0 = 0; d0 < D; d0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      const int d = d0 + i;
      const float sc = (d < D) ? (s ? s[d] : 1.f) : 0.f;
      As[i][tx] = (d < D && n0 + tx < NA) ? A[(size_t)d * lda + n0 + tx] * sc : 0.f;
      Bs[i][tx] = (d < D && j0 + tx < NJ) ? Bm[(size_t)d * ldb + j0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const double b = (double)Bs[i][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] += (double)As[i][ty + 8 * r] * b;
    }
    if (obias && blockIdx.x == 0 && ty == 0)
      for (int i = 0; i < 32; ++i)
        if (d0 + i < D) bacc += (double)As[i][tx] * (double)bv[d0 + i];
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int n = n0 + ty + 8 * r, j = j0 + tx;
    if (n < NA && j < NJ) out[(size_t)n * ldo + j] = (float)acc[r];
  }
  if (obias && blockIdx.x == 0 && ty == 0 && n0 + tx < NA)
    obias[n0 + tx] = (float)(bacc + (obias_add ? (double)obias_add[0] : 0.0));
}

// dst[c][r] = src[r][c]  (pack time only)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int lds, int R, int Cc,
                                                        float* __restrict__ dst, int ldd) {
  __shared__ float t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = ty; i < 32; i += 8)
    t[i][tx] = (r0 + i < R && c0 + tx < Cc) ? src[(size_t)(r0 + i) * lds + c0 + tx] : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < Cc && r0 + tx < R) dst[(size_t)(c0 + i) * ldd + r0 + tx] = t[tx][i];
}

int32_t launch_transpose(const float* src, int lds, int R, int Cc, float* dst, int ldd, cudaStream_t stream) {
  SFB_CHECK_ARG(src && dst && R >= 1 && Cc >= 1, "transpose: bad arguments");
  transpose_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32, 1), 256, 0, stream>>>(src, lds, R, Cc, dst, ldd);
  SFB_CHECK_LAUNCH();
  return 0;
}

int32_t launch_fold(const float* A, int lda, const float* s, const float* Bm, int ldb, const float* bv, int D, int NA,
                    int NJ, float* out, int ldo, float* obias, const float* obias_add, cudaStream_t stream) {
  SFB_CHECK_ARG(A && Bm && out && D >= 1 && NA >= 1 && NJ >= 1, "fold: bad arguments");
  SFB_CHECK_ARG(!obias || bv, "fold: bias vector missing");
  const dim3 grid((NJ + 31) / 32, (NA + 31) / 32, 1);
  fold_kernel<<<grid, 256, 0, stream>>>(A, lda, s, Bm, ldb, bv, D, NA, NJ, out, ldo, obias, obias_add);
  SFB_CHECK_LAUNCH();
  return 0;
}

}  // namespace sfb
