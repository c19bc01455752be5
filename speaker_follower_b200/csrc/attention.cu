// attention.cu — soft dot-product attention over a set of rows, one thread-block CLUSTER per batch element.
//
// Computes, for every batch element b:   s_r = rows[b,r,:] . q[b,:]   (masked rows -> -inf)
//                                        alpha = softmax_r(s)          out[b,:] = sum_r alpha_r rows[b,r,:]
// which is the core of VisualSoftDotAttention.forward (model.py:320-325, rows = the 36-view feature slab,
// q = W_v^T (W_h h + b_h)) and of SoftDotAttention.forward (model.py:132-139, rows = ctx, q = W_in h).
//
// B200 mapping: the R rows of one batch element are split over the CL CTAs of a cluster.  Each CTA pulls
// its rows HBM -> shared memory with cp.async.bulk (TMA engine, one mbarrier per row, all copies in
// flight at once), so every row is read from HBM exactly once; masked rows are never fetched.  A warp
// computes a row's score as soon as that row's barrier flips.  The CTAs then exchange (max, sum,
// partial weighted sum) through distributed shared memory and each finalises a column slice of `out`.
#include "kernels.h"

namespace sfb {

template <int CL, int NQ>
__global__ void __launch_bounds__(256) soft_dot_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int D = p.D, nvec = D >> 2, RPC = p.rows_per_cta;
  const int r0 = rank * RPC;
  const int nrows = max(0, min(RPC, p.R - r0));
  const int rpad = (RPC + 3) & ~3;

  float* rows = reinterpret_cast<float*>(smem_raw);          // [RPC][D]
  float* part = rows + (size_t)RPC * D;                      // [D]   partial weighted sum of this CTA
  float* scratch = part + D;                                 // [1024] row-group partials (small D)
  float* sc = scratch + 1024;                                // [rpad] raw scores
  float* ew = sc + rpad;                                     // [rpad] exp(s - m_local)
  float* stat = ew + rpad;                                   // [4]    m_local, Z_local
  uint64_t* bars = reinterpret_cast<uint64_t*>(stat + 4);    // [RPC]

  const uint8_t* mrow = p.mask ? p.mask + (size_t)b * p.ldmask : nullptr;

  pdl_launch_dependents();   // let the next kernel of the step start its own prologue / prefetch
  // ---- 1. warp 0 arms one mbarrier per row and launches every bulk copy of this CTA at once
  if (warp == 0) {
    for (int r = lane; r < nrows; r += 32) mbar_init(&bars[r], 1);
    mbar_fence_init();
    __syncwarp();
    const size_t ba = (size_t)(p.idxA ? p.idxA[b] : b) * p.strideA_b;
    const size_t bb = (size_t)(p.idxB ? p.idxB[b] : b) * p.strideB_b;
    const uint64_t pol = policy_evict_first();
    for (int r = lane; r < nrows; r += 32) {
      const int gr = r0 + r;
      if (mrow && mrow[gr]) continue;
      mbar_expect_tx(&bars[r], (uint32_t)D * 4u);
      bulk_g2s_hint(rows + (size_t)r * D, p.segA + ba + (size_t)gr * p.strideA_r, (uint32_t)p.lenA * 4u, &bars[r], pol);
      if (p.lenB > 0)
        bulk_g2s_hint(rows + (size_t)r * D + p.lenA, p.segB + bb + (size_t)gr * p.strideB_r, (uint32_t)p.lenB * 4u,
                      &bars[r], pol);
    }
  }
  // ---- 2. query slice of this lane to registers (lane-strided float4).  The rows above are step inputs and
  // were requested before this point; q is produced by the preceding kernel, so wait for it only now.
  pdl_wait();
  float4 qv[NQ];
  {
    const float4* q4 = reinterpret_cast<const float4*>(p.q + (size_t)b * p.ldq);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const int idx = lane + 32 * j;
      qv[j] = idx < nvec ? q4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();  // barriers are initialised before anyone waits on them

  // ---- 3. scores: one warp per row, as soon as the row has landed
  for (int r = warp; r < nrows; r += 8) {
    float s = -INFINITY;
    const bool masked = mrow && mrow[r0 + r];
    if (!masked) {
      mbar_wait(&bars[r], 0);
      const float4* row4 = reinterpret_cast<const float4*>(rows + (size_t)r * D);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const int idx = lane + 32 * j;
        if (idx < nvec) {
          const float4 v = row4[idx];
          acc = fmaf(v.x, qv[j].x, acc);
          acc = fmaf(v.y, qv[j].y, acc);
          acc = fmaf(v.z, qv[j].z, acc);
          acc = fmaf(v.w, qv[j].w, acc);
        }
      }
      s = warp_sum(acc);
    }
    if (lane == 0) sc[r] = s;
  }
  __syncthreads();

  // ---- 4. local softmax statistics
  float m = -INFINITY;
  for (int r = 0; r < nrows; ++r) m = fmaxf(m, sc[r]);
  if (tid < nrows) ew[tid] = (sc[tid] == -INFINITY) ? 0.f : __expf(sc[tid] - m);
  __syncthreads();
  float Z = 0.f;
  for (int r = 0; r < nrows; ++r) Z += ew[r];
  if (tid == 0) {
    stat[0] = m;
    stat[1] = Z;
  }

  // ---- 5. partial weighted sum of this CTA's rows (masked rows were never loaded: skip, don't scale)
  {
    float4* part4 = reinterpret_cast<float4*>(part);
    const int G = (nvec < 256 && (256 % nvec) == 0) ? 256 / nvec : 1;
    if (G == 1) {
      for (int j = tid; j < nvec; j += 256) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < nrows; ++r) {
          const float e = ew[r];
          if (e != 0.f) {
            const float4 v = reinterpret_cast<const float4*>(rows + (size_t)r * D)[j];
            a.x = fmaf(e, v.x, a.x);
            a.y = fmaf(e, v.y, a.y);
            a.z = fmaf(e, v.z, a.z);
            a.w = fmaf(e, v.w, a.w);
          }
        }
        part4[j] = a;
      }
    } else {
      const int g = tid / nvec, j = tid - g * nvec;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = g; r < nrows; r += G) {
        const float e = ew[r];
        if (e != 0.f) {
          const float4 v = reinterpret_cast<const float4*>(rows + (size_t)r * D)[j];
          a.x = fmaf(e, v.x, a.x);
          a.y = fmaf(e, v.y, a.y);
          a.z = fmaf(e, v.z, a.z);
          a.w = fmaf(e, v.w, a.w);
        }
      }
      reinterpret_cast<float4*>(scratch)[tid] = a;
      __syncthreads();
      if (tid < nvec) {
        float4 t = reinterpret_cast<float4*>(scratch)[tid];
        for (int gg = 1; gg < G; ++gg) {
          const float4 v = reinterpret_cast<float4*>(scratch)[gg * nvec + tid];
          t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        part4[tid] = t;
      }
    }
  }

  // ---- 6. merge across the cluster through distributed shared memory
  if (CL > 1) cluster_sync_all(); else __syncthreads();

  float mk[CL], wk[CL];
  float M = -INFINITY;
#pragma unroll
  for (int k = 0; k < CL; ++k) {
    mk[k] = (CL > 1) ? dsmem_ld_f32(dsmem_addr(stat, k)) : stat[0];
    M = fmaxf(M, mk[k]);
  }
  float Zt = 0.f;
#pragma unroll
  for (int k = 0; k < CL; ++k) {
    const float zk = (CL > 1) ? dsmem_ld_f32(dsmem_addr(stat + 1, k)) : stat[1];
    wk[k] = (mk[k] == -INFINITY) ? 0.f : __expf(mk[k] - M);
    Zt = fmaf(zk, wk[k], Zt);
  }
  const float inv = 1.0f / Zt;

  {
    const int j0 = (int)(((long long)rank * nvec) / CL), j1 = (int)(((long long)(rank + 1) * nvec) / CL);
    float4* out4 = reinterpret_cast<float4*>(p.out + (size_t)b * p.ldo);
    for (int j = j0 + tid; j < j1; j += 256) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CL; ++k) {
        if (wk[k] != 0.f) {
          const float4 v = (CL > 1) ? dsmem_ld_f32x4(dsmem_addr(part, k) + (uint32_t)j * 16u)
                                    : reinterpret_cast<const float4*>(part)[j];
          const float w = wk[k] * inv;
          o.x = fmaf(w, v.x, o.x);
          o.y = fmaf(w, v.y, o.y);
          o.z = fmaf(w, v.z, o.z);
          o.w = fmaf(w, v.w, o.w);
        }
      }
      out4[j] = o;
    }
  }
  if (p.alpha) {
    const float scale = (m == -INFINITY) ? 0.f : __expf(m - M) * inv;
    for (int r = tid; r < nrows; r += 256) p.alpha[(size_t)b * p.ldalpha + r0 + r] = ew[r] * scale;
  }
  if (CL > 1) cluster_sync_all();  // keep this CTA's shared memory alive until every peer has read it
}

// ------------------------------------------------------------------ host launcher

static size_t attn_smem_bytes(int rpc, int D) {
  const int rpad = (rpc + 3) & ~3;
  return ((size_t)rpc * D + D + 1024 + 2 * rpad + 4) * sizeof(float) + (size_t)rpc * sizeof(uint64_t);
}

template <int CL, int NQ>
static int32_t launch_attn_t(const AttnParams& p, int B, cudaStream_t stream) {
  auto kern = soft_dot_attn_kernel<CL, NQ>;
  const size_t smem = attn_smem_bytes(p.rows_per_cta, p.D);
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    SFB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  SFB_CHECK_CUDA(launch_ex(kern, dim3(CL, B, 1), dim3(256, 1, 1), smem, stream, dim3(CL, 1, 1), p));
  count_launch();
  return 0;
}

int32_t launch_soft_dot_attention(AttnParams p, int B, cudaStream_t stream) {
  SFB_CHECK_ARG(p.D > 0 && (p.D % 4) == 0, "attention row length must be a positive multiple of 4");
  SFB_CHECK_ARG(p.D <= 2176, "attention row length > 2176 is not supported");
  SFB_CHECK_ARG((p.lenA % 4) == 0 && (p.lenB % 4) == 0 && p.lenA + p.lenB == p.D, "bad row segments");
  SFB_CHECK_ARG(p.R >= 1, "need at least one row");
  // smallest cluster whose per-CTA row block is <= 96 KB (so that >= 2 CTAs share an SM), capped at 8;
  // at least 4 CTAs per batch element when there are enough rows, so that B=100 fills 148 SMs.
  const size_t row_bytes = (size_t)p.D * 4;
  int cl = 1;
  while (cl < 8 && (((size_t)((p.R + cl - 1) / cl) * row_bytes > 96 * 1024) || (cl < 4 && p.R >= 8 * cl))) cl *= 2;
  p.rows_per_cta = (p.R + cl - 1) / cl;
  SFB_CHECK_ARG(attn_smem_bytes(p.rows_per_cta, p.D) <= 200 * 1024, "too many attention rows for shared memory");
  const bool small = p.D <= 512;
  switch (cl) {
    case 1: return small ? launch_attn_t<1, 4>(p, B, stream) : launch_attn_t<1, 17>(p, B, stream);
    case 2: return small ? launch_attn_t<2, 4>(p, B, stream) : launch_attn_t<2, 17>(p, B, stream);
    case 4: return small ? launch_attn_t<4, 4>(p, B, stream) : launch_attn_t<4, 17>(p, B, stream);
    default: return small ? launch_attn_t<8, 4>(p, B, stream) : launch_attn_t<8, 17>(p, B, stream);
  }
}

}  // namespace sfb
