// attention.cu — soft dot-product attention over a set of rows, streamed once from HBM.
//
// For every batch element b:   s_r = rows[b,r,:] . q[b,:]   (masked rows -> excluded)
//                              alpha = softmax_r(s)          out[b,:] = sum_r alpha_r rows[b,r,:]
// which is the core of VisualSoftDotAttention.forward (model.py:320-325: rows = the 36-view feature slab,
// q = W_v^T (W_h h + b_h)) and of SoftDotAttention.forward (model.py:132-139: rows = ctx, q = W_in h).
//
// B200 mapping.  The rows of one batch element are split over SPLIT CTAs (grid = SPLIT x B, all co-resident:
// <= 40 KB of shared memory each).  A CTA streams its rows HBM -> shared memory through a ring of bulk async
// copies (cp.async.bulk, TMA engine, one mbarrier per stage, evict-first, masked rows never fetched; in gather
// mode a row is two copies: feature table + orientation table).  Each thread keeps its slice of the row in
// registers between the score and the accumulation, so a row is read from shared memory once: block-wide dot
// (warp shuffles + one __syncthreads), online softmax, FMA into the running weighted sum.  The SPLIT partial
// results (max, sum, weighted sum) meet in an L2-resident buffer; the last CTA to arrive (atomic ticket,
// self-resetting) merges them.  With PDL the ring is primed before the producer of q has finished.
#include "kernels.h"

namespace sfb {

namespace {
constexpr int ATT_THREADS = 256;
constexpr int ATT_MAX_ROWS = 64;   // rows per CTA
constexpr int ATT_MAX_STAGES = 32; // ring depth == rows processed per pass
}

template <int NJ>   // float4 slices per thread: D <= NJ * 1024
__global__ void __launch_bounds__(ATT_THREADS) soft_dot_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, split = blockIdx.x, SPLIT = gridDim.x;
  const int D = p.D, nvec = D >> 2, NSTG = p.stages;
  const int r0 = split * p.rows_per_cta;
  const int nrows = max(0, min(p.rows_per_cta, p.R - r0));

  float* ring = reinterpret_cast<float*>(smem_raw);                   // [NSTG][D]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)NSTG * D);   // [NSTG]
  float* red = reinterpret_cast<float*>(full + NSTG);                 // [8][ATT_MAX_STAGES] per-warp partial dots
  float* sc = red + 8 * ATT_MAX_STAGES;                               // [ATT_MAX_ROWS] raw scores of this CTA's rows
  int* list = reinterpret_cast<int*>(sc + ATT_MAX_ROWS);              // [ATT_MAX_ROWS] unmasked local row ids
  __shared__ int s_nvalid, s_last;

  const uint8_t* mrow = p.mask ? p.mask + (size_t)b * p.ldmask : nullptr;

  trace_mark(p.trace, 0);
  pdl_launch_dependents();   // let the next kernel of the step start its own prologue / prefetch

  // ---- 1. warp 0: compact the unmasked rows, arm the ring, launch the first NSTG copies
  size_t ba = 0, bb = 0;
  uint64_t pol = 0;
  if (warp == 0) {
    int n = 0;
    for (int base = 0; base < nrows; base += 32) {
      const int r = base + lane;
      const bool ok = r < nrows && !(mrow && mrow[r0 + r]);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) list[n + __popc(bal & ((1u << lane) - 1u))] = r;
      n += __popc(bal);
    }
    if (lane == 0) s_nvalid = n;
    if (lane < NSTG) mbar_init(&full[lane], 1);
    mbar_fence_init();
    __syncwarp();
    ba = (size_t)(p.idxA ? p.idxA[b] : b) * p.strideA_b;
    bb = (size_t)(p.idxB ? p.idxB[b] : b) * p.strideB_b;
    pol = policy_evict_first();
    if (lane < NSTG && lane < n) {
      const int gr = r0 + list[lane];
      float* dst = ring + (size_t)lane * D;
      mbar_expect_tx(&full[lane], (uint32_t)D * 4u);
      bulk_g2s_hint(dst, p.segA + ba + (size_t)gr * p.strideA_r, (uint32_t)p.lenA * 4u, &full[lane], pol);
      if (p.lenB > 0)
        bulk_g2s_hint(dst + p.lenA, p.segB + bb + (size_t)gr * p.strideB_r, (uint32_t)p.lenB * 4u, &full[lane], pol);
    }
  }
  for (int r = tid; r < nrows; r += ATT_THREADS) sc[r] = -INFINITY;
  // ---- 2. q is produced by the preceding kernel: wait for it only now (rows above are step inputs)
  pdl_wait();
  trace_mark(p.trace, 1);
  float4 qv[NJ], acc[NJ];
  {
    const float4* q4 = reinterpret_cast<const float4*>(p.q + (size_t)b * p.ldq);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int idx = tid + ATT_THREADS * j;
      qv[j] = idx < nvec ? q4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  const int nvalid = s_nvalid;
  trace_mark(p.trace, 4);

  // ---- 3. stream the rows in passes of up to NSTG rows (the whole ring): all rows of a pass are in flight
  // together, so a pass costs one memory round trip.  Pass = dots for every row (warp shuffles, per-warp partials
  // in smem) | scores | online-softmax update with the rows re-read from the ring | refill the ring.
  float m = -INFINITY, Z = 0.f;
  for (int i0 = 0; i0 < nvalid; i0 += NSTG) {
    const int nb = min(NSTG, nvalid - i0);
    const uint32_t parity = (uint32_t)(i0 / NSTG) & 1u;
    for (int r = 0; r < nb; ++r) {
      mbar_wait(&full[r], parity);
      const float4* row4 = reinterpret_cast<const float4*>(ring + (size_t)r * D);
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int idx = tid + ATT_THREADS * j;
        if (idx < nvec) {
          const float4 v = row4[idx];
          part = fmaf(v.x, qv[j].x, part);
          part = fmaf(v.y, qv[j].y, part);
          part = fmaf(v.z, qv[j].z, part);
          part = fmaf(v.w, qv[j].w, part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) red[warp * ATT_MAX_STAGES + r] = part;
    }
    __syncthreads();
    if (tid < nb) {
      float sr = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sr += red[w * ATT_MAX_STAGES + tid];
      sc[list[i0 + tid]] = sr;
    }
    __syncthreads();
    float mb = m;
    for (int r = 0; r < nb; ++r) mb = fmaxf(mb, sc[list[i0 + r]]);
    const float corr = __expf(m - mb);
    Z *= corr;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      acc[j].x *= corr; acc[j].y *= corr; acc[j].z *= corr; acc[j].w *= corr;
    }
    for (int r = 0; r < nb; ++r) {
      const float e = __expf(sc[list[i0 + r]] - mb);
      Z += e;
      const float4* row4 = reinterpret_cast<const float4*>(ring + (size_t)r * D);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int idx = tid + ATT_THREADS * j;
        if (idx < nvec) {
          const float4 v = row4[idx];
          acc[j].x = fmaf(e, v.x, acc[j].x);
          acc[j].y = fmaf(e, v.y, acc[j].y);
          acc[j].z = fmaf(e, v.z, acc[j].z);
          acc[j].w = fmaf(e, v.w, acc[j].w);
        }
      }
    }
    m = mb;
    if (i0 + NSTG < nvalid) {
      __syncthreads();   // every thread is done with the ring -> refill it (one lane per row)
      if (warp == 0) {
        const int i = i0 + NSTG + lane;
        if (lane < NSTG && i < nvalid) {
          const int gr = r0 + list[i];
          float* dst = ring + (size_t)lane * D;
          mbar_expect_tx(&full[lane], (uint32_t)D * 4u);
          bulk_g2s_hint(dst, p.segA + ba + (size_t)gr * p.strideA_r, (uint32_t)p.lenA * 4u, &full[lane], pol);
          if (p.lenB > 0)
            bulk_g2s_hint(dst + p.lenA, p.segB + bb + (size_t)gr * p.strideB_r, (uint32_t)p.lenB * 4u, &full[lane], pol);
        }
      }
    }
  }
  __syncthreads();   // sc[] complete
  trace_mark(p.trace, 5);

  if (SPLIT == 1) {
    const float inv = 1.0f / Z;
    float4* out4 = reinterpret_cast<float4*>(p.out + (size_t)b * p.ldo);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int idx = tid + ATT_THREADS * j;
      if (idx < nvec) out4[idx] = make_float4(acc[j].x * inv, acc[j].y * inv, acc[j].z * inv, acc[j].w * inv);
    }
    if (p.alpha)
      for (int r = tid; r < nrows; r += ATT_THREADS)
        p.alpha[(size_t)b * p.ldalpha + r0 + r] = sc[r] == -INFINITY ? 0.f : __expf(sc[r] - m) * inv;
    trace_mark(p.trace, 2);
    return;
  }

  // ---- 4. publish this CTA's partial (m, Z, weighted sum) and its raw scores; the last arriver merges
  const int PS = D + 4;   // floats per partial record
  float* mine = p.part + ((size_t)b * SPLIT + split) * PS;
  if (tid == 0) {
    mine[0] = m;
    mine[1] = Z;
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int idx = tid + ATT_THREADS * j;
    if (idx < nvec) __stcg(reinterpret_cast<float4*>(mine + 4) + idx, acc[j]);
  }
  if (p.alpha)
    for (int r = tid; r < nrows; r += ATT_THREADS) __stcg(p.alpha + (size_t)b * p.ldalpha + r0 + r, sc[r]);
  trace_mark(p.trace, 6);
  __threadfence();
  __syncthreads();
  trace_mark(p.trace, 7);
  if (tid == 0) {
    const unsigned int t = atomicAdd(p.ticket + b, 1u);
    s_last = (t == (unsigned int)SPLIT - 1u);
    if (s_last) atomicExch(p.ticket + b, 0u);   // re-arm for the next launch
    __threadfence();
  }
  __syncthreads();
  trace_mark(p.trace, 8);
  if (!s_last) { trace_mark(p.trace, 2); return; }

  // merge: every load of a phase is independent (two L2 round trips in total, not one per partial)
  const float* recs = p.part + (size_t)b * SPLIT * PS;
  float wk[16];
  {
    float mk[16], zk[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      mk[k] = k < SPLIT ? __ldcg(recs + (size_t)k * PS) : -INFINITY;
      zk[k] = k < SPLIT ? __ldcg(recs + (size_t)k * PS + 1) : 0.f;
    }
    float M = -INFINITY;
#pragma unroll
    for (int k = 0; k < 16; ++k) M = fmaxf(M, mk[k]);
    float Zt = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      wk[k] = (mk[k] == -INFINITY) ? 0.f : __expf(mk[k] - M);
      Zt = fmaf(zk[k], wk[k], Zt);
    }
    const float inv = 1.0f / Zt;
#pragma unroll
    for (int k = 0; k < 16; ++k) wk[k] *= inv;
    if (p.alpha) {   // raw scores -> probabilities (masked rows hold -inf -> 0)
      float* al = p.alpha + (size_t)b * p.ldalpha;
      for (int r = tid; r < p.R; r += ATT_THREADS) {
        const float sv = __ldcg(al + r);
        al[r] = (sv == -INFINITY) ? 0.f : __expf(sv - M) * inv;
      }
    }
  }
  float4* out4 = reinterpret_cast<float4*>(p.out + (size_t)b * p.ldo);
  for (int idx = tid; idx < nvec; idx += ATT_THREADS) {
    float4 a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      a[k] = k < SPLIT ? __ldcg(reinterpret_cast<const float4*>(recs + (size_t)k * PS + 4) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      o.x = fmaf(wk[k], a[k].x, o.x);
      o.y = fmaf(wk[k], a[k].y, o.y);
      o.z = fmaf(wk[k], a[k].z, o.z);
      o.w = fmaf(wk[k], a[k].w, o.w);
    }
    out4[idx] = o;
  }
  if (p.trace && tid == 0) p.trace[3] = globaltimer_ns();   // exit of a merging CTA (any block)
}

// ------------------------------------------------------------------ host launcher

static size_t attn_smem_bytes(int stages, int D) {
  return (size_t)stages * D * sizeof(float) + (size_t)stages * sizeof(uint64_t) + 8 * ATT_MAX_STAGES * sizeof(float) +
         ATT_MAX_ROWS * (sizeof(float) + sizeof(int));
}

AttnPlan attention_plan(int B, int R, int D, int num_sms) {
  AttnPlan pl{};
  // enough CTAs to cover the machine ~2.5x, at least 4 rows per CTA, at most ATT_MAX_ROWS rows per CTA
  int split = 1;
  while (split < 16 && ((long long)B * split < (long long)(5 * num_sms) / 2) && (R + split) / (split * 2) >= 4) split *= 2;
  while ((R + split - 1) / split > ATT_MAX_ROWS) split *= 2;
  pl.split = split;
  pl.rows_per_cta = (R + split - 1) / split;
  int stages = (int)((44 * 1024) / ((size_t)D * 4));     // <= 44 KB ring -> 5 CTAs per SM; a pass = the whole ring
  if (stages > pl.rows_per_cta) stages = pl.rows_per_cta;
  if (stages < 1) stages = 1;
  if (stages > ATT_MAX_STAGES) stages = ATT_MAX_STAGES;
  pl.stages = stages;
  pl.ticket_bytes = ((size_t)B * sizeof(unsigned int) + 255) & ~size_t(255);
  pl.bytes = pl.ticket_bytes + (((size_t)B * split * (D + 4) * sizeof(float) + 255) & ~size_t(255));
  return pl;
}

template <int NJ>
static int32_t launch_attn_t(const AttnParams& p, int B, int split, cudaStream_t stream) {
  auto kern = soft_dot_attn_kernel<NJ>;
  const size_t smem = attn_smem_bytes(p.stages, p.D);
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    SFB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  SFB_CHECK_CUDA(launch_ex(kern, dim3(split, B, 1), dim3(ATT_THREADS, 1, 1), smem, stream, dim3(1, 1, 1), p));
  count_launch();
  return 0;
}

int32_t launch_soft_dot_attention(AttnParams p, int B, void* ws, size_t ws_bytes, cudaStream_t stream) {
  SFB_CHECK_ARG(p.D > 0 && (p.D % 4) == 0, "attention row length must be a positive multiple of 4");
  SFB_CHECK_ARG(p.D <= 3072, "attention row length > 3072 is not supported");
  SFB_CHECK_ARG((p.lenA % 4) == 0 && (p.lenB % 4) == 0 && p.lenA + p.lenB == p.D, "bad row segments");
  SFB_CHECK_ARG(p.R >= 1, "need at least one row");
  const AttnPlan pl = attention_plan(B, p.R, p.D, device_num_sms());
  SFB_CHECK_ARG(ws && ws_bytes >= pl.bytes && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "attention: workspace");
  p.rows_per_cta = pl.rows_per_cta;
  p.stages = pl.stages;
  p.trace = next_trace_slot();
  p.ticket = static_cast<unsigned int*>(ws);
  p.part = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.ticket_bytes);
  SFB_CHECK_ARG(attn_smem_bytes(p.stages, p.D) <= 200 * 1024, "attention rows do not fit shared memory");
  return p.D <= 1024 ? launch_attn_t<1>(p, B, pl.split, stream) : launch_attn_t<3>(p, B, pl.split, stream);
}

}  // namespace sfb
