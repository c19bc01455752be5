// attention.cu — soft dot-product attention over a set of rows, streamed once from HBM.
//
// For every batch element b:   s_r = rows[b,r,:] . q[b,:]   (masked rows -> excluded)
//                              alpha = softmax_r(s)          out[b,:] = sum_r alpha_r rows[b,r,:]
// which is the core of VisualSoftDotAttention.forward (model.py:320-325: rows = the 36-view feature slab,
// q = W_v^T (W_h h + b_h)) and of SoftDotAttention.forward (model.py:132-139: rows = ctx, q = W_in h).
//
// B200 mapping.  One thread-block CLUSTER of CL CTAs per batch element (grid = CL x B, CL in {1,2,4,8}); CTA `rank`
// owns rows rank, rank+CL, ... (interleaved, so padding masks at the end of a sequence stay balanced).  A CTA streams
// its rows HBM -> shared memory through a ring of bulk async copies (cp.async.bulk = TMA engine, one mbarrier per
// stage, evict-first, masked rows never fetched; in gather mode a row is two copies: feature table + orientation
// table).  The ring is as deep as 52 KB allows (6 slab rows; 4 CTAs per SM -> ~200 KB of loads in flight per SM, and
// enough free CTA slots that all clusters of a B=100 launch are resident in ONE wave) and a stage is refilled the
// moment its row has been consumed, so the stream never drains.  Each thread keeps its slice
// of the row in registers between the score and the accumulation (a row is read from shared memory once): block-wide
// dot (warp shuffles + one __syncthreads), online softmax, FMA into the running weighted sum.  The CL partial results
// (max, sum, weighted sum) are merged INSIDE the cluster through distributed shared memory: every CTA parks its partial
// in its own ring, one cluster barrier, then the owner of a column block pulls the CL partials and (max, sum) pairs
// of its columns from the peers (ld.shared::cluster), combines them in rank order and writes the output — no global
// partial buffer, no atomics; a second barrier only guards the exit.  With PDL the
// ring is primed before the producer of q has finished.
#include <cuda_bf16.h>

#include "kernels.h"
#include "pack.cuh"

namespace sfb {

namespace {
constexpr int ATT_MAX_STAGES = 32;            // ring depth (one lane of warp 0 per stage when priming)
constexpr size_t ATT_RING_BUDGET = 52 * 1024;  // bytes of ring per CTA -> 4 CTAs per SM (6 slab rows each)

}  // namespace

// NJ float4 slices per thread, NT threads per CTA (D <= NJ * NT * 4), RB rows consumed per block barrier
template <int NJ, int NT, int RB, int MINB = (NT == 256 ? 4 : 8)>
__global__ void __launch_bounds__(NT, MINB) soft_dot_attn_kernel(const AttnParams p) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, rank = blockIdx.x, CL = gridDim.x;
  const int D = p.D, nvec = D >> 2, NSTG = p.stages;
  const int SD = p.keyA ? 2 * D : D;   // floats per stage: value row [+ separate key row]
  const int nmine = rank < p.R ? (p.R - rank + CL - 1) / CL : 0;   // rows rank, rank+CL, ...

  float* ring = reinterpret_cast<float*>(smem_raw);                          // [NSTG][D]; reused as the merge inbox
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)NSTG * SD);    // [NSTG]
  float* red = reinterpret_cast<float*>(full + NSTG);                        // [2][NW][RB] per-warp partial dots
  float* stat = red + 2 * NW * RB;                                                // [2] (max, sum) of this CTA, read by peers
  float* sc = stat + 2;                                                      // [rows_per_cta] raw scores (local order)
  int* list = reinterpret_cast<int*>(sc + p.rows_per_cta);                   // [rows_per_cta] unmasked local row ids
  __shared__ int s_nvalid;

  const uint8_t* mrow = p.mask ? p.mask + (size_t)b * p.ldmask : nullptr;

  trace_mark(p.trace, 0);
  pdl_launch_dependents();   // let the next kernel of the step start its own prologue / prefetch
  unsigned long long* ct = p.cta_trace ? p.cta_trace + (size_t)(b * CL + rank) * 8 : nullptr;
  if (ct && tid == 0) {
    ct[0] = globaltimer_ns();
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ct[4] = smid;
  }

  // ---- 1. warp 0: compact the unmasked rows, arm the ring, launch the first NSTG copies
  size_t ba = 0, bb = 0;
  uint64_t pol = 0;
  auto fetch = [&](int i, int stage) {   // local unmasked row #i -> ring[stage]
    const int gr = rank + CL * list[i];
    float* dst = ring + (size_t)stage * SD;
    mbar_expect_tx(&full[stage], (uint32_t)SD * 4u);
    if (p.keyA)   // scores use a separate (pre-projected) key row, the weighted sum the value row
      bulk_g2s_hint(dst + D, p.keyA + (size_t)b * p.strideK_b + (size_t)gr * p.strideK_r, (uint32_t)D * 4u, &full[stage], pol);
    bulk_g2s_hint(dst, p.segA + ba + (size_t)gr * p.strideA_r, (uint32_t)p.lenA * 4u, &full[stage], pol);
    if (p.lenB > 0)
      bulk_g2s_hint(dst + p.lenA, p.segB + bb + (size_t)gr * p.strideB_r, (uint32_t)p.lenB * 4u, &full[stage], pol);
  };
  if (p.idx_dependent) pdl_wait();
  if (warp == 0) {
    // the slab addresses hang on a dependent index load: request it first, use it after the barrier setup
    const long long ia = p.idxA ? (long long)p.idxA[b] : (long long)b;
    const long long ib = p.idxB ? (long long)p.idxB[b] : (long long)b;
    int n = 0;
    if (mrow) {
      for (int base = 0; base < nmine; base += 32) {
        const int i = base + lane;
        const bool ok = i < nmine && !mrow[rank + CL * i];
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (ok) list[n + __popc(bal & ((1u << lane) - 1u))] = i;
        n += __popc(bal);
      }
    } else {   // no mask: every row is live
      for (int i = lane; i < nmine; i += 32) list[i] = i;
      n = nmine;
    }
    if (lane == 0) s_nvalid = n;
    if (lane < NSTG) mbar_init(&full[lane], 1);
    mbar_fence_init();
    __syncwarp();
    ba = (size_t)ia * p.strideA_b;
    bb = (size_t)ib * p.strideB_b;
    pol = p.no_hint ? policy_evict_normal() : policy_evict_first();
    if (lane < NSTG && lane < n) fetch(lane, lane);
  }
  for (int i = tid; i < nmine; i += NT) sc[i] = -INFINITY;
  // ---- 2. q is produced by the preceding kernel: wait for it only now (rows above are step inputs)
  if (!p.defer_wait) pdl_wait();
  trace_mark(p.trace, 1);
  float4 qv[NJ], acc[NJ];
  {
    const float4* q4 = reinterpret_cast<const float4*>(p.q + (size_t)b * p.ldq);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int idx = tid + NT * j;
      qv[j] = idx < nvec ? q4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (p.has_side) {   // side job: pack step operands another kernel needs (one item per thread at B=100)
    const long long total = (long long)p.side.ntile * p.side.nkb * p.side.R * 8;
    const long long nthreads = (long long)gridDim.x * gridDim.y * NT;
    for (long long idx = ((long long)b * CL + rank) * NT + tid; idx < total; idx += nthreads) pack_item(p.side, idx);
  }
  __syncthreads();
  const int nvalid = s_nvalid;
  trace_mark(p.trace, 4);
  if (ct && tid == 0) ct[5] = globaltimer_ns();   // q available, ring primed

  // ---- 3. stream: RB rows per iteration (independent dot products -> ILP), one block barrier per iteration, every
  // stage refilled as soon as its row sits in registers
  float m = -INFINITY, Z = 0.f;
  for (int i0 = 0; i0 < nvalid; i0 += RB) {
    const int nb = min(RB, nvalid - i0);
    float4 v[RB][NJ];
    float part[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      part[r] = 0.f;
      if (r < nb) {
        const int i = i0 + r, s = i % NSTG;
        mbar_wait(&full[s], (uint32_t)(i / NSTG) & 1u);
        const float4* row4 = reinterpret_cast<const float4*>(ring + (size_t)s * SD);
        const float4* key4 = p.keyA ? row4 + nvec : row4;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int idx = tid + NT * j;
          v[r][j] = idx < nvec ? row4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 kx = p.keyA ? (idx < nvec ? key4[idx] : make_float4(0.f, 0.f, 0.f, 0.f)) : v[r][j];
          part[r] = fmaf(kx.x, qv[j].x, part[r]);
          part[r] = fmaf(kx.y, qv[j].y, part[r]);
          part[r] = fmaf(kx.z, qv[j].z, part[r]);
          part[r] = fmaf(kx.w, qv[j].w, part[r]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) v[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RB; ++r) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
    if (ct && tid == 0 && i0 == 0) ct[1] = globaltimer_ns();
    float* rb = red + ((i0 / RB) & 1) * NW * RB;
    if (lane == 0)
#pragma unroll
      for (int r = 0; r < RB; ++r) rb[warp * RB + r] = part[r];
    __syncthreads();   // every thread holds its slices in registers -> the stages are free
    if (tid == 0)
      for (int r = 0; r < nb; ++r)
        if (i0 + r + NSTG < nvalid) fetch(i0 + r + NSTG, (i0 + r) % NSTG);
    float sr[RB], mn = m;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      sr[r] = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) sr[r] += rb[w * RB + r];
      if (r >= nb) sr[r] = -INFINITY;
      mn = fmaxf(mn, sr[r]);
    }
    if (tid == 0)
      for (int r = 0; r < nb; ++r) sc[list[i0 + r]] = sr[r];
    const float corr = __expf(m - mn);
    float e[RB], esum = 0.f;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      e[r] = __expf(sr[r] - mn);   // rows beyond nb: exp(-inf) = 0
      esum += e[r];
    }
    Z = fmaf(Z, corr, esum);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float4 a = acc[j];
      a.x *= corr; a.y *= corr; a.z *= corr; a.w *= corr;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        a.x = fmaf(e[r], v[r][j].x, a.x);
        a.y = fmaf(e[r], v[r][j].y, a.y);
        a.z = fmaf(e[r], v[r][j].z, a.z);
        a.w = fmaf(e[r], v[r][j].w, a.w);
      }
      acc[j] = a;
    }
    m = mn;
  }
  trace_mark(p.trace, 5);
  if (ct && tid == 0) ct[2] = globaltimer_ns();

  // ---- 4. merge inside the cluster through distributed shared memory (pull): every CTA parks its un-normalised
  // partial in its OWN ring (all its rows are consumed), ONE cluster barrier, then the owner of a column block reads
  // the CL partials of its columns and the CL (max, sum) pairs straight out of the peers' shared memory
  {
    float4* mine = reinterpret_cast<float4*>(ring);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int idx = tid + NT * j;
      if (idx < nvec) mine[idx] = acc[j];
    }
    if (tid == 0) {
      stat[0] = m;
      stat[1] = Z;
    }
  }
  cluster_sync_all();   // partials, stats and scores of every CTA of the cluster are in place
  trace_mark(p.trace, 6);
  float M = -INFINITY, Zt = 0.f;
  float wk[8];
  {
    float mk[8], zk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mk[k] = -INFINITY;
      zk[k] = 0.f;
      if (k < CL) {
        const uint32_t a = dsmem_addr(stat, (uint32_t)k);
        mk[k] = dsmem_ld_f32(a);
        zk[k] = dsmem_ld_f32(a + 4u);
      }
      M = fmaxf(M, mk[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wk[k] = (mk[k] != -INFINITY) ? __expf(mk[k] - M) : 0.f;
      Zt = fmaf(zk[k], wk[k], Zt);
    }
  }
  const float inv = Zt > 0.f ? 1.0f / Zt : 0.f;      // every row masked -> zeros (the reference would give NaN)
#pragma unroll
  for (int k = 0; k < 8; ++k) wk[k] *= inv;
  const int cpo = (nvec + CL - 1) / CL;             // float4 columns per owner CTA
  trace_mark(p.trace, 7);
  if (p.alpha)
    for (int i = tid; i < nmine; i += NT)
      p.alpha[(size_t)b * p.ldalpha + rank + CL * i] = sc[i] == -INFINITY ? 0.f : __expf(sc[i] - M) * inv;
  // grid completion must imply the predecessor's completion; its output (post_add) is consumed right below
  if (p.defer_wait) pdl_wait();
  {
    float4* out4 = reinterpret_cast<float4*>(p.out + (size_t)b * p.ldo);
    for (int lc = tid; lc < cpo; lc += NT) {
      const int col = rank * cpo + lc;
      if (col >= nvec) break;
      float4 pk4[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)   // all peer loads in flight together (rank order = summation order)
        if (k < CL) pk4[k] = dsmem_ld_f32x4(dsmem_addr(ring, (uint32_t)k) + (uint32_t)col * 16u);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < CL) {
          o.x = fmaf(wk[k], pk4[k].x, o.x); o.y = fmaf(wk[k], pk4[k].y, o.y);
          o.z = fmaf(wk[k], pk4[k].z, o.z); o.w = fmaf(wk[k], pk4[k].w, o.w);
        }
      if (p.post_add) {   // out = f(attention output + post_add): h~ = tanh(W_out_c wc + W_out_h h), model.py:140-142
        const float4 a = *reinterpret_cast<const float4*>(p.post_add + (size_t)b * p.ld_post + col * 4);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        if (p.post_tanh) { o.x = tanhf(o.x); o.y = tanhf(o.y); o.z = tanhf(o.z); o.w = tanhf(o.w); }
      }
      out4[col] = o;
      if (p.pk_out) {   // the same 4 values as bf16 (hi, lo) in the gate GEMM's packed activation operand
        const int k = col * 4, zt = b / p.pk_rows_per_z, r = b - zt * p.pk_rows_per_z;
        if (p.pk_scale) {
          const float4 sc = *reinterpret_cast<const float4*>(p.pk_scale + (size_t)b * p.pk_ldscale + k);
          o.x *= sc.x; o.y *= sc.y; o.z *= sc.z; o.w *= sc.w;
        }
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
        const size_t half = (size_t)p.pk_NB * 128;
        unsigned char* dst = p.pk_out + ((size_t)zt * p.pk_nkb + p.pk_kb0 + (k >> 6)) * (2 * half) + (size_t)(r >> 3) * 1024 +
                             (size_t)((k & 63) >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(k & 7) * 2;
        *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(dst + half) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
    }
  }
  cluster_sync_all();   // nobody leaves while a peer may still read its shared memory
  trace_mark(p.trace, 2);
  if (ct && tid == 0) ct[3] = globaltimer_ns();
  if (p.trace && tid == 0 && b == gridDim.y - 1 && rank == CL - 1) p.trace[3] = globaltimer_ns();   // last cluster
}

// ------------------------------------------------------------------ host launcher

static size_t attn_smem_bytes(int stages, int D, int rows_per_cta) {   // D = floats per stage
  return (size_t)stages * D * sizeof(float) + (size_t)stages * sizeof(uint64_t) + (2 * 8 * 4 + 2) * sizeof(float) +
         (size_t)rows_per_cta * (sizeof(float) + sizeof(int));
}

AttnPlan attention_plan(int B, int R, int D, int num_sms, bool kv) {
  if (kv) D *= 2;   // a stage holds the value row and the key row
  AttnPlan pl{};
  // cluster size: the largest CL whose B x CL CTAs are still ONE resident wave, with >= 4 rows per CTA
  const int nt = (kv ? D / 2 : D) <= 512 ? 128 : 256;
  // the largest cluster whose B x CL CTAs are ONE resident wave; the ring may shrink (>= 6 rows or all rows) to fit
  int best = 1, best_st = 4;
  bool found = false;
  for (int cl = 8; cl >= 1 && !found; cl >>= 1) {
    if (cl > 1 && R / cl < 4) continue;
    if (cl == 8 && (long long)B * cl > 2LL * num_sms) continue;   // 8-CTA clusters place badly once the machine is full
    const int rows = (R + cl - 1) / cl;
    int st_hi = (int)(ATT_RING_BUDGET / ((size_t)D * 4));
    st_hi = st_hi > rows ? rows : st_hi;
    st_hi = st_hi > ATT_MAX_STAGES ? ATT_MAX_STAGES : st_hi;
    st_hi = st_hi < 4 ? 4 : st_hi;
    int st_lo = rows < 6 ? rows : 6;
    st_lo = st_lo < 4 ? 4 : st_lo;
    if (st_lo > st_hi) st_lo = st_hi;
    for (int st = st_hi; st >= st_lo; --st) {
      const size_t smem = attn_smem_bytes(st, D, rows) + 1024;
      long long per_sm = (long long)((220 * 1024) / smem);
      if (per_sm > 2048 / nt) per_sm = 2048 / nt;
      if (per_sm > 16) per_sm = 16;
      // 20 % of the slots stay free: clusters need all their CTAs inside one GPC at once, and a launch that almost fills
      // the machine leaves some clusters waiting for a second wave (measured: +10 us on the late ones)
      if (cl == 1 || (long long)B * cl * 5 <= per_sm * num_sms * 4) {
        best = cl; best_st = st; found = true;
        break;
      }
    }
  }
  pl.split = best;
  pl.rows_per_cta = (R + best - 1) / best;
  pl.stages = best_st;
  pl.ticket_bytes = 0;
  pl.bytes = 256;   // no global scratch any more; kept non-zero so workspace carving stays uniform
  return pl;
}

template <int NJ, int NT, int RB, int MINB = (NT == 256 ? 4 : 8)>
static int32_t launch_attn_t(const AttnParams& p, int B, int cl, cudaStream_t stream) {
  auto kern = soft_dot_attn_kernel<NJ, NT, RB, MINB>;
  const size_t smem = attn_smem_bytes(p.stages, p.keyA ? 2 * p.D : p.D, p.rows_per_cta);
  static SmemMarks marks;  // per instantiation, per device
  SFB_CHECK_CUDA(ensure_dynamic_smem(kern, smem, marks));
  SFB_CHECK_CUDA(launch_ex(kern, dim3(cl, B, 1), dim3(NT, 1, 1), smem, stream, dim3(cl, 1, 1), p));
  count_launch();
  return 0;
}

int32_t launch_soft_dot_attention(AttnParams p, int B, void* ws, size_t ws_bytes, cudaStream_t stream) {
  (void)ws; (void)ws_bytes;
  SFB_CHECK_ARG(p.D > 0 && (p.D % 4) == 0, "attention row length must be a positive multiple of 4");
  SFB_CHECK_ARG(p.D <= 3072, "attention row length > 3072 is not supported");
  SFB_CHECK_ARG((p.lenA % 4) == 0 && (p.lenB % 4) == 0 && p.lenA + p.lenB == p.D, "bad row segments");
  SFB_CHECK_ARG(p.R >= 1, "need at least one row");
  SFB_CHECK_ARG(B <= 65535, "attention: batch > 65535");
  SFB_CHECK_ARG(!p.post_add || ((reinterpret_cast<uintptr_t>(p.post_add) & 15u) == 0 && (p.ld_post % 4) == 0), "attention: post_add alignment");
  const int SD = p.keyA ? 2 * p.D : p.D;
  SFB_CHECK_ARG(!p.keyA || p.D <= 1024, "attention: separate key rows are supported for row length <= 1024");
  SFB_CHECK_ARG(!p.keyA || ((reinterpret_cast<uintptr_t>(p.keyA) & 15u) == 0 && (p.strideK_r % 4) == 0 && (p.strideK_b % 4) == 0),
                "attention: key rows must be 16-byte aligned");
  AttnPlan pl = attention_plan(B, p.R, p.D, device_num_sms(), p.keyA != nullptr);
  if (g_attn_force_cl > 0) {   // bring-up: force the cluster size
    pl.split = g_attn_force_cl;
    pl.rows_per_cta = (p.R + pl.split - 1) / pl.split;
    const size_t budget = g_attn_ring_kb > 0 ? (size_t)g_attn_ring_kb * 1024 : ATT_RING_BUDGET;
    int st = (int)(budget / ((size_t)SD * 4));
    st = st > pl.rows_per_cta ? pl.rows_per_cta : st;
    pl.stages = st < 4 ? 4 : (st > ATT_MAX_STAGES ? ATT_MAX_STAGES : st);
  }
  if (g_attn_force_stages > 0 && g_attn_force_stages <= pl.rows_per_cta) pl.stages = g_attn_force_stages < 4 ? 4 : g_attn_force_stages;
  p.no_hint = g_attn_no_hint;
  p.cta_trace = (size_t)B * pl.split <= 4096 ? cta_trace_buffer() : nullptr;
  p.rows_per_cta = pl.rows_per_cta;
  p.stages = pl.stages;
  p.trace = next_trace_slot();
  p.ticket = nullptr;
  p.part = nullptr;
  if (p.has_side) SFB_PROPAGATE(pack_prepare(p.side));
  SFB_CHECK_ARG(attn_smem_bytes(p.stages, SD, p.rows_per_cta) <= 200 * 1024, "attention rows do not fit shared memory");
  if (p.D <= 512) return launch_attn_t<1, 128, 4>(p, B, pl.split, stream);
  if (p.D <= 1024) return launch_attn_t<1, 256, 4>(p, B, pl.split, stream);
  if (g_attn_rb == 2) return launch_attn_t<3, 256, 2, 1>(p, B, pl.split, stream);
  if (g_attn_rb == 3) return launch_attn_t<3, 256, 3, 1>(p, B, pl.split, stream);
  if (g_attn_rb == 4) return launch_attn_t<3, 256, 4, 1>(p, B, pl.split, stream);
  if (g_attn_rb == 1) return launch_attn_t<3, 256, 1, 1>(p, B, pl.split, stream);
  return launch_attn_t<3, 256, 1>(p, B, pl.split, stream);
}

}  // namespace sfb
