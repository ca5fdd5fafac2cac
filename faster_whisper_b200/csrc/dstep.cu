// One persistent, cooperative kernel per decode step for R <= 8 rows (one chunk x beam 5, greedy, best_of <= 8).
//
// With so few rows every sub-layer streams 3-13 MB of weights: as separate launches each costs 6-15 us, and a large-v3
// step is 257 of them.  Here one CTA per SM stays resident for the whole step and walks the phases
//   embed | L x { QKV, self-attn, out-proj, cross-q, cross-attn, cross-out, FFN1, FFN2 } | logits
// separated by grid barriers.  Three things matter on B200 (measured, tools/bench/*.cu and profiles/r1_dstep_*.txt):
//   * a dependent hop through L2 costs ~0.6 us and a grid barrier ~1.4 us, so a phase has a ~2.4 us floor: every load a
//     phase needs that does not depend on the previous phase (weight tiles, biases, pointers) is issued earlier — the
//     weight tile of the *next* work item, whichever phase it belongs to, is always in flight into the other shared-memory
//     buffer (cp.async) while the CTA computes, waits at a barrier or runs an attention phase;
//   * the instruction cache: the first version of this kernel was 105 KB of SASS and every phase ran at instruction-fetch
//     speed (~10x slower than its arithmetic).  The per-layer loop is therefore written for code size: rolled loops,
//     cp.async staging instead of register batches, out-of-line phase functions;
//   * LayerNorms are recomputed by each consumer CTA from the fp32 residual stream instead of being phases of their own,
//     and the residual adds are fire-and-forget fp32 reductions (no read-modify-write hop).
//
// Work units and math are those of decode.cu (16-channel weight tiles feeding mma.sync fragments, 8 warps splitting K;
// self-attention through the beam ancestry table; beam-shared cross attention with flash-decoding splits).
//
// Replaces the per-token body of CTranslate2's Whisper.generate loop (reference call sites
// faster_whisper/transcribe.py:222-236, 1446-1459) — the "single persistent kernel per
// decode step" of BASELINE.json's north_star.
#include <math.h>

#include "common.cuh"
#include "decode.h"
#include "dstep.h"

namespace b2w {

constexpr int kDsThreads = 256;
constexpr int kDsWarps = 8;
constexpr int kDsXQ = 8;
constexpr int kDsNBuf = 3;      // weight-tile ring: the tile being consumed + two in flight
constexpr int kDsSelfKeys = 224;  // keys staged per self-attention pass

__device__ __forceinline__ void ds_mma(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ds_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier.  Unlike cp.async (LDGSTS), whose issue stalls once
// the SM's miss queue is full (measured: ~4400 cycles to issue one 40 KB tile), the issuing thread returns immediately.
__device__ __forceinline__ void ds_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ds_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ds_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ds_cp_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ unsigned ds_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ds_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// fine-grained cycle accounting for CTA 0 (B2W_DSTEP_PROF): slot = 3000 + kind*16 + point
#define DS_TICK(a, kind, point, tprev)                                                      \
  do {                                                                                      \
    if ((a).prof && blockIdx.x == 0 && threadIdx.x == 0) {                                  \
      const long long _now = clock64();                                                     \
      (a).prof[3000 + (kind) * 16 + (point)] += (unsigned long long)(_now - (tprev));       \
      (tprev) = _now;                                                                       \
    }                                                                                       \
  } while (0)

// per-CTA state that lives in shared memory (pointers/tables are never re-fetched from L2 inside the layer loop)
struct DsShared {
  DLayer lay[32];
  RowInfo rows[8];
  unsigned epoch;
  int prof_i;
  int flag;
  // weight pipeline: the next item to issue (sequence index, item index, buffer) and the buffer to consume next
  int p_s, p_item, p_buf, c_buf, ahead;  // ahead = tiles issued and not yet consumed
  int consumed;    // tiles consumed so far (buffer = consumed % kDsNBuf, mbarrier parity = (consumed / kDsNBuf) & 1)
  int x_parity;    // parity of the input-staging mbarrier
  uint64_t wbar[kDsNBuf];
  uint64_t xbar;
};

// Grid barrier: every CTA arrives once; sh.epoch is the running arrival target (host zeroes *bar before the launch).
// Arrive = red.release (orders the CTA's earlier writes, cumulative through bar.sync); wait = relaxed polling.  No acquire
// fence on purpose: it would invalidate the SM's L1 (CCTL.IVALL) and with it the stack, and every read of data produced by
// other CTAs in this kernel already bypasses L1 (ld.global.cg / cp.async.cg / atomics).
__device__ __forceinline__ void ds_barrier_arrive(const DStepArgs& a, DsShared& sh) {
  __syncthreads();
  if (threadIdx.x == 0) {
    sh.epoch += gridDim.x;
    if (a.prof && blockIdx.x == 0) a.prof[sh.prof_i] = ds_globaltimer();  // arrival of CTA 0
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
  }
}
__device__ __forceinline__ void ds_barrier_wait(const DStepArgs& a, DsShared& sh) {
  if (threadIdx.x == 0) {
    const unsigned target = sh.epoch;
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
    } while (v < target);
    if (a.prof && blockIdx.x == 0) a.prof[sh.prof_i + 1] = ds_globaltimer();  // release
    sh.prof_i += 2;
  }
  __syncthreads();
}

enum { DS_QKV = 0, DS_F16 = 1, DS_GELU = 2, DS_RESID = 3, DS_F32 = 4 };

// The GEMVs of a step form a static sequence s = 0 .. 6L (per layer: qkv, out, cross_q, cross_out, ffn1, ffn2; then logits).
// A work item is (16 output channels, one of `ksplit` K ranges of width kr <= d).
struct GemvDesc {
  const __half* W;
  const float* bias;
  int N, K, ksplit, mode;
  int ln;              // input = normalised x (LayerNorm affine is folded into W/bias at load) when set, else `src16`
  const __half* src16;
};

__device__ __forceinline__ GemvDesc ds_desc(const DStepArgs& a, const DsShared& sh, int s) {
  GemvDesc g;
  const int d = a.d;
  g.bias = nullptr; g.ln = 0; g.src16 = nullptr; g.ksplit = 1; g.K = d; g.N = d;
  if (s >= 6 * a.L) {
    g.W = a.logit_w; g.bias = a.logit_b; g.N = a.vpad; g.mode = DS_F32; g.ln = 1;
    return g;
  }
  const DLayer& W = sh.lay[s / 6];
  switch (s % 6) {
    case 0: g.W = W.wqkv; g.bias = W.bqkv; g.N = 3 * d; g.mode = DS_QKV; g.ln = 1; break;
    case 1: g.W = W.wo; g.bias = W.bo; g.mode = DS_RESID; g.src16 = a.ao; break;
    case 2: g.W = W.wq_x; g.bias = W.bq_x; g.mode = DS_F16; g.ln = 1; break;
    case 3: g.W = W.wo_x; g.bias = W.bo_x; g.mode = DS_RESID; g.src16 = a.ao; break;
    case 4: g.W = W.w1; g.bias = W.b1; g.N = 4 * d; g.mode = DS_GELU; g.ln = 1; break;
    default: g.W = W.w2; g.bias = W.b2; g.K = 4 * d; g.ksplit = 4; g.mode = DS_RESID; g.src16 = a.h; break;
  }
  return g;
}

// Issues the cp.async copies of the next work item's weight tile (+ its 16 bias values) into buffer sh.p_buf.
// Buffers: [16 rows][kr + 32 halves] + 16 floats  (row stride kr*2 + 64 bytes -> conflict-free 16-byte fragment reads)
__device__ __noinline__ bool ds_issue_next(const DStepArgs& a, DsShared& sh, __half* wbuf, int wbuf_halves) {
  const int last = 6 * a.L;
  long long tq = clock64();
  int s = sh.p_s, item = sh.p_item;
  const int buf = sh.p_buf;
  GemvDesc g;
  for (;;) {
    if (s > last) return false;
    g = ds_desc(a, sh, s);
    if (item < (g.N >> 4) * g.ksplit) break;
    s += 1;
    item = blockIdx.x;
  }
  const int kr = g.K / g.ksplit, tl = item / g.ksplit, ks = item - tl * g.ksplit;
  const __half* src = g.W + (long long)tl * 16 * g.K + ks * kr;
  __half* dst = wbuf + (long long)buf * wbuf_halves;
  const int ld = kr + 32;
  if (threadIdx.x < 32) {  // warp 0: one bulk copy per weight row (+ the item's 16 bias values)
    const bool with_bias = g.bias && ks == 0;
    DS_TICK(a, 7, 0, tq);
    if (threadIdx.x == 0) {
      fence_proxy_async();  // the buffer was last read through the generic proxy
      DS_TICK(a, 7, 1, tq);
      mbar_expect_tx(&sh.wbar[buf], 16u * (uint32_t)kr * 2u + (with_bias ? 64u : 0u));
      DS_TICK(a, 7, 2, tq);
    }
    __syncwarp();
    if (threadIdx.x < 16) ds_bulk_g2s(dst + threadIdx.x * ld, src + (long long)threadIdx.x * g.K, (uint32_t)kr * 2u, &sh.wbar[buf]);
    if (threadIdx.x == 16 && with_bias) ds_bulk_g2s(dst + 16 * ld, g.bias + tl * 16, 64u, &sh.wbar[buf]);
    DS_TICK(a, 7, 3, tq);
  }
  __syncthreads();  // every thread has read p_* before thread 0 advances them
  DS_TICK(a, 7, 4, tq);
  if (threadIdx.x == 0) {
    sh.p_s = s;
    sh.p_item = item + gridDim.x;
    sh.p_buf = (buf + 1 == kDsNBuf) ? 0 : buf + 1;
    sh.ahead += 1;
  }
  __syncthreads();
  DS_TICK(a, 7, 5, tq);
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[3000 + 7 * 16 + 15] += 1;
  return true;
}

// keep the ring full (called between a barrier's arrive and wait, and after every consumed tile)
__device__ __noinline__ void ds_fill_pipeline(const DStepArgs& a, DsShared& sh, __half* wbuf, int wbuf_halves) {
  while (sh.ahead < kDsNBuf) {
    if (!ds_issue_next(a, sh, wbuf, wbuf_halves)) break;
  }
}

// GEMV input -> xs [8][K + 32] halves (row stride K*2 + 64 bytes).  Rows arrive by TMA bulk copies (one per row) on an
// mbarrier; LayerNorm inputs land as fp32 in a staging area and are normalised one warp per row in a single pass
// (sum and sum of squares from registers); fp16 activations are copied straight into xs.
__device__ __noinline__ void ds_stage_input(const DStepArgs& a, DsShared& sh, int ln, const __half* src16, int K, __half* xs, float* stage32,
                                            int kind) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = K + 32;
  long long tp = clock64();
  if (warp == 0) {
    const uint32_t row_bytes = ln ? (uint32_t)K * 4u : (uint32_t)K * 2u;
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&sh.xbar, row_bytes * (uint32_t)a.R);
    }
    __syncwarp();
    if (lane < a.R) {
      if (ln)
        ds_bulk_g2s(stage32 + lane * K, a.x + (long long)lane * K, row_bytes, &sh.xbar);
      else
        ds_bulk_g2s(xs + lane * ld, src16 + (long long)lane * K, row_bytes, &sh.xbar);
    }
  }
  DS_TICK(a, kind, 0, tp);  // copies issued
  mbar_wait(&sh.xbar, (uint32_t)sh.x_parity);
  DS_TICK(a, kind, 1, tp);  // input landed
  if (ln && warp < a.R) {
    const float4* xr = reinterpret_cast<const float4*>(stage32 + warp * K);
    const int n4 = K >> 2;
    float4 v[10];
    float su = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int idx = lane + 32 * i;
      v[i] = idx < n4 ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
      su += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    su = warp_sum(su);
    sq = warp_sum(sq);
    const float mean = su / K;
    const float rstd = rsqrtf(fmaxf(sq / K - mean * mean, 0.f) + 1e-5f);
    uint2* o = reinterpret_cast<uint2*>(xs + warp * ld);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int idx = lane + 32 * i;
      if (idx < n4)
        o[idx] = make_uint2(pack_half2((v[i].x - mean) * rstd, (v[i].y - mean) * rstd), pack_half2((v[i].z - mean) * rstd, (v[i].w - mean) * rstd));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) sh.x_parity ^= 1;
  DS_TICK(a, kind, 2, tp);  // normalised / copied input ready
}

// One GEMV phase: y[R,N] = in[R,K] W[N,K]^T for this CTA's items, weights consumed from the shared-memory pipeline.
__device__ __noinline__ void ds_gemv_phase(const DStepArgs& a, DsShared& sh, int s, __half* wbuf, int wbuf_halves, __half* xs, float* stage32,
                                           float* red, __half* kc, __half* vc) {
  const GemvDesc gd = ds_desc(a, sh, s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int N = gd.N, K = gd.K, ksplit = gd.ksplit;
  const int nitems = (N >> 4) * ksplit, kr = K / ksplit, kchunks = kr >> 5;
  const int ldx = K + 32, ldw = kr + 32;
  int item = blockIdx.x;
  if (item >= nitems) return;
  const int kind = s >= 6 * a.L ? 6 : s % 6;
  ds_stage_input(a, sh, gd.ln, gd.src16, K, xs, stage32, kind);
  long long tp = clock64();
  bool first = true;
#pragma unroll 1
  for (; item < nitems; item += gridDim.x) {
    const int tl = item / ksplit, ks = item - tl * ksplit;
    const int n0 = tl * 16, kbase = ks * kr;
    const int cbuf = sh.c_buf;
    // this item's tile was issued earlier (possibly phases ago); wait for its mbarrier
    mbar_wait(&sh.wbar[cbuf], (uint32_t)((sh.consumed / kDsNBuf) & 1));
    if (first) DS_TICK(a, kind, 3, tp);  // weight tile landed
    const __half* wt = wbuf + (long long)cbuf * wbuf_halves;
    const __half* w_lo = wt + g * ldw + 8 * t;
    const __half* w_hi = w_lo + 8 * ldw;
    const __half* xb = xs + g * ldx + kbase + 8 * t;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = warp; c < kchunks; c += kDsWarps) {
      const uint4 wa = *reinterpret_cast<const uint4*>(w_lo + c * 32);
      const uint4 wb = *reinterpret_cast<const uint4*>(w_hi + c * 32);
      const uint4 xv = *reinterpret_cast<const uint4*>(xb + c * 32);
      ds_mma(acc, wa.x, wb.x, wa.y, wb.y, xv.x, xv.y);
      ds_mma(acc, wa.z, wb.z, wa.w, wb.w, xv.z, xv.w);
    }
    float* my = red + warp * 128;  // [16 ch][8 rows]
    my[g * 8 + 2 * t] = acc[0];
    my[g * 8 + 2 * t + 1] = acc[1];
    my[(g + 8) * 8 + 2 * t] = acc[2];
    my[(g + 8) * 8 + 2 * t + 1] = acc[3];
    __syncthreads();
    if (first) DS_TICK(a, kind, 4, tp);  // MMAs + partials stored
    if (threadIdx.x < 128) {
      const int ch = threadIdx.x & 15, r = threadIdx.x >> 4;
      if (r < a.R) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kDsWarps; ++w) v += red[w * 128 + ch * 8 + r];
        const int n = n0 + ch;
        if (gd.bias && ks == 0) v += reinterpret_cast<const float*>(wt + 16 * ldw)[ch];
        if (gd.mode == DS_QKV) {
          const int d = a.d;
          if (n < d) {
            a.q[(long long)r * d + n] = __float2half_rn(v);
          } else {
            const RowInfo ri = sh.rows[r];
            const int which = (n >= 2 * d) ? 1 : 0;
            (which ? vc : kc)[(((long long)ri.chunk * a.n_ctx + ri.pos) * a.slots + ri.slot) * d + (n - d - which * d)] = __float2half_rn(v);
          }
        } else if (gd.mode == DS_F16) {
          a.q[(long long)r * N + n] = __float2half_rn(v);
        } else if (gd.mode == DS_GELU) {
          a.h[(long long)r * N + n] = __float2half_rn(gelu_erf(v));
        } else if (gd.mode == DS_RESID) {
          atomicAdd(a.x + (long long)r * N + n, v);  // fire-and-forget reduction into the fp32 residual stream
        } else {
          a.logits[(long long)r * a.vpad + n] = v;
        }
      }
    }
    if (threadIdx.x == 0) {
      sh.c_buf = (cbuf + 1 == kDsNBuf) ? 0 : cbuf + 1;
      sh.ahead -= 1;
      sh.consumed += 1;
    }
    __syncthreads();
    if (first) DS_TICK(a, kind, 5, tp);  // epilogue
    ds_fill_pipeline(a, sh, wbuf, wbuf_halves);
    if (first) {
      DS_TICK(a, kind, 6, tp);  // ring topped up
      if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[3000 + kind * 16 + 15] += 1;
    }
    first = false;
  }
}

// masked self-attention for one (head, row) task: the row's K/V history is gathered through the ancestry table into
// shared memory with async copies (one L2/HBM round trip per pass of kDsSelfKeys keys), then scored from there with an
// online softmax across passes
__device__ __noinline__ void ds_self_attn_task(const DStepArgs& a, const DsShared& sh, int h, int r, const __half* kc, const __half* vc, float* sm) {
  float* sc = sm;                      // [kDsSelfKeys]
  float* red = sm + kDsSelfKeys;       // [16]
  float* oacc = red + 16;              // [4][64]
  float* qf = oacc + 256;              // [64]
  __half* kt = reinterpret_cast<__half*>(qf + 64);  // [kDsSelfKeys][72]  (offset 560 floats: 16-byte aligned)
  __half* vt = kt + kDsSelfKeys * 72;               // [kDsSelfKeys][64]
  const RowInfo ri = sh.rows[r];
  const int d = a.d, nk = ri.pos + 1, tid = threadIdx.x;
  const uint8_t* anc = a.anc + (ri.pos & 1) * a.anc_buf_stride + ((long long)ri.chunk * a.slots + ri.slot) * a.n_ctx;
  if (tid < 64) qf[tid] = __half2float(__ldcg(a.q + (long long)r * d + h * 64 + tid)) * 0.125f;
  const int e = tid & 63, part = tid >> 6;
  float m_run = -INFINITY, l_run = 0.f, acc = 0.f;
#pragma unroll 1
  for (int base = 0; base < nk; base += kDsSelfKeys) {
    const int n = min(kDsSelfKeys, nk - base);
    if (tid < n) {
      const int j = base + tid;
      const int slot = (j == ri.pos) ? ri.slot : anc[j];
      const long long off = (((long long)ri.chunk * a.n_ctx + j) * a.slots + slot) * d + h * 64;
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        ds_cp_async16(kt + tid * 72 + i * 8, kc + off + i * 8);
        ds_cp_async16(vt + tid * 64 + i * 8, vc + off + i * 8);
      }
    }
    ds_cp_commit();
    ds_cp_wait_all();
    __syncthreads();
    float mx = -INFINITY;
    if (tid < n) {
      const __half2* kp = reinterpret_cast<const __half2*>(kt + tid * 72);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
      for (int i = 0; i < 32; i += 2) {
        const float2 k0 = __half22float2(kp[i]), k1 = __half22float2(kp[i + 1]);
        s0 = fmaf(k0.x, qf[2 * i], fmaf(k0.y, qf[2 * i + 1], s0));
        s1 = fmaf(k1.x, qf[2 * i + 2], fmaf(k1.y, qf[2 * i + 3], s1));
      }
      mx = s0 + s1;
      sc[tid] = mx;
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < kDsWarps; ++i) mx = fmaxf(mx, red[i]);
    const float m_new = fmaxf(m_run, mx);
    float sum = 0.f;
    if (tid < n) {
      const float p = __expf(sc[tid] - m_new);
      sc[tid] = p;
      sum = p;
    }
    sum = warp_sum(sum);
    if ((tid & 31) == 0) red[8 + (tid >> 5)] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < kDsWarps; ++i) sum += red[8 + i];
    const float alpha = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
    float a0 = 0.f, a1 = 0.f;
    int j = part;  // this thread's keys: part, part + 4, ...; two independent chains
#pragma unroll 2
    for (; j + 4 < n; j += 8) {
      a0 = fmaf(sc[j], __half2float(vt[j * 64 + e]), a0);
      a1 = fmaf(sc[j + 4], __half2float(vt[(j + 4) * 64 + e]), a1);
    }
    if (j < n) a0 = fmaf(sc[j], __half2float(vt[j * 64 + e]), a0);
    acc = fmaf(acc, alpha, a0 + a1);
    l_run = fmaf(l_run, alpha, sum);
    m_run = m_new;
    __syncthreads();  // the staging buffers are reused by the next pass
  }
  oacc[part * 64 + e] = acc;
  __syncthreads();
  if (tid < 64) a.ao[(long long)r * d + h * 64 + tid] = __float2half_rn((oacc[tid] + oacc[64 + tid] + oacc[128 + tid] + oacc[192 + tid]) / l_run);
  __syncthreads();
}

// beam-shared cross attention: one (key split, head, chunk) task; the last split of a (chunk, head) group combines
__device__ __noinline__ void ds_cross_attn_task(const DStepArgs& a, DsShared& sh, int layer, int split, int h, int b, float* sm) {
  const int T = a.T, S = a.xsplits, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = a.rows_per_chunk, row0 = b * a.rows_per_chunk, d = a.d;
  const int k0 = (int)((long long)T * split / S), k1 = (int)((long long)T * (split + 1) / S), nk = k1 - k0;
  const int kmax = (T + S - 1) / S + 1;
  float* qs = sm;                              // [8][64]
  float* sc = qs + kDsXQ * 64;                 // [8][kmax]
  float* wred = sc + kDsXQ * kmax;             // [8 warps][8][64]
  float* stat = wred + kDsWarps * kDsXQ * 64;  // [8][2]
  __half* vt = reinterpret_cast<__half*>(stat + kDsXQ * 2);  // [kmax][64]
  __half* kt = vt + kmax * 64;                 // [kmax][72]
  const DecBindings bd = *a.bind;
  const long long per = (long long)bd.B_total * a.H * T * 64;
  const __half* Kb = bd.xkv + ((long long)layer * 2 + 0) * per + (((long long)(bd.chunk0 + b) * a.H + h) * T + k0) * 64;
  const __half* Vb = bd.xkv + ((long long)layer * 2 + 1) * per + (((long long)(bd.chunk0 + b) * a.H + h) * T + k0) * 64;
#pragma unroll 1
  for (int i = tid; i < nk * 8; i += kDsThreads) ds_cp_async16(kt + (i >> 3) * 72 + (i & 7) * 8, Kb + i * 8);
  ds_cp_commit();
#pragma unroll 1
  for (int i = tid; i < nk * 8; i += kDsThreads) ds_cp_async16(vt + i * 8, Vb + i * 8);
  ds_cp_commit();
  for (int i = tid; i < kDsXQ * 64; i += kDsThreads) {
    const int q = i >> 6, e = i & 63;
    qs[i] = (q < nq) ? __half2float(__ldcg(a.q + (long long)(row0 + q) * d + h * 64 + e)) * 0.125f : 0.f;
  }
  ds_cp_wait_1();
  __syncthreads();
  // scores: one key per thread, eight queries at a time in registers
#pragma unroll 1
  for (int j = tid; j < nk; j += kDsThreads) {
    const __half2* kp = reinterpret_cast<const __half2*>(kt + j * 72);
    float s[kDsXQ];
#pragma unroll
    for (int q = 0; q < kDsXQ; ++q) s[q] = 0.f;
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
      const float2 kf = __half22float2(kp[i]);
#pragma unroll
      for (int q = 0; q < kDsXQ; ++q) {
        const float2 qq = *reinterpret_cast<const float2*>(qs + q * 64 + 2 * i);
        s[q] = fmaf(kf.x, qq.x, fmaf(kf.y, qq.y, s[q]));
      }
    }
#pragma unroll
    for (int q = 0; q < kDsXQ; ++q) sc[q * kmax + j] = s[q];
  }
  __syncthreads();
  {  // one warp per query: partial softmax statistics
    const int q = warp;
    if (q < nq) {
      float mx = -INFINITY;
      for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, sc[q * kmax + j]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < nk; j += 32) {
        const float p = __expf(sc[q * kmax + j] - mx);
        sc[q * kmax + j] = p;
        sum += p;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        stat[q * 2] = mx;
        stat[q * 2 + 1] = sum;
      }
    } else {
      for (int j = lane; j < nk; j += 32) sc[q * kmax + j] = 0.f;
    }
  }
  ds_cp_wait_all();
  __syncthreads();
  float acc[kDsXQ][2];
#pragma unroll
  for (int q = 0; q < kDsXQ; ++q) acc[q][0] = acc[q][1] = 0.f;
#pragma unroll 1
  for (int j = warp; j < nk; j += kDsWarps) {
    const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(vt + j * 64 + 2 * lane));
#pragma unroll
    for (int q = 0; q < kDsXQ; ++q) {
      const float p = sc[q * kmax + j];
      acc[q][0] = fmaf(p, vf.x, acc[q][0]);
      acc[q][1] = fmaf(p, vf.y, acc[q][1]);
    }
  }
#pragma unroll
  for (int q = 0; q < kDsXQ; ++q) *reinterpret_cast<float2*>(wred + (warp * kDsXQ + q) * 64 + 2 * lane) = make_float2(acc[q][0], acc[q][1]);
  __syncthreads();
  const long long group = (long long)b * a.H + h;
  float* part = a.xpart + (group * S + split) * (kDsXQ * 66);
#pragma unroll 1
  for (int i = tid; i < nq * 64; i += kDsThreads) {
    const int q = i >> 6, e = i & 63;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDsWarps; ++w) v += wred[(w * kDsXQ + q) * 64 + e];
    __stcg(part + q * 66 + e, v);
  }
  if (tid < nq) {
    __stcg(part + tid * 66 + 64, stat[tid * 2]);
    __stcg(part + tid * 66 + 65, stat[tid * 2 + 1]);
  }
  __syncthreads();
  if (tid == 0) {
    int ticket;  // release: the CTA's partials are visible before the ticket; the combiner reads them with ld.cg (L2)
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(a.xcounters + group) : "memory");
    sh.flag = (ticket == S - 1);
    if (sh.flag) a.xcounters[group] = 0;
  }
  __syncthreads();
  if (sh.flag) {
    const float* pg = a.xpart + group * S * (kDsXQ * 66);
#pragma unroll 1
    for (int i = tid; i < nq * 64; i += kDsThreads) {
      const int q = i >> 6, e = i & 63;
      float pm[8], pl[8], pa[8];
#pragma unroll
      for (int s2 = 0; s2 < 8; ++s2) {  // all L2 loads in flight together
        const bool on = s2 < S;
        const float* base = pg + ((on ? s2 : 0) * kDsXQ + q) * 66;
        pm[s2] = on ? __ldcg(base + 64) : -INFINITY;
        pl[s2] = on ? __ldcg(base + 65) : 0.f;
        pa[s2] = on ? __ldcg(base + e) : 0.f;
      }
      float M = pm[0];
#pragma unroll
      for (int s2 = 1; s2 < 8; ++s2) M = fmaxf(M, pm[s2]);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < 8; ++s2) {
        const float w = (s2 < S) ? __expf(pm[s2] - M) : 0.f;
        num = fmaf(w, pa[s2], num);
        den = fmaf(w, pl[s2], den);
      }
      a.ao[(long long)(row0 + q) * d + h * 64 + e] = __float2half_rn(num / den);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kDsThreads, 1) dstep_kernel(const DStepArgs a_param) {
  extern __shared__ __align__(16) unsigned char ds_smem[];
  // the argument block is copied to shared memory: the out-of-line phase functions take it by reference, and a reference
  // to a kernel parameter would otherwise be materialised on the (L1-cached, local-memory) stack
  __shared__ DStepArgs a_sh;
  if (threadIdx.x == 0) a_sh = a_param;
  __syncthreads();
  const DStepArgs& a = a_sh;
  // [weight ring: kDsNBuf buffers][union: GEMV input xs (+ fp32 LayerNorm staging) | attention scratch][red]
  const int wbuf_halves = 16 * (a.d + 32) + 32;  // 16 padded rows + 16 fp32 bias values
  __half* wbuf = reinterpret_cast<__half*>(ds_smem);
  unsigned char* uni = ds_smem + kDsNBuf * (size_t)wbuf_halves * sizeof(__half);
  __half* xs = reinterpret_cast<__half*>(uni);
  float* stage32 = reinterpret_cast<float*>(uni + (size_t)8 * (a.d + 32) * sizeof(__half));  // only used by LayerNorm inputs (K = d)
  float* att = reinterpret_cast<float*>(uni);
  float* red = reinterpret_cast<float*>(uni + a.smem_xs_bytes);
  __shared__ DsShared sh;
  const int d = a.d, L = a.L;
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.layers);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(sh.lay);
    for (int i = threadIdx.x; i < L * (int)(sizeof(DLayer) / 8); i += kDsThreads) dst[i] = src[i];
    if (threadIdx.x < a.R) sh.rows[threadIdx.x] = a.rows[threadIdx.x];
    if (threadIdx.x == 0) {
      sh.epoch = 0;
      sh.prof_i = 1;
      sh.p_s = 0;
      sh.p_item = blockIdx.x;
      sh.p_buf = 0;
      sh.c_buf = 0;
      sh.ahead = 0;
      sh.consumed = 0;
      sh.x_parity = 0;
      for (int i = 0; i < kDsNBuf; ++i) mbar_init(&sh.wbar[i], 1);
      mbar_init(&sh.xbar, 1);
      fence_mbar_init();
      if (a.prof && blockIdx.x == 0) a.prof[0] = ds_globaltimer();
    }
  }
  __syncthreads();
  ds_fill_pipeline(a, sh, wbuf, wbuf_halves);  // the first weight tiles are in flight before anything else happens

  // ---- embed: x = tok_emb[token] + pos_emb[pos] (CTA r owns row r) ----
  if ((int)blockIdx.x < a.R) {
    const int r = blockIdx.x;
    int tok = a.tokens_in[r];
    tok = tok < 0 ? 0 : (tok >= a.n_vocab ? a.n_vocab - 1 : tok);
    const int pos = sh.rows[r].pos;
#pragma unroll 2
    for (int i = threadIdx.x; i < d; i += kDsThreads)
      __stcg(a.x + (long long)r * d + i, __half2float(a.tok_emb[(long long)tok * d + i]) + a.pos_emb[(long long)pos * d + i]);
  }
  ds_barrier_arrive(a, sh);
  ds_barrier_wait(a, sh);

#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    __half* kc = a.kcache + (long long)l * a.kv_layer_stride;
    __half* vc = a.vcache + (long long)l * a.kv_layer_stride;
#pragma unroll 1
    for (int ph = 0; ph < 8; ++ph) {
      if (ph == 1) {  // masked self-attention
#pragma unroll 1
        for (int task = blockIdx.x; task < a.H * a.R; task += gridDim.x) ds_self_attn_task(a, sh, task % a.H, task / a.H, kc, vc, att);
      } else if (ph == 4) {  // beam-shared cross attention
#pragma unroll 1
        for (int task = blockIdx.x; task < a.xsplits * a.H * a.n_chunks; task += gridDim.x) {
          const int split = task % a.xsplits, rest = task / a.xsplits;
          ds_cross_attn_task(a, sh, l, split, rest % a.H, rest / a.H, att);
        }
      } else {
        // GEMV sequence index inside the layer: ph 0 -> qkv(0), 2 -> out(1), 3 -> cross_q(2), 5 -> cross_out(3), 6 -> ffn1(4), 7 -> ffn2(5)
        const int j = ph == 0 ? 0 : (ph < 4 ? ph - 1 : ph - 2);
        ds_gemv_phase(a, sh, 6 * l + j, wbuf, wbuf_halves, xs, stage32, red, kc, vc);
      }
      // arrive, then use the barrier latency to top up the weight ring, then wait
      ds_barrier_arrive(a, sh);
      ds_fill_pipeline(a, sh, wbuf, wbuf_halves);
      ds_barrier_wait(a, sh);
    }
  }
  // ---- logits = LN_f(x) E^T ----
  ds_gemv_phase(a, sh, 6 * L, wbuf, wbuf_halves, xs, stage32, red, nullptr, nullptr);
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[sh.prof_i] = ds_globaltimer();
}

size_t dstep_smem_bytes(const DStepArgs& a, size_t* xs_bytes) {
  const size_t wbufs = kDsNBuf * ((size_t)16 * (a.d + 32) + 32) * sizeof(__half);
  const size_t xs_ffn2 = (size_t)8 * (4 * a.d + 32) * sizeof(__half);
  const size_t xs_ln = (size_t)8 * (a.d + 32) * sizeof(__half) + (size_t)8 * a.d * sizeof(float);  // fp16 rows + fp32 rows, gamma, beta
  const int kmax = (a.T + a.xsplits - 1) / a.xsplits + 1;
  const size_t xat = (size_t)(kDsXQ * 64 + kDsXQ * kmax + kDsWarps * kDsXQ * 64 + kDsXQ * 2) * sizeof(float) + (size_t)kmax * (64 + 72) * sizeof(__half);
  const size_t sat = (size_t)(kDsSelfKeys + 16 + 256 + 64) * sizeof(float) + (size_t)kDsSelfKeys * (72 + 64) * sizeof(__half);
  size_t region = xs_ffn2;
  if (xs_ln > region) region = xs_ln;
  if (xat > region) region = xat;
  if (sat > region) region = sat;
  region = (region + 127) & ~size_t(127);
  if (xs_bytes) *xs_bytes = region;
  return wbufs + region + kDsWarps * 128 * sizeof(float);
}

void dstep_configure() {
  B2W_CUDA(cudaFuncSetAttribute(dstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
}

int dstep_max_grid(int num_sms, size_t smem) {
  if (smem > 220 * 1024) return 0;
  int per_sm = 0;
  B2W_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dstep_kernel, kDsThreads, smem));
  return per_sm >= 1 ? num_sms : 0;
}

void dstep_launch(DStepArgs a, int grid, cudaStream_t s) {
  size_t xs = 0;
  const size_t smem = dstep_smem_bytes(a, &xs);
  a.smem_xs_bytes = (int)xs;
  B2W_CUDA(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), s));
  void* args[] = {&a};
  B2W_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(dstep_kernel), dim3(grid), dim3(kDsThreads), args, smem, s));
  count_launch();
}

}  // namespace b2w
