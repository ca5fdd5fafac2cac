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

__device__ __forceinline__ void ds_mma(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ds_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
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

// per-CTA state that lives in shared memory (pointers/tables are never re-fetched from L2 inside the layer loop)
struct DsShared {
  DLayer lay[32];
  RowInfo rows[8];
  unsigned epoch;
  int prof_i;
  int flag;
  // weight pipeline: the next item to issue (sequence index, item index, buffer) and the buffer to consume next
  int p_s, p_item, p_buf, c_buf;
};

// Grid barrier: every CTA arrives once; sh.epoch is the running arrival target (host zeroes *bar before the launch).
// Arrive = red.release (orders the CTA's earlier writes, cumulative through bar.sync); wait = relaxed polling.  No acquire
// fence on purpose: it would invalidate the SM's L1 (CCTL.IVALL) and with it the stack, and every read of data produced by
// other CTAs in this kernel already bypasses L1 (ld.global.cg / cp.async.cg / atomics).
__device__ __noinline__ void ds_grid_barrier(const DStepArgs& a, DsShared& sh) {
  __syncthreads();
  if (threadIdx.x == 0) {
    sh.epoch += gridDim.x;
    const unsigned target = sh.epoch;
    if (a.prof && blockIdx.x == 0) a.prof[sh.prof_i] = ds_globaltimer();  // arrival of CTA 0
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
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
  const float* ln_g;   // LayerNorm(x) input when non-null, else the fp16 activation `src16`
  const float* ln_b;
  const __half* src16;
};

__device__ __forceinline__ GemvDesc ds_desc(const DStepArgs& a, const DsShared& sh, int s) {
  GemvDesc g;
  const int d = a.d;
  g.bias = nullptr; g.ln_g = nullptr; g.ln_b = nullptr; g.src16 = nullptr; g.ksplit = 1; g.K = d; g.N = d;
  if (s >= 6 * a.L) {
    g.W = a.tok_emb; g.N = a.vpad; g.mode = DS_F32; g.ln_g = a.lnf_g; g.ln_b = a.lnf_b;
    return g;
  }
  const DLayer& W = sh.lay[s / 6];
  switch (s % 6) {
    case 0: g.W = W.wqkv; g.bias = W.bqkv; g.N = 3 * d; g.mode = DS_QKV; g.ln_g = W.ln1_g; g.ln_b = W.ln1_b; break;
    case 1: g.W = W.wo; g.bias = W.bo; g.mode = DS_RESID; g.src16 = a.ao; break;
    case 2: g.W = W.wq_x; g.bias = W.bq_x; g.mode = DS_F16; g.ln_g = W.ln2_g; g.ln_b = W.ln2_b; break;
    case 3: g.W = W.wo_x; g.bias = W.bo_x; g.mode = DS_RESID; g.src16 = a.ao; break;
    case 4: g.W = W.w1; g.bias = W.b1; g.N = 4 * d; g.mode = DS_GELU; g.ln_g = W.ln3_g; g.ln_b = W.ln3_b; break;
    default: g.W = W.w2; g.bias = W.b2; g.K = 4 * d; g.ksplit = 4; g.mode = DS_RESID; g.src16 = a.h; break;
  }
  return g;
}

// Issues the cp.async copies of the next work item's weight tile (+ its 16 bias values) into buffer sh.p_buf.
// Buffers: [16 rows][kr + 32 halves] + 16 floats  (row stride kr*2 + 64 bytes -> conflict-free 16-byte fragment reads)
__device__ __noinline__ bool ds_issue_next(const DStepArgs& a, DsShared& sh, __half* wbuf, int wbuf_halves) {
  const int last = 6 * a.L;
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
  const int per_row = kr >> 3, ld = kr + 32;
#pragma unroll 1
  for (int r = 0; r < 16; ++r)
    for (int c = threadIdx.x; c < per_row; c += kDsThreads) ds_cp_async16(dst + r * ld + c * 8, src + (long long)r * g.K + c * 8);
  if (g.bias && ks == 0 && threadIdx.x < 4) ds_cp_async16(dst + 16 * ld + threadIdx.x * 8, g.bias + tl * 16 + threadIdx.x * 4);
  ds_cp_commit();
  __syncthreads();  // every thread has read p_* before thread 0 advances them
  if (threadIdx.x == 0) {
    sh.p_s = s;
    sh.p_item = item + gridDim.x;
    sh.p_buf = buf ^ 1;
  }
  __syncthreads();
  return true;
}

// GEMV input -> xs [8][K + 32] halves (row stride K*2 + 64 bytes).  LayerNorm inputs go through an fp32 staging area
// (raw rows + gamma + beta fetched with one batch of async copies), fp16 activations are copied straight in.
__device__ __noinline__ void ds_stage_input(const DStepArgs& a, const float* ln_g, const float* ln_b, const __half* src16, int K, __half* xs,
                                            float* stage32) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = K + 32;
  if (ln_g) {
    const int n16 = K >> 2;  // 16-byte pieces per fp32 row
    float* gam = stage32 + 8 * K;
    float* bet = gam + K;
#pragma unroll 1
    for (int r = 0; r < a.R; ++r)
      for (int c = threadIdx.x; c < n16; c += kDsThreads) ds_cp_async16(stage32 + r * K + c * 4, a.x + (long long)r * K + c * 4);
    for (int c = threadIdx.x; c < n16; c += kDsThreads) {
      ds_cp_async16(gam + c * 4, ln_g + c * 4);
      ds_cp_async16(bet + c * 4, ln_b + c * 4);
    }
    ds_cp_commit();
    ds_cp_wait_all();
    __syncthreads();
    // All 8 warps work on all rows: thread t owns float4 columns t and t+256 of every row (K <= 2048), single-pass
    // statistics (sum, sum of squares), one cross-warp exchange, then normalise from registers.
    const int n4 = K >> 2;
    float* part = bet + K;  // [8 warps][8 rows][2]
    float4 v[8][2];
    float su[8], sq[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      su[r] = 0.f;
      sq[r] = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int idx = threadIdx.x + k * kDsThreads;
        v[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < a.R && idx < n4) v[r][k] = reinterpret_cast<const float4*>(stage32 + r * K)[idx];
        su[r] += v[r][k].x + v[r][k].y + v[r][k].z + v[r][k].w;
        sq[r] += v[r][k].x * v[r][k].x + v[r][k].y * v[r][k].y + v[r][k].z * v[r][k].z + v[r][k].w * v[r][k].w;
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      su[r] = warp_sum(su[r]);
      sq[r] = warp_sum(sq[r]);
    }
    if (lane < 8) {
      float s_l = 0.f, q_l = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (lane == r) {
          s_l = su[r];
          q_l = sq[r];
        }
      part[(warp * 8 + lane) * 2] = s_l;
      part[(warp * 8 + lane) * 2 + 1] = q_l;
    }
    __syncthreads();
    float mean[8], rstd[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float s_t = 0.f, q_t = 0.f;
#pragma unroll
      for (int w = 0; w < kDsWarps; ++w) {
        const float2 p2 = *reinterpret_cast<const float2*>(part + (w * 8 + r) * 2);
        s_t += p2.x;
        q_t += p2.y;
      }
      mean[r] = s_t / K;
      rstd[r] = rsqrtf(fmaxf(q_t / K - mean[r] * mean[r], 0.f) + 1e-5f);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int idx = threadIdx.x + k * kDsThreads;
      if (idx < n4) {
        const float4 gg = reinterpret_cast<const float4*>(gam)[idx], bb = reinterpret_cast<const float4*>(bet)[idx];
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < a.R) {
            const float4 x4 = v[r][k];
            *reinterpret_cast<uint2*>(xs + r * ld + idx * 4) =
                make_uint2(pack_half2((x4.x - mean[r]) * rstd[r] * gg.x + bb.x, (x4.y - mean[r]) * rstd[r] * gg.y + bb.y),
                           pack_half2((x4.z - mean[r]) * rstd[r] * gg.z + bb.z, (x4.w - mean[r]) * rstd[r] * gg.w + bb.w));
          }
      }
    }
  } else {
    const int n16 = K >> 3;
#pragma unroll 1
    for (int r = 0; r < a.R; ++r)
      for (int c = threadIdx.x; c < n16; c += kDsThreads) ds_cp_async16(xs + r * ld + c * 8, src16 + (long long)r * K + c * 8);
    ds_cp_commit();
    ds_cp_wait_all();
  }
  __syncthreads();
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
  ds_stage_input(a, gd.ln_g, gd.ln_b, gd.src16, K, xs, stage32);
#pragma unroll 1
  for (; item < nitems; item += gridDim.x) {
    const int tl = item / ksplit, ks = item - tl * ksplit;
    const int n0 = tl * 16, kbase = ks * kr;
    const int cbuf = sh.c_buf;
    // keep the pipeline one item ahead, then wait for this item's tile
    if (ds_issue_next(a, sh, wbuf, wbuf_halves))
      ds_cp_wait_1();
    else
      ds_cp_wait_all();
    __syncthreads();
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
    if (threadIdx.x == 0) sh.c_buf = cbuf ^ 1;
    __syncthreads();
  }
}

// masked self-attention for one (head, row) task: the row's K/V history is gathered through the ancestry table into
// shared memory with async copies (one L2/HBM round trip), then scored from there
__device__ __noinline__ void ds_self_attn_task(const DStepArgs& a, const DsShared& sh, int h, int r, const __half* kc, const __half* vc, float* sm) {
  float* sc = sm;                      // [n_ctx]
  float* red = sm + B2W_MAX_TEXT_CTX;  // [16]
  float* oacc = red + 16;              // [4][64]
  float* qf = oacc + 256;              // [64]
  __half* kt = reinterpret_cast<__half*>(qf + 64);  // [nk][72]  (offset 784 floats: 16-byte aligned)
  const RowInfo ri = sh.rows[r];
  const int d = a.d, nk = ri.pos + 1, tid = threadIdx.x;
  __half* vt = kt + (long long)a.n_ctx * 72;        // [nk][64]
  const uint8_t* anc = a.anc + (ri.pos & 1) * a.anc_buf_stride + ((long long)ri.chunk * a.slots + ri.slot) * a.n_ctx;
#pragma unroll 1
  for (int j = tid; j < nk; j += kDsThreads) {
    const int slot = (j == ri.pos) ? ri.slot : anc[j];
    const long long off = (((long long)ri.chunk * a.n_ctx + j) * a.slots + slot) * d + h * 64;
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      ds_cp_async16(kt + j * 72 + i * 8, kc + off + i * 8);
      ds_cp_async16(vt + j * 64 + i * 8, vc + off + i * 8);
    }
  }
  ds_cp_commit();
  if (tid < 64) qf[tid] = __half2float(__ldcg(a.q + (long long)r * d + h * 64 + tid)) * 0.125f;
  ds_cp_wait_all();
  __syncthreads();
  float mx = -INFINITY;
#pragma unroll 1
  for (int j = tid; j < nk; j += kDsThreads) {
    const __half2* kp = reinterpret_cast<const __half2*>(kt + j * 72);
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const float2 kf = __half22float2(kp[i]);
      s = fmaf(kf.x, qf[2 * i], s);
      s = fmaf(kf.y, qf[2 * i + 1], s);
    }
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kDsWarps; ++i) mx = fmaxf(mx, red[i]);
  float sum = 0.f;
#pragma unroll 1
  for (int j = tid; j < nk; j += kDsThreads) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if ((tid & 31) == 0) red[8 + (tid >> 5)] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < kDsWarps; ++i) sum += red[8 + i];
  const int e = tid & 63, part = tid >> 6;
  float acc = 0.f;
#pragma unroll 2
  for (int j = part; j < nk; j += 4) acc = fmaf(sc[j], __half2float(vt[j * 64 + e]), acc);
  oacc[part * 64 + e] = acc;
  __syncthreads();
  if (tid < 64) a.ao[(long long)r * d + h * 64 + tid] = __float2half_rn((oacc[tid] + oacc[64 + tid] + oacc[128 + tid] + oacc[192 + tid]) / sum);
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
  // [weight buffer 0][weight buffer 1][union: GEMV input xs (+ fp32 LayerNorm staging) | attention scratch][red]
  const int wbuf_halves = 16 * (a.d + 32) + 32;  // 16 padded rows + 16 fp32 bias values
  __half* wbuf = reinterpret_cast<__half*>(ds_smem);
  unsigned char* uni = ds_smem + 2 * (size_t)wbuf_halves * sizeof(__half);
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
      if (a.prof && blockIdx.x == 0) a.prof[0] = ds_globaltimer();
    }
  }
  __syncthreads();
  ds_issue_next(a, sh, wbuf, wbuf_halves);  // the first weight tile is in flight before anything else happens

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
  ds_grid_barrier(a, sh);

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
      ds_grid_barrier(a, sh);
    }
  }
  // ---- logits = LN_f(x) E^T ----
  ds_gemv_phase(a, sh, 6 * L, wbuf, wbuf_halves, xs, stage32, red, nullptr, nullptr);
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[sh.prof_i] = ds_globaltimer();
}

size_t dstep_smem_bytes(const DStepArgs& a, size_t* xs_bytes) {
  const size_t wbufs = 2 * ((size_t)16 * (a.d + 32) + 32) * sizeof(__half);
  const size_t xs_ffn2 = (size_t)8 * (4 * a.d + 32) * sizeof(__half);
  const size_t xs_ln = (size_t)8 * (a.d + 32) * sizeof(__half) + ((size_t)10 * a.d + 128) * sizeof(float);  // fp16 rows + fp32 rows, gamma, beta
  const int kmax = (a.T + a.xsplits - 1) / a.xsplits + 1;
  const size_t xat = (size_t)(kDsXQ * 64 + kDsXQ * kmax + kDsWarps * kDsXQ * 64 + kDsXQ * 2) * sizeof(float) + (size_t)kmax * (64 + 72) * sizeof(__half);
  const size_t sat = (size_t)(B2W_MAX_TEXT_CTX + 16 + 256 + 64) * sizeof(float) + (size_t)a.n_ctx * (72 + 64) * sizeof(__half);
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
