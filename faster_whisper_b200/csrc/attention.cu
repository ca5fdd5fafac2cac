// Encoder self-attention (non-causal, T = 1500, head_dim = 64) as a flash-attention forward on tcgen05.
//
// One CTA per (128-query tile, head, chunk).  TMA streams Q once and K/V tiles of 128 keys through a
// 2-deep ring; S = Q K^T lands in TMEM (UMMA 128x128x16), four softmax warps own one query row per thread
// (TMEM lane == row, so row max / row sum need no shuffles), write P as packed fp16 back into TMEM, and
// P V runs as a TMEM-A / shared-memory-B UMMA (128x64x16, V consumed MN-major straight from the fused QKV
// activation layout — no transpose pass).  The running output is rescaled in registers (online softmax).
// Two CTAs co-reside per SM (80 KB smem, 256 TMEM columns each) so one CTA's exponentials overlap the
// other's MMAs.
//
// Replaces CTranslate2's batched-GEMM + softmax kernel + batched-GEMM attention
// inside Whisper.encode (reference faster_whisper/transcribe.py:1391-1400).
#include <math.h>

#include "common.cuh"
#include "engine.h"

namespace b2w {

constexpr int kAttnThreads = 192;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: 128 rows x 64 halves, 128B-swizzled
constexpr int kKvStages = 2;

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm, __half* __restrict__ out, int T, int H) {
  constexpr uint32_t IDESC_S = umma_idesc_f16(128, 128, false);
  constexpr uint32_t IDESC_O = umma_idesc_f16(128, 64, true);
  constexpr int TMEM_COLS = 256;  // S: [0,128)  P(fp16x2): [128,192)  O_tile: [192,256)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTileBytes;  // [stage][K|V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBytes * (1 + 2 * kKvStages));
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kKvStages;
  uint64_t* s_full = kv_empty + kKvStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int d = H * 64;
  const int n_kv = (T + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tP = tmem_base + 128, tO = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm);
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tm, q_full, h * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        mbar_wait(&kv_empty[s], ((j / kKvStages) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * kTileBytes);
        uint8_t* k_dst = sKV + s * 2 * kTileBytes;
        tma_load_3d(k_dst, &tm, &kv_full[s], d + h * 64, j * 128, b);
        tma_load_3d(k_dst + kTileBytes, &tm, &kv_full[s], 2 * d + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      const uint64_t dq = umma_smem_desc_sw128(smem_u32(sQ));
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        mbar_wait(&kv_full[s], (j / kKvStages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + s * 2 * kTileBytes);
        const uint64_t dk = umma_smem_desc_sw128(k_addr);
        const uint64_t dv = umma_smem_desc_sw128(k_addr + kTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0 ? 1u : 0u);
        tc_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ts(tO, tP + 8 * k, dv + k * (2048 >> 4), IDESC_O, k != 0 ? 1u : 0u);
        tc_commit(o_full);
        tc_commit(&kv_empty[s]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_run = -INFINITY, l_run = 0.f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * 128;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kbase + c * 32 + i < T) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx * sc);
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_new);
      float l_new = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c * 32, v);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = (kbase + c * 32 + 2 * i < T) ? ex2_approx(__uint_as_float(v[2 * i]) * sc - m_new) : 0.f;
          float p1 = (kbase + c * 32 + 2 * i + 1 < T) ? ex2_approx(__uint_as_float(v[2 * i + 1]) * sc - m_new) : 0.f;
          const __half2 hh = __floats2half2_rn(p0, p1);
          // the row sum uses the rounded probabilities that the P*V product will see
          l_new += __low2float(hh) + __high2float(hh);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        tmem_st16(tP + lane_off + c * 16, pk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + lane_off + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, __uint_as_float(v[i]));
      }
      l_run = fmaf(l_run, alpha, l_new);
      m_run = m_new;
    }
    if (q0 + row < T) {
      const float inv = 1.0f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * T + q0 + row) * d + h * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_uint4(pack_half2(o[8 * i] * inv, o[8 * i + 1] * inv), pack_half2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                            pack_half2(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_half2(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int attn_smem_bytes() { return kTileBytes * (1 + 2 * kKvStages) + 1024 + 128; }

AttnPlan attn_plan(const __half* qkv, __half* out, int B, int T, int H) {
  AttnPlan p;
  p.qkv = qkv;
  p.out = out;
  p.B = B;
  p.T = T;
  p.H = H;
  const uint64_t d3 = 3ull * H * 64;
  uint64_t dims[3] = {d3, (uint64_t)T, (uint64_t)B};
  uint64_t strides[2] = {d3 * 2, d3 * 2 * (uint64_t)T};
  uint32_t box[3] = {64, 128, 1};
  p.tm = make_tmap_f16(qkv, 3, dims, strides, box);
  return p;
}

void attn_run(const AttnPlan& p, cudaStream_t stream) {
  static unsigned long long configured = 0;
  int dev = 0;
  B2W_CUDA(cudaGetDevice(&dev));
  if (!(configured >> dev & 1)) {
    B2W_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes()));
    configured |= 1ull << dev;
  }
  dim3 grid(ceil_div(p.T, 128), p.H, p.B);
  attn_tc_kernel<<<grid, kAttnThreads, attn_smem_bytes(), stream>>>(p.tm, p.out, p.T, p.H);
  B2W_LAUNCHED();
}

// ---- SIMT reference (debug / bisecting only) ------------------------------------------------------------------
__global__ void attn_ref_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int H) {
  extern __shared__ float sc[];  // [T] scores, then 128 scratch
  float* red = sc + T;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z, d = H * 64;
  const __half* base = qkv + (long long)b * T * 3 * d;
  const __half* qp = base + (long long)t * 3 * d + h * 64;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    const __half* kp = base + (long long)j * 3 * d + d + h * 64;
    float s = 0.f;
    for (int e = 0; e < 64; ++e) s = fmaf(__half2float(qp[e]), __half2float(kp[e]), s);
    s *= 0.125f;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < blockDim.x / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    const float p = expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < blockDim.x / 32; ++i) sum += red[i];
  if (threadIdx.x < 64) {
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(sc[j], __half2float(base[(long long)j * 3 * d + 2 * d + h * 64 + threadIdx.x]), acc);
    out[((long long)b * T + t) * d + h * 64 + threadIdx.x] = __float2half_rn(acc / sum);
  }
}

void attn_ref_run(const __half* qkv, __half* out, int B, int T, int H, cudaStream_t stream) {
  dim3 grid(T, H, B);
  attn_ref_kernel<<<grid, 128, (T + 128) * sizeof(float), stream>>>(qkv, out, T, H);
  B2W_LAUNCHED();
}

}  // namespace b2w
