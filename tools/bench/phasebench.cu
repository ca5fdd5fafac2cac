// SM-side cost of one GEMV phase of the persistent decode kernel, isolated: (a) stage input (LayerNorm of R rows from
// L2 through cp.async + smem), (b) MMA over a 16 x 1280 weight tile that is already in shared memory, (c) cross-warp
// reduction + epilogue.  Varies the CTA size to see how much is latency-chain (few warps) cost.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* d, const void* s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(s) : "memory"); }
__device__ __forceinline__ float warp_sum(float v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(~0u, v, o); return v; }
__device__ __forceinline__ void mma(float* c, unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NT, bool FENCE>
__global__ void __launch_bounds__(NT, 1) phase(const float* x, const float* gam, const float* bet, const __half* W, float* out, int iters, long long* cyc) {
  constexpr int K = 1280, R = 5, NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem[];
  __half* wt = reinterpret_cast<__half*>(smem);                 // [16][K+32]
  __half* xs = wt + 16 * (K + 32);                              // [8][K+32]
  float* st = reinterpret_cast<float*>(xs + 8 * (K + 32));      // [10][K]
  float* red = st + 10 * K;                                     // [NW][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // weight tile once
  for (int r = 0; r < 16; ++r) for (int c = threadIdx.x; c < K / 8; c += NT) cp16(wt + r * (K + 32) + c * 8, W + (size_t)(blockIdx.x * 16 + r) * K + c * 8);
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  long long t_stage = 0, t_mma = 0, t_epi = 0;
  for (int it = 0; it < iters; ++it) {
    if (FENCE && threadIdx.x == 0) __threadfence();
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < R; ++r) for (int c = threadIdx.x; c < K / 4; c += NT) cp16(st + r * K + c * 4, x + r * K + c * 4);
    for (int c = threadIdx.x; c < K / 4; c += NT) { cp16(st + 8 * K + c * 4, gam + c * 4); cp16(st + 9 * K + c * 4, bet + c * 4); }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // LayerNorm: NW/R warps... keep it simple: row r handled by warps r, r+R.. cooperatively? one warp per row (rows < R), others idle
    if (warp < R) {
      const float* xr = st + warp * K;
      float s = 0.f;
#pragma unroll 4
      for (int i = lane; i < K; i += 32) s += xr[i];
      const float mean = warp_sum(s) / K;
      float q = 0.f;
#pragma unroll 4
      for (int i = lane; i < K; i += 32) { float d = xr[i] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / K + 1e-5f);
      __half2* o = reinterpret_cast<__half2*>(xs + warp * (K + 32));
#pragma unroll 4
      for (int i = lane; i < K / 2; i += 32) {
        float2 v = reinterpret_cast<const float2*>(xr)[i], gg = reinterpret_cast<const float2*>(st + 8 * K)[i], bb = reinterpret_cast<const float2*>(st + 9 * K)[i];
        o[i] = __floats2half2_rn((v.x - mean) * rstd * gg.x + bb.x, (v.y - mean) * rstd * gg.y + bb.y);
      }
    }
    __syncthreads();
    long long t1 = clock64();
    float acc[4] = {0, 0, 0, 0};
    const __half* w_lo = wt + g * (K + 32) + 8 * t;
    const __half* w_hi = w_lo + 8 * (K + 32);
    const __half* xb = xs + g * (K + 32) + 8 * t;
    for (int c = warp; c < K / 32; c += NW) {
      const uint4 wa = *reinterpret_cast<const uint4*>(w_lo + c * 32), wb = *reinterpret_cast<const uint4*>(w_hi + c * 32), xv = *reinterpret_cast<const uint4*>(xb + c * 32);
      mma(acc, wa.x, wb.x, wa.y, wb.y, xv.x, xv.y);
      mma(acc, wa.z, wb.z, wa.w, wb.w, xv.z, xv.w);
    }
    float* my = red + warp * 128;
    my[g * 8 + 2 * t] = acc[0]; my[g * 8 + 2 * t + 1] = acc[1]; my[(g + 8) * 8 + 2 * t] = acc[2]; my[(g + 8) * 8 + 2 * t + 1] = acc[3];
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x < 128) {
      const int ch = threadIdx.x & 15, r = threadIdx.x >> 4;
      if (r < R) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += red[w * 128 + ch * 8 + r];
        atomicAdd(out + r * 2048 + blockIdx.x * 16 + ch, v);
      }
    }
    __syncthreads();
    long long t3 = clock64();
    t_stage += t1 - t0; t_mma += t2 - t1; t_epi += t3 - t2;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t_stage / iters; cyc[1] = t_mma / iters; cyc[2] = t_epi / iters; }
}

template <int NT, bool FENCE>
void run(const float* x, const float* g, const float* b, const __half* W, float* out, long long* cyc) {
  const int K = 1280;
  size_t smem = (size_t)24 * (K + 32) * 2 + 10 * K * 4 + (NT / 32) * 128 * 4;
  CK(cudaFuncSetAttribute(phase<NT, FENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  phase<NT, FENCE><<<80, NT, smem>>>(x, g, b, W, out, 2000, cyc);
  CK(cudaDeviceSynchronize());
  long long h[3];
  CK(cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost));
  printf("threads %4d fence %d: stage(LN) %5lld  mma %5lld  reduce+epilogue %5lld cycles\n", NT, (int)FENCE, h[0], h[1], h[2]);
}

int main() {
  float *x, *g, *b, *out; __half* W; long long* cyc;
  CK(cudaMalloc(&x, 8 * 1280 * 4)); CK(cudaMalloc(&g, 1280 * 4)); CK(cudaMalloc(&b, 1280 * 4)); CK(cudaMalloc(&out, 8 * 2048 * 4));
  CK(cudaMalloc(&W, 1280 * 1280 * 2)); CK(cudaMalloc(&cyc, 64));
  CK(cudaMemset(x, 0, 8 * 1280 * 4)); CK(cudaMemset(g, 0, 1280 * 4)); CK(cudaMemset(b, 0, 1280 * 4)); CK(cudaMemset(W, 0, 1280 * 1280 * 2)); CK(cudaMemset(out, 0, 8 * 2048 * 4));
  run<256, false>(x, g, b, W, out, cyc);
  run<256, true>(x, g, b, W, out, cyc);
  run<512, false>(x, g, b, W, out, cyc);
  run<1024, false>(x, g, b, W, out, cyc);
  return 0;
}
