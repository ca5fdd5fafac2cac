// Fused log-mel front end: framing (reflect pad) + Hann window + 400-point real DFT + |X|^2 + Slaney mel
// filterbank + log10/clamp in one kernel, then a per-chunk finalize pass for the global-max clamp.
//
// Replaces FeatureExtractor.__call__ (reference faster_whisper/feature_extractor.py:198-230) and the
// per-chunk `feature_extractor(chunk)[..., :-1]` + `pad_or_trim` of the batched pipeline
// (faster_whisper/transcribe.py:463-467, 514-516).
//
// HBM-bound by construction: 4 B/sample in, 4 B per (mel, frame) out; the DFT twiddles (321 KB) and the
// filterbank live in L2, the Hann window and the frame samples are staged in shared memory.
// The 400-point DFT is evaluated as two real [frames x 200] x [200 x 201] products after folding the
// windowed frame into its even/odd parts (y[n] +- y[400-n]), which halves the multiply count.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "engine.h"

namespace b2w {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFR = 32;  // frames per CTA
constexpr int kMelThreads = 224;
constexpr int kSigLen = (kFR - 1) * kHop + kNfft;  // 5360

__constant__ float c_hann[kNfft];  // np.hanning(401)[:-1] computed in float64, rounded once

// order-preserving float<->int mapping for atomicMax on possibly-negative floats
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

struct MelChunk {
  const float* pcm;   // device
  int64_t n_samples;  // real samples (before the `padding` zeros)
  int32_t n_frames;   // frames that exist (and take part in the global max) = 1 + n_samples/160
  int32_t n_emit;     // frames written
};

__global__ void __launch_bounds__(kMelThreads) logmel_power_kernel(
    const MelChunk* __restrict__ chunks, int padding, int n_mels, const float2* __restrict__ twiddle /*[200][201]*/,
    const int* __restrict__ filt_lo, const int* __restrict__ filt_n, const int* __restrict__ filt_off,
    const float* __restrict__ filt_w, float* __restrict__ out, int64_t out_chunk_stride, int out_ld,
    int* __restrict__ chunk_max) {
  extern __shared__ float smem[];
  float* sig = smem;                    // [kSigLen]
  float* Ye = sig + kSigLen;            // [201][32]
  float* Yo = Ye + kBins * kFR;         // [201][32]  (row 0 and 200 unused)
  float* P = Ye;                        // power [201][32] aliases Ye after the DFT

  const MelChunk ck = chunks[blockIdx.y];
  const int f0 = blockIdx.x * kFR;
  if (f0 >= ck.n_frames) return;
  const int tid = threadIdx.x;
  const int64_t L = ck.n_samples + padding;  // length of the zero-padded signal that gets reflect-padded
  const int64_t period = 2 * (L - 1);

  // ---- stage the samples this CTA's frames touch (np.pad(..., mode="reflect") semantics) ----
  for (int i = tid; i < kSigLen; i += kMelThreads) {
    int64_t j = (int64_t)f0 * kHop + i - kNfft / 2;
    float v = 0.f;
    if (L == 1) {
      j = 0;
    } else {
      j %= period;
      if (j < 0) j += period;
      if (j >= L) j = period - j;
    }
    if (j < ck.n_samples) v = __ldg(ck.pcm + j);
    sig[i] = v;
  }
  __syncthreads();

  // ---- windowed even/odd folds, layout [n][frame] ----
  for (int i = tid; i < kBins * kFR; i += kMelThreads) {
    const int n = i / kFR, f = i % kFR;
    const float* s = sig + f * kHop;
    float a = s[n] * c_hann[n];
    if (n == 0 || n == 200) {
      Ye[i] = a;
      Yo[i] = 0.f;
    } else {
      float b = s[400 - n] * c_hann[400 - n];
      Ye[i] = a + b;
      Yo[i] = a - b;
    }
  }
  __syncthreads();

  // ---- DFT: thread k owns bin k for all 32 frames ----
  float re[kFR], im[kFR];
  const int k = tid;
  if (k < kBins) {
    const float sgn = (k & 1) ? -1.f : 1.f;
#pragma unroll
    for (int f = 0; f < kFR; ++f) {
      re[f] = Ye[f] + sgn * Ye[200 * kFR + f];
      im[f] = 0.f;
    }
    for (int n = 1; n < 200; ++n) {
      const float2 cs = __ldg(twiddle + n * kBins + k);
      const float4* ye4 = reinterpret_cast<const float4*>(Ye + n * kFR);
      const float4* yo4 = reinterpret_cast<const float4*>(Yo + n * kFR);
#pragma unroll
      for (int q = 0; q < kFR / 4; ++q) {
        const float4 e = ye4[q], o = yo4[q];
        re[4 * q + 0] = fmaf(e.x, cs.x, re[4 * q + 0]);
        re[4 * q + 1] = fmaf(e.y, cs.x, re[4 * q + 1]);
        re[4 * q + 2] = fmaf(e.z, cs.x, re[4 * q + 2]);
        re[4 * q + 3] = fmaf(e.w, cs.x, re[4 * q + 3]);
        im[4 * q + 0] = fmaf(o.x, cs.y, im[4 * q + 0]);
        im[4 * q + 1] = fmaf(o.y, cs.y, im[4 * q + 1]);
        im[4 * q + 2] = fmaf(o.z, cs.y, im[4 * q + 2]);
        im[4 * q + 3] = fmaf(o.w, cs.y, im[4 * q + 3]);
      }
    }
  }
  __syncthreads();  // everyone is done reading Ye/Yo
  if (k < kBins) {
#pragma unroll
    for (int f = 0; f < kFR; ++f) P[k * kFR + f] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  // ---- mel filterbank (sparse triangles) + log10, lane = frame ----
  const int warp = tid >> 5, lane = tid & 31;
  const int frame = f0 + lane;
  float vmax = -INFINITY;
  for (int m = warp; m < n_mels; m += kMelThreads / 32) {
    const int lo = filt_lo[m], cnt = filt_n[m];
    const float* w = filt_w + filt_off[m];
    float acc = 0.f;
    for (int i = 0; i < cnt; ++i) acc = fmaf(__ldg(w + i), P[(lo + i) * kFR + lane], acc);
    const float lg = log10f(fmaxf(acc, 1e-10f));
    if (frame < ck.n_frames) {
      vmax = fmaxf(vmax, lg);
      if (frame < ck.n_emit) out[blockIdx.y * out_chunk_stride + (int64_t)m * out_ld + frame] = lg;
    }
  }
  vmax = warp_max(vmax);
  if (lane == 0 && vmax > -INFINITY) atomicMax(chunk_max + blockIdx.y, float_to_ordered(vmax));
}

// out = (max(x, gmax - 8) + 4) / 4 for emitted frames; zeros (pad_or_trim) beyond them up to out_ld.
__global__ void logmel_finalize_kernel(const MelChunk* __restrict__ chunks, int n_mels, float* __restrict__ out,
                                       int64_t out_chunk_stride, int out_ld, const int* __restrict__ chunk_max,
                                       int zero_fill) {
  const MelChunk ck = chunks[blockIdx.z];
  const int m = blockIdx.y;
  const float floor_v = ordered_to_float(chunk_max[blockIdx.z]) - 8.0f;
  float* row = out + blockIdx.z * out_chunk_stride + (int64_t)m * out_ld;
  const int limit = zero_fill ? out_ld : ck.n_emit;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < limit; f += gridDim.x * blockDim.x) {
    if (f < ck.n_emit)
      row[f] = (fmaxf(row[f], floor_v) + 4.0f) * 0.25f;
    else
      row[f] = 0.f;
  }
}

// ---- host side ---------------------------------------------------------------------------------------
static void build_filters(int n_mels, std::vector<int>& lo, std::vector<int>& cnt, std::vector<int>& off,
                          std::vector<float>& w) {
  // Slaney mel filterbank in float64, cast to float32 (feature_extractor.py:24-65, :20-22)
  const int nb = kBins;
  std::vector<double> fftfreqs(nb), mels(n_mels + 2), freqs(n_mels + 2);
  for (int i = 0; i < nb; ++i) fftfreqs[i] = i * (16000.0 / kNfft);
  const double max_mel = 45.245640471924965, f_sp = 200.0 / 3, min_log_hz = 1000.0;
  const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  for (int i = 0; i < n_mels + 2; ++i) {
    // np.linspace(0, max_mel, n) : start + i*step with the last point forced exact
    double step = max_mel / (n_mels + 1);
    mels[i] = (i == n_mels + 1) ? max_mel : i * step;
    freqs[i] = mels[i] >= min_log_mel ? min_log_hz * exp(logstep * (mels[i] - min_log_mel)) : f_sp * mels[i];
  }
  lo.assign(n_mels, 0);
  cnt.assign(n_mels, 0);
  off.assign(n_mels, 0);
  w.clear();
  for (int m = 0; m < n_mels; ++m) {
    const double fd0 = freqs[m + 1] - freqs[m], fd1 = freqs[m + 2] - freqs[m + 1];
    const double enorm = 2.0 / (freqs[m + 2] - freqs[m]);
    int first = -1, last = -1;
    std::vector<float> row(nb, 0.f);
    for (int k = 0; k < nb; ++k) {
      const double lower = -(freqs[m] - fftfreqs[k]) / fd0;
      const double upper = (freqs[m + 2] - fftfreqs[k]) / fd1;
      double v = fmax(0.0, fmin(lower, upper)) * enorm;
      row[k] = (float)v;
      if (row[k] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    off[m] = (int)w.size();
    if (first >= 0) {
      lo[m] = first;
      cnt[m] = last - first + 1;
      for (int k = first; k <= last; ++k) w.push_back(row[k]);
    }
  }
}

MelPlan::MelPlan(int n_mels_) : n_mels(n_mels_) {
  std::vector<float2> tw((size_t)200 * kBins);
  for (int n = 0; n < 200; ++n)
    for (int k = 0; k < kBins; ++k) {
      const int r = (n * k) % kNfft;  // exact argument reduction
      const double a = 2.0 * M_PI * r / kNfft;
      tw[(size_t)n * kBins + k] = make_float2((float)cos(a), (float)sin(a));
    }
  std::vector<int> lo, cnt, off;
  std::vector<float> w;
  build_filters(n_mels, lo, cnt, off, w);
  float hann[kNfft];
  for (int n = 0; n < kNfft; ++n) hann[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / 400.0));  // np.hanning(401)[n]
  B2W_CUDA(cudaMemcpyToSymbol(c_hann, hann, sizeof hann));
  B2W_CUDA(cudaMalloc(&d_twiddle, tw.size() * sizeof(float2)));
  B2W_CUDA(cudaMemcpy(d_twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  B2W_CUDA(cudaMalloc(&d_lo, n_mels * sizeof(int)));
  B2W_CUDA(cudaMalloc(&d_cnt, n_mels * sizeof(int)));
  B2W_CUDA(cudaMalloc(&d_off, n_mels * sizeof(int)));
  B2W_CUDA(cudaMalloc(&d_w, w.size() * sizeof(float)));
  B2W_CUDA(cudaMemcpy(d_lo, lo.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
  B2W_CUDA(cudaMemcpy(d_cnt, cnt.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
  B2W_CUDA(cudaMemcpy(d_off, off.data(), n_mels * sizeof(int), cudaMemcpyHostToDevice));
  B2W_CUDA(cudaMemcpy(d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  const int smem = (kSigLen + 2 * kBins * kFR) * sizeof(float);
  B2W_CUDA(cudaFuncSetAttribute(logmel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
}

MelPlan::~MelPlan() {
  cudaFree(d_twiddle);
  cudaFree(d_lo);
  cudaFree(d_cnt);
  cudaFree(d_off);
  cudaFree(d_w);
}

// chunks_dev: device array of MelChunkDesc (same layout as MelChunk); chunk_max_dev: int[n_chunks] scratch.
void MelPlan::run(const void* chunks_dev, int n_chunks, int max_frames, int padding, float* out, int64_t out_chunk_stride,
                  int out_ld, int* chunk_max_dev, bool zero_fill, cudaStream_t stream) const {
  static_assert(sizeof(MelChunk) == sizeof(MelChunkDesc), "layout");
  if (n_chunks == 0) return;
  // 0x80800000 orders below every finite float in the ordered-int encoding
  B2W_CUDA(cudaMemsetAsync(chunk_max_dev, 0x80, n_chunks * sizeof(int), stream));
  const int smem = (kSigLen + 2 * kBins * kFR) * sizeof(float);
  dim3 grid(ceil_div(max_frames, kFR), n_chunks);
  logmel_power_kernel<<<grid, kMelThreads, smem, stream>>>(
      reinterpret_cast<const MelChunk*>(chunks_dev), padding, n_mels, reinterpret_cast<const float2*>(d_twiddle), d_lo,
      d_cnt, d_off, d_w, out, out_chunk_stride, out_ld, chunk_max_dev);
  B2W_LAUNCHED();
  const int span = zero_fill ? out_ld : max_frames;
  dim3 g2(ceil_div(span, 256), n_mels, n_chunks);
  logmel_finalize_kernel<<<g2, 256, 0, stream>>>(reinterpret_cast<const MelChunk*>(chunks_dev), n_mels, out,
                                                 out_chunk_stride, out_ld, chunk_max_dev, zero_fill ? 1 : 0);
  B2W_LAUNCHED();
}

}  // namespace b2w
