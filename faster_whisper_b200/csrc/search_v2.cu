// Row search split over the vocabulary (opt-in: B2W_SEARCH_V2=1; the default is search.cu:search_rows_kernel).
//
// search_rows_kernel runs one 1024-thread CTA per row: 1.2 M warp-instructions land on R SMs (5 for one chunk x beam 5),
// which makes it issue-bound at ~39 us per step (profiles/r1_launches_single_v5_summary.txt).  Here every row is cut into
// kV2Slices vocabulary slices: kernel 1 (one CTA per (slice, row)) applies the same logits processors to its slice, reduces
// the soft-max statistics the row needs and extracts the slice's sorted top-2K; kernel 2 (one warp per row) combines the
// statistics, takes the Whisper "timestamp mass beats every text token" decision, merges the slice lists and writes the same
// cand_score / cand_tok that search_update_kernel consumes.  Semantics are those of search.cu (CTranslate2 4.x processors,
// SURVEY.md §8a rows D4-D6); the only numerical difference is the summation order of the soft-max denominator.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "decode.h"

namespace b2w {

namespace {

constexpr int kV2Slices = 8;
constexpr int kV2Threads = 512;
constexpr float kLowest = -FLT_MAX;
constexpr int kNoIdx = 0x7fffffff;

struct KeyIdx {
  float v;
  int i;
};
__device__ __forceinline__ KeyIdx pick(KeyIdx a, KeyIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }
__device__ __forceinline__ KeyIdx warp_pick(KeyIdx a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    KeyIdx b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = pick(a, b);
  }
  return a;
}
__device__ float cta_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ float cta_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}
// the same counter-based Gumbel noise as search.cu:gumbel_u32 (mirrored by oracle/whisper_oracle.py:gumbel_noise)
__device__ __forceinline__ float gumbel(unsigned long long seed, int row, int step, int idx) {
  unsigned long long x = (unsigned long long)idx * 0x9E3779B97F4A7C15ull + seed * 0xBF58476D1CE4E5B9ull +
                         (unsigned long long)row * 0x94D049BB133111EBull + (unsigned long long)step * 0xD6E8FEB86659FD93ull;
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  const float u = ((float)(unsigned)(x >> 41) + 0.5f) * (1.0f / 8388608.0f);
  return -logf(-logf(u));
}

}  // namespace

// per (row, slice): [0] max over the slice after the processors, [1] sum exp(x - max), [2] max over timestamp tokens,
// [3] sum over timestamp tokens of exp(x - that max), [4] max over text tokens, [5] raw max, [6] raw sum exp (first step)
constexpr int kV2Stats = 8;

__global__ void __launch_bounds__(kV2Threads) search_part_kernel(const float* __restrict__ logits, const SearchBuffers bf,
                                                                 const SearchPartBuffers pb) {
  const SearchParams p = *bf.params;
  extern __shared__ float s[];  // [slice length]
  __shared__ float red[32];
  __shared__ float wl_v[(kV2Threads / 32) * kMaxCand];
  __shared__ int wl_i[(kV2Threads / 32) * kMaxCand];
  __shared__ int sh_flags[4];
  const int slice = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
  const int k = r % p.K;
  const int V = p.n_vocab;
  const int SL = (((p.vpad + kV2Slices - 1) / kV2Slices) + 3) & ~3;  // slice length, a multiple of 4
  const int v0 = slice * SL, v1 = min(V, v0 + SL);                   // this CTA owns token ids [v0, v1)
  const int n = max(0, v1 - v0);
  const RowInfo ri = bf.rows[r];
  const int step = ri.pos - (p.prompt_len - 1);
  const int cur = ri.pos & 1;
  const int* hist = bf.hist + ((long long)cur * p.B * p.K + r) * bf.n_ctx;
  const int len = step;
  const float* row = logits + (long long)r * p.vpad;
  float* st = pb.stats + ((long long)r * kV2Slices + slice) * kV2Stats;

  // ---- raw soft-max statistics for no_speech_prob (first step, first row of the chunk) ----
  if (p.want_no_speech_first && step == 0 && k == 0) {
    float mx = -INFINITY;
    for (int i = tid; i < n; i += kV2Threads) mx = fmaxf(mx, row[v0 + i]);
    mx = cta_max(mx, red);
    float sm = 0.f;
    for (int i = tid; i < n; i += kV2Threads) sm += __expf(row[v0 + i] - mx);
    sm = cta_sum(sm, red);
    if (tid == 0) {
      st[5] = mx;
      st[6] = sm;
    }
  }

  // ---- load + static suppression ----
  for (int i = tid; i < n; i += kV2Threads) s[i] = bf.suppress[v0 + i] ? kLowest : row[v0 + i];
  __syncthreads();
  // repetition penalty on the raw value of every distinct generated token (tokens of this slice only)
  if (p.repetition_penalty != 1.0f) {
    for (int i = tid; i < len; i += kV2Threads) {
      const int tok = hist[i];
      if (tok < v0 || tok >= v1) continue;
      bool first = true;
      for (int j = 0; j < i; ++j)
        if (hist[j] == tok) {
          first = false;
          break;
        }
      if (first && !bf.suppress[tok]) {
        const float x = row[tok];
        s[tok - v0] = x < 0.f ? x * p.repetition_penalty : x / p.repetition_penalty;
      }
    }
    __syncthreads();
  }
  const int ng = p.no_repeat_ngram;
  if (ng > 0 && len >= ng) {
    for (int s0 = tid; s0 + ng <= len; s0 += kV2Threads) {
      bool same = true;
      for (int j = 0; j < ng - 1; ++j)
        if (hist[s0 + j] != hist[len - ng + 1 + j]) {
          same = false;
          break;
        }
      const int tok = hist[s0 + ng - 1];
      if (same && tok >= v0 && tok < v1) s[tok - v0] = kLowest;
    }
    __syncthreads();
  }
  if (p.suppress_blank && step == 0 && tid < p.n_suppress_begin) {
    const int tok = p.suppress_begin[tid];
    if (tok >= v0 && tok < v1) s[tok - v0] = kLowest;
  }
  __syncthreads();

  // ---- Whisper timestamp rules: the range masks are slice-local, the probability rule is decided by the merge ----
  const int ts0 = p.timestamp_begin;
  const bool rule = p.timestamp_rules && step > 0;  // "timestamp mass beats every text token" can force a timestamp
  if (p.timestamp_rules) {
    if (tid == 0) {
      int last_ts = 0, penult_ts = 0, t_last = -1;
      if (step > 0) {
        last_ts = hist[len - 1] >= ts0;
        penult_ts = (len < 2) || (hist[len - 2] >= ts0);
        for (int i = len - 1; i >= 0; --i)
          if (hist[i] >= ts0) {
            t_last = hist[i];
            break;
          }
        if (t_last >= 0 && !(last_ts && !penult_ts)) t_last += 1;
      }
      sh_flags[0] = last_ts;
      sh_flags[1] = penult_ts;
      sh_flags[2] = t_last;
    }
    __syncthreads();
    const int last_ts = sh_flags[0], penult_ts = sh_flags[1], t_last = sh_flags[2];
    if (p.no_timestamps >= v0 && p.no_timestamps < v1 && tid == 0) s[p.no_timestamps - v0] = kLowest;
    if (step == 0) {
      const int hi = ts0 + p.max_initial_ts;
      for (int i = tid; i < n; i += kV2Threads) {
        const int v = v0 + i;
        if (v < ts0 || (p.max_initial_ts >= 0 && v > hi)) s[i] = kLowest;
      }
    } else {
      int lo_a = 0, hi_a = 0;
      if (last_ts) {
        if (penult_ts) {
          lo_a = ts0;
          hi_a = V;
        } else {
          lo_a = 0;
          hi_a = p.eot;
        }
      }
      for (int i = tid; i < n; i += kV2Threads) {
        const int v = v0 + i;
        if ((v >= lo_a && v < hi_a) || (t_last >= 0 && v >= ts0 && v < t_last)) s[i] = kLowest;
      }
    }
    __syncthreads();
  }

  // ---- soft-max statistics of the slice (all tokens; timestamp tokens; best text token) ----
  {
    float m_all = -INFINITY, m_ts = -INFINITY, m_text = -INFINITY;
    for (int i = tid; i < n; i += kV2Threads) {
      const float x = s[i];
      m_all = fmaxf(m_all, x);
      if (v0 + i >= ts0)
        m_ts = fmaxf(m_ts, x);
      else
        m_text = fmaxf(m_text, x);
    }
    m_all = cta_max(m_all, red);
    float e_all = 0.f, e_ts = 0.f;
    if (rule) {
      m_ts = cta_max(m_ts, red);
      m_text = cta_max(m_text, red);
    }
    for (int i = tid; i < n; i += kV2Threads) {
      const float x = s[i];
      e_all += __expf(x - m_all);
      if (rule && v0 + i >= ts0) e_ts += __expf(x - m_ts);
    }
    e_all = cta_sum(e_all, red);
    if (rule) e_ts = cta_sum(e_ts, red);
    if (tid == 0) {
      st[0] = n > 0 ? m_all : -INFINITY;
      st[1] = n > 0 ? e_all : 0.f;
      st[2] = m_ts;
      st[3] = e_ts;
      st[4] = m_text;
    }
  }

  // ---- ranking keys; masked entries keep "lowest" in beam/greedy mode and drop out (-inf) when sampling ----
  const bool sampling = (p.mode == 1 && p.sampling_topk != 1);
  const float inv_t = 1.0f / p.temperature;
  const int ncand = p.ncand;
  const int warp = tid >> 5, lane = tid & 31;
  const int n_lists = rule ? 2 : 1;
  for (int list = 0; list < n_lists; ++list) {
    // list 0: every token of the slice; list 1: timestamp tokens only.  Keys are recomputed per list from s[] (kept intact);
    // an extracted element is remembered in `taken` instead of being overwritten.
    const int lo = list == 0 ? 0 : max(0, ts0 - v0);
    auto key_of = [&](int i) -> float {
      const float x = s[i];
      if (x <= kLowest * 0.5f) return sampling ? -INFINITY : x;
      return sampling ? x * inv_t + gumbel(p.seed, r, step, v0 + i) : x;
    };
    // every thread tracks the best two of its strided share (strict > keeps the lowest index among equals)
    KeyIdx b1{-INFINITY, kNoIdx}, b2{-INFINITY, kNoIdx};
    for (int i = lo + tid; i < n; i += kV2Threads) {
      const float kv = key_of(i);
      if (kv > b1.v) {
        b2 = b1;
        b1 = KeyIdx{kv, v0 + i};
      } else if (kv > b2.v) {
        b2 = KeyIdx{kv, v0 + i};
      }
    }
    bool have_second = true;
    KeyIdx mine = b1;
    for (int c = 0; c < ncand; ++c) {
      const KeyIdx w = warp_pick(mine);
      if (lane == 0) {
        wl_v[warp * kMaxCand + c] = w.v;
        wl_i[warp * kMaxCand + c] = w.i;
      }
      if (w.i != kNoIdx && w.i == mine.i) {  // this thread's element won: advance to its next one in (value desc, index asc) order
        if (have_second) {
          mine = b2;
          have_second = false;
        } else {
          // rare (a third winner from the same thread): rescan the share for the successor of the element just taken
          const float bound_v = mine.v;
          const int bound_i = mine.i;
          KeyIdx nb{-INFINITY, kNoIdx};
          for (int i = lo + tid; i < n; i += kV2Threads) {
            const float kv = key_of(i);
            const int id = v0 + i;
            const bool after = kv < bound_v || (kv == bound_v && id > bound_i);
            if (after && kv > -INFINITY && (kv > nb.v || (kv == nb.v && id < nb.i))) nb = KeyIdx{kv, id};
          }
          mine = nb;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      int head = 0;  // lane l walks warp l's sorted list
      const int nw = kV2Threads / 32;
      float* o_key = pb.key + (((long long)r * kV2Slices + slice) * 2 + list) * kMaxCand;
      float* o_x = pb.x + (((long long)r * kV2Slices + slice) * 2 + list) * kMaxCand;
      int* o_idx = pb.idx + (((long long)r * kV2Slices + slice) * 2 + list) * kMaxCand;
      for (int c = 0; c < ncand; ++c) {
        KeyIdx curv{-INFINITY, kNoIdx};
        if (lane < nw && head < ncand) curv = KeyIdx{wl_v[lane * kMaxCand + head], wl_i[lane * kMaxCand + head]};
        if (curv.v == -INFINITY) curv.i = kNoIdx;
        const KeyIdx best = warp_pick(curv);
        if (best.i != kNoIdx && curv.i == best.i) ++head;
        if (lane == 0) {
          o_key[c] = best.v;
          o_idx[c] = best.i;
          o_x[c] = best.i == kNoIdx ? kLowest : s[best.i - v0];
        }
      }
    }
    __syncthreads();
  }
}

// one warp per row: combine the slices
__global__ void __launch_bounds__(32) search_merge_kernel(const float* __restrict__ logits, const SearchBuffers bf, const SearchPartBuffers pb) {
  const SearchParams p = *bf.params;
  const int r = blockIdx.x, lane = threadIdx.x;
  const int b = r / p.K, k = r % p.K;
  const RowInfo ri = bf.rows[r];
  const int step = ri.pos - (p.prompt_len - 1);
  const int cur = ri.pos & 1;
  const bool rule = p.timestamp_rules && step > 0;
  const bool sampling = (p.mode == 1 && p.sampling_topk != 1);
  const float inv_t = 1.0f / p.temperature;
  const int ncand = p.ncand;
  const float* st = pb.stats + ((long long)r * kV2Slices + (lane < kV2Slices ? lane : 0)) * kV2Stats;
  const bool on = lane < kV2Slices;
  // soft-max denominators from the slice statistics
  const float m_all = on ? st[0] : -INFINITY, e_all = on ? st[1] : 0.f;
  const float M_all = warp_max(m_all);
  const float S_all = warp_sum(e_all > 0.f ? e_all * __expf(m_all - M_all) : 0.f);
  bool forced = false;
  float lse = M_all + logf(S_all);
  if (rule) {
    const float m_ts = on ? st[2] : -INFINITY, e_ts = on ? st[3] : 0.f, m_text = on ? st[4] : -INFINITY;
    const float M_ts = warp_max(m_ts);
    const float S_ts_all = warp_sum(e_ts > 0.f ? e_ts * __expf(m_ts - M_all) : 0.f);  // relative to the row maximum, as search.cu sums it
    const float max_text = warp_max(m_text);
    const float ts_lp = (S_ts_all > 0.f) ? (M_all + logf(S_ts_all) - lse) : -INFINITY;
    const float text_lp = max_text - lse;
    forced = ts_lp > text_lp;
    if (forced) {
      // every text token becomes "lowest": the soft-max runs over the timestamp tokens only
      const float S_ts = warp_sum(e_ts > 0.f ? e_ts * __expf(m_ts - M_ts) : 0.f);
      lse = M_ts + logf(S_ts);
    }
  }
  if (p.want_no_speech_first && step == 0 && k == 0) {
    const float m_raw = on ? st[5] : -INFINITY, e_raw = on ? st[6] : 0.f;
    const float M = warp_max(m_raw);
    const float S = warp_sum(e_raw > 0.f ? e_raw * __expf(m_raw - M) : 0.f);
    if (lane == 0) bf.no_speech[b] = __expf(logits[(long long)r * p.vpad + p.no_speech] - M) / S;
  }
  const float cum = bf.cum[(long long)cur * p.B * p.K + r];
  // merge the sorted slice lists (list 1 when a timestamp is forced)
  const int list = forced ? 1 : 0;
  const long long base = (((long long)r * kV2Slices + (on ? lane : 0)) * 2 + list) * kMaxCand;
  int head = 0, filler = 0;
  for (int c = 0; c < ncand; ++c) {
    KeyIdx curv{-INFINITY, kNoIdx};
    float x = kLowest;
    if (on && head < ncand) {
      curv = KeyIdx{pb.key[base + head], pb.idx[base + head]};
      x = pb.x[base + head];
      // when forced, masked timestamp tokens tie with every (now masked) text token, which wins on index: drop them here
      if (forced && !(x > kLowest * 0.5f)) curv = KeyIdx{-INFINITY, kNoIdx};
    }
    if (curv.v == -INFINITY) curv.i = kNoIdx;
    const KeyIdx best = warp_pick(curv);
    const bool mine = best.i != kNoIdx && curv.i == best.i;
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const float bx = __shfl_sync(0xffffffffu, x, who ? __ffs(who) - 1 : 0);  // the winner's post-processor logit
    if (mine) ++head;
    if (lane == 0) {
      float sc;
      int tok;
      if (best.i == kNoIdx) {
        if (sampling) {
          sc = kLowest;
          tok = 0;
        } else {
          // fewer live candidates than requested: search.cu continues with the masked ("lowest") entries in index order
          sc = kLowest - lse;
          tok = filler++;
        }
      } else if (sampling) {
        const float t = (bx - lse) * inv_t;
        const float g = gumbel(p.seed, r, step, best.i);
        sc = (t + g) - g;  // the draw's tempered log-prob, with search.cu's rounding
        tok = best.i;
      } else {
        sc = bx - lse;
        tok = best.i;
      }
      bf.cand_score[(long long)r * kMaxCand + c] = (p.mode == 0) ? cum + sc : sc;
      bf.cand_tok[(long long)r * kMaxCand + c] = tok;
    }
  }
}

void search_v2_configure() {
  B2W_CUDA(cudaFuncSetAttribute(search_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
}

// (re)allocates the partial buffers for up to R rows; call outside stream capture
void search_v2_reserve(SearchPartBuffers& pb, int& capacity_rows, int R) {
  if (R <= capacity_rows) return;
  if (pb.stats) {
    cudaFree(pb.stats);
    cudaFree(pb.key);
    cudaFree(pb.x);
    cudaFree(pb.idx);
  }
  const size_t lists = (size_t)R * kV2Slices * 2 * kMaxCand;
  B2W_CUDA(cudaMalloc(&pb.stats, (size_t)R * kV2Slices * kV2Stats * sizeof(float)));
  B2W_CUDA(cudaMalloc(&pb.key, lists * sizeof(float)));
  B2W_CUDA(cudaMalloc(&pb.x, lists * sizeof(float)));
  B2W_CUDA(cudaMalloc(&pb.idx, lists * sizeof(int)));
  B2W_CUDA(cudaMemset(pb.stats, 0, (size_t)R * kV2Slices * kV2Stats * sizeof(float)));
  capacity_rows = R;
}

void search_rows_v2(const float* logits, int R, int vpad, const SearchBuffers& b, const SearchPartBuffers& pb, cudaStream_t s) {
  const int SL = (((vpad + kV2Slices - 1) / kV2Slices) + 3) & ~3;
  B2W_CHECK(SL * (int)sizeof(float) <= 64 * 1024, "vocabulary slice too large");
  search_part_kernel<<<dim3(kV2Slices, R), kV2Threads, SL * sizeof(float), s>>>(logits, b, pb);
  B2W_LAUNCHED();
  search_merge_kernel<<<R, 32, 0, s>>>(logits, b, pb);
  B2W_LAUNCHED();
}

}  // namespace b2w
