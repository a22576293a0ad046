// K5  exact k-th smallest value (0-based rank k) of n floats by MSB-first radix
// select on the order-preserving uint32 key.  Replaces the reference's full
// sort (qsparse/util.py:113-116: values = sort(flat); thr = values[idx + 1]) —
// O(n) reads instead of O(n log n), result bit-identical to sort()[k].
//
// Three streaming passes over v (digits of 8 / 12 / 12 bits):
//   pass 0: the top 8 bits (sign + 7 exponent bits) of real data land in a
//           handful of bins, so the shared-memory histogram is laid out
//           [256 bins][32 lanes]: lane l only ever touches bank l — no bank
//           conflicts and no same-address serialisation inside a warp.
//   pass 1/2: only elements whose higher bits equal the selected prefix are
//           counted; inside one 8-bit bucket the next 12 bits are close to
//           uniform, so a plain 4096-bin shared histogram is contention free.
// Each later pass re-derives the prefix from the global histograms of the
// earlier passes in its prologue (4-16 KB from L2), so there is no separate
// scan launch.  Counters are 64-bit: n may exceed 2^32.
//
// Large inputs (n >= 2^22) take a ~1-pass route first:
//   sample   16384 evenly spaced keys, select two pivots lo <= hi around rank k of the
//            sample (+-5 sigma of the binomial rank error), one CTA;
//   partition ONE streaming pass over v: count keys < lo, append lo <= key <= hi
//            (~2 % of n) to a candidate buffer (warp-aggregated atomics);
//   select   the exact 3-pass radix select on the candidates with rank k - count_lt.
// If rank k is not inside [lo, hi] or the candidates overflow the buffer (heavy ties)
// the full 3-pass select over v runs instead.  Both continuations are always
// launched — the decision is taken on the device from the counters, every CTA of
// the route not taken exits at once — so there is no host synchronisation.
#include "qsb_common.cuh"

namespace qsb {

constexpr int kBits0 = 8, kBits1 = 12, kBits2 = 12;
constexpr int kBins0 = 1 << kBits0, kBins1 = 1 << kBits1, kBins2 = 1 << kBits2;
static_assert(kBins0 == QSB_THREADS, "pass-0 merge maps one thread to one bin");

struct SelectWs {
  unsigned long long *hist0;  // [256]
  unsigned long long *hist1;  // [4096]
  unsigned long long *hist2;  // [4096]
};

// fast-route state (device memory, zeroed before every call)
struct SelState {
  unsigned long long count_lt;  // keys < lo
  unsigned long long n_cand;    // keys in [lo, hi] (attempted appends, may exceed cap)
  uint32_t lo_key, hi_key;
};

enum { kModePlain = 0, kModeCandidates = 1, kModeFallback = 2 };

// Is rank k inside the candidate set, and did every candidate fit?  Evaluated by every
// CTA of both continuations from the same counters, so they always agree.
__device__ __forceinline__ bool sel_fast_valid(const SelState *st, unsigned long long k,
                                               unsigned long long cap,
                                               unsigned long long *k_in_cand,
                                               unsigned long long *n_cand) {
  const unsigned long long lt = st->count_lt, nc = st->n_cand;
  *k_in_cand = k - lt;
  *n_cand = nc;
  return nc <= cap && k >= lt && (k - lt) < nc;
}

// Find the bin where the running count first exceeds k.  All threads of the CTA
// call it; result broadcast through shared memory.  k_inout becomes the rank
// inside the chosen bin.
template <int BINS>
__device__ void find_bin(const unsigned long long *hist, unsigned long long k,
                         uint32_t *bin_out, unsigned long long *k_out) {
  constexpr int PER = (BINS + QSB_THREADS - 1) / QSB_THREADS;
  __shared__ unsigned long long s_part[QSB_THREADS];
  __shared__ uint32_t s_bin;
  __shared__ unsigned long long s_k;
  const int tid = threadIdx.x;
  unsigned long long local[PER];
  unsigned long long sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int b = tid * PER + i;
    local[i] = (b < BINS) ? hist[b] : 0ull;
    sum += local[i];
  }
  s_part[tid] = sum;
  if (tid == 0) {
    s_bin = BINS - 1;
    s_k = 0;
  }
  __syncthreads();
  // exclusive prefix of this thread's chunk (256 sequential adds: negligible)
  unsigned long long before = 0;
  for (int j = 0; j < tid; ++j) before += s_part[j];
  if (k >= before && k < before + sum) {
    unsigned long long run = before;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (k < run + local[i]) {
        s_bin = tid * PER + i;
        s_k = k - run;
        break;
      }
      run += local[i];
    }
  }
  __syncthreads();
  *bin_out = s_bin;
  *k_out = s_k;
  __syncthreads();
}

// The prefix tests run on the raw IEEE bits instead of the converted key: inside
// one 8-bit key bucket every value has the same sign, so "top bits of the key ==
// prefix" is "top bits of the raw word == a per-launch constant", and the next
// digit is the raw field XOR a per-launch constant.  Only negative NaNs need a fix
// up front (they must order last, like every NaN): they are remapped to 0x7fffffff.
struct PassConst {
  uint32_t raw_prefix;  // PASS 1: wanted raw bits >> 24; PASS 2: wanted raw bits >> 12
  uint32_t flip;        // 0xfff for a negative bucket (key = ~bits), else 0
};

__device__ __forceinline__ PassConst make_pass_const(int pass, uint32_t key_prefix) {
  PassConst pc;
  if (pass == 1) {
    const bool neg = (key_prefix & 0x80u) == 0;          // key top bit clear <=> negative
    pc.raw_prefix = neg ? (key_prefix ^ 0xffu) : (key_prefix ^ 0x80u);
    pc.flip = neg ? 0xfffu : 0u;
  } else {
    const bool neg = (key_prefix & 0x80000u) == 0;
    pc.raw_prefix = neg ? (~key_prefix & 0xfffffu) : (key_prefix ^ 0x80000u);
    pc.flip = neg ? 0xfffu : 0u;
  }
  return pc;
}

template <int PASS, bool ABS>
__device__ __forceinline__ void count_value(float f, const PassConst &pc,
                                            uint32_t *s_hist, int lane) {
  uint32_t b = __float_as_uint(f);
  if constexpr (ABS) b &= 0x7fffffffu;
  else b = (b > 0xff800000u) ? 0x7fffffffu : b;
  if constexpr (PASS == 0) {
    const uint32_t t = b >> 24;
    const uint32_t digit = ((int32_t)b < 0) ? (t ^ 0xffu) : (t | 0x80u);
    atomicAdd(&s_hist[digit * 32 + lane], 1u);
  } else if constexpr (PASS == 1) {
    if ((b >> 24) == pc.raw_prefix)
      atomicAdd(&s_hist[((b >> 12) & 0xfffu) ^ pc.flip], 1u);
  } else {
    if ((b >> 12) == pc.raw_prefix) atomicAdd(&s_hist[(b & 0xfffu) ^ pc.flip], 1u);
  }
}

template <int PASS, int V, bool ABS>
__global__ void __launch_bounds__(QSB_THREADS)
    select_hist_kernel(const float *__restrict__ v, int64_t n, int64_t k,
                       SelectWs ws, const SelState *st, int mode, int64_t cap) {
  if (mode != kModePlain) {
    unsigned long long kc, nc;
    const bool valid = sel_fast_valid(st, (unsigned long long)k, (unsigned long long)cap, &kc, &nc);
    if ((mode == kModeCandidates) != valid) return;  // the other continuation runs
    if (mode == kModeCandidates) {
      n = (int64_t)nc;
      k = (int64_t)kc;
    }
  }
  constexpr int kSmemWords = (PASS == 0) ? kBins0 * 32 : kBins1;
  __shared__ uint32_t s_hist[kSmemWords];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < kSmemWords; i += QSB_THREADS) s_hist[i] = 0;

  PassConst pc{0, 0};
  if constexpr (PASS >= 1) {
    unsigned long long kk;
    uint32_t b0;
    find_bin<kBins0>(ws.hist0, (unsigned long long)k, &b0, &kk);
    uint32_t prefix = b0;
    if constexpr (PASS == 2) {
      uint32_t b1;
      find_bin<kBins1>(ws.hist1, kk, &b1, &kk);
      prefix = (b0 << 12) | b1;
    }
    pc = make_pass_const(PASS, prefix);
  }
  __syncthreads();

  constexpr int U = (V == 8) ? 2 : 4;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  const int64_t n_main = (n / V) * V;
  for (int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)tid * V;
       base < n_main; base += (int64_t)gridDim.x * kTile) {
    VecF<V> x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = base + (int64_t)u * QSB_THREADS * V;
      if (e < n_main) x[u] = ld_vec<V, Hint::KEEP>(v + e);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = base + (int64_t)u * QSB_THREADS * V;
      if (e < n_main) {
#pragma unroll
        for (int j = 0; j < V; ++j)
          count_value<PASS, ABS>(x[u].v[j], pc, s_hist, lane);
      }
    }
  }
  if (blockIdx.x == 0) {
    const int64_t e = n_main + tid;
    if (e < n) count_value<PASS, ABS>(v[e], pc, s_hist, lane);
  }
  __syncthreads();

  if constexpr (PASS == 0) {
    // bin = tid; rotate the lane index so the 32 reads of a warp hit 32 banks
    unsigned long long sum = 0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) sum += s_hist[tid * 32 + ((l + tid) & 31)];
    if (sum) atomicAdd(&ws.hist0[tid], sum);
  } else {
    unsigned long long *g = (PASS == 1) ? ws.hist1 : ws.hist2;
    for (int b = tid; b < kBins1; b += QSB_THREADS) {
      const uint32_t c = s_hist[b];
      if (c) atomicAdd(&g[b], (unsigned long long)c);
    }
  }
}

__global__ void __launch_bounds__(QSB_THREADS)
    select_final_kernel(int64_t k, SelectWs ws, SelectWs ws_cand, const SelState *st,
                        int64_t cap, float *thr_out) {
  if (st) {
    unsigned long long kc, nc;
    if (sel_fast_valid(st, (unsigned long long)k, (unsigned long long)cap, &kc, &nc)) {
      ws = ws_cand;
      k = (int64_t)kc;
    }
  }
  unsigned long long kk;
  uint32_t b0, b1, b2;
  find_bin<kBins0>(ws.hist0, (unsigned long long)k, &b0, &kk);
  find_bin<kBins1>(ws.hist1, kk, &b1, &kk);
  find_bin<kBins2>(ws.hist2, kk, &b2, &kk);
  if (threadIdx.x == 0)
    *thr_out = key_to_float((b0 << 24) | (b1 << 12) | b2);
}

template <int PASS, int V, bool ABS>
static int launch_hist(const float *v, int64_t n, int64_t k, const SelectWs &ws,
                       cudaStream_t stream, const SelState *st = nullptr,
                       int mode = kModePlain, int64_t cap = 0) {
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    QSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &o, select_hist_kernel<PASS, V, ABS>, QSB_THREADS, 0));
    occ = o > 0 ? o : 1;
  }
  constexpr int U = (V == 8) ? 2 : 4;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  int64_t grid = (int64_t)device_props().sm_count * occ;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  select_hist_kernel<PASS, V, ABS>
      <<<(unsigned)grid, QSB_THREADS, 0, stream>>>(v, n, k, ws, st, mode, cap);
  QSB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// fast route, step 1: pivots from a sample (one CTA, 1024 threads, 64 KB of keys)
// ---------------------------------------------------------------------------
constexpr int kSampleSize = 16384;

__device__ __forceinline__ uint32_t sel_key(float f, bool take_abs) {
  return float_to_key(take_abs ? fabsf(f) : f);
}

// rank-r key (0-based) of keys[0..m) in shared memory; all 1024 threads call it
__device__ uint32_t smem_select(const uint32_t *keys, int m, int r, uint32_t *hist,
                                uint32_t *bcast) {
  uint32_t prefix = 0, mask_hi = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const uint32_t key = keys[i];
      if ((key & mask_hi) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t local[8], sum = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        local[b] = hist[threadIdx.x * 8 + b];
        sum += local[b];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)threadIdx.x >= o) incl += t;
      }
      uint32_t run = incl - sum;  // keys in the bins before this lane's
      if ((uint32_t)r >= run && (uint32_t)r < incl) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if ((uint32_t)r < run + local[b]) {
            bcast[0] = threadIdx.x * 8 + b;
            bcast[1] = (uint32_t)r - run;
            break;
          }
          run += local[b];
        }
      }
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask_hi |= 0xffu << shift;
    r = (int)bcast[1];
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(1024)
    select_sample_kernel(const float *__restrict__ v, int64_t n, int take_abs, int r_lo,
                         int r_hi, SelState *st) {
  extern __shared__ uint32_t s_keys[];  // kSampleSize keys
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_bcast[2];
  const int64_t stride = n / kSampleSize;
  for (int i = threadIdx.x; i < kSampleSize; i += blockDim.x)
    s_keys[i] = sel_key(v[(int64_t)i * stride + (stride >> 1)], take_abs != 0);
  __syncthreads();
  const uint32_t lo = (r_lo < 0) ? 0u : smem_select(s_keys, kSampleSize, r_lo, s_hist, s_bcast);
  const uint32_t hi =
      (r_hi >= kSampleSize) ? 0xffffffffu : smem_select(s_keys, kSampleSize, r_hi, s_hist, s_bcast);
  if (threadIdx.x == 0) {
    st->lo_key = lo;
    st->hi_key = hi;
  }
}

// ---------------------------------------------------------------------------
// fast route, step 2: count keys < lo, compact keys in [lo, hi].  One CTA per tile.
// The candidates are stored as values (|v| already applied), so the candidate select
// runs on plain floats.
// ---------------------------------------------------------------------------
template <int V, bool ABS>
__global__ void __launch_bounds__(QSB_THREADS)
    select_partition_kernel(const float *__restrict__ v, int64_t n, SelState *st,
                            float *__restrict__ cand, int64_t cap) {
  constexpr int U = 2;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  __shared__ unsigned long long s_lt[QSB_THREADS / 32];
  const uint32_t lo = st->lo_key, span = st->hi_key - st->lo_key;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t n_main = (n / V) * V;
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)tid * V;
  VecF<V> x[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * QSB_THREADS * V;
    if (e < n_main) x[u] = ld_vec<V, Hint::KEEP>(v + e);
  }
  uint32_t lt = 0, nc = 0;
  uint32_t cmask = 0;  // bit (u * V + j) set: element is a candidate
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * QSB_THREADS * V;
    if (e < n_main) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const uint32_t key = float_to_key(ABS ? fabsf(x[u].v[j]) : x[u].v[j]);
        lt += key < lo;
        const bool c = (key - lo) <= span;
        nc += c;
        cmask |= (uint32_t)c << (u * V + j);
      }
    }
  }
  // the last n % V elements: the owning CTA's first threads, scalar
  float tail_val = 0.f;
  bool tail_c = false;
  if (base - (int64_t)tid * V <= n_main && n_main < base - (int64_t)tid * V + kTile) {
    const int64_t e = n_main + tid;
    if (e < n) {
      tail_val = ABS ? fabsf(v[e]) : v[e];
      const uint32_t key = float_to_key(tail_val);
      lt += key < lo;
      tail_c = (key - lo) <= span;
      nc += tail_c;
    }
  }
  // warp-aggregated append: one atomic per warp that has candidates
  uint32_t incl = nc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total) {
    unsigned long long wbase = 0;
    if (lane == 31) wbase = atomicAdd(&st->n_cand, (unsigned long long)total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    unsigned long long pos = wbase + (incl - nc);
    if (nc) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (cmask & (1u << (u * V + j))) {
            if (pos < (unsigned long long)cap) cand[pos] = ABS ? fabsf(x[u].v[j]) : x[u].v[j];
            ++pos;
          }
      if (tail_c && pos < (unsigned long long)cap) cand[pos] = tail_val;
    }
  }
  // keys < lo: one atomic per CTA
  unsigned long long wl = warp_reduce((unsigned long long)lt,
                                      [](unsigned long long a, unsigned long long b) { return a + b; });
  if (lane == 0) s_lt[tid >> 5] = wl;
  __syncthreads();
  if (tid == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < QSB_THREADS / 32; ++w) t += s_lt[w];
    if (t) atomicAdd(&st->count_lt, t);
  }
}

constexpr int64_t kHistBytes =
    (int64_t)(kBins0 + kBins1 + kBins2) * sizeof(unsigned long long);
// [hist x3 | candidate hist x3 | SelState | pad] then the candidate buffer
constexpr int64_t kSelectHeaderBytes = 2 * kHistBytes + 256;
constexpr int64_t kFastMinN = 1 << 22;
static int g_select_fast = 1;  // tuning key 4

static int64_t select_cap(int64_t n) { return n >= kFastMinN ? n / 8 + 4096 : 0; }

void set_select_fast(int v) { g_select_fast = v; }

}  // namespace qsb

using namespace qsb;

extern "C" int64_t qsb_kth_workspace_bytes(int64_t n) {
  return kSelectHeaderBytes + 256 + select_cap(n) * (int64_t)sizeof(float) + 32;
}

template <int V, bool ABS>
static int run_passes(const float *v, int64_t n, int64_t k, const SelectWs &ws,
                      cudaStream_t stream, const SelState *st = nullptr,
                      int mode = kModePlain, int64_t cap = 0) {
  int rc;
  if ((rc = launch_hist<0, V, ABS>(v, n, k, ws, stream, st, mode, cap))) return rc;
  if ((rc = launch_hist<1, V, ABS>(v, n, k, ws, stream, st, mode, cap))) return rc;
  return launch_hist<2, V, ABS>(v, n, k, ws, stream, st, mode, cap);
}

extern "C" int qsb_kth_value(const float *v, int64_t n, int64_t k, int take_abs,
                             float *thr_out_dev, void *workspace,
                             int64_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0 || k < 0 || k >= n) return QSB_E_BADARG;
  if (!v || !thr_out_dev || !workspace) return QSB_E_BADARG;
  if (!aligned_to(v, 4)) return QSB_E_ALIGN;
  if (workspace_bytes < qsb_kth_workspace_bytes(n)) return QSB_E_WORKSPACE;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256;
  SelectWs ws, wc;
  ws.hist0 = reinterpret_cast<unsigned long long *>(base);
  ws.hist1 = ws.hist0 + kBins0;
  ws.hist2 = ws.hist1 + kBins1;
  wc.hist0 = ws.hist2 + kBins2;
  wc.hist1 = wc.hist0 + kBins0;
  wc.hist2 = wc.hist1 + kBins1;
  SelState *st = reinterpret_cast<SelState *>(wc.hist2 + kBins2);
  float *cand = reinterpret_cast<float *>((base + kSelectHeaderBytes + 31) / 32 * 32);
  QSB_CUDA_TRY(cudaMemsetAsync(ws.hist0, 0, kSelectHeaderBytes, stream));
  const bool v32 = aligned_to(v, 32);
  const bool fast = g_select_fast && n >= kFastMinN && v32;
  int rc;
  if (fast) {
    const int64_t cap = select_cap(n);
    // sample ranks bracketing k: +-5 sigma of the binomial rank error, +3
    const double m = (double)kSampleSize, p = (double)k / (double)n;
    const double delta = 5.0 * sqrt(m * p * (1.0 - p)) + 3.0;
    const int r_lo = (int)floor(p * m - delta), r_hi = (int)ceil(p * m + delta);
    static bool smem_set = false;
    if (!smem_set) {
      QSB_CUDA_TRY(cudaFuncSetAttribute(select_sample_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kSampleSize * (int)sizeof(uint32_t)));
      smem_set = true;
    }
    select_sample_kernel<<<1, 1024, kSampleSize * sizeof(uint32_t), stream>>>(
        v, n, take_abs, r_lo, r_hi, st);
    QSB_LAUNCH_CHECK();
    constexpr int64_t kTile = (int64_t)QSB_THREADS * 8 * 2;
    const int64_t tiles = (n + kTile - 1) / kTile;
    if (take_abs)
      select_partition_kernel<8, true><<<(unsigned)tiles, QSB_THREADS, 0, stream>>>(v, n, st, cand, cap);
    else
      select_partition_kernel<8, false><<<(unsigned)tiles, QSB_THREADS, 0, stream>>>(v, n, st, cand, cap);
    QSB_LAUNCH_CHECK();
    // continuation A: exact select on the candidates (runs only if rank k is inside)
    if ((rc = run_passes<8, false>(cand, cap, k, wc, stream, st, kModeCandidates, cap))) return rc;
    // continuation B: the full select (runs only if A is not valid)
    rc = take_abs ? run_passes<8, true>(v, n, k, ws, stream, st, kModeFallback, cap)
                  : run_passes<8, false>(v, n, k, ws, stream, st, kModeFallback, cap);
    if (rc) return rc;
    select_final_kernel<<<1, QSB_THREADS, 0, stream>>>(k, ws, wc, st, cap, thr_out_dev);
    QSB_LAUNCH_CHECK();
    return 0;
  }
  if (v32)
    rc = take_abs ? run_passes<8, true>(v, n, k, ws, stream)
                  : run_passes<8, false>(v, n, k, ws, stream);
  else
    rc = take_abs ? run_passes<1, true>(v, n, k, ws, stream)
                  : run_passes<1, false>(v, n, k, ws, stream);
  if (rc) return rc;
  select_final_kernel<<<1, QSB_THREADS, 0, stream>>>(k, ws, ws, nullptr, 0, thr_out_dev);
  QSB_LAUNCH_CHECK();
  return 0;
}
