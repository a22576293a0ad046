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
  unsigned long long n_cand;    // keys in [lo, hi]
  uint32_t lo_key, hi_key;
  uint32_t overflow;            // some tile had more candidates than its region holds
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
  (void)cap;
  return st->overflow == 0 && k >= lt && (k - lt) < nc;
}

// Find the bin where the running count first exceeds k.  All threads of the CTA
// call it; result broadcast through shared memory.  k_inout becomes the rank
// inside the chosen bin.
template <int BINS>
__device__ void find_bin(const unsigned long long *hist, unsigned long long k,
                         uint32_t *bin_out, unsigned long long *k_out) {
  constexpr int PER = (BINS + QSB_THREADS - 1) / QSB_THREADS;
  __shared__ unsigned long long s_warp[QSB_THREADS / 32];
  __shared__ uint32_t s_bin;
  __shared__ unsigned long long s_k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long local[PER];
  unsigned long long sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int b = tid * PER + i;
    local[i] = (b < BINS) ? hist[b] : 0ull;
    sum += local[i];
  }
  // block-wide exclusive scan of the per-thread sums (warp shuffles + one smem hop)
  unsigned long long incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  if (tid == 0) {
    s_bin = BINS - 1;
    s_k = 0;
  }
  __syncthreads();
  unsigned long long before = incl - sum;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
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
// fast route, step 1: pivots from a sample (one CTA, 1024 threads): 16384 evenly
// spaced keys, bitonic sort in 64 KB of shared memory, pivots = two order statistics
// ---------------------------------------------------------------------------
constexpr int kSampleSize = 16384;

__global__ void __launch_bounds__(1024)
    select_sample_kernel(const float *__restrict__ v, int64_t n, int take_abs, int r_lo,
                         int r_hi, SelState *st) {
  extern __shared__ uint32_t s_keys[];  // kSampleSize keys
  const int64_t stride = n / kSampleSize;
  for (int i = threadIdx.x; i < kSampleSize; i += blockDim.x) {
    const float f = v[(int64_t)i * stride + (stride >> 1)];
    s_keys[i] = float_to_key(take_abs ? fabsf(f) : f);
  }
  __syncthreads();
  for (int size = 2; size <= kSampleSize; size <<= 1) {
    for (int step = size >> 1; step > 0; step >>= 1) {
      for (int t = threadIdx.x; t < kSampleSize / 2; t += blockDim.x) {
        const int i = ((t / step) * step * 2) + (t % step);  // lower index of the pair
        const int j = i + step;
        const bool up = ((i & size) == 0);
        const uint32_t a = s_keys[i], b = s_keys[j];
        if ((a > b) == up) {
          s_keys[i] = b;
          s_keys[j] = a;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    st->lo_key = (r_lo < 0) ? 0u : s_keys[r_lo];
    st->hi_key = (r_hi >= kSampleSize) ? 0xffffffffu : s_keys[r_hi];
  }
}

// ---------------------------------------------------------------------------
// fast route, step 2: ONE streaming pass.  Each CTA owns a tile of 4096 elements and
// a private region of kRegion candidate slots: it counts its keys < lo and compacts
// its keys in [lo, hi] into the region — no atomics, no ordering between CTAs.
// The candidates are stored as values (|v| already applied).
// ---------------------------------------------------------------------------
constexpr int kRegion = 512;  // candidate slots per 4096-element tile (12.5 %; ~4 % expected)

template <int V, bool ABS>
__global__ void __launch_bounds__(QSB_THREADS)
    select_partition_kernel(const float *__restrict__ v, int64_t n, const SelState *st,
                            float *__restrict__ cand, uint32_t *__restrict__ cnt,
                            uint32_t *__restrict__ lt_arr) {
  constexpr int U = 2;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  static_assert(kTile == 4096, "region bookkeeping assumes 4096-element tiles");
  __shared__ uint32_t s_nc[QSB_THREADS / 32], s_lt[QSB_THREADS / 32];
  const uint32_t lo = st->lo_key, span = st->hi_key - st->lo_key;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_main = (n / V) * V;
  const int64_t t0 = (int64_t)blockIdx.x * kTile;
  const int64_t base = t0 + (int64_t)tid * V;
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
  if (t0 <= n_main && n_main < t0 + kTile) {
    const int64_t e = n_main + tid;
    if (e < n) {
      tail_val = ABS ? fabsf(v[e]) : v[e];
      const uint32_t key = float_to_key(tail_val);
      lt += key < lo;
      tail_c = (key - lo) <= span;
      nc += tail_c;
    }
  }
  // CTA-wide exclusive scan of the candidate counts
  uint32_t incl = nc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t wlt = warp_reduce(lt, [](uint32_t a, uint32_t b) { return a + b; });
  if (lane == 31) s_nc[warp] = incl;
  if (lane == 0) s_lt[warp] = wlt;
  __syncthreads();
  uint32_t pos = incl - nc;
  for (int w = 0; w < warp; ++w) pos += s_nc[w];
  if (tid == 0) {
    uint32_t total = 0, tlt = 0;
    for (int w = 0; w < QSB_THREADS / 32; ++w) {
      total += s_nc[w];
      tlt += s_lt[w];
    }
    cnt[blockIdx.x] = total;  // > kRegion marks an overflow
    lt_arr[blockIdx.x] = tlt;
  }
  if (nc) {
    float *region = cand + (int64_t)blockIdx.x * kRegion;
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (cmask & (1u << (u * V + j))) {
          if (pos < (uint32_t)kRegion) region[pos] = ABS ? fabsf(x[u].v[j]) : x[u].v[j];
          ++pos;
        }
    if (tail_c && pos < (uint32_t)kRegion) region[pos] = tail_val;
  }
}

// step 3: totals (one CTA): count_lt, n_cand, overflow
__global__ void __launch_bounds__(1024)
    select_decide_kernel(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ lt_arr,
                         int64_t tiles, SelState *st) {
  __shared__ unsigned long long s_a[32], s_b[32];
  __shared__ uint32_t s_o[32];
  unsigned long long a = 0, b = 0;
  uint32_t o = 0;
  for (int64_t i = threadIdx.x; i < tiles; i += blockDim.x) {
    const uint32_t c = cnt[i];
    a += lt_arr[i];
    b += c;
    o |= (c > (uint32_t)kRegion);
  }
  a = warp_reduce(a, [](unsigned long long x, unsigned long long y) { return x + y; });
  b = warp_reduce(b, [](unsigned long long x, unsigned long long y) { return x + y; });
  o = __any_sync(0xffffffffu, o);
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = a;
    s_b[threadIdx.x >> 5] = b;
    s_o[threadIdx.x >> 5] = o;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = 0;
    o = 0;
    for (int w = 0; w < 32; ++w) {
      a += s_a[w];
      b += s_b[w];
      o |= s_o[w];
    }
    st->count_lt = a;
    st->n_cand = b;
    st->overflow = o;
  }
}

// step 4: radix-select histogram passes over the candidate regions, a warp per region
template <int PASS>
__global__ void __launch_bounds__(QSB_THREADS)
    select_hist_regions_kernel(const float *__restrict__ cand, const uint32_t *__restrict__ cnt,
                               int64_t tiles, int64_t k, SelectWs ws, const SelState *st) {
  unsigned long long kc, nc;
  if (!sel_fast_valid(st, (unsigned long long)k, 0, &kc, &nc)) return;
  constexpr int kSmemWords = (PASS == 0) ? kBins0 * 32 : kBins1;
  __shared__ uint32_t s_hist[kSmemWords];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < kSmemWords; i += QSB_THREADS) s_hist[i] = 0;
  PassConst pc{0, 0};
  if constexpr (PASS >= 1) {
    unsigned long long kk;
    uint32_t b0;
    find_bin<kBins0>(ws.hist0, kc, &b0, &kk);
    uint32_t prefix = b0;
    if constexpr (PASS == 2) {
      uint32_t b1;
      find_bin<kBins1>(ws.hist1, kk, &b1, &kk);
      prefix = (b0 << 12) | b1;
    }
    pc = make_pass_const(PASS, prefix);
  }
  __syncthreads();
  const int64_t warps_total = (int64_t)gridDim.x * (QSB_THREADS / 32);
  for (int64_t r = (int64_t)blockIdx.x * (QSB_THREADS / 32) + (tid >> 5); r < tiles; r += warps_total) {
    const uint32_t c = cnt[r];
    const float *region = cand + r * kRegion;
    for (uint32_t i = lane; i < c; i += 32) count_value<PASS, false>(region[i], pc, s_hist, lane);
  }
  __syncthreads();
  if constexpr (PASS == 0) {
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

constexpr int64_t kHistBytes =
    (int64_t)(kBins0 + kBins1 + kBins2) * sizeof(unsigned long long);
// [hist x3 | candidate hist x3 | SelState | pad] then the candidate buffer
constexpr int64_t kSelectHeaderBytes = 2 * kHistBytes + 256;
constexpr int64_t kFastMinN = 1 << 22;
static int g_select_fast = 1;  // tuning key 4

static int64_t select_tiles(int64_t n) { return n >= kFastMinN ? (n + 4095) / 4096 : 0; }

void set_select_fast(int v) { g_select_fast = v; }

}  // namespace qsb

using namespace qsb;

extern "C" int64_t qsb_kth_workspace_bytes(int64_t n) {
  // header | cnt[tiles] | lt[tiles] | candidate regions [tiles][512]
  const int64_t tiles = select_tiles(n);
  return kSelectHeaderBytes + 256 + tiles * 8 + 64 + tiles * kRegion * (int64_t)sizeof(float) + 32;
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
  const int64_t tiles = select_tiles(n);
  uint32_t *cnt = reinterpret_cast<uint32_t *>(base + kSelectHeaderBytes);
  uint32_t *lt_arr = cnt + tiles;
  float *cand = reinterpret_cast<float *>(
      (reinterpret_cast<uintptr_t>(lt_arr + tiles) + 31) / 32 * 32);
  QSB_CUDA_TRY(cudaMemsetAsync(ws.hist0, 0, kSelectHeaderBytes, stream));
  const bool v32 = aligned_to(v, 32);
  const bool fast = g_select_fast && n >= kFastMinN && v32;
  int rc;
  if (fast) {
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
    if (take_abs)
      select_partition_kernel<8, true><<<(unsigned)tiles, QSB_THREADS, 0, stream>>>(v, n, st, cand, cnt, lt_arr);
    else
      select_partition_kernel<8, false><<<(unsigned)tiles, QSB_THREADS, 0, stream>>>(v, n, st, cand, cnt, lt_arr);
    QSB_LAUNCH_CHECK();
    select_decide_kernel<<<1, 1024, 0, stream>>>(cnt, lt_arr, tiles, st);
    QSB_LAUNCH_CHECK();
    // continuation A: exact select on the candidate regions (runs only if rank k is inside)
    {
      int64_t grid = (int64_t)device_props().sm_count * 4;
      const int64_t need = (tiles + QSB_THREADS / 32 - 1) / (QSB_THREADS / 32);
      if (grid > need) grid = need;
      select_hist_regions_kernel<0><<<(unsigned)grid, QSB_THREADS, 0, stream>>>(cand, cnt, tiles, k, wc, st);
      select_hist_regions_kernel<1><<<(unsigned)grid, QSB_THREADS, 0, stream>>>(cand, cnt, tiles, k, wc, st);
      select_hist_regions_kernel<2><<<(unsigned)grid, QSB_THREADS, 0, stream>>>(cand, cnt, tiles, k, wc, st);
      QSB_LAUNCH_CHECK();
    }
    // continuation B: the full select (runs only if A is not valid)
    rc = take_abs ? run_passes<8, true>(v, n, k, ws, stream, st, kModeFallback, 0)
                  : run_passes<8, false>(v, n, k, ws, stream, st, kModeFallback, 0);
    if (rc) return rc;
    select_final_kernel<<<1, QSB_THREADS, 0, stream>>>(k, ws, wc, st, 0, thr_out_dev);
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
