// K5  exact k-th smallest value (0-based rank k) of n floats.  Replaces the reference's
// full sort (qsparse/util.py:113-116: values = sort(flat); thr = values[idx + 1]) —
// O(n) reads instead of O(n log n); the result equals sort()[k] (as a float; the sign
// of a zero result and the payload of a NaN result are not defined by sort either).
//
// Core: MSB-first radix select on the order-preserving uint32 key, digits 8 / 12 / 12:
//   pass 0: the top 8 bits (sign + 7 exponent bits) of real data land in a handful of
//           bins, so the shared-memory histogram is laid out [256 bins][32 lanes]:
//           lane l only ever touches bank l — no bank conflicts, no same-address
//           serialisation inside a warp.
//   pass 1/2: only elements whose higher bits equal the selected prefix are counted;
//           the tests run on the raw IEEE bits (inside one 8-bit key bucket every value
//           has the same sign, so "top bits of the key == prefix" is a compare of the raw
//           word against a per-launch constant).
// Each pass re-derives the prefix from the global histograms of the earlier passes in
// its prologue; the last CTA of pass 2 (atomic ticket) turns the histograms into the
// answer, so a select is memset + 3 launches.  Counters are 64-bit (n may exceed 2^32).
//
// Large inputs (n >= 2^22) first try a ~1-pass route:
//   sample    4096 evenly spaced values; their exact ranks inside the sample are
//             counted by 128 CTAs; two pivots lo <= hi = the order statistics +-5 sigma
//             (binomial rank error) around rank k * 4096 / n;
//   partition ONE streaming pass over v: every CTA counts its values < lo and compacts
//             its values in [lo, hi] (~8 %) into a private 512-slot region of the
//             candidate buffer — no atomics on the critical path, totals by RED;
//   passes    the same 3 radix passes, but over the candidate regions (a warp per
//             region) with rank k - count_lt.
// If rank k is not inside [lo, hi] or a region overflows (heavy ties), the 3 passes run
// over v itself.  The route is chosen ON THE DEVICE in the prologue of each pass from
// the counters — one launch per pass either way, no host synchronisation.
#include <math.h>

#include "qsb_common.cuh"

namespace qsb {

constexpr int kBits0 = 8, kBits1 = 12, kBits2 = 12;
constexpr int kBins0 = 1 << kBits0, kBins1 = 1 << kBits1, kBins2 = 1 << kBits2;
static_assert(kBins0 == QSB_THREADS, "pass-0 merge maps one thread to one bin");

constexpr int kSampleSize = 4096;
constexpr int kSampleParts = 8;    // threads that share the rank count of one sample
constexpr int kRegion = 512;       // candidate slots per 4096-element tile (12.5 %)
constexpr int64_t kFastMinN = 1 << 22;

struct SelectWs {
  unsigned long long *hist0;  // [256]
  unsigned long long *hist1;  // [4096]
  unsigned long long *hist2;  // [4096]
};

// device-side state of one select call (zeroed by the memset that starts it)
struct SelState {
  unsigned long long count_lt;  // values < lo            (fast route)
  unsigned long long n_cand;    // values in [lo, hi]     (fast route)
  float lo, hi;                 // pivots
  uint32_t overflow;            // a tile had more candidates than its region holds
  uint32_t done;                // CTAs of pass 2 that have finished (ticket)
};

// Is rank k inside the candidate set, and did every candidate fit?  Evaluated by every
// CTA of every pass from the same counters, so they always agree.
__device__ __forceinline__ bool fast_route_valid(const SelState *st, unsigned long long k,
                                                 unsigned long long *k_in_cand) {
  const unsigned long long lt = st->count_lt, nc = st->n_cand;
  *k_in_cand = k - lt;
  return st->overflow == 0 && k >= lt && (k - lt) < nc;
}

// Find the bin where the running count first exceeds k; result broadcast to the CTA.
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
    local[i] = (b < BINS) ? __ldcg(hist + b) : 0ull;  // L2: written by other CTAs' atomics
    sum += local[i];
  }
  unsigned long long incl = sum;  // block-wide exclusive scan: shuffles + one smem hop
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

struct PassConst {
  uint32_t raw_prefix;  // PASS 1: wanted raw bits >> 24; PASS 2: wanted raw bits >> 12
  uint32_t flip;        // 0xfff for a negative bucket (key = ~bits), else 0
};

__device__ __forceinline__ PassConst make_pass_const(int pass, uint32_t key_prefix) {
  PassConst pc;
  if (pass == 1) {
    const bool neg = (key_prefix & 0x80u) == 0;  // key top bit clear <=> negative
    pc.raw_prefix = neg ? (key_prefix ^ 0xffu) : (key_prefix ^ 0x80u);
    pc.flip = neg ? 0xfffu : 0u;
  } else {
    const bool neg = (key_prefix & 0x80000u) == 0;
    pc.raw_prefix = neg ? (~key_prefix & 0xfffffu) : (key_prefix ^ 0x80000u);
    pc.flip = neg ? 0xfffu : 0u;
  }
  return pc;
}

// Only negative NaNs need a fix up front (every NaN must order last): remapped to 0x7fffffff.
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

// ---------------------------------------------------------------------------
// one radix pass; route chosen on the device
// ---------------------------------------------------------------------------
struct FastBufs {
  const SelState *st;     // nullptr: there is no fast route (small n)
  const float *cand;      // [tiles][kRegion]
  const uint32_t *cnt;    // [tiles]
  int64_t tiles;
  SelectWs ws_cand;
};

template <int PASS, int V, bool ABS>
__global__ void __launch_bounds__(QSB_THREADS)
    select_pass_kernel(const float *__restrict__ v, int64_t n, int64_t k, SelectWs ws,
                       FastBufs fb, SelState *st_rw, float *thr_out) {
  constexpr int kSmemWords = (PASS == 0) ? kBins0 * 32 : kBins1;
  __shared__ uint32_t s_hist[kSmemWords];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < kSmemWords; i += QSB_THREADS) s_hist[i] = 0;

  bool on_candidates = false;
  unsigned long long kk = (unsigned long long)k;
  if (fb.st) {
    unsigned long long kc;
    if (fast_route_valid(fb.st, kk, &kc)) {
      on_candidates = true;
      kk = kc;
      ws = fb.ws_cand;
    }
  }
  PassConst pc{0, 0};
  if constexpr (PASS >= 1) {
    unsigned long long k1;
    uint32_t b0;
    find_bin<kBins0>(ws.hist0, kk, &b0, &k1);
    uint32_t prefix = b0;
    if constexpr (PASS == 2) {
      uint32_t b1;
      find_bin<kBins1>(ws.hist1, k1, &b1, &k1);
      prefix = (b0 << 12) | b1;
    }
    pc = make_pass_const(PASS, prefix);
  }
  __syncthreads();

  if (on_candidates) {
    // candidate regions hold values with |.| already applied: a warp per region
    const int64_t warps_total = (int64_t)gridDim.x * (QSB_THREADS / 32);
    for (int64_t r = (int64_t)blockIdx.x * (QSB_THREADS / 32) + (tid >> 5); r < fb.tiles;
         r += warps_total) {
      const uint32_t c = fb.cnt[r];
      const float *region = fb.cand + r * kRegion;
      for (uint32_t i = lane; i < c; i += 32)
        count_value<PASS, false>(region[i], pc, s_hist, lane);
    }
  } else {
    constexpr int U = (V == 8) ? 2 : 4;
    constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
    const int64_t n_main = (n / V) * V;
    for (int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)tid * V; base < n_main;
         base += (int64_t)gridDim.x * kTile) {
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
          for (int j = 0; j < V; ++j) count_value<PASS, ABS>(x[u].v[j], pc, s_hist, lane);
        }
      }
    }
    if (blockIdx.x == 0) {
      const int64_t e = n_main + tid;
      if (e < n) count_value<PASS, ABS>(v[e], pc, s_hist, lane);
    }
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

  if constexpr (PASS == 2) {
    // the last CTA to finish turns the three histograms into the answer
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&st_rw->done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
      unsigned long long k1;
      uint32_t b0, b1, b2;
      find_bin<kBins0>(ws.hist0, kk, &b0, &k1);
      find_bin<kBins1>(ws.hist1, k1, &b1, &k1);
      find_bin<kBins2>(ws.hist2, k1, &b2, &k1);
      if (tid == 0) *thr_out = key_to_float((b0 << 24) | (b1 << 12) | b2);
    }
  }
}

// ---------------------------------------------------------------------------
// fast route, step 1: pivots.  kSampleSize evenly spaced values; 128 CTAs count the
// exact rank range [#less, #less-or-equal) of every sample inside the sample
// (kSampleParts threads per sample, each scanning 1/8 of the keys from shared memory);
// the samples whose range contains r_lo / r_hi are the pivots.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(QSB_THREADS)
    select_sample_kernel(const float *__restrict__ v, int64_t n, int take_abs, int r_lo,
                         int r_hi, SelState *st) {
  __shared__ __align__(16) uint32_t s_keys[kSampleSize];
  const int64_t stride = n / kSampleSize;
  for (int i = threadIdx.x; i < kSampleSize; i += QSB_THREADS) {
    const float f = v[(int64_t)i * stride + (stride >> 1)];
    s_keys[i] = float_to_key(take_abs ? fabsf(f) : f);
  }
  __syncthreads();
  constexpr int kPerCta = QSB_THREADS / kSampleParts;  // samples ranked by one CTA
  const int part = threadIdx.x % kSampleParts;
  const int i = blockIdx.x * kPerCta + threadIdx.x / kSampleParts;
  const uint32_t ki = s_keys[i];
  const uint4 *k4 = reinterpret_cast<const uint4 *>(s_keys);
  constexpr int kVecPerPart = kSampleSize / 4 / kSampleParts;
  int lt = 0, le = 0;
#pragma unroll 4
  for (int j = 0; j < kVecPerPart; ++j) {
    const uint4 q = k4[part * kVecPerPart + j];
    lt += (q.x < ki) + (q.y < ki) + (q.z < ki) + (q.w < ki);
    le += (q.x <= ki) + (q.y <= ki) + (q.z <= ki) + (q.w <= ki);
  }
#pragma unroll
  for (int o = kSampleParts >> 1; o > 0; o >>= 1) {
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
    le += __shfl_xor_sync(0xffffffffu, le, o);
  }
  if (part == 0) {
    // the r-th order statistic equals this key iff lt <= r < le (ties write the same value)
    if (r_lo >= 0 && lt <= r_lo && r_lo < le) st->lo = key_to_float(ki);
    if (r_hi < kSampleSize && lt <= r_hi && r_hi < le) st->hi = key_to_float(ki);
    if (i == 0) {
      if (r_lo < 0) st->lo = -INFINITY;
      if (r_hi >= kSampleSize) st->hi = INFINITY;
    }
  }
}

// ---------------------------------------------------------------------------
// fast route, step 2: ONE streaming pass.  Plain float compares against the pivots
// (NaNs compare false and so count as "above hi"); each CTA owns a 4096-element tile
// and a private region of kRegion candidate slots; totals by fire-and-forget RED.
// ---------------------------------------------------------------------------
template <int V, bool ABS>
__global__ void __launch_bounds__(QSB_THREADS)
    select_partition_kernel(const float *__restrict__ v, int64_t n, SelState *st,
                            float *__restrict__ cand, uint32_t *__restrict__ cnt) {
  constexpr int U = 2;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  static_assert(kTile == 4096, "region bookkeeping assumes 4096-element tiles");
  __shared__ uint32_t s_nc[QSB_THREADS / 32], s_lt[QSB_THREADS / 32];
  const float lo = st->lo, hi = st->hi;
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
  uint32_t lt = 0, nc = 0, cmask = 0;  // cmask bit (u * V + j): element is a candidate
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * QSB_THREADS * V;
    if (e < n_main) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float a = ABS ? fabsf(x[u].v[j]) : x[u].v[j];
        lt += a < lo;
        const bool c = (a >= lo) && (a <= hi);
        nc += c;
        cmask |= (uint32_t)c << (u * V + j);
      }
    }
  }
  float tail_val = 0.f;  // the last n % V elements: the owning CTA's first threads
  bool tail_c = false;
  if (t0 <= n_main && n_main < t0 + kTile) {
    const int64_t e = n_main + tid;
    if (e < n) {
      tail_val = ABS ? fabsf(v[e]) : v[e];
      lt += tail_val < lo;
      tail_c = (tail_val >= lo) && (tail_val <= hi);
      nc += tail_c;
    }
  }
  uint32_t incl = nc;  // CTA-wide exclusive scan of the candidate counts
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
    cnt[blockIdx.x] = total < (uint32_t)kRegion ? total : (uint32_t)kRegion;
    if (total > (uint32_t)kRegion) st->overflow = 1;
    if (tlt) atomicAdd(&st->count_lt, (unsigned long long)tlt);    // result unused: RED
    if (total) atomicAdd(&st->n_cand, (unsigned long long)total);  // result unused: RED
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

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
constexpr int64_t kHistBytes = (int64_t)(kBins0 + kBins1 + kBins2) * sizeof(unsigned long long);
// [hist x3 | candidate hist x3 | SelState | pad] | cnt[tiles] | candidate regions
constexpr int64_t kSelectHeaderBytes = 2 * kHistBytes + 256;
static int g_select_fast = 1;  // tuning key 4

static int64_t select_tiles(int64_t n) { return n >= kFastMinN ? (n + 4095) / 4096 : 0; }

void set_select_fast(int v) { g_select_fast = v; }

template <int PASS, int V, bool ABS>
static int launch_pass(const float *v, int64_t n, int64_t k, const SelectWs &ws,
                       const FastBufs &fb, SelState *st, float *thr_out,
                       cudaStream_t stream) {
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    QSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &o, select_pass_kernel<PASS, V, ABS>, QSB_THREADS, 0));
    occ = o > 0 ? o : 1;
  }
  constexpr int U = (V == 8) ? 2 : 4;
  constexpr int64_t kTile = (int64_t)QSB_THREADS * V * U;
  int64_t grid = (int64_t)device_props().sm_count * occ;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  select_pass_kernel<PASS, V, ABS>
      <<<(unsigned)grid, QSB_THREADS, 0, stream>>>(v, n, k, ws, fb, st, thr_out);
  QSB_LAUNCH_CHECK();
  return 0;
}

template <int V, bool ABS>
static int run_passes(const float *v, int64_t n, int64_t k, const SelectWs &ws,
                      const FastBufs &fb, SelState *st, float *thr_out,
                      cudaStream_t stream) {
  int rc;
  if ((rc = launch_pass<0, V, ABS>(v, n, k, ws, fb, st, thr_out, stream))) return rc;
  if ((rc = launch_pass<1, V, ABS>(v, n, k, ws, fb, st, thr_out, stream))) return rc;
  return launch_pass<2, V, ABS>(v, n, k, ws, fb, st, thr_out, stream);
}

}  // namespace qsb

using namespace qsb;

extern "C" int64_t qsb_kth_workspace_bytes(int64_t n) {
  const int64_t tiles = select_tiles(n);
  return kSelectHeaderBytes + 256 + tiles * 4 + 64 + tiles * kRegion * (int64_t)sizeof(float) + 32;
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
  static_assert(sizeof(SelState) <= 256, "SelState must fit the header pad");
  QSB_CUDA_TRY(cudaMemsetAsync(ws.hist0, 0, kSelectHeaderBytes, stream));

  const bool v32 = aligned_to(v, 32);
  FastBufs fb{nullptr, nullptr, nullptr, 0, wc};
  if (g_select_fast && n >= kFastMinN && v32) {
    const int64_t tiles = select_tiles(n);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(base + kSelectHeaderBytes);
    float *cand = reinterpret_cast<float *>(
        (reinterpret_cast<uintptr_t>(cnt + tiles) + 31) / 32 * 32);
    // sample ranks bracketing k: +-5 sigma of the binomial rank error, +3
    const double m = (double)kSampleSize, p = (double)k / (double)n;
    const double delta = 5.0 * sqrt(m * p * (1.0 - p)) + 3.0;
    const int r_lo = (int)floor(p * m - delta), r_hi = (int)ceil(p * m + delta);
    select_sample_kernel<<<kSampleSize / (QSB_THREADS / kSampleParts), QSB_THREADS, 0, stream>>>(
        v, n, take_abs, r_lo, r_hi, st);
    QSB_LAUNCH_CHECK();
    if (take_abs)
      select_partition_kernel<8, true><<<(unsigned)tiles, QSB_THREADS, 0, stream>>>(v, n, st, cand, cnt);
    else
      select_partition_kernel<8, false><<<(unsigned)tiles, QSB_THREADS, 0, stream>>>(v, n, st, cand, cnt);
    QSB_LAUNCH_CHECK();
    fb = FastBufs{st, cand, cnt, tiles, wc};
  }
  if (v32)
    return take_abs ? run_passes<8, true>(v, n, k, ws, fb, st, thr_out_dev, stream)
                    : run_passes<8, false>(v, n, k, ws, fb, st, thr_out_dev, stream);
  return take_abs ? run_passes<1, true>(v, n, k, ws, fb, st, thr_out_dev, stream)
                  : run_passes<1, false>(v, n, k, ws, fb, st, thr_out_dev, stream);
}
