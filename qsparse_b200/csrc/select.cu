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
                       SelectWs ws) {
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
    select_final_kernel(int64_t k, SelectWs ws, float *thr_out) {
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
                       cudaStream_t stream) {
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
      <<<(unsigned)grid, QSB_THREADS, 0, stream>>>(v, n, k, ws);
  QSB_LAUNCH_CHECK();
  return 0;
}

constexpr int64_t kSelectWsBytes =
    (int64_t)(kBins0 + kBins1 + kBins2) * sizeof(unsigned long long);

}  // namespace qsb

using namespace qsb;

extern "C" int64_t qsb_kth_workspace_bytes(int64_t n) {
  (void)n;
  return kSelectWsBytes + 256;
}

template <int V, bool ABS>
static int run_passes(const float *v, int64_t n, int64_t k, const SelectWs &ws,
                      cudaStream_t stream) {
  int rc;
  if ((rc = launch_hist<0, V, ABS>(v, n, k, ws, stream))) return rc;
  if ((rc = launch_hist<1, V, ABS>(v, n, k, ws, stream))) return rc;
  return launch_hist<2, V, ABS>(v, n, k, ws, stream);
}

extern "C" int qsb_kth_value(const float *v, int64_t n, int64_t k, int take_abs,
                             float *thr_out_dev, void *workspace,
                             int64_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0 || k < 0 || k >= n) return QSB_E_BADARG;
  if (!v || !thr_out_dev || !workspace) return QSB_E_BADARG;
  if (!aligned_to(v, 4)) return QSB_E_ALIGN;
  if (workspace_bytes < kSelectWsBytes + 256) return QSB_E_WORKSPACE;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256;
  SelectWs ws;
  ws.hist0 = reinterpret_cast<unsigned long long *>(base);
  ws.hist1 = ws.hist0 + kBins0;
  ws.hist2 = ws.hist1 + kBins1;
  QSB_CUDA_TRY(cudaMemsetAsync(ws.hist0, 0, kSelectWsBytes, stream));
  int rc;
  if (aligned_to(v, 32))
    rc = take_abs ? run_passes<8, true>(v, n, k, ws, stream)
                  : run_passes<8, false>(v, n, k, ws, stream);
  else
    rc = take_abs ? run_passes<1, true>(v, n, k, ws, stream)
                  : run_passes<1, false>(v, n, k, ws, stream);
  if (rc) return rc;
  select_final_kernel<<<1, QSB_THREADS, 0, stream>>>(k, ws, thr_out_dev);
  QSB_LAUNCH_CHECK();
  return 0;
}
