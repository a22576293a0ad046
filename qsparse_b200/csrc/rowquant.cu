// K8  row-resident fused scale estimation + fake-quantization for tensors whose
// quantization channel is the leading axis (the `channelwise=0` weight case:
// `quantize(nn.Linear(4096, 4096), bits=4, channelwise=0, callback=AdaptiveQuantizer())`,
// BASELINE config 3).
//
// The reference runs, per training access of the weight,
//     optimize: min / max (or abs-max) per row  -> EMA into the layer's parameter
//               qsparse/quantize.py:327-349 (Decimal/Scaler), :393-430 (Adaptive)
//     forward : fake-quantize every row with its updated parameter
//               qsparse/quantize.py:44-63, :100-117, :148-181
// i.e. read (reduce) + tiny kernel + read + write = 12 B/elem in three launches with a
// grid-wide dependency in the middle.  A channel here is ONE contiguous row, so nothing
// crosses CTAs: a CTA (rows of 1 Ki .. 16 Ki elements) or a warp (shorter rows) loads its
// row ONCE into registers, reduces it, updates the row's parameter, and quantizes the row
// from the registers: 8 B/elem, one launch, no dependency between CTAs.
//
// Two kernels:
//   row_quant_kernel      the row lives in registers: a CTA per row (1 Ki .. 16 Ki elements, one
//       barrier: every thread merges the warps' partials and derives the new parameter
//       itself) or a warp per row (shorter rows).  Measured on [4096, 4096] (134 MB of
//       traffic): 25.6 us scaler / 28.7 us line — the same as the plain fake-quant kernel
//       alone on that tensor (28.6 us), against 50.2 us for reduce -> EMA -> quantize.
//   row_quant_tma_kernel  alternative for the CTA-per-row range (tuning key 9): persistent CTAs,
//       ONE thread streams the CTA's rows into a ring of shared-memory stages with
//       cp.async.bulk (TMA engine, completion on an mbarrier) 2-4 rows ahead, so the bytes
//       in flight are bounded by shared memory, not registers.  Measured 30.7 us for every
//       (CTAs per SM, stages) in {2..8} x {2..4}: the row's dependent chain (wait -> LDS ->
//       reduce -> barrier -> quantize -> store) bounds a persistent CTA, and one CTA per row
//       with 5 CTAs per SM overlaps those chains better.  Kept selectable, not the default.
//
// Arithmetic is the same device code as the unfused kernels (fq_ops.cuh, param_math.cuh),
// so the results are bit-identical to reduce_stats -> *_ema -> fq_*_fwd.
#include <math.h>

#include "fq_ops.cuh"
#include "map_kernel.cuh"
#include "param_math.cuh"

namespace qsb {

constexpr int kRowDecimal = 0, kRowScaler = 1, kRowLine = 2;

struct RowConsts {
  float limit;     // 2^(bits-1)                      (decimal / scaler)
  float n_levels;  // 2^bits                          (line)
  float q_max;     // 2^bits - 1                      (line)
  int64_t t;       // scale EMA: calls so far (0-based); lines EMA: this call's number (1-based)
  const long long *t_dev;  // optional device step counter (CUDA graphs): the index is *t_dev + t
};

struct RowStat {
  uint32_t amax;  // bits of max |x|; NaN patterns order above inf (as in reduce.cu)
  float mn, mx;
  int nan;
};

template <int KIND>
__device__ __forceinline__ void stat_add(RowStat &s, float v) {
  if constexpr (KIND == kRowLine) {
    s.mn = fminf(s.mn, v);
    s.mx = fmaxf(s.mx, v);
    s.nan |= (v != v);
  } else {
    const uint32_t b = __float_as_uint(v) & 0x7fffffffu;
    s.amax = b > s.amax ? b : s.amax;
  }
}

template <int KIND>
__device__ __forceinline__ void stat_merge(RowStat &s, const RowStat &o) {
  if constexpr (KIND == kRowLine) {
    s.mn = fminf(s.mn, o.mn);
    s.mx = fmaxf(s.mx, o.mx);
    s.nan |= o.nan;
  } else {
    s.amax = o.amax > s.amax ? o.amax : s.amax;
  }
}

template <int KIND>
__device__ __forceinline__ RowStat stat_warp(RowStat s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    RowStat t;
    if constexpr (KIND == kRowLine) {
      t.mn = __shfl_xor_sync(0xffffffffu, s.mn, o);
      t.mx = __shfl_xor_sync(0xffffffffu, s.mx, o);
      t.nan = __shfl_xor_sync(0xffffffffu, s.nan, o);
    } else {
      t.amax = __shfl_xor_sync(0xffffffffu, s.amax, o);
    }
    stat_merge<KIND>(s, t);
  }
  return s;
}

// statistics + the row's old parameter -> its updated parameter (x = scale or lo, y = hi)
template <int KIND>
__device__ __forceinline__ float2 ema_param(const RowStat &s, float2 w, const RowConsts &k) {
  float2 r;
  if constexpr (KIND == kRowLine) {
    const float mn = s.nan ? __uint_as_float(0x7fc00000u) : s.mn;
    const float mx = s.nan ? __uint_as_float(0x7fc00000u) : s.mx;
    const int64_t t = k.t_dev ? (int64_t)__ldg(k.t_dev) + k.t : k.t;
    const float tm1 = (float)(t - 1), tf = (float)t;
    r.x = lines_ema_step(w.x, mn, tm1, tf);
    r.y = lines_ema_step(w.y, mx, tm1, tf);
  } else {
    const int64_t t = k.t_dev ? (int64_t)__ldg(k.t_dev) + k.t : k.t;
    r.x = scale_ema_step(w.x, __uint_as_float(s.amax), k.limit, t);
    r.y = 0.f;
  }
  return r;
}

template <int KIND>
__device__ __forceinline__ float2 load_param(const float *prm) {
  if constexpr (KIND == kRowLine) return *reinterpret_cast<const float2 *>(prm);
  return make_float2(prm[0], 0.f);
}

template <int KIND>
__device__ __forceinline__ void store_param(float *prm, float2 v) {
  if constexpr (KIND == kRowLine) *reinterpret_cast<float2 *>(prm) = v;
  else prm[0] = v.x;
}

template <int KIND, bool FZP>
struct RowOp;
template <bool FZP>
struct RowOp<kRowDecimal, FZP> {
  using Op = Pow2Op<QSB_MASK_NONE>;
  __device__ static Op make(const RowConsts &) { return Op{nullptr, 0, 1.f, 1.f, nullptr}; }
  __device__ static typename Op::P derive(const Op &, float2 v) { return Op::derive(v.y); }
};
template <bool FZP>
struct RowOp<kRowScaler, FZP> {
  using Op = ScalerOp<QSB_MASK_NONE>;
  __device__ static Op make(const RowConsts &) { return Op{nullptr, 0, 1.f, nullptr}; }
  __device__ static typename Op::P derive(const Op &, float2 v) { return Op::derive(v.x); }
};
template <bool FZP>
struct RowOp<kRowLine, FZP> {
  using Op = LineOp<QSB_MASK_NONE, FZP>;
  __device__ static Op make(const RowConsts &k) {
    return Op{nullptr, 0, 0.f, 0.f, k.n_levels, k.q_max, nullptr};
  }
  __device__ static typename Op::P derive(const Op &op, float2 v) { return op.derive(v.x, v.y); }
};

// G threads own one row: G = QSB_THREADS (a CTA per row) or 32 (a warp per row, 8 rows per CTA).
// MASKED: an element prune mask is applied first (the weight chain quantize(prune(layer)),
// ref qsparse/imitation.py:61-71): statistics and quantization both see x * mask — a real multiply by
// 0.0 / 1.0 like the reference's — in the same single read of the row (+1 B/elem for the mask).
// 6 CTAs / SM (<= 42 registers) for the unmasked kernels with <= 2 vectors per thread: the line quantizer otherwise
// takes 54 registers = 4 CTAs / SM, whose load / reduce / quantize phases then overlap too little.  A/B on one box
// (benchmarks/time_c3.py, 20 launches per graph on [4096,4096]): line 26.8 -> 21.9 us, scaler 21.0 -> 20.1 us,
// decimal 22.4 -> 21.0 us; ncu, cold and isolated: line 24.5 -> 21.8 us.  (Single event-bracketed launches vary by
// +-1.5 us from box to box, which hid this in bench.py --config 3.)
#ifndef QSB_K8_MIN_CTAS
#define QSB_K8_MIN_CTAS 6
#endif
template <int KIND, bool FZP, int U, int G, bool MASKED>
__global__ void __launch_bounds__(QSB_THREADS, (U <= 2 && !MASKED) ? QSB_K8_MIN_CTAS : 1)
    row_quant_kernel(const float *__restrict__ x, float *__restrict__ y, float *__restrict__ param,
                     float *__restrict__ decimal_out, const uint8_t *__restrict__ mask, int64_t rows, int inner,
                     RowConsts k) {
  constexpr int V = 8;
  constexpr int kRowsPerCta = QSB_THREADS / G;
  constexpr int kWsz = (KIND == kRowLine) ? 2 : 1;
  const int tid = threadIdx.x, g = tid % G, lane = tid & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowsPerCta + tid / G;
  const bool live = row < rows;  // uniform per warp (G is a multiple of 32)
  const float *xr = x + row * inner;
  float *yr = y + row * inner;
  const int nvec = inner / V;

  VecF<V> a[U];
  float2 w_old = make_float2(0.f, 0.f);
  if (live) {
    VecB<V> mb[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (u * G + g < nvec) {
        a[u] = ld_vec<V, Hint::KEEP>(xr + (int64_t)(u * G + g) * V);
        if constexpr (MASKED) mb[u] = ld_bytes<V>(mask + row * inner + (int64_t)(u * G + g) * V);
      }
    w_old = load_param<KIND>(param + row * kWsz);  // in flight together with the row
    if constexpr (MASKED) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u * G + g < nvec) {
#pragma unroll
          for (int j = 0; j < V; ++j) a[u].v[j] = __fmul_rn(a[u].v[j], mb[u].b[j] ? 1.0f : 0.0f);
        }
    }
  }
  RowStat s{0u, INFINITY, -INFINITY, 0};
  if (live) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (u * G + g < nvec) {
#pragma unroll
        for (int j = 0; j < V; ++j) stat_add<KIND>(s, a[u].v[j]);
      }
  }
  s = stat_warp<KIND>(s);
  if constexpr (G != 32) {  // one barrier: every thread merges the warps' partials itself
    __shared__ RowStat s_part[QSB_THREADS / 32];
    if (lane == 0) s_part[tid >> 5] = s;
    __syncthreads();
    s = s_part[0];
#pragma unroll
    for (int w = 1; w < QSB_THREADS / 32; ++w) stat_merge<KIND>(s, s_part[w]);
  }
  // every thread derives the new parameter (a few flops, no second barrier); one writes it
  float2 pv = ema_param<KIND>(s, w_old, k);
  if (live && g == 0) store_param<KIND>(param + row * kWsz, pv);
  if constexpr (KIND == kRowDecimal) {
    // the fp64 log2 of scale -> decimal is worth computing once per row
    if constexpr (G == 32) {
      if (lane == 0) pv.y = scale_to_decimal(pv.x);
      pv.y = __shfl_sync(0xffffffffu, pv.y, 0);
    } else {
      __shared__ float s_dec;
      if (tid == 0) s_dec = scale_to_decimal(pv.x);
      __syncthreads();
      pv.y = s_dec;
    }
    if (live && g == 0 && decimal_out) decimal_out[row] = pv.y;
  }
  if (!live) return;

  using RO = RowOp<KIND, FZP>;
  const typename RO::Op op = RO::make(k);
  const typename RO::Op::P p = RO::derive(op, pv);
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u * G + g < nvec) {
      VecF<V> o0, o1;
      VecB<V> mb, ob;
      apply_vec<typename RO::Op, V>(op, a[u], a[u], mb, false, p, o0, o1, ob);
      st_vec<V, Hint::KEEP>(yr + (int64_t)(u * G + g) * V, o0);
    }
}

// ---------------------------------------------------------------------------
// TMA-pipelined persistent variant (a CTA per row, rows staged in shared memory)
// ---------------------------------------------------------------------------
constexpr int kTmaMaxStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// one elected thread: expect `bytes` on the barrier, then bulk-copy global -> shared
__device__ __forceinline__ void tma_load_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <int KIND, bool FZP, int U>
__global__ void __launch_bounds__(QSB_THREADS)
    row_quant_tma_kernel(const float *__restrict__ x, float *__restrict__ y, float *__restrict__ param,
                         float *__restrict__ decimal_out, int64_t rows, int inner, RowConsts k,
                         int stages) {
  constexpr int V = 8;
  constexpr int kWsz = (KIND == kRowLine) ? 2 : 1;
  extern __shared__ __align__(128) unsigned char s_rows[];  // [stages][inner] floats
  __shared__ uint64_t s_full[kTmaMaxStages];
  __shared__ RowStat s_part[2][QSB_THREADS / 32];
  __shared__ float s_dec[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int nvec = inner / V;
  const uint32_t row_bytes = (uint32_t)inner * 4u;
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t n_mine = first < rows ? (rows - first + step - 1) / step : 0;

  if (tid == 0) {
    for (int st = 0; st < stages; ++st) mbar_init(&s_full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int st = 0; st < stages && st < n_mine; ++st)
      tma_load_row(s_rows + (size_t)st * row_bytes, x + (first + st * step) * inner, row_bytes, &s_full[st]);
  }

  using RO = RowOp<KIND, FZP>;
  const typename RO::Op op = RO::make(k);
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t i = 0; i < n_mine; ++i) {
    const int64_t row = first + i * step;
    const float2 w_old = load_param<KIND>(param + row * kWsz);  // in flight while we wait for the row
    mbar_wait(&s_full[stage], parity);
    const float4 *srow = reinterpret_cast<const float4 *>(s_rows + (size_t)stage * row_bytes);
    VecF<V> a[U];
    RowStat s{0u, INFINITY, -INFINITY, 0};
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (u * QSB_THREADS + tid < nvec) {
        const float4 q0 = srow[(u * QSB_THREADS + tid) * 2], q1 = srow[(u * QSB_THREADS + tid) * 2 + 1];
        a[u].v[0] = q0.x, a[u].v[1] = q0.y, a[u].v[2] = q0.z, a[u].v[3] = q0.w;
        a[u].v[4] = q1.x, a[u].v[5] = q1.y, a[u].v[6] = q1.z, a[u].v[7] = q1.w;
#pragma unroll
        for (int j = 0; j < V; ++j) stat_add<KIND>(s, a[u].v[j]);
      }
    s = stat_warp<KIND>(s);
    RowStat *part = s_part[i & 1];  // alternate: a warp may run one iteration ahead of another
    if (lane == 0) part[tid >> 5] = s;
    __syncthreads();  // every thread has its part of the row in registers: the stage is free
    if (tid == 0 && i + stages < n_mine)
      tma_load_row(s_rows + (size_t)stage * row_bytes, x + (row + stages * step) * inner, row_bytes,
                   &s_full[stage]);
    s = part[0];
#pragma unroll
    for (int w = 1; w < QSB_THREADS / 32; ++w) stat_merge<KIND>(s, part[w]);
    float2 pv = ema_param<KIND>(s, w_old, k);
    if (tid == 0) store_param<KIND>(param + row * kWsz, pv);
    if constexpr (KIND == kRowDecimal) {
      if (tid == 0) s_dec[i & 1] = scale_to_decimal(pv.x);  // fp64 log2: once per row
      __syncthreads();
      pv.y = s_dec[i & 1];
      if (tid == 0 && decimal_out) decimal_out[row] = pv.y;
    }
    const typename RO::Op::P p = RO::derive(op, pv);
    float *yr = y + row * inner;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (u * QSB_THREADS + tid < nvec) {
        VecF<V> o0, o1;
        VecB<V> mb, ob;
        apply_vec<typename RO::Op, V>(op, a[u], a[u], mb, false, p, o0, o1, ob);
        st_vec<V, Hint::KEEP>(yr + (int64_t)(u * QSB_THREADS + tid) * V, o0);
      }
    if (++stage == stages) {
      stage = 0;
      parity ^= 1u;
    }
  }
}

static int g_row_tma = 0;       // tuning key 9: 1 = TMA-pipelined kernel for long rows, 0 = register kernel (default: measured faster)
static int g_row_ctas_per_sm = 0;  // tuning key 10: persistent CTAs per SM (0 = as many as fit, at most 3)
static int g_row_stages = 0;       // tuning key 11: ring stages (0 = auto)
void set_row_tma(int v) { g_row_tma = v != 0; }
void set_row_ctas_per_sm(int v) { g_row_ctas_per_sm = v; }
void set_row_stages(int v) { g_row_stages = v; }

template <int KIND, bool FZP, int U>
static int launch_rows_tma(const float *x, float *y, float *param, float *decimal_out, int64_t rows,
                           int64_t inner, const RowConsts &k, cudaStream_t stream) {
  const int64_t row_bytes = inner * 4;
  int stages = (int)(96 * 1024 / row_bytes);  // ~64-96 KB of rows in flight per CTA
  if (g_row_stages > 0) stages = g_row_stages;
  if (stages > kTmaMaxStages) stages = kTmaMaxStages;
  if (stages < 2) stages = 2;
  const size_t smem = (size_t)stages * row_bytes;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    QSB_CUDA_TRY(cudaFuncSetAttribute(row_quant_tma_kernel<KIND, FZP, U>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bytes = smem;
  }
  int occ = 0;
  QSB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, row_quant_tma_kernel<KIND, FZP, U>,
                                                             QSB_THREADS, smem));
  if (occ < 1) return QSB_E_UNSUPPORTED;
  int per_sm = g_row_ctas_per_sm > 0 ? g_row_ctas_per_sm : 3;  // occupancy caps it below
  if (per_sm > occ) per_sm = occ;
  int64_t grid = (int64_t)device_props().sm_count * per_sm;
  if (grid > rows) grid = rows;
  row_quant_tma_kernel<KIND, FZP, U><<<(unsigned)grid, QSB_THREADS, smem, stream>>>(
      x, y, param, decimal_out, rows, (int)inner, k, stages);
  QSB_LAUNCH_CHECK();
  return 0;
}

template <int KIND, bool FZP, int U, int G>
static int launch_rows(const float *x, float *y, float *param, float *decimal_out, const uint8_t *mask,
                       int64_t rows, int64_t inner, const RowConsts &k, cudaStream_t stream) {
  constexpr int kRowsPerCta = QSB_THREADS / G;
  const int64_t grid = (rows + kRowsPerCta - 1) / kRowsPerCta;
  if (grid > 0x7fffffffll) return QSB_E_UNSUPPORTED;
  if (mask)
    row_quant_kernel<KIND, FZP, U, G, true>
        <<<(unsigned)grid, QSB_THREADS, 0, stream>>>(x, y, param, decimal_out, mask, rows, (int)inner, k);
  else
    row_quant_kernel<KIND, FZP, U, G, false>
        <<<(unsigned)grid, QSB_THREADS, 0, stream>>>(x, y, param, decimal_out, nullptr, rows, (int)inner, k);
  QSB_LAUNCH_CHECK();
  return 0;
}

template <int KIND, bool FZP>
static int dispatch_rows(const float *x, float *y, float *param, float *decimal_out, const uint8_t *mask,
                         int64_t rows, int64_t inner, const RowConsts &k, cudaStream_t stream) {
  // a warp per row up to 1 Ki elements, a CTA per row up to 16 Ki
  if (inner <= 256) return launch_rows<KIND, FZP, 1, 32>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  if (inner <= 512) return launch_rows<KIND, FZP, 2, 32>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  if (inner <= 1024) return launch_rows<KIND, FZP, 4, 32>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  if (g_row_tma && !mask) {
    if (inner <= 2048) return launch_rows_tma<KIND, FZP, 1>(x, y, param, decimal_out, rows, inner, k, stream);
    if (inner <= 4096) return launch_rows_tma<KIND, FZP, 2>(x, y, param, decimal_out, rows, inner, k, stream);
    if (inner <= 8192) return launch_rows_tma<KIND, FZP, 4>(x, y, param, decimal_out, rows, inner, k, stream);
    if (inner <= 16384) return launch_rows_tma<KIND, FZP, 8>(x, y, param, decimal_out, rows, inner, k, stream);
  }
  if (inner <= 2048) return launch_rows<KIND, FZP, 1, QSB_THREADS>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  if (inner <= 4096) return launch_rows<KIND, FZP, 2, QSB_THREADS>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  if (inner <= 8192) return launch_rows<KIND, FZP, 4, QSB_THREADS>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  if (inner <= 16384) return launch_rows<KIND, FZP, 8, QSB_THREADS>(x, y, param, decimal_out, mask, rows, inner, k, stream);
  return QSB_E_UNSUPPORTED;
}

}  // namespace qsb

using namespace qsb;

static int row_quant_fused_impl(const float *x, float *y, float *param, float *decimal_out,
                                const uint8_t *mask, int kind, int bits, int float_zero_point,
                                int64_t rows, int64_t inner, int64_t t, void *stream_,
                                const int64_t *t_dev = nullptr) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || inner < 0 || bits < 0 || bits > 62) return QSB_E_BADARG;
  if (kind < kRowDecimal || kind > kRowLine) return QSB_E_BADARG;
  if (!t_dev && (kind == kRowLine ? t < 1 : t < 0)) return QSB_E_BADARG;
  if (rows == 0 || inner == 0) return 0;
  if (!x || !y || !param) return QSB_E_BADARG;
  if (kind == kRowLine && !aligned_to(param, 8)) return QSB_E_ALIGN;
  // rows must be whole 256-bit vectors: anything else takes the two-kernel path
  if (inner % 8 != 0 || inner > 16384 || !aligned_to(x, 32) || !aligned_to(y, 32))
    return QSB_E_UNSUPPORTED;
  RowConsts k;
  k.limit = (float)pow(2.0, (double)bits - 1.0);
  const double N = ldexp(1.0, bits);
  k.n_levels = (float)N;
  k.q_max = (float)(N - 1.0);
  k.t = t;
  k.t_dev = reinterpret_cast<const long long *>(t_dev);
  switch (kind) {
    case kRowDecimal:
      return dispatch_rows<kRowDecimal, true>(x, y, param, decimal_out, mask, rows, inner, k, stream);
    case kRowScaler:
      return dispatch_rows<kRowScaler, true>(x, y, param, nullptr, mask, rows, inner, k, stream);
    default:
      return float_zero_point
                 ? dispatch_rows<kRowLine, true>(x, y, param, nullptr, mask, rows, inner, k, stream)
                 : dispatch_rows<kRowLine, false>(x, y, param, nullptr, mask, rows, inner, k, stream);
  }
}

extern "C" int qsb_row_quant_fused(const float *x, float *y, float *param, float *decimal_out,
                                   int kind, int bits, int float_zero_point, int64_t rows,
                                   int64_t inner, int64_t t, void *stream) {
  return row_quant_fused_impl(x, y, param, decimal_out, nullptr, kind, bits, float_zero_point, rows, inner, t,
                              stream);
}

// the weight chain quantize(prune(layer)) with a frozen mask: y = Q(x * mask), parameters estimated
// on x * mask, one read of x (4) + mask (1) + one write (4) = 9 B/elem instead of mask-apply (9) +
// estimate/quantize (8).  mask_dev: uint8 [rows * inner], 8-byte aligned.
extern "C" int qsb_row_quant_fused_masked(const float *x, float *y, float *param, float *decimal_out,
                                          const uint8_t *mask_dev, int kind, int bits,
                                          int float_zero_point, int64_t rows, int64_t inner,
                                          int64_t t, void *stream) {
  if (!mask_dev) return QSB_E_BADARG;
  if (!aligned_to(mask_dev, 8)) return QSB_E_UNSUPPORTED;
  return row_quant_fused_impl(x, y, param, decimal_out, mask_dev, kind, bits, float_zero_point, rows, inner,
                              t, stream);
}

// CUDA-graph form of the two entry points above: the EMA index is *t_dev + t_offset, read by the kernel
// (launch arguments of a captured graph are frozen); mask_dev may be NULL.  The caller advances the counter.
extern "C" int qsb_row_quant_fused_at(const float *x, float *y, float *param, float *decimal_out,
                                      const uint8_t *mask_dev, int kind, int bits, int float_zero_point,
                                      int64_t rows, int64_t inner, const int64_t *t_dev, int64_t t_offset,
                                      void *stream) {
  if (!t_dev) return QSB_E_BADARG;
  if (mask_dev && !aligned_to(mask_dev, 8)) return QSB_E_UNSUPPORTED;
  return row_quant_fused_impl(x, y, param, decimal_out, mask_dev, kind, bits, float_zero_point, rows, inner,
                              t_offset, stream, t_dev);
}
