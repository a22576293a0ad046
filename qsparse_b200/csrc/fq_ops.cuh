// Fake-quantization Op functors (K1a / K1b / K1c) shared by the streaming map kernels
// (elementwise.cu) and the row-resident fused estimate+quantize kernel (rowquant.cu).
//
// Arithmetic is written with explicit round-to-nearest intrinsics so that no
// FMA contraction can happen (the reference rounds after every ATen op).
#pragma once
#include <math.h>

#include "qsb_common.cuh"

namespace qsb {

// mask multiply exactly like `x * mask.float()`
__device__ __forceinline__ float mask_mul(float x, bool keep) {
  return __fmul_rn(x, keep ? 1.0f : 0.0f);
}

// min(max(v, lo), hi) with NaN-PROPAGATING FMNMX (one instruction each instead of compare +
// select): the same value as clamp_torch for finite bounds — a NaN v stays NaN (canonical
// payload), and the one sign-of-zero difference (v = -0, lo = +0 gives +0) cannot reach the
// output of the line quantizer (q * step + lo rounds both to the same value).
__device__ __forceinline__ float clamp_fmnmx_nan(float v, float lo, float hi) {
  float t, r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(t) : "f"(v), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(t), "f"(hi));
  return r;
}

// torch.clamp(v, min=Tensor, max=Tensor): a NaN bound makes the result NaN.
__device__ __forceinline__ float clamp_torch_tensor(float v, float lo,
                                                    float hi) {
  if (lo != lo || hi != hi) return __uint_as_float(0x7fc00000u);
  return clamp_torch(v, lo, hi);
}

// ---------------------------------------------------------------------------
// K1a  pow2 fake-quant   ref qsparse/quantize.py:44-63
// ---------------------------------------------------------------------------
template <int MASK>
struct Pow2Op {
  static constexpr bool kIn1 = false, kInB = (MASK == QSB_MASK_ELEMENT),
                        kOut0 = true, kOut1 = false, kOutB = false,
                        kCanSkip = (MASK == QSB_MASK_CHANNEL),
                        kHasFast = false;
  const float *dec;  // device decimals or nullptr
  int dec_stride;    // 0: one decimal for the tensor, 1: per channel
  float toi_host, tof_host;
  const uint8_t *cmask;
  struct P {
    float toi, tof, m;
  };
  // constants of one decimal (also used by the row-resident fused kernel, rowquant.cu)
  __device__ __forceinline__ static P derive(float d) {
    P p;
    p.toi = pow2f_exact(d);
    p.tof = pow2f_exact(-d);
    p.m = 1.0f;
    return p;
  }
  __device__ __forceinline__ P params(int32_t c) const {
    P p;
    if (dec) {
      p = derive(__ldg(dec + (int64_t)c * dec_stride));
    } else {
      p.toi = toi_host;
      p.tof = tof_host;
    }
    p.m = 1.0f;
    if constexpr (MASK == QSB_MASK_CHANNEL) p.m = __ldg(cmask + c) ? 1.0f : 0.0f;
    return p;
  }
  __device__ __forceinline__ bool skip(const P &p) const { return p.m == 0.0f; }
  __device__ __forceinline__ void apply(float a, float, uint8_t mb, const P &p,
                                        float &o0, float &, uint8_t &) const {
    float t = a;
    if constexpr (MASK == QSB_MASK_CHANNEL) t = __fmul_rn(a, p.m);
    if constexpr (MASK == QSB_MASK_ELEMENT) t = mask_mul(a, mb != 0);
    // (x * toi).int() : truncation toward zero (cvt.rzi.s32.f32);
    // q.float() * tof : the int round trip also erases the sign of zero.
    const int q = __float2int_rz(__fmul_rn(t, p.toi));
    o0 = __fmul_rn(__int2float_rn(q), p.tof);
  }
};

// ---------------------------------------------------------------------------
// K1b  float-scale fake-quant   ref qsparse/quantize.py:100-117
// ---------------------------------------------------------------------------
template <int MASK>
struct ScalerOp {
  static constexpr bool kIn1 = false, kInB = (MASK == QSB_MASK_ELEMENT),
                        kOut0 = true, kOut1 = false, kOutB = false,
                        kCanSkip = (MASK == QSB_MASK_CHANNEL),
                        kHasFast = true;
  const float *scale;
  int scale_stride;
  float scale_host;
  const uint8_t *cmask;
  struct P {
    float s, r, m;
    bool ok;
  };
  __device__ __forceinline__ static P derive(float s) {
    P p;
    p.s = s;
    p.r = __frcp_rn(p.s);
    p.ok = fastdiv_divisor_ok(p.s);
    p.m = 1.0f;
    return p;
  }
  __device__ __forceinline__ P params(int32_t c) const {
    P p = derive(scale ? __ldg(scale + (int64_t)c * scale_stride) : scale_host);
    if constexpr (MASK == QSB_MASK_CHANNEL) p.m = __ldg(cmask + c) ? 1.0f : 0.0f;
    return p;
  }
  __device__ __forceinline__ bool skip(const P &p) const { return p.m == 0.0f; }
  __device__ __forceinline__ void apply(float a, float, uint8_t mb, const P &p,
                                        float &o0, float &, uint8_t &) const {
    float t = a;
    if constexpr (MASK == QSB_MASK_CHANNEL) t = __fmul_rn(a, p.m);
    if constexpr (MASK == QSB_MASK_ELEMENT) t = mask_mul(a, mb != 0);
    // (x / s).round().int() : IEEE divide, then round-half-even + cast in one
    // cvt.rni.s32.f32 (== trunc(rint(v)) including the saturating edge cases)
    const int q = __float2int_rn(__fdiv_rn(t, p.s));
    o0 = __fmul_rn(__int2float_rn(q), p.s);
  }
  // fast path: the scale is in the reciprocal-division range and every |x| of the
  // vector is below 2^64 (NaNs pass the max and end as q = 0 on both paths)
  template <int V>
  __device__ __forceinline__ bool fast(const P &p, const float *a) const {
    float m = fabsf(a[0]);
#pragma unroll
    for (int j = 1; j < V; ++j) m = fmaxf(m, fabsf(a[j]));
    return p.ok && m < 1.8446744e19f;
  }
  __device__ __forceinline__ void apply_fast(float a, float, uint8_t mb, const P &p,
                                             float &o0, float &, uint8_t &) const {
    float t = a;
    if constexpr (MASK == QSB_MASK_CHANNEL) t = __fmul_rn(a, p.m);
    if constexpr (MASK == QSB_MASK_ELEMENT) t = mask_mul(a, mb != 0);
    const int q = __float2int_rn(div_rn_by_unchecked(t, p.s, p.r));
    o0 = __fmul_rn(__int2float_rn(q), p.s);
  }
};

// ---------------------------------------------------------------------------
// K1c  asymmetric ("line") fake-quant   ref qsparse/quantize.py:148-181
// ---------------------------------------------------------------------------
template <int MASK, bool FZP>
struct LineOp {
  static constexpr bool kIn1 = false, kInB = (MASK == QSB_MASK_ELEMENT),
                        kOut0 = true, kOut1 = false, kOutB = false,
                        kCanSkip = false,  // output depends on lo even for x = 0
                        kHasFast = true;
  const float *lines;  // [n][2] or nullptr
  int lines_stride;    // 0 or 1 (in rows)
  float lo_host, hi_host;
  float n_levels;  // float(2^bits)
  float q_max;     // float(2^bits - 1)
  const uint8_t *cmask;
  struct P {
    float lo, hi, step, rstep, qstart, m;
    // fast: step is in the fast-division range, the bounds are finite, every
    // quotient is below 2^21 in magnitude -> unchecked reciprocal division and
    // rint via the 1.5*2^23 trick; otherwise the generic IEEE path.
    bool fast;
  };
  __device__ __forceinline__ P derive(float lo, float hi) const {
    P p;
    p.lo = lo;
    p.hi = hi;
    // step = (end - start) / N ; step[step == 0] = 0.0001   (:159-160)
    p.step = __fdiv_rn(__fsub_rn(p.hi, p.lo), n_levels);
    if (p.step == 0.0f) p.step = 0.0001f;
    p.rstep = __frcp_rn(p.step);
    p.qstart = FZP ? 0.0f : rintf(__fdiv_rn(p.lo, p.step));  // (:163)
    // after the clamp |x| <= max(|lo|, |hi|) (and x - lo in [0, hi - lo]): when that
    // bound / step stays below 2^21, every quotient is in the range where the
    // reciprocal division needs no guard and (q + 1.5*2^23) - 1.5*2^23 == rint(q)
    // (round half to even) — two FADDs instead of an XU-pipe FRND.  NaN / inf
    // bounds fail the comparison and take the generic path.
    {
      const float span = FZP ? fabsf(__fsub_rn(p.hi, p.lo)) : fmaxf(fabsf(p.lo), fabsf(p.hi));
      p.fast = fastdiv_divisor_ok(p.step) && (__fmul_rn(span, fabsf(p.rstep)) < 2097152.0f) &&
               (fabsf(p.lo) < 1.8446744e19f) && (fabsf(p.hi) < 1.8446744e19f);
    }
    p.m = 1.0f;
    return p;
  }
  __device__ __forceinline__ P params(int32_t c) const {
    P p;
    if (lines) {
      const float2 l =
          __ldg(reinterpret_cast<const float2 *>(lines) + (int64_t)c * lines_stride);
      p = derive(l.x, l.y);
    } else {
      p = derive(lo_host, hi_host);
    }
    p.m = 1.0f;
    if constexpr (MASK == QSB_MASK_CHANNEL) p.m = __ldg(cmask + c) ? 1.0f : 0.0f;
    return p;
  }
  __device__ __forceinline__ bool skip(const P &) const { return false; }
  __device__ __forceinline__ void apply(float a, float, uint8_t mb, const P &p,
                                        float &o0, float &, uint8_t &) const {
    float t = a;
    if constexpr (MASK == QSB_MASK_CHANNEL) t = __fmul_rn(a, p.m);
    if constexpr (MASK == QSB_MASK_ELEMENT) t = mask_mul(a, mb != 0);
    const float xc = clamp_torch_tensor(t, p.lo, p.hi);  // (:158)
    if constexpr (FZP) {
      float q = __fdiv_rn(__fsub_rn(xc, p.lo), p.step);      // (:176-177)
      q = clamp_torch(rintf(q), 0.0f, q_max);                // (:178)
      o0 = __fadd_rn(__fmul_rn(q, p.step), p.lo);            // (:179-180)
    } else {
      float q = rintf(__fdiv_rn(xc, p.step));                         // (:162)
      q = clamp_torch(__fsub_rn(q, p.qstart), 0.0f, q_max);           // (:164)
      o0 = __fmul_rn(__fadd_rn(q, p.qstart), p.step);                 // (:165)
    }
  }
  template <int V>
  __device__ __forceinline__ bool fast(const P &p, const float *) const {
    return p.fast;
  }
  __device__ __forceinline__ void apply_fast(float a, float, uint8_t mb, const P &p,
                                             float &o0, float &, uint8_t &) const {
    float t = a;
    if constexpr (MASK == QSB_MASK_CHANNEL) t = __fmul_rn(a, p.m);
    if constexpr (MASK == QSB_MASK_ELEMENT) t = mask_mul(a, mb != 0);
    const float kMagic = 12582912.0f;                  // 1.5 * 2^23
    const float xc = clamp_fmnmx_nan(t, p.lo, p.hi);   // (:158) bounds are finite here
    if constexpr (FZP) {
      // xc - lo >= +0: the zero-sign fix-up of the exact division is dead here
      const float d = __fsub_rn(xc, p.lo);
      const float q0 = __fmul_rn(d, p.rstep);                                 // (:176-177)
      float e = __fmaf_rn(-p.step, q0, d);
      float q = __fmaf_rn(e, p.rstep, q0);
      e = __fmaf_rn(-p.step, q, d);
      q = __fmaf_rn(e, p.rstep, q);
      q = __fsub_rn(__fadd_rn(q, kMagic), kMagic);                            // round_ (:178)
      q = clamp_fmnmx_nan(q, 0.0f, q_max);
      o0 = __fadd_rn(__fmul_rn(q, p.step), p.lo);                             // (:179-180)
    } else {
      float q = div_rn_by_unchecked(xc, p.step, p.rstep);
      q = __fsub_rn(__fadd_rn(q, kMagic), kMagic);                            // (:162)
      q = clamp_fmnmx_nan(__fsub_rn(q, p.qstart), 0.0f, q_max);               // (:164)
      o0 = __fmul_rn(__fadd_rn(q, p.qstart), p.step);                         // (:165)
    }
  }
};

}  // namespace qsb
