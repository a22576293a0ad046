"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy front end of ``oracle/liboracle.so`` (the plain-C restatement of the
reference's quantize/prune arithmetic, see ``qsparse_oracle.c``) plus the
reference's host-side schedule arithmetic restated in Python.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  Parity status: PINNED against golden vectors generated
from the imported reference (``oracle/gen_golden.py`` -> ``tests/golden/``).
"""
from __future__ import annotations

import ctypes
import subprocess
from ctypes import c_double, c_float, c_int, c_int64, c_void_p
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"
_lib = None

MASK_NONE, MASK_CHANNEL, MASK_ELEMENT = 0, 1, 2


def build(force: bool = False) -> Path:
    src = _HERE / "qsparse_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.orc_kth_value.restype = c_float
        _lib.orc_squeeze_mean_abs.restype = c_int
        _lib.orc_mask_given_importance.restype = c_int
    return _lib


def _p(a):
    return c_void_p(0 if a is None else a.ctypes.data)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def layout(shape, channel_index):
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if channel_index is None or channel_index < 0:
        return 1, 1, n
    outer = int(np.prod(shape[:channel_index], dtype=np.int64))
    inner = int(np.prod(shape[channel_index + 1:], dtype=np.int64))
    return outer, int(shape[channel_index]), inner


def _mask_args(mask, x_shape, channel_index):
    if mask is None:
        return None, MASK_NONE
    m = np.ascontiguousarray(mask).astype(np.uint8).reshape(-1)
    n = int(np.prod(x_shape, dtype=np.int64))
    if m.size == n:
        return m, MASK_ELEMENT
    return m, MASK_CHANNEL


# ---------------------------------------------------------------- fake-quant
def fq_pow2_fwd(x, decimal, channel_index=-1, mask=None):
    """ref DecimalQuantization.forward qsparse/quantize.py:30-63 (optionally after x*mask)."""
    x = _f32(x)
    d = _f32(np.asarray(decimal)).reshape(-1)
    o, c, i = layout(x.shape, channel_index if (d.size > 1 or mask is not None) else -1)
    m, mk = _mask_args(mask, x.shape, channel_index)
    if mk == MASK_ELEMENT and d.size == 1:
        o, c, i = 1, 1, x.size
    y = np.empty_like(x)
    lib().orc_fq_pow2_fwd(_p(x), _p(y), c_int64(o), c_int64(c), c_int64(i), _p(d), c_int64(d.size), _p(m), c_int(mk))
    return y


def fq_scaler_fwd(x, scaler, channel_index=-1, mask=None):
    """ref ScalerQuantization.forward qsparse/quantize.py:86-117."""
    x = _f32(x)
    s = _f32(np.asarray(scaler)).reshape(-1)
    o, c, i = layout(x.shape, channel_index if (s.size > 1 or mask is not None) else -1)
    m, mk = _mask_args(mask, x.shape, channel_index)
    if mk == MASK_ELEMENT and s.size == 1:
        o, c, i = 1, 1, x.size
    y = np.empty_like(x)
    lib().orc_fq_scaler_fwd(_p(x), _p(y), c_int64(o), c_int64(c), c_int64(i), _p(s), c_int64(s.size), _p(m), c_int(mk))
    return y


def fq_line_fwd(x, lines, bits=8, channel_index=-1, float_zero_point=True, mask=None):
    """ref LineQuantization.forward qsparse/quantize.py:140-181."""
    x = _f32(x)
    l = _f32(np.asarray(lines)).reshape(-1, 2)
    o, c, i = layout(x.shape, channel_index if (l.shape[0] > 1 or mask is not None) else -1)
    m, mk = _mask_args(mask, x.shape, channel_index)
    if mk == MASK_ELEMENT and l.shape[0] == 1:
        o, c, i = 1, 1, x.size
    y = np.empty_like(x)
    lib().orc_fq_line_fwd(_p(x), _p(y), c_int64(o), c_int64(c), c_int64(i), _p(l), c_int64(l.shape[0]),
                          c_int(bits), c_int(1 if float_zero_point else 0), _p(m), c_int(mk))
    return y


def ste_bwd(g, scale, bits=8, channel_index=-1, is_decimal=True, flip_axis=False, mask=None):
    """ref Decimal/ScalerQuantization.backward qsparse/quantize.py:65-77.
    Returns (g_clamped, gx) — gx is None without a mask."""
    g = _f32(g).copy()
    s = _f32(np.asarray(scale)).reshape(-1)
    o, c, i = layout(g.shape, channel_index if (s.size > 1 or mask is not None) else -1)
    m, mk = _mask_args(mask, g.shape, channel_index)
    if mk == MASK_ELEMENT and s.size == 1:
        o, c, i = 1, 1, g.size
    gx = np.empty_like(g) if mask is not None else None
    lib().orc_ste_bwd(_p(g), _p(gx), c_int64(o), c_int64(c), c_int64(i), _p(s), c_int64(s.size),
                      c_int(1 if is_decimal else 0), c_int(bits), c_int(1 if flip_axis else 0), _p(m), c_int(mk))
    return g, gx


def mask_apply(x, mask, channel_index=-1):
    """ref `x * mask` qsparse/sparse.py:66."""
    x = _f32(x)
    m, mk = _mask_args(mask, x.shape, channel_index)
    o, c, i = layout(x.shape, channel_index) if mk == MASK_CHANNEL else (1, 1, x.size)
    y = np.empty_like(x)
    lib().orc_mask_apply(_p(x), _p(y), c_int64(o), c_int64(c), c_int64(i), _p(m), c_int(mk))
    return y


# ---------------------------------------------------------------- statistics
def absmax(x, channel_index=-1):
    """ref DecimalQuantizer.optimize qsparse/quantize.py:329-340 (before the /2^(b-1))."""
    x = _f32(x)
    o, c, i = layout(x.shape, channel_index)
    out = np.empty(c, dtype=np.float32)
    lib().orc_absmax(_p(x), c_int64(o), c_int64(c), c_int64(i), _p(out))
    return out


def minmax(x, channel_index=-1):
    """ref AdaptiveQuantizer.optimize qsparse/quantize.py:396-418."""
    x = _f32(x)
    o, c, i = layout(x.shape, channel_index)
    mn = np.empty(c, dtype=np.float32)
    mx = np.empty(c, dtype=np.float32)
    lib().orc_minmax(_p(x), c_int64(o), c_int64(c), c_int64(i), _p(mn), _p(mx))
    return mn, mx


def squeeze_mean_abs(x, target_shape, use_l0=False):
    """ref squeeze_tensor_to_shape(x.abs(), shape) qsparse/util.py:79-99."""
    x = _f32(x)
    assert x.ndim == len(target_shape)
    keep = []
    for sx, sm in zip(x.shape, target_shape):
        if sx != sm and sm != 1:
            raise ValueError("mismatch between the input tensor and mask")
        keep.append(1 if sx == sm else 0)
    dims = (c_int64 * x.ndim)(*x.shape)
    keep_a = (c_int * x.ndim)(*keep)
    out = np.empty(tuple(target_shape), dtype=np.float32)
    rc = lib().orc_squeeze_mean_abs(_p(x), dims, c_int(x.ndim), keep_a, c_int(1 if use_l0 else 0), _p(out))
    assert rc == 0
    return out


def scale_ema(w, amax, bits, t):
    """ref qsparse/quantize.py:340,344-348."""
    w = _f32(w).copy().reshape(-1)
    a = _f32(amax).reshape(-1)
    lib().orc_scale_ema(_p(w), _p(a), c_int64(w.size), c_int(bits), c_int64(t))
    return w


def scale_to_decimal(s):
    """ref qsparse/quantize.py:316."""
    s = _f32(s)
    d = np.empty_like(s)
    lib().orc_scale_to_decimal(_p(s), _p(d), c_int64(s.size))
    return d


def lines_ema(lines, mn, mx, t):
    """ref qsparse/quantize.py:428-430 (t already incremented)."""
    l = _f32(lines).copy().reshape(-1, 2)
    lib().orc_lines_ema(_p(l), _p(_f32(mn)), _p(_f32(mx)), c_int64(l.shape[0]), c_int64(t))
    return l


def magnitude_ema(mag, m, t):
    """ref qsparse/sparse.py:89."""
    mag = _f32(mag).copy()
    m = _f32(m)
    lib().orc_magnitude_ema(_p(mag), _p(m), c_int64(mag.size), c_int64(t))
    return mag


# ---------------------------------------------------------------- prune mask
def mask_given_importance(importance, sparsity):
    """ref calculate_mask_given_importance qsparse/util.py:103-117.  Returns (mask, thr)."""
    imp = _f32(importance)
    mask = np.empty(imp.size, dtype=np.uint8)
    thr = c_float(0)
    rc = lib().orc_mask_given_importance(_p(imp), c_int64(imp.size), c_double(float(sparsity)), _p(mask),
                                         ctypes.byref(thr))
    if rc != 0:
        raise IndexError("index out of range in calculate_mask_given_importance")
    return mask.reshape(imp.shape).astype(bool), np.float32(thr.value)


def kth_value(x, k, take_abs=False):
    x = _f32(x)
    return np.float32(lib().orc_kth_value(_p(x), c_int64(x.size), c_int64(k), c_int(1 if take_abs else 0)))


def kth_index(sparsity: float, n: int) -> int:
    """rank (0-based, ascending) of the threshold: idx + 1, idx = max(int(s*n - 1), 0).
    ref qsparse/util.py:114-115."""
    return max(int(sparsity * n - 1), 0) + 1


# ---------------------------------------------------------------- host-side schedule
def prune_schedule(start, interval, repetition, rampup=False):
    """ref PruneLayer.__init__ qsparse/sparse.py:186-195 -> (schedules, rampup_interval)."""
    schedules = [start + interval * ((1 if rampup else 0) + i) for i in range(repetition)]
    return schedules, (0 if rampup else interval)


def ramp_sparsity(n_updates, sparsity, start, interval, repetition, rampup_interval):
    """ref PruneLayer.forward qsparse/sparse.py:252-257; the value is stored in an fp32
    parameter and read back with .item(), hence the float32 round trip."""
    ratio = (1.0 - (n_updates - start + rampup_interval) / (interval * repetition)) ** 3
    return float(np.float32(sparsity * (1 - ratio)))


def refresh_gate(t, sparsity, mask_refresh_interval, stop_mask_refresh, running_average):
    """ref MagnitudePruningCallback.forward qsparse/sparse.py:110-113."""
    return (sparsity >= 0 and (t % mask_refresh_interval == 0 and t <= stop_mask_refresh)
            and (t > 0 or not running_average))
