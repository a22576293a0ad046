"""Multi-GPU plumbing of the quantize/prune path.

The path shards along the batch: every GPU reduces, applies and back-propagates its own
activation shard; the only exchange is the tiny per-channel statistics row
``[C x fp64 sum|x|  |  C x fp32 max|x|]`` (768 B for 64 channels).  One
``all_gather`` moves the rows; the parameter kernel then combines them in rank order
(SUM / MAX), so every rank derives bit-identical magnitudes, masks and scales — equal
to a single process running on the concatenated batch (MAX exactly; SUM to the
last fp64 bit of a fixed-order sum).  The bulk tensors never cross NVLink.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def stat_row_bytes(channels: int) -> int:
    """bytes of one rank's statistics row, padded to 256 B"""
    return (channels * 12 + 255) // 256 * 256


class StatRow:
    """A rank-local statistics row plus typed views into it.  The reduction kernel
    writes its results straight into these views (no packing launch)."""

    def __init__(self, channels: int, device):
        self.channels = channels
        self.nbytes = stat_row_bytes(channels)
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.abssum = self.buf[: 8 * channels].view(torch.float64)
        self.absmax = self.buf[8 * channels: 12 * channels].view(torch.float32)


class StatExchange:
    """all-gather of statistics rows over a process group (None = default group).

    ``gather(row)`` returns ``(rows_buffer, n_rows, row_stride_bytes)``; with a single
    process it is the row itself and no collective runs."""

    def __init__(self, channels: int, device, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.row = StatRow(channels, device)
        self.gathered = (torch.zeros(self.world * self.row.nbytes, dtype=torch.uint8, device=device)
                         if self.world > 1 else self.row.buf)

    def gather(self) -> Tuple[torch.Tensor, int, int]:
        if self.world == 1:
            return self.row.buf, 1, self.row.nbytes
        dist.all_gather_into_tensor(self.gathered, self.row.buf, group=self.group)
        return self.gathered, self.world, self.row.nbytes

    def views(self, buf: torch.Tensor):
        """(abssum row 0, absmax row 0) views of a gathered buffer, as the kernel wants them"""
        c = self.row.channels
        return buf[: 8 * c].view(torch.float64), buf[8 * c: 12 * c].view(torch.float32)


class P2PExchange:
    """Peer-memory statistics exchange: every rank maps every other rank's small exchange
    buffer through CUDA IPC (same node, NVLink / NVSwitch); the parameter step (the tail of
    the fused reduction kernel) pushes its row to all peers as 8-byte {data, stamp} packets,
    polls its own buffer for theirs and combines the rows in rank order — no collective launch
    on the critical path.  ``handle`` is what ``ops.reduce_prune_quant_step(group=...)`` takes.

    The constructor is collective.  Every local CUDA step is followed by a vote (all-reduce of
    an ok flag), so a failure on ONE rank makes every rank raise at the same point instead of
    leaving the others inside a mismatched collective.

    A step whose peers do not show up within ``timeout_ms`` does not continue on stale data: the
    kernel writes NaN into the layer's scale / decimal / magnitude and raises the group's error
    flag; ``check()`` (called by ``PruneQuantize`` every ``check_every`` steps) turns that into a
    ``RuntimeError``."""

    def __init__(self, channels: int, device, group: Optional[dist.ProcessGroup] = None,
                 timeout_ms: int = 30000):
        import ctypes
        from ctypes import byref, c_int, c_int64, c_void_p

        from . import _native as N

        self._N = N
        self.lib = N.load_library()
        self.group = group
        self.device = torch.device(device)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local = c_void_p()
        self.peers = []
        self.handle = None
        self.stamp = 0
        self._err_host = None
        self._err_event = None
        if self.world > 16:
            raise RuntimeError("P2PExchange supports up to 16 ranks of one node")
        # the vote travels on the group's own backend: CPU tensor for gloo, device tensor for NCCL
        vote_dev = self.device if dist.get_backend(group) == "nccl" else torch.device("cpu")

        def vote(err: Optional[BaseException], what: str):
            ok = torch.tensor([0.0 if err is not None else 1.0], device=vote_dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 0:
                self.close()
                raise RuntimeError(f"peer-memory exchange: {what} failed on "
                                   f"{'this rank: ' + repr(err) if err is not None else 'another rank'}")

        nbytes = self.lib.qsb_p2p_group_bytes(c_int(self.world), c_int64(channels))
        handle = ctypes.create_string_buffer(64)
        err = None
        with torch.cuda.device(self.device):
            try:
                N.check(self.lib.qsb_p2p_alloc(c_int64(nbytes), byref(self.local), handle), "qsb_p2p_alloc")
            except Exception as exc:
                err = exc
            vote(err, "qsb_p2p_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            bufs = (c_void_p * self.world)()
            try:
                for r in range(self.world):
                    if r == self.rank:
                        bufs[r] = self.local
                        continue
                    p = c_void_p()
                    N.check(self.lib.qsb_p2p_open(ctypes.create_string_buffer(handles[r], 64), byref(p)),
                            "qsb_p2p_open")
                    self.peers.append(p)
                    bufs[r] = p
                h = c_void_p()
                N.check(self.lib.qsb_p2p_group_create(byref(h), c_int(self.rank), c_int(self.world),
                                                      c_int64(channels), bufs), "qsb_p2p_group_create")
                self.handle = h
                N.check(self.lib.qsb_p2p_group_set_timeout_ms(self.handle, c_int64(int(timeout_ms))),
                        "qsb_p2p_group_set_timeout_ms")
            except Exception as exc:
                err = exc
            # every peer has mapped every buffer before the first kernel (the vote is the barrier)
            vote(err, "qsb_p2p_open / qsb_p2p_group_create")

    def next_stamp(self) -> int:
        self.stamp += 1
        return self.stamp

    def error(self) -> int:
        """the group's error flag (synchronises with the device)"""
        from ctypes import byref, c_int
        e = c_int(0)
        self._N.check(self.lib.qsb_p2p_group_error(self.handle, byref(e)), "qsb_p2p_group_error")
        return e.value

    def check(self):
        """Raise if a PREVIOUS poll saw the error flag, then queue the next asynchronous read of
        the flag behind the work already on the current stream — no synchronisation per step;
        a timed-out step is reported one poll later (and its outputs are NaN in the meantime)."""
        from ctypes import c_void_p
        if self._err_host is None:
            self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._err_event = torch.cuda.Event()
        elif self._err_event.query() and int(self._err_host[0]) != 0:
            raise RuntimeError(
                "qsparse_b200: a peer GPU's statistics did not arrive within the exchange timeout; the step was "
                "poisoned (scale / decimal / magnitude are NaN) instead of continuing on stale data")
        self._N.check(self.lib.qsb_p2p_group_error_async(self.handle, c_void_p(self._err_host.data_ptr()),
                                                         self._N.stream_ptr(self.device)),
                      "qsb_p2p_group_error_async")
        self._err_event.record(torch.cuda.current_stream(self.device))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.qsb_p2p_group_destroy(self.handle)
        self.handle = None
        for p in getattr(self, "peers", []):
            self.lib.qsb_p2p_close(p)
        self.peers = []
        if getattr(self, "local", None):
            self.lib.qsb_p2p_free(self.local)
            self.local = None

    def __del__(self):   # the cudaMalloc'ed buffer and the IPC mappings do not belong to torch's allocator
        try:
            self.close()
        except Exception:
            pass


def make_exchange(channels: int, device, group=None, timeout_ms: int = 30000):
    """P2PExchange when a multi-rank group is initialised and peer mapping works on EVERY rank, else
    None (single process, or the caller falls back to StatExchange / NCCL).  Collective: every rank
    returns the same kind of object (P2PExchange's constructor votes after each phase)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    try:
        return P2PExchange(channels, device, group, timeout_ms=timeout_ms)
    except RuntimeError as exc:  # pragma: no cover - depends on the box
        print(f"[qsparse_b200] peer-memory exchange unavailable ({exc}); using NCCL all-gather")
        return None


def assert_equal_shards(n_local: int, group=None):
    """The exchange combines per-channel SUMS and divides by ``count = n_local * world``: every rank
    must hold the same number of elements per channel (the reference's mean over the concatenated
    batch).  One all-gather of an integer at layer initialisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    sizes = [None] * dist.get_world_size(group)
    dist.all_gather_object(sizes, int(n_local), group=group)
    if len(set(sizes)) != 1:
        raise RuntimeError(f"qsparse_b200: ranks hold different shard sizes per channel {sizes}; the batch-sharded "
                           "prune/quantize step needs equal per-rank batches")


def combine_rows_host(gathered: torch.Tensor, n_rows: int, row_bytes: int, channels: int):
    """Host restatement of how the parameter kernel combines rows (rank order: SUM of the
    fp64 sums, MAX of the maxima).  Used by the CPU (gloo) tests of the exchange."""
    total = torch.zeros(channels, dtype=torch.float64)
    amax = torch.zeros(channels, dtype=torch.float32)
    g = gathered.cpu()
    for r in range(n_rows):
        row = g[r * row_bytes: (r + 1) * row_bytes]
        total = total + row[: 8 * channels].view(torch.float64)
        amax = torch.maximum(amax, row[8 * channels: 12 * channels].view(torch.float32).abs())
    return total, amax


# ----------------------------------------------------------------------------- weights
def layer_owner(layer_index: int, world: int) -> int:
    """Round-robin assignment of prune layers to ranks: the thresholds of different layers are
    independent, so the expensive part of an unstructured prune step (the k-th value of every
    layer's importance) shards by layer."""
    return layer_index % world


def combine_thresholds(local: torch.Tensor, group=None) -> torch.Tensor:
    """``local[i]`` holds the threshold of layer i on its owner and 0 elsewhere; one all-reduce
    (SUM of one non-zero term: exact) leaves every threshold on every rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local


def sharded_layer_thresholds(importances, ks, group=None, take_abs=False) -> torch.Tensor:
    """Thresholds ``sorted(importances[i])[ks[i]]`` for a set of (replicated) layers: every rank
    selects only on the layers it owns, then one all-reduce of L floats.  Returns a float32
    tensor [L] on the importances' device, identical on every rank (and equal to the
    single-process thresholds: the select is exact)."""
    from . import ops

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    dev = importances[0].device
    thr = torch.zeros(len(importances), dtype=torch.float32, device=dev)
    owned = [i for i in range(len(importances)) if layer_owner(i, world) == rank]
    if owned:   # every owned layer in the same launches (blockIdx.y = layer)
        got = ops.kth_value_batched([importances[i] for i in owned], [ks[i] for i in owned], take_abs=take_abs)
        if len(owned) == len(importances):
            return got                                  # a single rank owns everything
        for j, i in enumerate(owned):                   # device-to-device, capturable
            thr[i:i + 1] = got[j:j + 1]
    return combine_thresholds(thr, group)


def sharded_kth_value(v_local: torch.Tensor, k_global: int, group=None, take_abs=False) -> torch.Tensor:
    """k-th smallest value (0-based, global rank) of a tensor SHARDED over the ranks of ``group``
    by element range: three local histogram passes, each followed by an all-reduce (SUM) of the
    integer histogram (<= 32 KB) — exact, identical on every rank, and the bulk tensor never
    leaves its GPU.  ``qsb_kth_dist_*`` in include/qsparse_b200.h."""
    import ctypes
    from ctypes import byref, c_int, c_int64, c_void_p

    from . import _native as N

    lib = N.load_library()
    N.require_cuda(v_local, "v_local")
    v = N.as_f32_contiguous(v_local.detach()).reshape(-1)
    n = v.numel()
    dev = v.device
    nbytes = lib.qsb_kth_workspace_bytes(c_int64(n))
    ws = N.workspace(dev, nbytes)
    stream = N.stream_ptr(dev)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    N.check(lib.qsb_kth_dist_begin(N.ptr(ws), c_int64(ws.numel()), stream), "qsb_kth_dist_begin")
    for p in range(3):
        hist, count = c_void_p(), c_int64()
        N.check(lib.qsb_kth_dist_pass(N.ptr(v) if n else c_void_p(0), c_int64(n), c_int64(int(k_global)), c_int(p),
                                      c_int(1 if take_abs else 0), N.ptr(ws), c_int64(ws.numel()), byref(hist),
                                      byref(count), stream), "qsb_kth_dist_pass")
        if multi:
            off = hist.value - ws.data_ptr()
            h = ws[off: off + 8 * count.value].view(torch.int64)   # 64-bit counters (< 2^63)
            dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    thr = torch.empty(1, dtype=torch.float32, device=dev)
    N.check(lib.qsb_kth_dist_final(c_int64(int(k_global)), N.ptr(ws), N.ptr(thr), stream), "qsb_kth_dist_final")
    return thr


FUSE_PRUNE_STEP = True   # False: EMA / select / mask as three (multi-tensor) stages — same results


def prune_weight_set_step(weights, magnitudes, masks, outs, t: int, sparsity: float, group=None,
                          shard_by_layer: bool = False, hints=None):
    """One unstructured, running-average prune step over a set of replicated weight tensors
    (BASELINE config 4) on every rank of ``group``:

        magnitude EMA (replicated, 12 B/elem)  ref sparse.py:82-89
        k-th value per layer, sharded by layer + one all-reduce of L floats  ref util.py:103-117
        mask = magnitude >= thr, out = w * mask (replicated, 13 B/elem)       ref sparse.py:65-66,116

    ``t`` is the callback's step counter.  Results equal ``MagnitudePruningCallback`` run on one
    GPU, bit for bit.

    Default: every rank runs the one-pass kernel K9 on all (replicated) layers — 236 us for 64 Mi
    elements, no communication at all, and bit-identical on every rank because the threshold is exact.
    ``shard_by_layer=True`` takes the three-stage route in which only the select is sharded:
    worthwhile when the select dominates, e.g. very large layers on many ranks."""
    from . import ops
    from .util import kth_rank

    ks = [kth_rank(sparsity, m.numel()) for m in magnitudes]
    if not shard_by_layer and FUSE_PRUNE_STEP and ops.prune_step_supported(magnitudes, weights, masks, outs):
        # EMA, select and mask/apply of every layer in ONE streaming pass (K9, ~17.5 B/elem)
        # hints (ops.new_select_hints(len(weights), device), kept by the caller across steps): warm-started
        # pivots — from the third step on no sampler pass and a few thousand candidates per layer
        return ops.prune_unstructured_step_batched_(magnitudes, weights, masks, outs, ks, t, hints=hints)
    ops.magnitude_ema_full_multi_(magnitudes, weights, t)            # one launch for the whole set
    thr = sharded_layer_thresholds(magnitudes, ks, group)            # one launch sequence per rank
    ops.mask_build_apply_multi(magnitudes, thr, weights, masks, outs)  # one launch
    return thr
