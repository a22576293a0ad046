"""N > 1 host path on CPU: two gloo ranks exchange their statistics rows; combining them in
rank order must equal the oracle on the concatenated batch (MAX exactly, SUM to fp64 rounding;
masks and scales identical on every rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qsparse_b200.parallel import StatExchange, combine_rows_host, stat_row_bytes
    C, per_rank = 12, 3
    rng = np.random.default_rng(100)                       # the same global batch on every rank
    full = np.maximum(rng.standard_normal((world * per_rank, C, 5, 7)), 0).astype(np.float32)
    shard = full[rank * per_rank:(rank + 1) * per_rank]
    ex = StatExchange(C, torch.device("cpu"))
    assert ex.world == world and ex.row.nbytes == stat_row_bytes(C) == 256
    # what the reduction kernel writes into the row views (here: the oracle as a stand-in)
    ex.row.abssum.copy_(torch.from_numpy(np.abs(shard.astype(np.float64)).sum(axis=(0, 2, 3))))
    ex.row.absmax.copy_(torch.from_numpy(orc.absmax(shard, 1)))
    rows, n_rows, stride = ex.gather()
    assert n_rows == world and stride == 256 and rows.numel() == world * 256
    total, amax = combine_rows_host(rows, n_rows, stride, C)
    count = world * per_rank * 35
    mean = (total / count).float().numpy()
    ref_mean = orc.squeeze_mean_abs(full, (1, C, 1, 1)).reshape(-1)
    assert np.array_equal(amax.numpy(), orc.absmax(full, 1))                       # MAX: exact
    assert np.max(np.abs(mean - ref_mean) / ref_mean) < 3e-7                       # SUM: few ulp
    mag = orc.magnitude_ema(np.zeros(C, np.float32), mean, 0)
    mask, _ = orc.mask_given_importance(mag, 0.5)
    ref_mask, _ = orc.mask_given_importance(orc.magnitude_ema(np.zeros(C, np.float32), ref_mean, 0), 0.5)
    assert np.array_equal(mask, ref_mask)
    scale = orc.scale_ema(np.zeros(1, np.float32), np.array([np.max(amax.numpy() * mask)], np.float32), 8, 0)
    # every rank must hold bit-identical parameters
    gathered = [None] * world
    dist.all_gather_object(gathered, (mean.tobytes(), mask.tobytes(), scale.tobytes()))
    assert all(g == gathered[0] for g in gathered)
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_stat_exchange_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"


def _thr_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qsparse_b200.parallel import combine_thresholds, layer_owner
    L = 7
    truth = torch.tensor([0.5, -0.0, 3e-7, float("inf"), 1.25, 0.0, 9.0])
    local = torch.zeros(L)
    owned = [i for i in range(L) if layer_owner(i, world) == rank]
    for i in owned:
        local[i] = truth[i]
    got = combine_thresholds(local)
    assert torch.equal(got, truth), (rank, got)           # (-0.0 == 0.0: the mask compare is unaffected)
    assert sorted(owned) == list(range(rank, L, world))
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_layer_sharded_thresholds_gloo():
    """host logic of the layer-sharded prune step: round-robin ownership, one all-reduce"""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_thr_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"
