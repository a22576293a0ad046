"""Multi-GPU parity (needs >= 2 GPUs, skipped otherwise): two ranks run fused.PruneQuantize on
their batch shards with the peer-memory statistics exchange; masks / scales / outputs must equal
a single process running on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_batch(world, per_rank, C):
    rng = np.random.default_rng(9)
    x = np.maximum(rng.standard_normal((world * per_rank, C, 12, 12)), 0).astype(np.float32)
    return x * np.linspace(0.2, 2.0, C, dtype=np.float32).reshape(1, C, 1, 1)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from qsparse_b200.fused import PruneQuantize
    C, per_rank = 48, 6
    full = _make_batch(world, per_rank, C)
    layer = PruneQuantize(sparsity=0.5, bits=8)
    layer.train()
    ys = []
    for t in range(4):
        x = torch.from_numpy(full[rank * per_rank:(rank + 1) * per_rank] * (1 + 0.1 * t)).cuda()
        ys.append(layer(x).cpu().numpy())
    res = dict(rank=rank, p2p=layer._p2p is not None, err=layer._p2p.error() if layer._p2p else 0,
               mask=layer.mask.cpu().numpy(), scale=layer.scale.cpu().numpy(), mag=layer.magnitude.cpu().numpy(), y=ys)
    out.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_peer_memory_exchange_equals_single_process():
    import torch.multiprocessing as mp
    from qsparse_b200.fused import PruneQuantize
    world, C, per_rank = 2, 48, 6
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=180) for _ in range(world)], key=lambda r: r["rank"])
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs)
    full = _make_batch(world, per_rank, C)
    single = PruneQuantize(sparsity=0.5, bits=8)
    single.train()
    ys = [single(torch.from_numpy(full * (1 + 0.1 * t)).cuda()).cpu().numpy() for t in range(4)]
    for r in results:
        assert r["err"] == 0
        assert np.array_equal(r["mask"], single.mask.cpu().numpy())
        assert np.array_equal(r["scale"], single.scale.cpu().numpy())          # MAX statistics: exact
        assert np.allclose(r["mag"], single.magnitude.cpu().numpy(), rtol=1e-6)  # SUM: fixed-order fp64
        for t in range(4):
            lo, hi = r["rank"] * per_rank, (r["rank"] + 1) * per_rank
            assert np.array_equal(r["y"][t], ys[t][lo:hi]), t
    assert results[0]["p2p"] == results[1]["p2p"]
