"""Multi-process stream sharding (world size 2, gloo, CPU): each rank enhances its own stream range
with no data-path collective; gathering the shards reproduces the single-process result."""
import os

import numpy as np
import pytest

from dpdfnet_b200 import dist as ddist


def test_partition_covers_all_streams():
    for n, w in ((1024, 8), (10, 3), (2, 4), (0, 2)):
        parts = ddist.partition(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n and len(parts) == w
        assert all(a1 == b0 for (_, b0), (a1, _) in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ddist.partition(4, 0)


def _worker(rank, world, port, pcm, ret):
    import torch
    import torch.distributed as dist
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import pack_tensors, random_checkpoint
    from oracle.oracle_np import OracleEngine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = get_spec("baseline")
    mine = ddist.shard(pcm, rank, world)
    eng = OracleEngine(spec, pack_tensors(spec, random_checkpoint(spec, 0)), max(len(mine), 1))
    hop = spec.hop
    outs = [eng.step_pcm(mine[:, t * hop:(t + 1) * hop]) for t in range(pcm.shape[1] // hop)] if len(mine) else []
    local = torch.from_numpy(np.concatenate(outs, 1) if outs else np.zeros((0, pcm.shape[1]), np.float32))
    full = ddist.gather_rows(local, pcm.shape[0])
    if rank == 0:
        ret["out"] = full.numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_streams_equal_single_process():
    import torch.multiprocessing as mp
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import pack_tensors, random_checkpoint
    from oracle.oracle_np import OracleEngine
    spec = get_spec("baseline")
    pcm = (np.random.default_rng(0).standard_normal((5, 6 * spec.hop)) * 0.1).astype(np.float32)
    ref_eng = OracleEngine(spec, pack_tensors(spec, random_checkpoint(spec, 0)), 5)
    ref = np.concatenate([ref_eng.step_pcm(pcm[:, t * spec.hop:(t + 1) * spec.hop]) for t in range(6)], 1)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, 29000 + os.getpid() % 2000, pcm, ret), nprocs=2, join=True)
        out = ret["out"]
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() < 1e-5


def _worker_bcast(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import pack_checkpoint, random_checkpoint
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = get_spec("baseline")
    blob = pack_checkpoint(spec, random_checkpoint(spec, 3)) if rank == 0 else None      # only rank 0 has the checkpoint
    got = ddist.broadcast_weights(blob)
    n = 5
    full = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4) if rank == 0 else torch.empty(0, 4)
    mine = ddist.scatter_rows(full, n)
    back = ddist.gather_rows(mine * 2, n)
    ret[rank] = (len(got), hash(got), mine.numpy().copy(), back.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_pcm_scatter_gather():
    """Start-up broadcast of the packed weights from the rank that has the checkpoint, scatter of a front rank's
    rows and gather of the results, world size 2 on gloo."""
    import torch.multiprocessing as mp
    from dpdfnet_b200.spec import get_spec
    from dpdfnet_b200.weights import pack_checkpoint, random_checkpoint
    spec = get_spec("baseline")
    ref = pack_checkpoint(spec, random_checkpoint(spec, 3))
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_bcast, args=(2, 31000 + os.getpid() % 2000, ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0] == r1[0] == len(ref)
    full = np.arange(20, dtype=np.float32).reshape(5, 4)
    assert np.array_equal(r0[2], full[:3]) and np.array_equal(r1[2], full[3:])
    assert np.array_equal(r0[3], full * 2) and np.array_equal(r1[3], full * 2)
