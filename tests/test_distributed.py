"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: the pair sharding rule and the gather/reassembly of
fixed-capacity result slabs.  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jetson_slam_b200 import distributed as jd


def test_shard_pairs_is_a_partition():
    for n in (0, 1, 5, 8, 17):
        for world in (1, 2, 3, 8):
            got = sorted(p for r in range(world) for p in jd.shard_pairs(n, world, r))
            assert got == list(range(n))
            assert max((len(jd.shard_pairs(n, world, r)) for r in range(world)), default=0) <= jd.local_capacity(n, world)


def _fake_slabs(pairs, cap_pairs, cap=7):
    """Deterministic slabs for the given global pair ids, padded to cap_pairs pairs."""
    n_slots = 2 * cap_pairs
    out = {"n": torch.zeros(n_slots, dtype=torch.int32), "kps": torch.zeros(n_slots, 6, cap, dtype=torch.int32),
           "desc": torch.zeros(n_slots, cap, 32, dtype=torch.uint8), "u_right": torch.full((n_slots, cap), -1.0),
           "depth": torch.full((n_slots, cap), -1.0)}
    for j, p in enumerate(pairs):
        for eye in (0, 1):
            s = 2 * j + eye
            out["n"][s] = 3 + (p + eye) % 4
            out["kps"][s] = torch.arange(6 * cap, dtype=torch.int32).reshape(6, cap) + 1000 * p + 100 * eye
            out["desc"][s] = (torch.arange(cap * 32).reshape(cap, 32) + p + eye) % 251
            out["u_right"][s] = torch.arange(cap, dtype=torch.float32) + p
            out["depth"][s] = torch.arange(cap, dtype=torch.float32) * 0.5 + p
    return out


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = jd.shard_pairs(n_pairs, world, rank)
        local = _fake_slabs(mine, jd.local_capacity(n_pairs, world))
        got = jd.gather_slabs(local, n_pairs, dst=0)
        if rank == 0:
            want = _fake_slabs(list(range(n_pairs)), n_pairs)
            ok = all(torch.equal(got[k], want[k]) for k in jd.SLAB_KEYS)
            q.put(("ok" if ok else "mismatch", {k: tuple(got[k].shape) for k in got}))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [4, 5])
def test_gather_slabs_world2_gloo(n_pairs):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    status, shapes = q.get(timeout=10)
    assert status == "ok", shapes
    assert shapes["n"] == (2 * n_pairs,)


def test_reference_arm_prints_only_on_rank0(monkeypatch, capsys):
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints; other ranks exit without work."""
    import bench
    from jetson_slam_b200.configs import CONFIGS

    class A:
        gpus, steps, warmup, cpu_threads = 2, 1, 0, 2
    monkeypatch.setenv("RANK", "1")
    bench.run_reference_arm(A, CONFIGS["tiny"])
    assert capsys.readouterr().out == ""
