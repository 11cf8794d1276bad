"""the gathered result bytes are a pure function of the inputs -- identical for G in {1, 2, 4, 8} ranks (as many as the
box has GPUs), for both transports (peer-memory stores over NVLink, NCCL send/recv), over several gathers in a row (double buffer).
One process per GPU under torch.distributed.run, NCCL; every run also checks the gathered bytes against the staged API on the
owning rank.  CPU part: the wire format reader against a numpy packer."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from jetson_slam_b200 import distributed as jd


def _pack(rank, seq, slots, capacity):
    """numpy mirror of k_gather_pack (include/jsfe.h layout): slots = [(kps[6,n], desc[n,32], u_right?, depth?), ...] in L,R order"""
    n_pairs = len(slots) // 2
    body = b""
    for s, sl in enumerate(slots):
        sec = sl[0].astype(np.int32).tobytes() + sl[1].astype(np.uint8).tobytes()
        if s % 2 == 0:
            sec += sl[2].astype(np.float32).tobytes() + sl[3].astype(np.float32).tobytes()
        body += sec + b"\0" * (-len(sec) % 16)
    hdr = np.array([jd.GATHER_MAGIC, rank, n_pairs, capacity], np.int32).tobytes() + np.array([len(body), seq], np.int64).tobytes()
    hdr += np.array([sl[0].shape[1] for sl in slots], np.int32).tobytes()
    hdr += b"\0" * (-len(hdr) % 16)
    return np.frombuffer(hdr + body, np.uint8)


def test_unpack_region_reads_the_documented_layout():
    rng = np.random.default_rng(0)
    slots = []
    for s, n in enumerate((5, 0, 3, 7)):
        sl = [rng.integers(-9, 9999, size=(6, n)), rng.integers(0, 256, size=(n, 32))]
        if s % 2 == 0:
            sl += [rng.random(n), rng.random(n)]
        slots.append(sl)
    buf = _pack(3, 11, slots, 40)
    assert len(buf) == jd.header_bytes(2) + sum(jd.slot_bytes(sl[0].shape[1], s % 2 == 0) for s, sl in enumerate(slots))
    u = jd.unpack_region(np.concatenate([buf, np.zeros(64, np.uint8)]))       # trailing bytes of the capacity bound are ignored
    assert (u["rank"], u["n_pairs"], u["capacity"], u["sequence"]) == (3, 2, 40, 11) and list(u["n"]) == [5, 0, 3, 7]
    for s, sl in enumerate(slots):
        assert np.array_equal(u["slots"][s]["kps"], sl[0].astype(np.int32)) and np.array_equal(u["slots"][s]["desc"], sl[1].astype(np.uint8))
        if s % 2 == 0:
            assert np.array_equal(u["slots"][s]["u_right"], sl[2].astype(np.float32)) and np.array_equal(u["slots"][s]["depth"], sl[3].astype(np.float32))
    bad = buf.copy()
    bad[0] ^= 1
    with pytest.raises(ValueError):
        jd.unpack_region(bad)


def _run(world, transport, tmp_path, pairs=8):
    out = tmp_path / f"g{world}_{transport}.json"
    cmd = [sys.executable]
    if world > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(29530 + world)]
    cmd += [os.path.join(ROOT, "tests", "gather_worker.py"), "--pairs", str(pairs), "--transport", transport, "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return json.load(open(out))


@pytest.mark.gpu
def test_gathered_bytes_are_identical_for_every_world_size(tmp_path):
    import torch
    ndev = torch.cuda.device_count()
    base = _run(1, "p2p", tmp_path)
    assert base["transport"] == "single" and len(base["rounds"]) == 3 and len(set(base["rounds"][0])) == 8
    assert base["rounds"][0] != base["rounds"][1]
    seen = {1}
    for world in (2, 4, 8):
        if world > ndev:
            continue
        for transport in ("p2p", "nccl"):
            got = _run(world, transport, tmp_path)
            assert got["transport"] == transport
            assert got["rounds"] == base["rounds"], f"world {world} / {transport}: gathered bytes differ from the single-GPU run"
        seen.add(world)
    print("world sizes checked:", sorted(seen))
