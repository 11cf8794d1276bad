"""Multi-GPU layer: stereo pairs are independent units (the reference's extract + stereo match carry no state from
frame to frame, src/cuda/orb_gpu.cpp:489-841, src/cuda/orb_stereo_match.cu:105-580), so a batch shards
pair k -> rank k mod G with NO data-path collective.  The only collective is the optional gather of the
fixed-capacity result slabs to one rank, issued when a batch consumer asks for the whole batch in one place
(BASELINE config C5).  torch.distributed is plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

SLAB_KEYS = ("n", "kps", "desc", "u_right", "depth")


def shard_pairs(n_pairs: int, world: int, rank: int) -> list[int]:
    """Global pair indices owned by `rank` (round-robin)."""
    return list(range(rank, n_pairs, world))


def local_capacity(n_pairs: int, world: int) -> int:
    """Pairs every rank must be able to hold (ceil); ranks with fewer pairs pad their slabs."""
    return (n_pairs + world - 1) // world


class _DevArray:
    """Zero-copy view of handle-owned device memory for torch.as_tensor (CUDA array interface v3)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def slab_tensors(fe, first_slot: int, n_slots: int) -> dict[str, torch.Tensor]:
    """Torch tensors aliasing a Frontend's device result slabs for slots [first_slot, first_slot+n_slots)."""
    v = fe.slot_view(first_slot)
    cap = v.capacity
    dev = torch.device("cuda", fe.device)
    mk = lambda ptr, shape, ts: torch.as_tensor(_DevArray(ptr, shape, ts), device=dev)
    return {
        "n": mk(v.n_keypoints, (n_slots,), "<i4"),
        "kps": mk(v.kps, (n_slots, 6, cap), "<i4"),
        "desc": mk(v.desc, (n_slots, cap, 32), "|u1"),
        "u_right": mk(v.u_right, (n_slots, cap), "<f4"),
        "depth": mk(v.depth, (n_slots, cap), "<f4"),
    }


_PERM_CACHE: dict = {}


def _slot_permutation(n_pairs: int, world: int, device) -> torch.Tensor:
    """Index into the rank-major stack [world * 2*cap_pairs] that yields global pair order [2*n_pairs]."""
    key = (n_pairs, world, str(device))
    if key not in _PERM_CACHE:
        cap_pairs = local_capacity(n_pairs, world)
        perm = np.empty(2 * n_pairs, np.int64)
        for r in range(world):
            for j, p in enumerate(shard_pairs(n_pairs, world, r)):
                base = (r * cap_pairs + j) * 2
                perm[2 * p], perm[2 * p + 1] = base, base + 1
        _PERM_CACHE[key] = torch.from_numpy(perm).to(device)
    return _PERM_CACHE[key]


def gather_slabs(local: dict[str, torch.Tensor], n_pairs: int, dst: int = 0, group=None) -> dict[str, torch.Tensor] | None:
    """Gather per-rank slabs (leading dim = 2 * local_capacity slots, pair-major L,R) to rank `dst` and put them in
    global pair order: one collective and one index_select per array.  Returns the dict on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cap_pairs = local_capacity(n_pairs, world)
    out = {}
    for k in SLAB_KEYS:
        t = local[k]
        assert t.shape[0] == 2 * cap_pairs, f"{k}: leading dim {t.shape[0]} != 2*{cap_pairs}"
        stacked = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device) if rank == dst else None
        bufs = list(stacked.unbind(0)) if rank == dst else None
        dist.gather(t.contiguous(), bufs, dst=dst, group=group)
        if rank == dst:
            flat = stacked.reshape((world * 2 * cap_pairs,) + tuple(t.shape[1:]))
            out[k] = flat.index_select(0, _slot_permutation(n_pairs, world, t.device))
    return out if rank == dst else None


def run_sharded(fe, images: np.ndarray, cfg, n_pairs: int, gather_to: int | None = None, group=None, stream=None):
    """Process this rank's share of a global batch `images` [2*n_pairs, H, W] (every rank sees the same host array,
    e.g. from a shared loader) on its own GPU; optionally gather the result slabs to `gather_to`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_pairs(n_pairs, world, rank)
    if mine:
        idx = np.array([[2 * p, 2 * p + 1] for p in mine]).ravel()
        fe.set_images(np.ascontiguousarray(images[idx]), 0, stream)
        fe.extract(0, 2 * len(mine), stream)
        fe.stereo_match(cfg.mb, cfg.mbf, 0, len(mine), stream=stream)
    if gather_to is None or world == 1:
        return None
    torch.cuda.current_stream().synchronize() if stream is None else stream.synchronize()
    cap_pairs = local_capacity(n_pairs, world)
    local = slab_tensors(fe, 0, 2 * cap_pairs)
    if len(mine) < cap_pairs:   # slots beyond this rank's share hold stale data: report 0 keypoints for them
        local = dict(local)
        n = local["n"].clone()
        n[2 * len(mine):] = 0
        local["n"] = n
    return gather_slabs(local, n_pairs, dst=gather_to, group=group)
