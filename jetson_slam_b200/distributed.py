"""Multi-GPU layer: stereo pairs are independent units (the reference's extract + stereo match carry no state from
frame to frame, src/cuda/orb_gpu.cpp:489-841, src/cuda/orb_stereo_match.cu:105-580), so a batch shards
pair k -> rank k mod G with NO data-path collective.  The only exchange step is the gather of the results to one rank,
issued when a batch consumer asks for the whole batch in one place (BASELINE config C5):

  Gatherer        binding of the C ABI's jsfe_gather_* (include/jsfe.h): every rank packs its results trimmed to the keypoint
                  counts and stores them into the root's landing buffer through a CUDA-IPC peer mapping (NVLink), or sends them
                  with one NCCL send/recv group; runs on its own stream so the transfer overlaps the next extraction.
  unpack_region   host-side reader of a gathered region (the wire format documented in jsfe.h).
  gather_slabs    the round-1 torch.distributed.gather of fixed-capacity slabs (kept for the gloo host-logic tests).

torch.distributed is plumbing: it owns the NCCL communicator and exchanges the IPC handles.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

SLAB_KEYS = ("n", "kps", "desc", "u_right", "depth")
GATHER_MAGIC = 0x3147534A   # 'JSG1'


def header_bytes(n_pairs: int) -> int:
    return (32 + 8 * n_pairs + 15) & ~15


def slot_bytes(n: int, left: bool) -> int:
    return ((64 if left else 56) * n + 15) & ~15


def unpack_region(buf: np.ndarray) -> dict:
    """Decode one rank's region of a gathered buffer (uint8 array, at least header + payload long; layout: include/jsfe.h).
    -> dict(rank, n_pairs, capacity, payload_bytes, sequence, n[2*n_pairs], slots=[dict(kps[6,n], desc[n,32], u_right?, depth?)])."""
    buf = np.ascontiguousarray(buf, np.uint8)
    magic, rank, n_pairs, capacity = (int(v) for v in buf[:16].view(np.int32))
    if magic != GATHER_MAGIC:
        raise ValueError(f"bad region magic {magic:#x}")
    payload, seq = (int(v) for v in buf[16:32].view(np.int64))
    n = buf[32:32 + 8 * n_pairs].view(np.int32).copy()
    off = header_bytes(n_pairs)
    slots = []
    for s in range(2 * n_pairs):
        k = int(n[s])
        left = s % 2 == 0
        sec = buf[off:off + slot_bytes(k, left)]
        d = dict(kps=sec[:24 * k].view(np.int32).reshape(6, k).copy(), desc=sec[24 * k:56 * k].reshape(k, 32).copy())
        if left:
            d["u_right"] = sec[56 * k:60 * k].view(np.float32).copy()
            d["depth"] = sec[60 * k:64 * k].view(np.float32).copy()
        slots.append(d)
        off += slot_bytes(k, left)
    if off - header_bytes(n_pairs) != payload:
        raise ValueError(f"payload_bytes {payload} does not match the sections ({off - header_bytes(n_pairs)})")
    return dict(rank=rank, n_pairs=n_pairs, capacity=capacity, payload_bytes=payload, sequence=seq, n=n, slots=slots)


def nccl_comm_ptr(group=None) -> int:
    """The raw ncclComm_t of torch.distributed's NCCL process group on the current device (0 if there is none)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    backend = pg._get_backend(torch.device("cuda", torch.cuda.current_device()))
    return int(backend._comm_ptr())


class Gatherer:
    """jsfe_gather_* for one rank: results of `max_pairs` pairs per rank, gathered on `root`.

    transport="p2p": the root exports its two landing buffers (CUDA IPC), every other rank maps them and stores its trimmed region
    there directly (falls back to "nccl" on every rank if any mapping fails); transport="nccl": one send/recv group per gather."""

    def __init__(self, fe, max_pairs: int, root: int = 0, group=None, transport: str = "p2p"):
        import ctypes as C
        from . import frontend
        self.fe, self.max_pairs, self.root = fe, max_pairs, root
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._L = frontend.lib()
        self._chk = frontend._check
        self._stream_ptr = frontend._stream_ptr
        self._Gathered = frontend.Gathered
        self.region_bytes = int(self._L.jsfe_gather_region_bytes(fe._h, max_pairs))
        if self.world > 1:   # make sure the communicator exists before asking for its pointer
            dist.barrier(group=group, device_ids=[torch.cuda.current_device()])
        h = C.c_void_p()
        self._chk(self._L.jsfe_gather_create(fe._h, C.c_void_p(nccl_comm_ptr(group)), self.rank, self.world, root, max_pairs, C.byref(h)))
        self._g = h
        self.transport = "single" if self.world == 1 else "nccl"
        if self.world > 1 and transport == "p2p":
            handles = [None, None]
            if self.rank == root:
                for b in range(2):
                    raw = (C.c_uint8 * 64)()
                    self._chk(self._L.jsfe_gather_ipc_export(self._g, b, raw))
                    handles[b] = bytes(raw)
            dist.broadcast_object_list(handles, src=root, group=group)
            ok = 1
            if self.rank != root:
                for b in range(2):
                    raw = (C.c_uint8 * 64).from_buffer_copy(handles[b])
                    if self._L.jsfe_gather_ipc_import(self._g, b, raw) != 0:
                        ok = 0
            flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.transport = "p2p"
                if self.rank == root:
                    self._chk(self._L.jsfe_gather_set_peers_mapped(self._g, 1))
            elif self.rank != root:
                raise RuntimeError("CUDA IPC mapping of the root's landing buffer failed on some rank; use transport='nccl'")

    def begin(self, first_pair: int, n_pairs: int, stream=None):
        self._chk(self._L.jsfe_gather_begin(self._g, first_pair, n_pairs, self._stream_ptr(stream)))

    def end(self):
        """Wait for the gather; on the root returns (device pointer, region stride, world, n_pairs, transport code), else None."""
        import ctypes as C
        g = self._Gathered()
        self._chk(self._L.jsfe_gather_end(self._g, C.byref(g)))
        self.last = g
        return g if self.rank == self.root else None

    STAGES = ("pack", "credit_allreduce", "put", "completion_allreduce")

    def profile(self, on=True):
        """Time the stages of every following gather on its own stream (jsfe_gather_profile)."""
        self._chk(self._L.jsfe_gather_profile(self._g, int(bool(on))))

    def stage_times(self) -> dict:
        """Microseconds per stage of the gather `end()` returned last (this rank's side)."""
        import ctypes as C
        us = (C.c_float * 4)()
        self._chk(self._L.jsfe_gather_stage_times(self._g, us))
        return {k: float(us[i]) for i, k in enumerate(self.STAGES)}

    def regions_to_host(self, g) -> list[np.ndarray]:
        """Root: copy every rank's region (header + payload only) to the host."""
        out = []
        for r in range(g.world):
            base = int(g.data) + r * int(g.region_stride)
            t = torch.as_tensor(_DevArray(base, (int(g.region_stride),), "|u1"), device=torch.device("cuda", self.fe.device))
            hdr = t[:32].cpu().numpy()
            payload = int(hdr[16:24].view(np.int64)[0])
            out.append(t[:header_bytes(int(g.n_pairs)) + payload].cpu().numpy())
        return out

    def close(self):
        if getattr(self, "_g", None):
            self._L.jsfe_gather_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
