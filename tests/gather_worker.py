"""Worker of tests/test_gather_multigpu.py and of bench.py's gather check: run under torchrun (or alone for one rank).
Processes a fixed global batch of stereo pairs sharded pair k -> rank k mod G, gathers the results on rank 0 with the C ABI's
jsfe_gather_* and writes, per global pair, the SHA-256 of its two slot sections (the bytes that crossed the wire), plus a check
of those bytes against the staged single-handle API on the owning rank."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--workload", default="C1")
    ap.add_argument("--transport", default="p2p")
    ap.add_argument("--rounds", type=int, default=3, help="gathers in a row (exercises the double buffer and its credit protocol)")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from jetson_slam_b200 import distributed as jd, frontend, synth
    from jetson_slam_b200.configs import CONFIGS
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = CONFIGS[a.workload]
    mine = jd.shard_pairs(a.pairs, world, rank)
    cap_pairs = jd.local_capacity(a.pairs, world)
    fe = frontend.Frontend(**cfg.extractor_kwargs(), device=local, max_images=2 * cap_pairs)
    stream = torch.cuda.Stream()
    g = jd.Gatherer(fe, cap_pairs, root=0, transport=a.transport)
    digests, per_round = None, []
    for rnd in range(a.rounds):
        # round r processes the batch with seeds shifted by r, so that consecutive gathers carry different bytes
        imgs = np.stack([im for p in mine for im in synth.stereo_pair(cfg.height, cfg.width, 500 + p + 100 * rnd)]) if mine else None
        if mine:
            fe.set_images(imgs, 0, stream)
            fe.extract(0, 2 * len(mine), stream)
            fe.stereo_match(cfg.mb, cfg.mbf, 0, len(mine), stream=stream)
        if len(mine) != cap_pairs:
            raise SystemExit("this worker needs pairs % world == 0")
        g.begin(0, cap_pairs, stream)
        res = g.end()
        # the owner's own view of its pairs through the staged API (what the gathered bytes must equal)
        stream.synchronize()
        own = {}
        for j, p in enumerate(mine):
            kl, dl = fe.get_keypoints(2 * j)
            kr, dr = fe.get_keypoints(2 * j + 1)
            ur, dp, _, _ = fe.get_stereo(j)
            own[p] = hashlib.sha256(b"".join(x.tobytes() for x in (kl, dl, ur, dp, kr, dr))).hexdigest()
        gathered_own = [None] * world
        if world > 1:
            dist.all_gather_object(gathered_own, own)
        else:
            gathered_own = [own]
        if rank == 0:
            regions = g.regions_to_host(res)
            dig = {}
            for r, reg in enumerate(regions):
                u = jd.unpack_region(reg)
                assert u["rank"] == r and u["n_pairs"] == cap_pairs and u["sequence"] == rnd, (u["rank"], u["n_pairs"], u["sequence"])
                for j, p in enumerate(jd.shard_pairs(a.pairs, world, r)):
                    L, R = u["slots"][2 * j], u["slots"][2 * j + 1]
                    dig[p] = hashlib.sha256(b"".join(x.tobytes() for x in (L["kps"], L["desc"], L["u_right"], L["depth"], R["kps"], R["desc"]))).hexdigest()
            want = {}
            for o in gathered_own:
                want.update(o)
            assert dig == want, "gathered bytes differ from the owners' results"
            per_round.append([dig[p] for p in range(a.pairs)])
    if rank == 0:
        json.dump({"world": world, "pairs": a.pairs, "transport": g.transport, "rounds": per_round}, open(a.out, "w"))
    g.close()
    fe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
