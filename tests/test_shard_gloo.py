"""N > 1 path on CPU: two processes over gloo partition one picture's job list (tile columns) and a lookahead's frame
pairs, run the shards (the oracle stands in for the GPU here — this tests the host-side partition / merge logic, the
GPU path is covered by tests/test_gpu_parity.py), gather the MV fields on rank 0 and compare with the unsharded run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SR = 192, 128, 8


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from video_codecs_b200 import HMB200, shard, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        hm = HMB200()                 # host logic only: job lists, tile columns (no GPU call)
        O = Oracle(fen=1, hadme=1)
        lam = 40000
        frames = [synth.luma_frame(W, H, t, seed=9) for t in range(4)]
        pad = [synth.pad_plane(f, 80, 80) for f in frames]
        stride, o0 = pad[0].shape[1], 80 * pad[0].shape[1] + 80

        # --- tile columns of one picture --------------------------------------------------------------------------
        jobs = shard.tile_column_jobs(hm, W, H, world, rank, SR, lam)
        res, _ = O.run_jobs((pad[1], o0, stride), (pad[0], o0, stride), jobs, 8, True)
        got = shard.gather_results(jobs, res, dst=0)
        # --- frame pairs of a lookahead ---------------------------------------------------------------------------
        mine = list(shard.frame_pairs_of_rank(len(frames) - 1, world, rank))
        all_jobs = hm.build_canonical_jobs(W, H, SR, lam)[::5]
        pair_res = {p: O.run_jobs((pad[p + 1], o0, stride), (pad[p], o0, stride), all_jobs, 8, True)[0] for p in mine}
        box = [None] * world if rank == 0 else None
        dist.gather_object(pair_res, box, dst=0)
        if rank == 0:
            full_jobs = hm.build_canonical_jobs(W, H, SR, lam)
            merged = shard.merge_shards(got[0], got[1], full_jobs)
            exp, _ = O.run_jobs((pad[1], o0, stride), (pad[0], o0, stride), full_jobs, 8, True)
            ok_tiles = bool(np.array_equal(merged, exp))
            pairs = {}
            for d in box:
                assert not (set(d) & set(pairs)), "a frame pair was processed twice"
                pairs.update(d)
            ok_pairs = sorted(pairs) == list(range(len(frames) - 1)) and all(
                np.array_equal(pairs[p], O.run_jobs((pad[p + 1], o0, stride), (pad[p], o0, stride), all_jobs, 8, True)[0]) for p in pairs)
            q.put((ok_tiles, ok_pairs, len(full_jobs), [len(j) for j in got[0]]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_columns_and_frame_pairs():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_tiles, ok_pairs, n_full, n_shards = q.get(timeout=300)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok_tiles and ok_pairs
    assert sum(n_shards) == n_full and all(n > 0 for n in n_shards)


def test_partition_helpers():
    sys.path.insert(0, ROOT)
    from video_codecs_b200 import HMB200, shard
    hm = HMB200()
    for world in (1, 2, 3, 4, 8):
        got = [i for r in range(world) for i in shard.frame_pairs_of_rank(255, world, r)]
        assert got == list(range(255))
        cols = [hm.tile_column_range(3840, world, c) for c in range(world)]
        assert cols[0][0] == 0 and cols[-1][1] == 60 and all(cols[i][1] == cols[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in cols) - min(b - a for a, b in cols) <= 1
    full = hm.build_canonical_jobs(416, 240, 64, 7)
    parts = [shard.tile_column_jobs(hm, 416, 240, 3, c, 64, 7) for c in range(3)]
    merged = shard.merge_shards(parts, [np.arange(len(p), dtype=np.int64) + 1000000 * i for i, p in enumerate(parts)], full)
    assert len(merged) == len(full) and len(set(merged.tolist())) == len(full)
    with pytest.raises(ValueError):
        shard.merge_shards(parts[:2], [np.zeros(len(p)) for p in parts[:2]], full)
