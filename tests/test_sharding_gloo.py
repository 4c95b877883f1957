"""N > 1 host logic on CPU: frame sharding and the max-over-ranks timing reduction with a world_size-2 gloo group
(frames shard with no data-path collective; only a barrier and a MAX reduction exist)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import corpus
    from swf_renderer_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a 6-frame morph sweep sharded over 2 ranks; each rank renders its frames with the oracle (no GPU here):
    ratios = [0, 13107, 26214, 39321, 52428, 65535]
    mine = sharding.frames_of_rank(len(ratios), rank, world)
    sc = corpus.morph_scene(ratios)
    sums = {f: int(corpus.render_oracle(sc, frame=f).astype(np.int64).sum()) for f in mine}
    ms = sharding.reduce_max_time(10.0 * (rank + 1))
    dist.barrier()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([[f, s] for f, s in sums.items()] + [[-1, int(ms)]]))
    dist.destroy_process_group()


def test_frame_sharding_is_a_partition():
    from swf_renderer_b200 import sharding

    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            fr = sharding.frames_of_rank(67, r, world)
            assert all(sharding.owner_of_frame(f, world) == (r, i) for i, f in enumerate(fr))
            seen += fr
        assert sorted(seen) == list(range(67))
    assert sharding.total_throughput([100, 100], 50.0) == 4000.0
    with pytest.raises(ValueError):
        sharding.frames_of_rank(4, 2, 2)


def test_two_rank_gloo_sharded_sweep(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import corpus

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = {}
    for r in range(world):
        arr = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert arr[-1, 1] == 20  # MAX over ranks of (10, 20)
        for f, s in arr[:-1]:
            assert f % world == r
            got[int(f)] = int(s)
    ratios = [0, 13107, 26214, 39321, 52428, 65535]
    sc = corpus.morph_scene(ratios)
    assert sorted(got) == list(range(6))
    for f in range(6):
        assert got[f] == int(corpus.render_oracle(sc, frame=f).astype(np.int64).sum())


def _gather_worker(rank, world, port, out_dir, n_frames):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from swf_renderer_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.frames_of_rank(n_frames, rank, world)
    # frame f is filled with the value f (3 x 5 "pixels")
    local = torch.stack([torch.full((3, 5, 4), f, dtype=torch.uint8) for f in mine]) if mine else torch.zeros((0, 3, 5, 4), dtype=torch.uint8)
    out = sharding.gather_frames(local, n_frames, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), out.numpy())
    else:
        assert out is None
    if rank == 1:  # a rank that holds the wrong number of frames is an error, not a silent misplacement
        wrong = torch.zeros((len(mine) + 1, 3, 5, 4), dtype=torch.uint8)
        with pytest.raises(ValueError):
            sharding.gather_frames(wrong, n_frames, dst=0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [6, 7, 1])
def test_two_rank_gloo_frame_gather_restores_frame_order(tmp_path, n_frames):
    world, port = 2, _free_port()
    mp.spawn(_gather_worker, args=(world, port, str(tmp_path), n_frames), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    assert got.shape == (n_frames, 3, 5, 4)
    for f in range(n_frames):
        assert (got[f] == f).all()


def test_frame_gather_single_process_is_the_identity():
    import torch

    from swf_renderer_b200 import sharding

    x = torch.arange(2 * 3 * 5 * 4, dtype=torch.uint8).reshape(2, 3, 5, 4)
    assert sharding.gather_frames(x, 2) is x
    with pytest.raises(ValueError):
        sharding.gather_frames(x, 3)
