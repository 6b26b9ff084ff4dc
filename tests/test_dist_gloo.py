"""N > 1 host logic on CPU: two gloo ranks shard images / windows and rank 0 reassembles reference order.
(The compute inside each rank is a stand-in -- the oracle flow -- because this container has no GPU.)"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import nodes as onodes
    from pyfaceanalysis_b200 import shard, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- batch mode: images round robin, per-image detection lists gathered on the host
        n_images = 7
        mine = shard.image_shard(n_images, rank, world)
        local = [np.full((k % 3, 10), float(k)) for k in mine]           # "detections" of image k
        allr = shard.gather_detections(local, mine, n_images)
        assert [len(a) for a in allr] == [k % 3 for k in range(n_images)]
        assert all((a == k).all() for k, a in enumerate(allr))
        # ---- one big window batch: contiguous tile-aligned ranges, concatenation == unsharded result
        flow = synthetic.make_flow("tiny", seed=3)
        x = synthetic.synthetic_patches(700, (16, 16), 1).astype(np.float64)
        a, b = shard.window_shard(len(x), rank, world)
        assert a % 128 == 0
        part = onodes.flow_execute(flow, x[a:b]) if b > a else np.zeros((0, 16))
        parts = [None] * world
        dist.all_gather_object(parts, (a, b, part))
        if rank == 0:
            parts.sort(key=lambda t: t[0])
            assert parts[0][0] == 0 and parts[-1][1] == len(x) and all(p[1] == q[0] for p, q in zip(parts, parts[1:]))
            full = np.concatenate([p[2] for p in parts])
            assert np.array_equal(full, onodes.flow_execute(flow, x))
            open(os.path.join(out_dir, "ok"), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_shard_arithmetic():
    from pyfaceanalysis_b200 import shard
    for n in (0, 1, 127, 128, 129, 1000, 1 << 20):
        for world in (1, 2, 4, 8):
            spans = [shard.window_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(a % 128 == 0 for a, _ in spans if a < n)
    assert shard.image_shard(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard.image_shard(10, r, 4) for r in range(4)), [])) == list(range(10))
    assert shard.gather_detections(["b", "a"], [1, 0], 2) == ["a", "b"]
    with pytest.raises(RuntimeError):
        shard.gather_detections(["a"], [0], 2)
