"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch of independent
codewords with no data-path collective, decode their shards (the CPU oracle stands in for
the per-rank decoder -- no GPU here) and sum only the counters.  The union of the shards
must equal the single-process result frame for frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from labrador_ldpc_b200.sharding import DecodeStats, shard_bounds, shard_view

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_tile_exactly():
    for batch in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(batch, world, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == batch
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, llrs, tmpdir):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = pyoracle.Oracle()
    mine = shard_view(llrs, world, rank)
    out, ok, iters = oracle.decode_ms_batch(5, mine, 40)
    stats = DecodeStats().add(ok, iters).all_reduce(dist)
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), out=out, ok=ok, iters=iters,
             stats=stats.to_vector(), bounds=np.array(shard_bounds(len(llrs), world, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_decode(oracle, tmp_path):
    from frames import make_frames
    world = 2
    _, _, llrs = make_frames(oracle, 5, 37, 2.2, seed=77, ty="i8")     # odd batch: ragged shards
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, llrs, str(tmp_path)), nprocs=world, join=True)
    want = oracle.decode_ms_batch(5, llrs, 40)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert parts[0]["bounds"][1] == parts[1]["bounds"][0] and parts[1]["bounds"][1] == 37
    assert np.array_equal(np.concatenate([p["out"] for p in parts]), want[0])
    assert np.array_equal(np.concatenate([p["ok"] for p in parts]), want[1])
    assert np.array_equal(np.concatenate([p["iters"] for p in parts]), want[2])
    total = DecodeStats().add(want[1], want[2])
    for p in parts:     # every rank holds the same reduced counters
        assert np.array_equal(p["stats"], total.to_vector())
    assert abs(total.fer - (1 - want[1].mean())) < 1e-12
