"""N>1 host logic on CPU: world_size-2 gloo processes agree on the sharded-argmax winner."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, D, seed, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bore_b200 import distributed as bd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    bd.init_process_group(backend="gloo")
    rs = np.random.RandomState(seed)
    fun = rs.normal(size=total).astype(np.float32).astype(np.float64)
    fun[5] = fun[total - 3] = fun.min() - 1.0           # a tie across shards -> lowest index wins
    status = rs.choice([0, 1, 2], size=total, p=[0.6, 0.1, 0.3]).astype(np.int32)
    status[5] = status[total - 3] = 0
    X = rs.uniform(size=(total, D))
    lo, hi = bd.shard_bounds(total, rank, world)
    key = torch.tensor([bd.pack_key_numpy(fun[lo:hi], status[lo:hi], idx_offset=lo)], dtype=torch.int64)
    rec_fn = lambda i: torch.from_numpy(np.concatenate([X[lo + i], [fun[lo + i]]]))
    gidx, rec = bd.global_winner(key, rec_fn, total, D + 1)
    ok = (status == 0) | (status == 1)
    want = int(np.flatnonzero(ok)[np.argmin(fun[ok])])  # first minimum among qualifying
    assert gidx == want == 5, (gidx, want)
    assert np.array_equal(rec.numpy()[:D], X[want]) and rec.numpy()[D] == fun[want]
    # nobody qualifies anywhere -> None on every rank
    key0 = torch.tensor([bd.pack_key_numpy(fun[lo:hi], np.full(hi - lo, 2), idx_offset=lo)], dtype=torch.int64)
    assert bd.global_winner(key0, rec_fn, total, D + 1) == (None, None)
    dist.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1")
    dist.destroy_process_group()


def test_two_rank_gloo_winner(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 101, 4, 3, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_helpers_and_key_order():
    from bore_b200 import distributed as bd
    for total, world in ((65536, 8), (101, 2), (7, 4), (3, 8)):
        spans = [bd.shard_bounds(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        for g in range(total):
            r, i = bd.owner_of(g, total, world)
            assert spans[r][0] + i == g
    f = np.array([0.5, -1.25, -1.25, 3.0, -0.0, 0.0])
    st = np.zeros(6, np.int32)
    assert bd.decode_key(bd.pack_key_numpy(f, st)) == 1          # first minimum on ties
    assert bd.decode_key(bd.pack_key_numpy(f, st, keep=[1, 0, 1, 1, 1, 1])) == 2
    assert bd.decode_key(bd.pack_key_numpy(f[4:], st[4:], idx_offset=10)) == 10  # -0 == +0
    st2 = np.array([2, 2, 2, 1, 2, 2], np.int32)
    assert bd.decode_key(bd.pack_key_numpy(f, st2)) == 3          # status 1 qualifies
    assert bd.decode_key(bd.pack_key_numpy(f, np.full(6, 2))) is None
    # monotone: smaller fun <=> larger key
    vals = np.sort(np.random.RandomState(0).normal(size=50))
    keys = [bd.pack_key_numpy([v], [0], idx_offset=7) for v in vals]
    assert all(a > b for a, b in zip(keys, keys[1:]))
