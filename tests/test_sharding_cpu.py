"""World-size-2 gloo test (CPU) of the host side of the point-sharded ICP: contiguous sharding,
exchange of opaque handle bytes, and the property the device exchange relies on -- per-rank normal-equation rows added
in rank order reproduce the single-rank sums (checked with the C oracle's align_clouds on each shard)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as g
    g.load_package()
    from importlib import import_module
    sh = import_module("limu_b200.sharding")
    import oracle
    port_api = oracle.load_port()
    # 1. opaque bytes travel unchanged and come back in rank order
    mine = bytes([rank]) * 64
    got = sh.dist_all_gather_bytes(mine)
    ok_gather = got == [bytes([r]) * 64 for r in range(world)]
    idb = sh.dist_broadcast_bytes(b"\x07" * 128 if rank == 0 else None)
    ok_bcast = idb == b"\x07" * 128
    # 2. shards tile the index range
    n = 10007
    lo, hi = sh.shard_range(n, rank, world)
    spans = sh.dist_all_gather_bytes(np.array([lo, hi], np.int64).tobytes())
    spans = [np.frombuffer(b, np.int64) for b in spans]
    ok_tile = spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    # 3. rank-ordered sum of per-shard rows == single-rank normal equations
    rng = np.random.default_rng(3)
    src = rng.normal(size=(n, 3)) * 15
    tgt = src + rng.normal(size=(n, 3)) * 0.05
    r = port_api.align(src[lo:hi], tgt[lo:hi], 0.5)
    row = np.concatenate([r["H"].ravel(), r["g"]])
    rows = [torch.zeros(42, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(rows, torch.from_numpy(row))
    total = sum(x.numpy() for x in rows)            # rank order, like the device fold
    full = port_api.align(src, tgt, 0.5)
    ref = np.concatenate([full["H"].ravel(), full["g"]])
    ok_sum = np.allclose(total, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    q.put((rank, ok_gather, ok_bcast, ok_tile, ok_sum))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert len(res) == 2
    for r in res:
        assert all(r[1:]), r


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from importlib import import_module
    sh = import_module("limu_b200.sharding")
    for n in (0, 1, 7, 8, 4194304, 4194305):
        for w in (1, 2, 4, 8):
            spans = [sh.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
