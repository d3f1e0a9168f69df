"""Head-sharding exchange protocol (idqn_b200.parallel) on CPU: world_size 2 and 3 over gloo, checked against the
single-process oracle shift/sync applied to the global head axis."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import networks as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k_total, stride, events, q):
    from idqn_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = parallel.head_partition(k_total, world)
    start, cnt = parts[rank]
    rng = np.random.default_rng(0)
    online_g = rng.standard_normal((k_total, stride)).astype(np.float32)
    target_g = rng.standard_normal((k_total, stride)).astype(np.float32)
    online = torch.from_numpy(online_g[start:start + cnt].copy())
    target = torch.from_numpy(target_g[start:start + cnt].copy())
    if cnt:
        parallel.warm_up_links(online[0], rank, parts, dist)  # scratch tensors only: must not change any head
    # odd ranks hand the in-shard part to callbacks, the way the CUDA agent hands it to the engine's own
    # shift_params / sync_target (which also move the bf16 planes); even ranks use the default tensor copies
    def local_shift():
        for k in range(cnt - 1):
            online[k].copy_(online[k + 1])

    def local_sync():
        if cnt > 1:
            target[1:].copy_(online[:-1])

    cb = rank % 2 == 1
    for ev in events:
        if ev == "T":
            target.copy_(online)
            parallel.exchange_for_shift(online, rank, parts, dist, local_shift=local_shift if cb else None)
        elif ev == "D":
            parallel.exchange_for_sync(online, target, rank, parts, dist, local_sync=local_sync if cb else None)
        else:  # "grad": heads drift independently
            online += (rank + 1) * 0.5 + torch.arange(cnt, dtype=torch.float32)[:, None]
    q.put((rank, start, online.numpy().copy(), target.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _expected(world, k_total, stride, events):
    from idqn_b200 import parallel
    parts = parallel.head_partition(k_total, world)
    rng = np.random.default_rng(0)
    online = rng.standard_normal((k_total, stride)).astype(np.float32)
    target = rng.standard_normal((k_total, stride)).astype(np.float32)
    tree = lambda a: {"params": {"L": {"w": a}}}
    for ev in events:
        if ev == "T":
            target = online.copy()
            online = O.shift_params(tree(online))["params"]["L"]["w"]
        elif ev == "D":
            target = O.sync_target_params(tree(online), tree(target))["params"]["L"]["w"]
        else:
            for r, (s, c) in enumerate(parts):
                online[s:s + c] += (r + 1) * 0.5 + np.arange(c, dtype=np.float32)[:, None]
    return online, target


@pytest.mark.parametrize("world,k_total", [(2, 5), (2, 2), (3, 4)])
def test_sharded_target_events_match_global_semantics(world, k_total):
    stride = 37
    events = ["grad", "D", "grad", "T", "grad", "D", "D", "T", "grad"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k_total, stride, events, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_online, exp_target = _expected(world, k_total, stride, events)
    for rank, start, online, target in got:
        np.testing.assert_array_equal(online, exp_online[start:start + online.shape[0]])
        np.testing.assert_array_equal(target, exp_target[start:start + target.shape[0]])


def test_head_partition():
    from idqn_b200.parallel import head_partition
    assert head_partition(5, 2) == [(0, 3), (3, 2)]
    assert head_partition(8, 8) == [(i, 1) for i in range(8)]
    assert head_partition(5, 8)[5:] == [(5, 0)] * 3
