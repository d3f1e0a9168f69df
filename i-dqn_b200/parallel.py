"""Head sharding of one i-DQN chain over several B200s — one process per GPU.

The K heads share no parameters and see the same batch, so rank r simply owns a contiguous block of heads and
runs the ordinary single-GPU step on it: there is NO per-step collective.  The only cross-GPU traffic is the
neighbour exchange of one head's parameters (16.2 MB for NatureCNN) at the target events of idqn.py:74-94:

* D-sync  target[k] <- online[k-1]:  rank r receives the LAST online head of rank r-1 into its target[0];
* T-shift online[k] <- online[k+1]:  rank r receives the FIRST online head of rank r+1 (its value before that
  rank's own shift — the read-before-overwrite hazard is avoided by staging the send first) into its last slot.

Transfers are NCCL send/recv pairs over NVLink issued on the learner's own CUDA stream, so they are ordered
with the step's kernels without host synchronisation.  The functions below work on any torch tensors, which is
how the protocol is tested on CPU with the gloo backend (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Tuple

import numpy as np


def head_partition(n_heads_total: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous (start, count) block of heads per rank; the first ``K % world`` ranks hold one more."""
    base, extra = divmod(n_heads_total, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def _neighbours(rank: int, parts: List[Tuple[int, int]]):
    """Previous / next rank that actually owns heads (ranks with an empty block are skipped)."""
    prev = next((r for r in range(rank - 1, -1, -1) if parts[r][1] > 0), None)
    nxt = next((r for r in range(rank + 1, len(parts)) if parts[r][1] > 0), None)
    return prev, nxt


def exchange_for_sync(online, target, rank: int, parts, dist, group=None, local_sync=None) -> None:
    """sync_target_params (idqn.py:20-24) across shards.  ``online``/``target``: [K_local, stride] tensors.
    ``local_sync``: callable that performs the in-shard part ``target[1:] <- online[:-1]`` (the engine's own
    ``sync_target``, which also moves the bf16 planes); default: a tensor copy."""
    import torch

    k_local = parts[rank][1]
    if k_local == 0:
        return
    prev, nxt = _neighbours(rank, parts)
    ops = []
    if nxt is not None:
        ops.append(dist.P2POp(dist.isend, online[k_local - 1], nxt, group))
    recv = None
    if prev is not None:
        recv = torch.empty_like(target[0])
        ops.append(dist.P2POp(dist.irecv, recv, prev, group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if local_sync is not None:
        local_sync()
    elif k_local > 1:
        target[1:].copy_(online[:-1])
    for r in reqs:
        r.wait()
    if recv is not None:
        target[0].copy_(recv)


def exchange_for_shift(online, rank: int, parts, dist, group=None, local_shift=None) -> None:
    """shift_params (idqn.py:13-17) across shards: online[k] <- online[k+1] over the GLOBAL head index.
    ``local_shift``: callable for the in-shard part (the engine's ``shift_params``); default: tensor copies."""
    import torch

    k_local = parts[rank][1]
    if k_local == 0:
        return
    prev, nxt = _neighbours(rank, parts)
    ops = []
    send_buf = None
    if prev is not None:
        send_buf = online[0].clone()  # stage the pre-shift value: the slot is overwritten below
        ops.append(dist.P2POp(dist.isend, send_buf, prev, group))
    recv = None
    if nxt is not None:
        recv = torch.empty_like(online[0])
        ops.append(dist.P2POp(dist.irecv, recv, nxt, group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if local_shift is not None:
        local_shift()
    else:
        for k in range(k_local - 1):
            online[k].copy_(online[k + 1])
    for r in reqs:
        r.wait()
    if recv is not None:
        online[k_local - 1].copy_(recv)


def warm_up_links(like, rank: int, parts, dist, group=None) -> None:
    """Run both neighbour-exchange patterns once on scratch tensors.  NCCL sets up its point-to-point connections
    lazily, per peer and direction, on first use (~100 ms): without this the first D-sync and the first T-shift of a
    run pay for it inside the training loop."""
    import torch

    if parts[rank][1] == 0:
        return
    prev, nxt = _neighbours(rank, parts)
    for send_to, recv_from in ((nxt, prev), (prev, nxt)):  # D-sync direction, then T-shift direction
        ops = []
        if send_to is not None:
            ops.append(dist.P2POp(dist.isend, torch.zeros_like(like), send_to, group))
        if recv_from is not None:
            ops.append(dist.P2POp(dist.irecv, torch.empty_like(like), recv_from, group))
        for r in (dist.batch_isend_irecv(ops) if ops else []):
            r.wait()


class _CudaView:
    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def arena_tensor(engine, which: int):
    """Zero-copy torch view [K_local, stride] of one of the engine's arenas (for NCCL / peer copies)."""
    import torch

    return torch.as_tensor(_CudaView(engine.arena_ptr(which), (engine.K, engine.stride)),
                           device=f"cuda:{engine.device}")


def engine_stream(engine):
    import torch

    return torch.cuda.ExternalStream(int(engine.lib.idqn_stream(engine.h)), device=f"cuda:{engine.device}")


def make_sharded_idqn(key, observation_dim, n_actions, n_networks_total: int, features, architecture_type,
                      learning_rate, gamma, update_horizon, update_to_data, target_update_frequency,
                      target_sync_frequency, adam_eps=1e-8, *, rank: int, world_size: int, device: int,
                      batch_size: int = 32, flags: int = 0, group=None):
    """This rank's shard of a ``n_networks_total``-head i-DQN as an ``iDQN`` whose target events exchange the
    boundary heads with the neighbouring ranks.  Requires an initialised NCCL process group."""
    import torch
    import torch.distributed as dist

    from . import _lib as L
    from . import _prng
    from .networks.idqn import iDQN

    parts = head_partition(n_networks_total, world_size)
    start, k_local = parts[rank]
    if k_local == 0:
        raise ValueError(f"rank {rank} owns no head: use world_size <= n_networks ({n_networks_total})")

    import ctypes as C
    import os

    use_peer = os.environ.get("IDQN_EXCHANGE", "peer") == "peer" and world_size > 1

    class ShardediDQN(iDQN):
        def update_target_params(self, step: int):
            eng = self._engine
            if self._peer is not None:
                # NVLink peer-memory exchange (csrc/peer.cu): everything is enqueued on the learner's stream
                if step % self.target_update_frequency == 0:
                    L.check(eng.lib.idqn_peer_shift_params(self._peer))
                    cumulated = eng.cumulated_losses(reset=True)
                    denom = self.target_update_frequency / self.update_to_data
                    logs = {"loss": np.mean(cumulated) / denom}
                    for i in range(self.n_networks):
                        logs[f"networks/{start + i}_loss"] = cumulated[i] / denom
                    return True, logs
                if step % self.target_sync_frequency == 0:
                    L.check(eng.lib.idqn_peer_sync_target(self._peer))
                return False, {}
            if step % self.target_update_frequency == 0:
                eng.copy_online_to_target()
                with torch.cuda.stream(self._stream):
                    exchange_for_shift(self._online, rank, parts, dist, group, local_shift=eng.shift_params)
                if _neighbours(rank, parts)[1] is not None:  # the last slot was rewritten behind the library's back
                    eng.mark_head_planes_dirty(L.ONLINE, k_local - 1)
                cumulated = eng.cumulated_losses(reset=True)
                denom = self.target_update_frequency / self.update_to_data
                logs = {"loss": np.mean(cumulated) / denom}
                for i in range(self.n_networks):
                    logs[f"networks/{start + i}_loss"] = cumulated[i] / denom
                return True, logs
            if step % self.target_sync_frequency == 0:
                with torch.cuda.stream(self._stream):
                    exchange_for_sync(self._online, self._target, rank, parts, dist, group, local_sync=eng.sync_target)
                if _neighbours(rank, parts)[0] is not None:
                    eng.mark_head_planes_dirty(L.TARGET, 0)
            return False, {}

    torch.cuda.set_device(device)
    # every rank derives the same K_total head keys and keeps its own block
    keys = _prng.split(key, n_networks_total)
    agent = ShardediDQN(0, observation_dim, n_actions, k_local, features, architecture_type, learning_rate, gamma,
                        update_horizon, update_to_data, target_update_frequency, target_sync_frequency, adam_eps,
                        batch_size=batch_size, device=device, flags=flags)
    obs = tuple(int(d) for d in np.atleast_1d(observation_dim))
    heads = [agent.network.init(k, np.zeros(obs, np.float32)) for k in keys[start:start + k_local]]
    from .networks.idqn import _map_stack
    agent._engine.upload_tree(L.ONLINE, _map_stack(heads))
    agent._engine.copy_online_to_target()
    agent._online = arena_tensor(agent._engine, L.ONLINE)
    agent._target = arena_tensor(agent._engine, L.TARGET)
    agent._stream = engine_stream(agent._engine)
    agent.head_offset, agent.n_networks_total = start, n_networks_total
    agent._peer = None
    if use_peer:
        # CUDA-IPC handles of every rank's arenas travel once through the process group; from then on the target events
        # are kernels storing into the neighbour's memory over NVLink
        lib = agent._engine.lib
        blob = C.create_string_buffer(int(lib.idqn_peer_export_size()))
        peer = C.c_void_p()
        L.check(lib.idqn_peer_create(agent._engine.h, C.byref(peer), blob))
        blobs = [None] * world_size
        dist.all_gather_object(blobs, bytes(blob.raw), group=group)
        prev, nxt = _neighbours(rank, parts)
        keep = [C.create_string_buffer(blobs[r], len(blobs[r])) if r is not None else None for r in (prev, nxt)]
        L.check(lib.idqn_peer_connect(peer, keep[0], keep[1]))
        agent._peer = peer
        agent._peer_finalizer = weakref.finalize(agent, lib.idqn_peer_destroy, peer)
        dist.barrier(group=group)  # every rank has mapped its neighbours before the first event
    else:
        with torch.cuda.stream(agent._stream):
            warm_up_links(agent._online[0], rank, parts, dist, group)
    torch.cuda.synchronize()
    return agent
