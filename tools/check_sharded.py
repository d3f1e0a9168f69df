"""Multi-GPU check (torchrun, one rank per GPU) of the head-sharded agent (idqn_b200/parallel.py).

Every rank runs two copies of its shard of a K_total-head chain on the same batches:
  * `sh`  -- the product path: make_sharded_idqn, NCCL send/recv of the boundary head at the D / T events, in-shard part
             through the engine, planes of the received head rebuilt alone;
  * `ref` -- an independent statement of the same events: a plain K_local-head agent whose boundary heads travel through
             an all_gather of every rank's first / last online head, in-shard part with tensor copies on the arena views,
             all planes rebuilt.
Both have the same K_local, hence the same kernels and summation orders: they must stay BIT-IDENTICAL.  (The unsharded
K_total-head agent is only a loose reference: its reduction groupings depend on K, the 1e-8 rounding differences are
amplified by Adam's early steps to 1e-3 within 25 steps -- printed for information.)
Exits non-zero on any difference between `sh` and `ref`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200 import _lib as L
from idqn_b200 import _prng
from idqn_b200.networks.idqn import iDQN, _map_stack
from idqn_b200.parallel import arena_tensor, engine_stream, head_partition, make_sharded_idqn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
obs, feats, A, B = (84, 84, 4), [32, 64, 64, 512], 6, 32
k_total = 2 * world
T, D, steps = 8, 4, 26
start, cnt = head_partition(k_total, world)[rank]
keys = _prng.split(0, k_total)

sh = make_sharded_idqn(0, obs, A, k_total, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4, rank=rank, world_size=world, device=local)

ref = iDQN(0, obs, A, cnt, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4, device=local)
ref._engine.upload_tree(L.ONLINE, _map_stack([ref.network.init(k, np.zeros(obs, np.float32)) for k in keys[start:start + cnt]]))
ref._engine.copy_online_to_target()
r_on, r_tg, r_stream = arena_tensor(ref._engine, L.ONLINE), arena_tensor(ref._engine, L.TARGET), engine_stream(ref._engine)

full = iDQN(0, obs, A, k_total, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4, device=local)
full._engine.upload_tree(L.ONLINE, _map_stack([full.network.init(k, np.zeros(obs, np.float32)) for k in keys]))
full._engine.copy_online_to_target()


def ref_events(step):
    """idqn.py:74-94 on the shard, boundary heads through all_gather."""
    eng = ref._engine
    with torch.cuda.stream(r_stream):
        if step % T == 0:
            firsts = [torch.empty_like(r_on[0]) for _ in range(world)]
            dist.all_gather(firsts, r_on[0].clone())  # pre-shift first online head of every rank
            r_tg.copy_(r_on)
            for k in range(cnt - 1):
                r_on[k].copy_(r_on[k + 1])
            if rank + 1 < world:
                r_on[cnt - 1].copy_(firsts[rank + 1])
            eng.mark_planes_dirty(L.ONLINE), eng.mark_planes_dirty(L.TARGET)
            eng.cumulated_losses(reset=True)
        elif step % D == 0:
            lasts = [torch.empty_like(r_on[0]) for _ in range(world)]
            dist.all_gather(lasts, r_on[cnt - 1].clone())
            if cnt > 1:
                r_tg[1:].copy_(r_on[:-1])
            if rank > 0:
                r_tg[0].copy_(lasts[rank - 1])
            eng.mark_planes_dirty(L.TARGET)
    torch.cuda.synchronize()


rng = np.random.default_rng(5)
exact_bad, loose = 0, 0.0
for step in range(1, steps + 1):
    batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
                 action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
                 is_terminal=(rng.random(B) < 0.1))
    ls = sh._engine.learn_host(batch, want_losses=True)
    lr_ = ref._engine.learn_host(batch, want_losses=True)
    lf = full._engine.learn_host(batch, want_losses=True)
    sh.update_target_params(step)
    ref_events(step)
    full.update_target_params(step)
    if not np.array_equal(ls, lr_):
        exact_bad += 1
        print(f"rank {rank} step {step}: losses differ from the all_gather reference: {ls} vs {lr_}", flush=True)
    loose = max(loose, float(np.max(np.abs(ls - lf[start:start + cnt]) / np.maximum(np.abs(lf[start:start + cnt]), 1e-12))))
for which, name in ((L.ONLINE, "online"), (L.TARGET, "target"), (L.MU, "mu"), (L.NU, "nu")):
    a, b = sh._engine.download_tree(which), ref._engine.download_tree(which)
    for m in a["params"]:
        for kind in a["params"][m]:
            if not np.array_equal(np.asarray(a["params"][m][kind]), np.asarray(b["params"][m][kind])):
                exact_bad += 1
                print(f"rank {rank}: {name} {m}/{kind} differs from the all_gather reference", flush=True)
print(f"rank {rank}: heads [{start},{start + cnt}) of {k_total}, {steps} steps (T={T}, D={D}): "
      f"{'bit-identical to' if exact_bad == 0 else 'DIFFERENT from'} the all_gather reference; "
      f"loss vs the unsharded K={k_total} agent within {loose:.1e} relative", flush=True)
ok = torch.tensor([1 if exact_bad == 0 else 0], device=f"cuda:{local}")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
