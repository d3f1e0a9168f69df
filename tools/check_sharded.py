"""Multi-GPU check (torchrun, one rank per GPU): a K_total-head chain sharded over the ranks follows the same
trajectory as the unsharded agent on one GPU -- gradient steps, D-syncs (neighbour exchange of the boundary head) and
T-shifts included.  Prints max relative differences per rank; exits non-zero on mismatch."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN, _map_stack
from idqn_b200.parallel import make_sharded_idqn, head_partition
from idqn_b200 import _prng

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
obs, feats, A, B = (84, 84, 4), [32, 64, 64, 512], 6, 32
k_total = 2 * world
T, D, steps = 8, 4, 26
sh = make_sharded_idqn(0, obs, A, k_total, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4, rank=rank, world_size=world, device=local)
full = iDQN(0, obs, A, k_total, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4, device=local)
keys = _prng.split(0, k_total)
full._engine.upload_tree(L.ONLINE, _map_stack([full.network.init(k, np.zeros(obs, np.float32)) for k in keys]))
full._engine.copy_online_to_target()
start, cnt = head_partition(k_total, world)[rank]
rng = np.random.default_rng(5)
worst = 0.0
for step in range(1, steps + 1):
    batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
                 action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
                 is_terminal=(rng.random(B) < 0.1))
    ls = sh._engine.learn_host(batch, want_losses=True)
    lf = full._engine.learn_host(batch, want_losses=True)
    sh.update_target_params(step)
    full.update_target_params(step)
    worst = max(worst, float(np.max(np.abs(ls - lf[start:start + cnt]) / np.maximum(np.abs(lf[start:start + cnt]), 1e-12))))
for which, name in ((L.ONLINE, "online"), (L.TARGET, "target")):
    a = sh._engine.download_tree(which)
    b = full._engine.download_tree(which)
    for m in a["params"]:
        for kind in a["params"][m]:
            x, y = np.asarray(a["params"][m][kind], np.float64), np.asarray(b["params"][m][kind], np.float64)[start:start + cnt]
            worst = max(worst, float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-30)))
print(f"rank {rank}: heads [{start},{start + cnt}) of {k_total}, {steps} steps (T={T}, D={D}): max relative difference {worst:.3e}", flush=True)
ok = torch.tensor([1 if worst < 1e-4 else 0], device=f"cuda:{local}")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(ok.item()) == 1 else 1)
