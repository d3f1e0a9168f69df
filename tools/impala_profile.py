"""Per-launch CUDA-event times of one impala learning step at the Atari sizes (idqn_profile_step: un-graphed launches)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN
import ctypes as C

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(0)
obs, A, B = (84, 84, 4), 6, 32
agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "impala", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4)
batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
             action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32), is_terminal=(rng.random(B) < 0.1))
eng = agent._engine
for _ in range(3):
    eng.learn_host(batch, want_losses=True)
names_buf, ms_buf, acc, order = np.zeros(128 * 32, np.uint8), np.zeros(128, np.float32), {}, []
reps = 3
for rep in range(reps + 1):
    n = C.c_int(0)
    L.check(eng.lib.idqn_profile_step(eng.h, 1, 128, L.ptr(ms_buf), L.ptr(names_buf), C.byref(n)))
    if rep == 0:
        continue
    for i in range(n.value):
        nm = bytes(names_buf[32 * i:32 * i + 32]).split(b"\0")[0].decode()
        if nm not in acc:
            order.append(nm)
        acc[nm] = acc.get(nm, 0.0) + float(ms_buf[i]) / reps
tot = sum(acc.values())
print(f"impala K={K}: {len(order)} launches, sum {tot:.3f} ms")
for nm in order:
    print(f"  {nm:28s} {acc[nm] * 1e3:9.1f} us  {100 * acc[nm] / tot:5.1f} %")
