"""IDQN_TL=<tag> python tools/timeline.py [K]: one un-graphed step, then print the pipeline timeline of CTA 0."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(0)
obs, A, B = (84, 84, 4), 6, 32
agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4, flags=L.F_NO_GRAPH)
batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
             action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
             is_terminal=(rng.random(B) < 0.1))
for i in range(3):
    agent._engine.learn_host(batch, want_losses=True)
buf = np.zeros(4096, np.uint64)
lib = agent._engine.lib
lib.idqn_debug_timeline.restype = C.c_int
n = lib.idqn_debug_timeline(buf.ctypes.data_as(C.c_void_p), 4096)
ev = sorted(((int(v) >> 16, int(v) & 0xffff) for v in buf[:n] if v))
t0 = ev[0][0] if ev else 0


def name(tag):
    if tag == 1:
        return "start"
    if 1000 <= tag < 1200:
        return f"X  image load issued, unit {tag - 1000}"
    if 1200 <= tag < 2000:
        return f"E  pass {(tag - 1200) // 2} " + ("done" if tag & 1 else "acc_full seen")
    if 2000 <= tag < 3000:
        return f"W  pass {(tag - 2000) // 16} tap {(tag - 2000) % 16} load issued (slot free)"
    k = tag - 3000
    return f"M  pass {k // 32} tap {k % 16} " + ("issued+commit" if k % 32 >= 16 else "w_full seen")


for t, tag in ev:
    print(f"{(t - t0) / 1.965e3:9.2f} us  {name(tag)}")
