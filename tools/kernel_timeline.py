"""True in-graph schedule of one learning step (IDQN_F_TIMELINE): start of the first CTA / end of the last CTA of every
kernel, from the global timer, relative to the step's first kernel.

    python tools/kernel_timeline.py [K] [steps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
flags = (int(sys.argv[3]) if len(sys.argv) > 3 else 0) | L.F_TIMELINE
rng = np.random.default_rng(0)
obs, A, B = (84, 84, 4), 6, 32
agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4, flags=flags)
eng = agent._engine
batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
             action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
             is_terminal=(rng.random(B) < 0.1))
out, names, n = np.zeros(128, np.uint64), np.zeros(64 * 32, np.uint8), C.c_int(0)
for i in range(steps):
    # two steps back to back, the timeline of the SECOND one is kept (steady state: its first kernel follows a step)
    eng.learn_host(batch, want_losses=False)
    L.check(eng.lib.idqn_kernel_timeline(eng.h, L.ptr(out), L.ptr(names), 64, C.byref(n)))
    eng.learn_host(batch, want_losses=False)
    eng.learn_host(batch, want_losses=True)
    L.check(eng.lib.idqn_kernel_timeline(eng.h, L.ptr(out), L.ptr(names), 64, C.byref(n)))
# NOTE: the slots hold min(begin) / max(end) over the steps since the last read: read after ONE step only
eng.learn_host(batch, want_losses=True)
L.check(eng.lib.idqn_kernel_timeline(eng.h, L.ptr(out), L.ptr(names), 64, C.byref(n)))
t0 = min(int(out[2 * i]) for i in range(n.value))
print(f"K={K} flags={flags}: kernel  start_us  end_us  dur_us   (gap to the previous end)")
prev_end = None
for i in range(n.value):
    nm = bytes(names[32 * i:32 * i + 32]).split(b"\0")[0].decode()
    b, e = (int(out[2 * i]) - t0) / 1e3, (int(out[2 * i + 1]) - t0) / 1e3
    gap = "" if prev_end is None else f"{b - prev_end:+7.2f}"
    print(f"{i:2d} {nm:22s} {b:8.2f} {e:8.2f} {e - b:7.2f}  {gap}")
    prev_end = e if prev_end is None else max(prev_end, e)
print(f"step span {max((int(out[2 * i + 1]) - t0) for i in range(n.value)) / 1e3:.2f} us")
