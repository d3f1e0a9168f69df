"""Where the CTAs of ONE conv kernel of the graph-replayed step spend their time (IDQN_F_TIMELINE): per CTA the global-timer
stamps of kernel entry, first operands landed, last MMA committed and exit, relative to the step's first kernel.

    python tools/cta_timeline.py [K] [slot ...]      (slot = launch index inside the step, see tools/kernel_timeline.py)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
slots = [int(s) for s in sys.argv[2:]] or [2]
rng = np.random.default_rng(0)
obs, A, B = (84, 84, 4), 6, 32
agent = iDQN(0, obs, A, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4, flags=L.F_TIMELINE)
eng = agent._engine
batch = dict(state=rng.integers(0, 256, (B,) + obs).astype(np.uint8), next_state=rng.integers(0, 256, (B,) + obs).astype(np.uint8),
             action=rng.integers(0, A, B).astype(np.int32), reward=rng.integers(-1, 2, B).astype(np.float32),
             is_terminal=(rng.random(B) < 0.1))
kt, names, n = np.zeros(128, np.uint64), np.zeros(64 * 32, np.uint8), C.c_int(0)
for slot in slots:
    for _ in range(5):
        eng.learn_host(batch, want_losses=True)
    L.check(eng.lib.idqn_kernel_timeline(eng.h, L.ptr(kt), L.ptr(names), 64, C.byref(n)))   # reset the kernel slots
    L.check(eng.lib.idqn_cta_timeline(eng.h, slot, None, 0, None))                          # select, clear
    eng.learn_host(batch, want_losses=True)
    L.check(eng.lib.idqn_kernel_timeline(eng.h, L.ptr(kt), L.ptr(names), 64, C.byref(n)))
    out, m = np.zeros(160 * 4, np.uint64), C.c_int(0)
    L.check(eng.lib.idqn_cta_timeline(eng.h, -1, L.ptr(out), 160, C.byref(m)))
    t = out.reshape(-1, 4).astype(np.float64)
    t = t[t[:, 3] > 0]
    if os.environ.get("CTL_ROWS"):
        for b in range(0, len(t), int(os.environ["CTL_ROWS"])):
            print(f"    cta {b:3d}: " + "  ".join(f"{(x - float(kt[2 * slot])) / 1e3:6.2f}" for x in t[b]))
    nm = bytes(names[32 * slot:32 * slot + 32]).split(b"\0")[0].decode()
    k0, k1 = float(kt[2 * slot]), float(kt[2 * slot + 1])
    prev_end = float(kt[2 * (slot - 1) + 1]) if slot > 0 else k0
    r = (t - k0) / 1e3
    print(f"K={K} slot {slot} {nm}: {len(t)} CTAs, kernel span {(k1 - k0) / 1e3:.2f} us, previous kernel ends at {(prev_end - k0) / 1e3:+.2f} us")
    for j, what in enumerate(("entry", "first operands", "last MMA commit", "exit")):
        c = r[:, j]
        print(f"  {what:16s} min {c.min():6.2f}  p10 {np.percentile(c, 10):6.2f}  median {np.median(c):6.2f}  p90 {np.percentile(c, 90):6.2f}  max {c.max():6.2f}")
    d = r[:, 3] - r[:, 0]
    print(f"  CTA lifetime     min {d.min():6.2f}  median {np.median(d):6.2f}  max {d.max():6.2f};  first operands - entry: median {np.median(r[:, 1] - r[:, 0]):5.2f};"
          f"  exit - last commit: median {np.median(r[:, 3] - r[:, 2]):5.2f}")
