"""Resident-path step time and per-kernel profile of the NatureCNN i-DQN step for several head counts K on cuda:0.

    python tools/k_sweep.py [--ks 1,2,3,5,8] [--steps 300] [--flags 0]

One JSON line per K: {"K", "ms_per_step", "steps_per_s", "kernel_us": {...}} (graph replay timed with CUDA events on the
engine's stream; kernel_us from idqn_profile_step = un-graphed, single-branch launches)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN
from idqn_b200.sample_collection.replay_buffer import ReplayBuffer, TransitionElement
from idqn_b200.sample_collection.samplers import UniformSamplingDistribution

OBS, FEATS, A, B = (84, 84, 4), [32, 64, 64, 512], 6, 32


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", default="1,2,3,5,8")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--events", type=int, default=1, help="1: include the T=200 / D=10 target events")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    rb = ReplayBuffer(UniformSamplingDistribution(seed=0), batch_size=B, max_capacity=2048, stack_size=4,
                      clipping=lambda r: np.clip(r, -1, 1), device=0)
    frames = rng.integers(0, 256, (2100, 84, 84), dtype=np.uint8)
    for t in range(2100):
        rb.add(TransitionElement(frames[t], int(rng.integers(0, A)), float(rng.integers(-1, 2)), bool(rng.random() < 0.01), False))
    for K in [int(k) for k in args.ks.split(",")]:
        agent = iDQN(0, OBS, A, K, FEATS, "cnn", 3e-4, 0.99, 1, 1, 200, 10, 1.5e-4, flags=args.flags)
        eng = agent._engine
        stream = torch.cuda.ExternalStream(int(eng.lib.idqn_stream(eng.h)), device="cuda:0")
        step = 1
        for _ in range(20):
            agent.update_online_params(step, rb)
            if args.events:
                agent.update_target_params(step)
            step += 1
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            agent.update_online_params(step, rb)
            if args.events:
                agent.update_target_params(step)
            step += 1
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        names_buf, ms_buf, acc, reps = np.zeros(64 * 32, np.uint8), np.zeros(64, np.float32), {}, 5
        for rep in range(reps + 1):
            n = C.c_int(0)
            L.check(eng.lib.idqn_profile_step(eng.h, 1, 64, L.ptr(ms_buf), L.ptr(names_buf), C.byref(n)))
            if rep == 0:
                continue
            for i in range(n.value):
                nm = bytes(names_buf[32 * i:32 * i + 32]).split(b"\0")[0].decode()
                acc[nm] = acc.get(nm, 0.0) + float(ms_buf[i]) * 1e3 / reps
        print(json.dumps({"K": K, "ms_per_step": round(ms, 5), "steps_per_s": round(1e3 / ms, 1), "flags": args.flags,
                          "kernel_us": {k: round(v, 1) for k, v in acc.items()}, "sum_kernel_us": round(sum(acc.values()), 1)}),
              flush=True)
        eng.close()
        del agent, eng


if __name__ == "__main__":
    main()
