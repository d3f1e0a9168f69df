#!/usr/bin/env python
"""Benchmark of the i-DQN learning step (BASELINE.json metric: gradient steps/s, NatureCNN K=5, batch 32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Ours, N GPUs: one process per GPU (torchrun for N>1); each rank owns 5 heads of one 5N-head i-DQN chain
(weak scaling, head-sharded, the only collective is the neighbour target exchange every D / T steps).
`value`  : steps/s with the replay store resident in HBM (sampler indices are the only H2D traffic);
`e2e`    : steps/s through the host-buffer C-ABI call (H2D of the batch + D2H of the losses every step).
Reference: the CPU oracle restatement of the reference's learn_on_batch (jax is not installable here) on the
host cores of the box (`--impl reference`, and the `cpu_baseline` leg of the default run)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBS, FEATS, A, B, HEADS_PER_GPU = (84, 84, 4), [32, 64, 64, 512], 6, 32, 5
LR, GAMMA, EPS, T_FREQ, D_FREQ = 3e-4, 0.99, 1.5e-4, 200, 10
P_HEAD = 4_046_502
F_MAC, DG_MAC, WG_MAC = 16_006_144, 12_393_472, 16_006_144
FLOP_PER_HEAD_STEP = B * 2 * (2 * F_MAC + DG_MAC + WG_MAC)
BYTES_PER_HEAD_STEP = 7 * P_HEAD * 4
BATCH_BYTES = B * (2 * 28224 + 9)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf_sus=float(p["bf16_tflops_sustained"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (pynvml, 50 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "sw_power_cap": 0x4, "hw_power_brake": 0x80}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def synthetic_batch(rng):
    return dict(state=rng.integers(0, 256, (B,) + OBS, dtype=np.uint8),
                next_state=rng.integers(0, 256, (B,) + OBS, dtype=np.uint8),
                action=rng.integers(0, A, B).astype(np.int32),
                reward=rng.integers(-1, 2, B).astype(np.float32), is_terminal=(rng.random(B) < 0.1))


# ------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_heads: int, min_seconds: float, max_steps: int, warmup: int = 1):
    """steps/s of the CPU oracle restatement of learn_on_batch (fp32, all host threads)."""
    import torch
    from oracle import networks as O

    rng = np.random.default_rng(0)
    params = O.init_params(rng, OBS, FEATS, "cnn", A, n_networks=n_heads)
    target = O.init_params(np.random.default_rng(1000), OBS, FEATS, "cnn", A, n_networks=n_heads)
    learner = O.CpuLearner(params, target, "cnn", GAMMA, 1, LR, EPS)
    batch = synthetic_batch(rng)
    for _ in range(warmup):
        learner.step(batch)
    times = []
    t_start = time.perf_counter()
    # min_seconds > 0: a time-bounded sample (cpu_baseline leg); otherwise exactly max_steps steps (reference arm)
    while len(times) < max_steps and (min_seconds <= 0 or time.perf_counter() - t_start < min_seconds or len(times) < 3):
        t0 = time.perf_counter()
        learner.step(batch)
        times.append(time.perf_counter() - t0)
    return 1.0 / float(np.median(times)), len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_heads = HEADS_PER_GPU * args.gpus
    steps = max(1, min(args.steps, 40))  # bounded sample: at most 40 CPU steps (~1.5-4 s each at N = 1..8)
    warm = min(max(args.warmup, 1), 3)
    rate, n, threads = cpu_reference_rate(n_heads, min_seconds=0.0, max_steps=steps, warmup=warm)
    value = rate * args.gpus  # same head-normalised unit as our arm (steps/s of a K=5 agent)
    print(json.dumps({
        "impl": "reference", "metric": "i-DQN grad steps/sec (NatureCNN K=5, batch 32)", "value": value,
        "unit": "steps/s", "n_gpus": args.gpus, "steps": n, "warmup": warm, "ms_per_step": 1e3 / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Atari NatureCNN i-DQN, {HEADS_PER_GPU} heads per GPU ({n_heads} total), batch 32, "
                               "84x84x4 uint8, A=6", "heads_total": n_heads},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n} full learn_on_batch steps of the torch-CPU fp32 oracle (jax not installable)"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from idqn_b200 import _lib as L
    from idqn_b200.networks.idqn import iDQN
    from idqn_b200.sample_collection.replay_buffer import ReplayBuffer, TransitionElement
    from idqn_b200.sample_collection.samplers import UniformSamplingDistribution

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    k_total = HEADS_PER_GPU * world
    if world > 1:
        from idqn_b200.parallel import make_sharded_idqn
        agent = make_sharded_idqn(0, OBS, A, k_total, FEATS, "cnn", LR, GAMMA, 1, 1, T_FREQ, D_FREQ, EPS, rank=rank,
                                  world_size=world, device=local)
    else:
        agent = iDQN(0, OBS, A, k_total, FEATS, "cnn", LR, GAMMA, 1, 1, T_FREQ, D_FREQ, EPS, device=local, flags=args.flags)
    eng = agent._engine
    # independent target draw so theta_bar != theta
    tgt = iDQN.__new__(iDQN)
    from idqn_b200.networks.architectures.dqn import DQNNet
    net = DQNNet(FEATS, "cnn", A)
    from idqn_b200.networks.idqn import _map_stack
    eng.upload_tree(L.TARGET, _map_stack([net.init(1000 + rank * 100 + k, np.zeros(OBS, np.float32))
                                          for k in range(eng.K)]))
    del tgt

    # device-resident replay: 4096 synthetic stacked elements = 231 MB (> 126 MB L2), same content on every rank
    rb = ReplayBuffer(UniformSamplingDistribution(seed=0), batch_size=B, max_capacity=4096, stack_size=4,
                      clipping=lambda r: np.clip(r, -1, 1), device=local)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (4200, 84, 84), dtype=np.uint8)
    for t in range(4200):
        rb.add(TransitionElement(frames[t], int(rng.integers(0, A)), float(rng.integers(-1, 2)),
                                 bool(rng.random() < 0.01), False))
    stream = torch.cuda.ExternalStream(int(eng.lib.idqn_stream(eng.h)), device=f"cuda:{local}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, first_step, drain=None):
        step = first_step
        for _ in range(warmup):
            fn(step)
            step += 1
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        sampler.start()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn(step)
            step += 1
        if drain:
            drain()  # the last step's losses are read inside the timed region
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        sampler.stop_flag = True
        sampler.join()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, sampler.summary(), step

    # (1) resident path: sampler -> device gather -> step (+ the T/D target events of the schedule)
    def resident_step(step):
        agent.update_online_params(step, rb)
        agent.update_target_params(step)

    ms, wall, clocks, nxt = timed(resident_step, args.steps, args.warmup, 1)
    steps_per_s = args.steps / (ms / 1e3)

    # (2) end to end through the host-buffer C-ABI call: pinned host batch -> H2D -> step -> D2H losses
    pool = []
    for i in range(8):
        b = synthetic_batch(np.random.default_rng(100 + i))
        pinned = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.uint8) if v.dtype == bool else v)).pin_memory()
                  for k, v in b.items()}
        pool.append({k: t.numpy() for k, t in pinned.items()} | {"_keep": pinned})

    # (2a) pipelined the way the reference's own call is (async dispatch): submit enqueues the pinned H2D copies of
    # this step on the copy stream + the step, the losses of the previous step are read back (D2H) right after
    pending = [None]
    loss_log = []

    def e2e_step(step):
        b = pool[step % len(pool)]
        t = eng.submit_host(b)
        if pending[0] is not None:
            loss_log.append(eng.wait_losses(pending[0]))
        pending[0] = t
        agent.update_target_params(step)

    def e2e_drain():
        if pending[0] is not None:
            loss_log.append(eng.wait_losses(pending[0]))
            pending[0] = None

    e2e_steps = max(args.steps // 2, 5)
    ms_e2e, _, _, nxt = timed(e2e_step, e2e_steps, max(3, args.warmup // 2), nxt, drain=e2e_drain)
    e2e_rate = e2e_steps / (ms_e2e / 1e3)
    assert len(loss_log) >= e2e_steps and all(np.isfinite(l).all() for l in loss_log)

    # (2b) the blocking call: H2D, step and D2H of the losses strictly one after the other
    def e2e_sync_step(step):
        eng.learn_host(pool[step % len(pool)], want_losses=True)
        agent.update_target_params(step)

    ms_sync, _, _, nxt = timed(e2e_sync_step, e2e_steps, 3, nxt)
    e2e_sync_rate = e2e_steps / (ms_sync / 1e3)

    # (2c) configs[4]: the same resident step fed by the prioritised sampler -- 1 M-leaf float64 SumTree on the device
    # (inverse-CDF descent per sample), keys mapped to replay slots on the host
    prio = None
    if world == 1 and not args.no_prioritized:
        from idqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution
        rb_p = ReplayBuffer(PrioritizedSamplingDistribution(seed=0, max_capacity=1 << 20, priority_exponent=1.0, device=local),
                            batch_size=B, max_capacity=4096, stack_size=4, clipping=lambda r: np.clip(r, -1, 1), device=local)
        rng_p = np.random.default_rng(1)
        for t in range(4200):
            rb_p.add(TransitionElement(frames[t], int(rng_p.integers(0, A)), float(rng_p.integers(-1, 2)),
                                       bool(rng_p.random() < 0.01), False), priority=float(rng_p.random() + 0.1))

        def prio_step(step):
            agent.update_online_params(step, rb_p)
            agent.update_target_params(step)

        p_steps = max(args.steps // 2, 5)
        ms_p, _, _, nxt = timed(prio_step, p_steps, 3, nxt)
        prio = {"value": p_steps / (ms_p / 1e3), "unit": "steps/s", "sum_tree_capacity": 1 << 20,
                "how": "resident step with PrioritizedSamplingDistribution: host PCG64 uniforms -> device SumTree descent "
                       "(float64, depth 21) -> keys -> replay slots; one D2H of the 32 leaves per step"}

    # (3) live per-kernel timing (CUDA events after every launch, un-graphed) for the roofline of the top kernel
    names_buf = (np.zeros(64 * 32, np.uint8))
    ms_buf = np.zeros(64, np.float32)
    import ctypes as C
    acc = {}
    reps = 5
    for rep in range(reps + 1):
        n = C.c_int(0)
        L.check(eng.lib.idqn_profile_step(eng.h, 1, 64, L.ptr(ms_buf), L.ptr(names_buf), C.byref(n)))
        if rep == 0:
            continue
        for i in range(n.value):
            nm = bytes(names_buf[32 * i:32 * i + 32]).split(b"\0")[0].decode()
            acc[nm] = acc.get(nm, 0.0) + float(ms_buf[i]) / reps
    kernels_per_step = int(eng.lib.idqn_kernels_per_step(eng.h))  # the replay gather is part of the step's first kernel
    top = max(acc, key=acc.get)
    pk = peaks()
    rl = kernel_roofline(top, acc[top], eng.K, pk)
    rl["traffic"] = None
    try:  # per-launch DRAM bytes of this kernel from the committed ncu --set full capture (K = 5 per GPU)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(top)
        if tr and eng.K == HEADS_PER_GPU:
            rl["traffic"] = tr
            rl["algorithmic_bytes"] = int(round(rl["achieved"] * acc[top] * 1e6)) if rl.get("achieved") else None
    except Exception:
        pass
    rl["kernel"] = top
    rl["peak_source"] = pk["src"]

    if rank == 0:
        cpu_rate, cpu_n, cpu_threads = (None, 0, 0)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_rate, cpu_n, cpu_threads = cpu_reference_rate(k_total, min_seconds=12.0, max_steps=60)
            cpu = {"value": cpu_rate, "unit": "steps/s", "cores": cpu_threads, "kind": "port",
                   "sample": f"{cpu_n} full learn_on_batch steps (K={k_total}, B=32) of the torch-CPU fp32 oracle, median"}
        value = steps_per_s * world  # head-normalised: each rank processes one K=5 agent's worth of heads per step
        out = {
            "metric": "i-DQN grad steps/sec (NatureCNN K=5, batch 32)", "value": value, "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"precision": "fp32 parameters/activations/accumulation; GEMM products as bf16x3 split on tcgen05 "
                                    "(fp32-faithful, parity 1e-4)",
                       "workload": f"Atari NatureCNN i-DQN, {HEADS_PER_GPU} heads per GPU ({k_total} total, head-sharded), "
                                   "batch 32, 84x84x4 uint8, A=6, T=200, D=10, uniform replay resident in HBM",
                       "heads_total": k_total, "unit_note": "value = chain steps/s x (heads_total / 5)",
                       "l2": "working set per step (5 arenas x 81 MB + 231 MB replay) exceeds the 126 MB L2; no flush"},
            "frames_per_s": value * B * 8, "transitions_per_s": value * B,
            "step_hbm_frac": (BYTES_PER_HEAD_STEP * eng.K + BATCH_BYTES) / (ms / args.steps * 1e-3) / 1e9 / pk["hbm"],
            "step_tflops": FLOP_PER_HEAD_STEP * eng.K / (ms / args.steps * 1e-3) / 1e12,
            "wall_s": wall, "clocks": clocks,
            "e2e": {"value": e2e_rate * world, "unit": "steps/s", "h2d_bytes_per_step": BATCH_BYTES,
                    "d2h_bytes_per_step": 4 * eng.K,
                    "how": "idqn_submit_batch_host / idqn_wait_losses: pinned H2D of step t+1 on a copy stream while "
                           "step t computes, losses of every step read back one step behind (two steps in flight)",
                    "blocking_call_value": e2e_sync_rate * world},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": rl, "kernel_ms": {k: round(v, 5) for k, v in acc.items()},
        }
        if prio is not None:
            out["prioritized_replay"] = prio
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_roofline(name: str, ms: float, k_heads: int, pk):
    """Algorithmic work of one launch of kernel `name` (DESIGN.md "Kernels") over its live-measured duration."""
    mac = {"fwd_L0": 2 * 3_612_672, "fwd_L1": 2 * 3_964_928, "fwd_L2": 2 * 4_460_544, "fwd_L3": 2 * 3_964_928,
           "wgrad_L0": 3_612_672, "wgrad_L1": 3_964_928, "wgrad_L2": 4_460_544, "wgrad_L3": 3_964_928,
           "dgrad_L1": 3_964_928, "dgrad_L2": 4_460_544, "dgrad_L3": 3_964_928}
    base = name
    for prefix in ("tc_", "img_", "dense_"):
        if base.startswith(prefix):
            base = base[len(prefix):]
    tensor_path = base != name
    dense0 = 7744 * 512 + 512
    if base.startswith("wgrad_adam"):
        # reads W, mu, nu and writes W, mu, nu of Dense_0 (+ the 1 MB of activations feeding the outer product)
        gb = (6 * dense0 * 4 + (32 * 7744 + 32 * 512) * 4) * k_heads / 1e9
        ach = gb / (ms * 1e-3)
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    if base.startswith("adam"):
        n = P_HEAD - (dense0 if base != "adam" else 0)
        gb = 7 * n * 4 * k_heads / 1e9
        ach = gb / (ms * 1e-3)
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    if base in ("fwd_L3", "dgrad_L3") and tensor_path:
        # weight streaming (M or N = batch 32): bound by reading the 15.9 MB Dense_0 kernel per net
        nets = 2 * k_heads if base == "fwd_L3" else k_heads
        gb = dense0 * 4 * nets / 1e9
        ach = gb / (ms * 1e-3)
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    if base in mac:
        tf = 2 * mac[base] * B * k_heads / 1e12
        ach = tf / (ms * 1e-3)
        return {"bound": "tensor", "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": ach / pk["tf_sus"]}
    return {"bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prioritized", action="store_true", help="skip the prioritised-replay (configs[4]) leg")
    ap.add_argument("--flags", type=int, default=0, help="IDQN_F_* engine flags (A/B experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
