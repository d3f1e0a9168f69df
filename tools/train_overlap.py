"""Overlap of environment stepping with the learning step in the training-loop mirror (idqn_b200/experiments/dqn.py):
environment steps per second of train() for an emulator costing `step_cost` us of host time per step, next to the two
serial bounds.   python tools/train_overlap.py [K] [step_cost_us ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from idqn_b200.experiments.dqn import SyntheticAtari, train
from idqn_b200.networks.idqn import iDQN
from idqn_b200.sample_collection.replay_buffer import ReplayBuffer
from idqn_b200.sample_collection.samplers import UniformSamplingDistribution

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
costs = [float(c) for c in sys.argv[2:]] or [0.0, 100.0, 200.0, 400.0]
for utd in (1, 4):
    for cost in costs:
        agent = iDQN(5, (84, 84, 4), 6, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, utd, 200, 10, 1.5e-4)
        rb = ReplayBuffer(UniformSamplingDistribution(seed=1), batch_size=32, max_capacity=4096, stack_size=4,
                          clipping=lambda r: np.clip(r, -1, 1))
        env = SyntheticAtari(episode_length=500, step_cost_s=cost * 1e-6)
        p = dict(n_epochs=1, n_training_steps_per_epoch=600, n_initial_samples=100, epsilon_end=0.01, epsilon_duration=300, horizon=10 ** 6)
        train(3, p, agent, env, rb)  # warm-up epoch (graph capture, allocator)
        env.actions.clear()
        p = dict(p, n_training_steps_per_epoch=3000, n_initial_samples=0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        train(4, p, agent, env, rb)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = len(env.actions)
        print(f"K={K} update_to_data={utd} emulator {cost:5.0f} us/step: {n / dt:8.1f} env steps/s ({dt / n * 1e6:6.1f} us per step, "
              f"{n / utd / dt:7.1f} learning steps/s)", flush=True)
        del agent, rb
