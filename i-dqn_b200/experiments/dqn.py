"""``train`` — the reference's only driver (experiments/base/dqn.py:12-69) over the B200 agent, with the learning step
overlapped with environment stepping.

Same signature, same control flow, same bookkeeping as the reference.  What differs is *when the host waits*:

* ``agent.update_online_params(step, rb)`` only ENQUEUES the step (sample keys -> CUDA graph of the whole learn_on_batch)
  and returns; the host goes straight on to the next ``collect_single_sample`` -- environment step, replay ``add``,
  the three PRNG draws -- while the GPU computes.  The reference gets the same overlap from jax's asynchronous dispatch
  (idqn.py:72 accumulates device futures).
* losses are summed on the device (``cumulated_losses``) and read only at a T-update (idqn.py:82-87), the one host
  synchronisation per ``target_update_frequency`` learning steps besides the 4-byte action of a greedy acting step.
* ``update_to_data`` > 1 (idqn.py:66): the learner runs every ``update_to_data``-th environment step, so up to that many
  acting steps hide behind one learning step.

``p`` needs: n_epochs, n_training_steps_per_epoch, n_initial_samples, epsilon_end, epsilon_duration, horizon; optional
``wandb`` (anything with ``.log(dict)``) and ``save`` (callable(p, returns, lengths, model), the reference's save_data).
"""
from __future__ import annotations

import numpy as np

from .. import _prng
from ..sample_collection.utils import collect_single_sample


def linear_schedule(init_value: float, end_value: float, transition_steps: int):
    """optax.linear_schedule(init, end, steps) (experiments/base/dqn.py:19): init + (end - init) * clip(count / steps, 0, 1)."""

    def schedule(count):
        if transition_steps <= 0:
            return init_value
        frac = min(max(count / transition_steps, 0.0), 1.0)
        return (init_value - end_value) * (1.0 - frac) + end_value  # optax's polynomial_schedule form, power 1

    return schedule


def train(key, p: dict, agent, env, rb):
    epsilon_schedule = linear_schedule(1.0, p["epsilon_end"], p["epsilon_duration"])
    log = p["wandb"].log if p.get("wandb") is not None else (lambda d: None)

    n_training_steps = 0
    env.reset()
    episode_returns_per_epoch = [[0]]
    episode_lengths_per_epoch = [[0]]

    for idx_epoch in range(p["n_epochs"]):
        n_training_steps_epoch = 0
        has_reset = False

        while n_training_steps_epoch < p["n_training_steps_per_epoch"] or not has_reset:
            key, exploration_key = _prng.split(key)
            reward, has_reset = collect_single_sample(exploration_key, env, agent, rb, p, epsilon_schedule, n_training_steps)

            n_training_steps_epoch += 1
            n_training_steps += 1

            episode_returns_per_epoch[idx_epoch][-1] += reward
            episode_lengths_per_epoch[idx_epoch][-1] += 1
            if has_reset and n_training_steps_epoch < p["n_training_steps_per_epoch"]:
                episode_returns_per_epoch[idx_epoch].append(0)
                episode_lengths_per_epoch[idx_epoch].append(0)

            if n_training_steps > p["n_initial_samples"]:
                agent.update_online_params(n_training_steps, rb)  # enqueued; the next acting step overlaps it
                target_updated, logs = agent.update_target_params(n_training_steps)
                if target_updated:
                    log({"n_training_steps": n_training_steps, **logs})

        avg_return = np.mean(episode_returns_per_epoch[idx_epoch])
        avg_length_episode = np.mean(episode_lengths_per_epoch[idx_epoch])
        log({"epoch": idx_epoch, "n_training_steps": n_training_steps, "avg_return": avg_return,
             "avg_length_episode": avg_length_episode})
        if idx_epoch < p["n_epochs"] - 1:
            episode_returns_per_epoch.append([0])
            episode_lengths_per_epoch.append([0])
        if p.get("save") is not None:
            p["save"](p, episode_returns_per_epoch, episode_lengths_per_epoch, agent.get_model())
    return episode_returns_per_epoch, episode_lengths_per_epoch


class SyntheticAtari:
    """A stand-in environment with the interface the loop uses (reference: slimdqn/environments/atari.py): 84x84 uint8
    frames, a 4-frame ``state`` as float32 holding 0..255 (atari.py:43-45), ``observation`` = newest frame, a fixed
    host cost per ``step`` and episodes of ``episode_length`` steps.  Used by the tests and the overlap measurement; the
    real ALE wrapper is outside the scope of this package."""

    def __init__(self, n_actions: int = 6, episode_length: int = 97, seed: int = 0, step_cost_s: float = 0.0):
        self.n_actions, self.episode_length, self.step_cost_s = n_actions, episode_length, step_cost_s
        self._rng = np.random.default_rng(seed)
        self._frames = self._rng.integers(0, 256, (64, 84, 84), dtype=np.uint8)
        self.actions = []
        self.reset()

    def reset(self):
        self.n_steps = 0
        self._stack = np.zeros((84, 84, 4), np.uint8)
        self._push(self._frames[0])

    def _push(self, frame):
        self._stack = np.concatenate([self._stack[..., 1:], frame[..., None]], axis=-1)

    @property
    def observation(self):
        return self._stack[..., -1]

    @property
    def state(self):
        return self._stack.astype(np.float32)

    def step(self, action: int):
        import time
        if self.step_cost_s > 0:  # emulator time: the host is busy, the GPU keeps learning
            t_end = time.perf_counter() + self.step_cost_s
            while time.perf_counter() < t_end:
                pass
        self.actions.append(int(action))
        self.n_steps += 1
        self._push(self._frames[(self.n_steps * 7 + int(action)) % len(self._frames)])
        reward = float((action + self.n_steps) % 3 - 1)
        absorbing = self.n_steps >= self.episode_length
        return reward, absorbing
