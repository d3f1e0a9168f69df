"""``DQN`` — drop-in for slimdqn/networks/dqn.py: the K = 1 case of the same CUDA step, parameter leaves
without the leading K axis, no D-sync (dqn.py:41-58)."""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .idqn import iDQN


class DQN(iDQN):
    def __init__(
        self,
        key,
        observation_dim,
        n_actions,
        features: list,
        architecture_type: str,
        learning_rate: float,
        gamma: float,
        update_horizon: int,
        update_to_data: int,
        target_update_frequency: int,
        adam_eps: float = 1e-8,
        *,
        batch_size: int = 32,
        device: int = 0,
        flags: int = 0,
    ):
        super().__init__(key, observation_dim, n_actions, 1, features, architecture_type, learning_rate, gamma,
                         update_horizon, update_to_data, target_update_frequency, target_update_frequency,
                         adam_eps, batch_size=batch_size, device=device, flags=flags)
        self._squeeze = True
        # dqn.py:24 initialises from `key` itself, not from split(key, 1)
        obs = tuple(int(d) for d in np.atleast_1d(observation_dim))
        self._engine.upload_tree(L.ONLINE, self.network.init(key, np.zeros(obs, np.float32)), squeezed=True)
        self._engine.copy_online_to_target()
        del self.target_sync_frequency

    @property
    def cumulated_loss(self):  # dqn.py:39,48
        return float(self._engine.cumulated_losses(reset=False)[0])

    def update_target_params(self, step: int):  # dqn.py:50-58
        if step % self.target_update_frequency == 0:
            self._engine.copy_online_to_target()
            cumulated = self._engine.cumulated_losses(reset=True)[0]
            logs = {"loss": cumulated / (self.target_update_frequency / self.update_to_data)}
            return True, logs
        return False, {}

    def _out_loss(self, losses):
        return losses[0]

    def best_action(self, params, state, **kwargs):  # dqn.py:89-92
        return super().best_action(params, state, idx_params=0)
