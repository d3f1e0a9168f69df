"""``DQNNet`` — the Q-network description and its stand-alone ``init`` / ``apply``
(mirrors slimdqn/networks/architectures/dqn.py:32-70; 'cnn' :39-53 and 'fc' :61-63 run on the tcgen05 kernels,
'impala' :7-29,54-60 on the fp32 CUDA-core kernels of the same engine)."""
from __future__ import annotations

import math
from typing import Dict, Sequence

import numpy as np

from ... import _lib as L
from ... import _prng


class DQNNet:
    def __init__(self, features: Sequence[int], architecture_type: str, n_actions: int):
        self.features = [int(f) for f in features]
        self.architecture_type = architecture_type
        self.n_actions = int(n_actions)
        self._engines: Dict[tuple, "object"] = {}

    # -- parameter creation ----------------------------------------------------------------------
    def init(self, key, x) -> Dict[str, Dict[str, Dict[str, np.ndarray]]]:
        """``network.init(key, zeros(observation_dim))`` (idqn.py:48-50): flax layout and initialiser
        distributions (xavier-uniform for cnn incl. its dense trunk :40,68,70; lecun-normal for fc :62; zero
        biases).  The reference draws from jax's threefry stream; here the key seeds numpy's PCG64."""
        from ._shapes import layer_shapes  # local import keeps module import light

        obs = tuple(np.asarray(x).shape)
        rng = np.random.default_rng([int(v) for v in _prng.as_key(key)])
        tree: Dict = {}
        for name, kshape, bshape in layer_shapes(obs, self.features, self.architecture_type, self.n_actions):
            rf = int(np.prod(kshape[:-2]))
            fan_in, fan_out = rf * kshape[-2], rf * kshape[-1]
            # impala: xavier-uniform for a Stack's first conv (:14-19) and the dense trunk (:55,68,70); the convs of the
            # residual blocks (:25-26) pass no kernel_init = flax's lecun-normal
            xavier = self.architecture_type == "cnn" or (
                self.architecture_type == "impala" and (name.endswith("/Conv_0") or name.startswith("Dense_")))
            if xavier:
                bound = math.sqrt(6.0 / (fan_in + fan_out))
                kernel = rng.uniform(-bound, bound, kshape)
            else:
                std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
                kernel = _truncated_normal(rng, kshape) * std
            node = tree
            parts = name.split("/")
            for part in parts[:-1]:  # flax sub-module nesting: "Stack_0/Conv_1" -> tree["Stack_0"]["Conv_1"]
                node = node.setdefault(part, {})
            node[parts[-1]] = {"kernel": kernel.astype(np.float32), "bias": np.zeros(bshape, np.float32)}
        return {"params": tree}

    # -- stand-alone forward (tests, compute_target/loss helpers) -------------------------------------
    def _engine_for(self, obs_shape, device: int):
        from .._engine import Engine

        key = (tuple(obs_shape), device)
        if key not in self._engines:
            self._engines[key] = Engine(obs_shape, self.n_actions, 1, self.features, self.architecture_type,
                                        0.0, 0.0, 1, 1e-8, batch_size=32, device=device, flags=L.F_NO_GRAPH)
        return self._engines[key]

    def apply(self, params, x, device: int = 0) -> np.ndarray:
        """Q-values of ONE head for one unbatched input (architectures/dqn.py:44,65 add/remove N=1) -> [A],
        or for a leading batch axis -> [N, A].  ``params`` leaves carry no K axis."""
        from ._shapes import obs_shape_of

        x = np.asarray(x)
        obs = obs_shape_of(x, self.architecture_type, params)
        eng = self._engine_for(obs, device)
        eng.upload_tree(L.ONLINE, _host_tree(params), squeezed=True)
        q = eng.apply(L.ONLINE, 0, x)
        return q[0] if x.size == eng.in_elems else q  # unbatched input -> [A] (the squeeze of :65)


def _truncated_normal(rng: np.random.Generator, shape) -> np.ndarray:
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2
    return out


def _host_tree(params):
    if hasattr(params, "to_host"):
        return params.to_host()
    return params
