"""Layer shapes of DQNNet in flax creation order (architectures/dqn.py:39-70)."""
from __future__ import annotations

import numpy as np

CNN_SPECS = ((8, 4), (4, 2), (3, 1))  # (kernel, stride): architectures/dqn.py:43,48,51


def same_out(size: int, stride: int) -> int:
    return -(-size // stride)


def layer_shapes(observation_dim, features, architecture_type, n_actions):
    layers = []
    if architecture_type == "cnn":
        h, w, c = observation_dim
        for i, (k, s) in enumerate(CNN_SPECS):
            layers.append((f"Conv_{i}", (k, k, c, features[i]), (features[i],)))
            h, w, c = same_out(h, s), same_out(w, s), features[i]
        fan_in, start = h * w * c, 3
    elif architecture_type == "impala":
        # architectures/dqn.py:54-60 / 7-29: per Stack Conv_0, 3x3 / 2 SAME max-pool, two residual blocks (Conv_1..Conv_4)
        h, w, c = observation_dim
        for i in range(3):
            f = features[i]
            layers.append((f"Stack_{i}/Conv_0", (3, 3, c, f), (f,)))
            for j in range(1, 5):
                layers.append((f"Stack_{i}/Conv_{j}", (3, 3, f, f), (f,)))
            h, w, c = same_out(h, 2), same_out(w, 2), f
        fan_in, start = h * w * c, 3
    elif architecture_type == "fc":
        fan_in, start = int(np.prod(observation_dim)), 0
    else:
        raise ValueError(f"architecture_type={architecture_type!r}: expected 'cnn', 'impala' or 'fc'")
    d = 0
    for f in list(features[start:]) + [n_actions]:
        layers.append((f"Dense_{d}", (fan_in, f), (f,)))
        fan_in, d = f, d + 1
    return layers


def obs_shape_of(x: np.ndarray, architecture_type: str, params):
    """Recover observation_dim from an input and the first layer's kernel."""
    inner = params["params"]
    if architecture_type in ("cnn", "impala"):
        first = inner["Conv_0"] if architecture_type == "cnn" else inner["Stack_0"]["Conv_0"]
        c = np.asarray(first["kernel"]).shape[-2]
        if x.ndim < 3 or x.shape[-1] != c:
            raise ValueError(f"{architecture_type} input must end in (H, W, {c}), got {x.shape}")
        return tuple(x.shape[-3:])
    return (int(np.asarray(inner["Dense_0"]["kernel"]).shape[-2]),)
