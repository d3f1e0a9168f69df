"""Host restatement of the jax PRNG calls on the acting path (idqn.py:128, sample_collection/utils.py:10-13).

threefry2x32 is pinned by the Random123 known-answer vectors (tests/test_prng.py).  The derived
``split`` / ``randint`` / ``uniform`` follow jax 0.4.30's ``jax._src.random`` / ``jax._src.prng``
(non-partitionable threefry, the default) but cannot be checked against a jax install here: *unpinned*."""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32


def threefry2x32(key, x0, x1):
    """20-round Threefry-2x32 on arrays of uint32 counters; returns (y0, y1)."""
    k0, k1 = np.uint64(int(key[0])), np.uint64(int(key[1]))
    ks = (k0, k1, k0 ^ k1 ^ np.uint64(0x1BD11BDA))
    x0 = (np.asarray(x0, np.uint64) + ks[0]) & _M32
    x1 = (np.asarray(x1, np.uint64) + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + np.uint64(i + 1)) & _M32
    return x0.astype(np.uint32), x1.astype(np.uint32)


def _threefry_scalar(k0: int, k1: int, x0: int, x1: int):
    """The same 20 rounds on Python ints -- the acting path draws ONE head per environment step, and numpy's
    per-call overhead on one-element arrays (3 block evaluations per draw) costs more than the forward pass."""
    M = 0xFFFFFFFF
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & M
    x1 = (x1 + ks[1]) & M
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & M
            x1 = (((x1 << r) | (x1 >> (32 - r))) & M) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & M
    return x0, x1


def as_key(key) -> np.ndarray:
    """jax.random.PRNGKey(seed) -> uint32[2] = [seed >> 32, seed & 0xffffffff]; uint32[2] arrays pass through."""
    if isinstance(key, (int, np.integer)):
        s = int(key) & 0xFFFFFFFFFFFFFFFF
        return np.asarray([s >> 32, s & 0xFFFFFFFF], np.uint32)
    a = np.asarray(key)
    if a.shape == (2,):
        return a.astype(np.uint32)
    raise TypeError("key must be an int seed or a uint32[2] threefry key")


def split(key, num: int = 2) -> np.ndarray:
    key = as_key(key)
    counts = np.arange(2 * num, dtype=np.uint32)
    y0, y1 = threefry2x32(key, counts[:num], counts[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)


def _key_ints(key):
    if isinstance(key, (int, np.integer)):
        s = int(key) & 0xFFFFFFFFFFFFFFFF
        return s >> 32, s & 0xFFFFFFFF
    a = as_key(key)
    return int(a[0]), int(a[1])


def _bits32(key) -> int:
    k0, k1 = _key_ints(key)
    return _threefry_scalar(k0, k1, 0, 0)[0]


def randint(key, minval: int, maxval: int) -> int:
    """jax.random.randint(key, (), minval, maxval) for int32."""
    k0, k1 = _key_ints(key)
    # split(key, 2): counters [0, 1 | 2, 3] -> blocks (0, 2) and (1, 3); keys = rows of [y0_0, y0_1, y1_0, y1_1].reshape(2, 2)
    a0, b0 = _threefry_scalar(k0, k1, 0, 2)
    a1, b1 = _threefry_scalar(k0, k1, 1, 3)
    hi = _threefry_scalar(a0, a1, 0, 0)[0]
    lo = _threefry_scalar(b0, b1, 0, 0)[0]
    span = max(int(maxval) - int(minval), 1) if maxval > minval else 1
    mult = (2 ** 16) % span
    mult = (mult * mult) % span
    off = (((hi % span) * mult) & 0xFFFFFFFF) + (lo % span)
    off = (off & 0xFFFFFFFF) % span
    return int(minval) + off


def uniform(key) -> np.float32:
    """jax.random.uniform(key) in [0, 1) as float32."""
    bits = np.uint32(_bits32(key))
    f = np.asarray((bits >> np.uint32(9)) | np.uint32(0x3F800000), np.uint32).view(np.float32)
    return np.float32(f - np.float32(1.0))
