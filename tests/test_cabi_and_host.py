"""CPU-only checks: the C-ABI library loads and exports every symbol include/idqn_b200.h declares (no compute
calls), and the host-side logic (accumulator, uniform sampler, PRNG restatement, numpy identities the device
path relies on) matches the golden vectors / known answers."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "idqn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(idqn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import idqn_b200
    from idqn_b200 import _lib
    path = idqn_b200.library_path()
    assert os.path.exists(path), "libidqn_b200.so missing: run __graft_entry__.build()"
    handle = ctypes.CDLL(path)
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/idqn_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding table and header disagree"
    assert idqn_b200.lib().idqn_version() >= 100


def test_no_silent_fallback_without_library(monkeypatch, tmp_path):
    from idqn_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("IDQN_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryError):
        _lib.lib()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "i-dqn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_accumulator_matches_reference_golden(golden_dir):
    """ReplayBuffer.accumulate is pure host logic: replay the golden transition streams and compare every
    emitted element with what the reference stored (FIFO tail = the last `capacity` emissions)."""
    from idqn_b200.sample_collection.replay_buffer import ReplayBuffer, TransitionElement
    g = np.load(os.path.join(golden_dir, "replay_buffer.npz"))
    for ci in range(4):
        p = f"cfg{ci}_"
        stack, n, gamma, cap = g[p + "cfg"]
        rb = ReplayBuffer(None, 8, int(cap), int(stack), int(n), float(gamma), compress=False)
        emitted = []
        for t in range(g[p + "obs"].shape[0]):
            tr = TransitionElement(g[p + "obs"][t], int(g[p + "act"][t]), float(g[p + "rew"][t]), bool(g[p + "term"][t]),
                                   bool(g[p + "trunc"][t]))
            emitted.extend(rb.accumulate(tr))
        assert len(emitted) == int(g[p + "add_count"])
        tail = emitted[-int(cap):]
        np.testing.assert_array_equal(np.stack([e.state for e in tail]), g[p + "state"])
        np.testing.assert_array_equal(np.stack([e.next_state for e in tail]), g[p + "next_state"])
        np.testing.assert_array_equal([e.action for e in tail], g[p + "action"])
        assert np.asarray([e.reward for e in tail], np.float64).tobytes() == g[p + "reward"].tobytes()
        np.testing.assert_array_equal([e.is_terminal for e in tail], g[p + "is_terminal"])


def test_uniform_sampler_matches_reference_golden(golden_dir):
    from idqn_b200.sample_collection.samplers import UniformSamplingDistribution
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    u = UniformSamplingDistribution(seed=11)
    for key in range(200):
        u.add(key)
        if key >= 37:
            u.remove(key - 37)
        np.testing.assert_array_equal(u.sample(6), g["uni_samples"][key])
    np.testing.assert_array_equal(np.asarray(u._index_to_key), g["uni_index_to_key"])


def test_uniform_scaled_equals_generator_uniform():
    """samplers.py:110 `rng.uniform(0.0, root, n)` == root * rng.random(n) bit for bit (one rounded multiply):
    the identity the device-side `idqn_sumtree_sample` relies on."""
    for seed, root in [(0, 3.7), (1, 524288.123456789), (2, 1e-3), (3, 549_876.5)]:
        a = np.random.default_rng(seed).uniform(0.0, root, 4096)
        b = root * np.random.default_rng(seed).random(4096)
        assert a.tobytes() == b.tobytes()


def test_element_pack_unpack_roundtrip():
    """reference tests/test_replay_buffer.py:21-49"""
    from idqn_b200.sample_collection.replay_buffer import ReplayElement
    state = np.zeros((84, 84, 4), np.uint8)
    next_state = np.ones((84, 84, 4), np.uint8)
    el = ReplayElement(state=state, action=1, reward=1.0, next_state=next_state, is_terminal=False, episode_end=False)
    packed = el.pack()
    assert packed.action == 1 and packed.reward == 1.0 and packed.is_terminal == packed.episode_end == False  # noqa: E712
    un = packed.unpack()
    np.testing.assert_array_equal(un.state, state)
    np.testing.assert_array_equal(un.next_state, next_state)


def test_threefry_known_answers():
    """Random123 / jax random_test.py testThreefry2x32 vectors."""
    from idqn_b200._prng import threefry2x32
    def run(key, ctr):
        y0, y1 = threefry2x32(np.asarray(key, np.uint32), np.asarray([ctr[0]], np.uint32), np.asarray([ctr[1]], np.uint32))
        return (int(y0[0]), int(y1[0]))
    assert run((0x0, 0x0), (0x0, 0x0)) == (0x6B200159, 0x99BA4EFE)
    assert run((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF)) == (0x1CB996FC, 0xBB002BE7)
    assert run((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3)) == (0xC4923A9C, 0x483DF7A0)


def test_prng_helpers_are_well_formed():
    from idqn_b200 import _prng
    keys = _prng.split(42, 3)
    assert keys.shape == (3, 2) and keys.dtype == np.uint32 and len({tuple(k) for k in keys}) == 3
    draws = [_prng.randint(k, 0, 5) for k in _prng.split(7, 200)]
    assert min(draws) == 0 and max(draws) == 4
    us = [_prng.uniform(k) for k in _prng.split(9, 200)]
    assert 0.0 <= min(us) and max(us) < 1.0 and 0.3 < float(np.mean(us)) < 0.7


def test_layer_shapes_match_survey():
    from idqn_b200.networks.architectures._shapes import layer_shapes
    shapes = layer_shapes((84, 84, 4), [32, 64, 64, 512], "cnn", 6)
    assert [s[0] for s in shapes] == ["Conv_0", "Conv_1", "Conv_2", "Dense_0", "Dense_1"]
    assert shapes[3][1] == (7744, 512)
    assert sum(int(np.prod(k)) + int(np.prod(b)) for _, k, b in shapes) == 4_046_502
    assert [s[1] for s in layer_shapes((8,), [100, 100], "fc", 4)] == [(8, 100), (100, 100), (100, 4)]


def test_engine_flag_constants_match_the_header():
    """The F_* constants of the ctypes binding are the IDQN_F_* bits of include/idqn_b200.h."""
    import re
    from idqn_b200 import _lib
    with open(os.path.join(ROOT, "include", "idqn_b200.h")) as f:
        header = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"#define IDQN_F_(\w+)\s+(\d+)", f.read()))
    assert header, "no IDQN_F_* flags found in the header"
    bits = sorted(header.values())
    assert len(set(bits)) == len(bits) and all(b & (b - 1) == 0 for b in bits), "flags must be distinct single bits"
    for name, value in header.items():
        assert getattr(_lib, "F_" + name) == value, f"F_{name} != IDQN_F_{name}"


def test_scalar_threefry_equals_the_pinned_array_version():
    """The acting path draws its head with a Python-int threefry (27 us instead of 240 us per draw); it must agree bit
    for bit with the array implementation that the Random123 vectors pin."""
    from idqn_b200 import _prng
    rng = np.random.default_rng(3)
    for _ in range(200):
        k = rng.integers(0, 2 ** 32, 2).astype(np.uint32)
        c = rng.integers(0, 2 ** 32, 2).astype(np.uint32)
        y0, y1 = _prng.threefry2x32(k, np.asarray([c[0]], np.uint32), np.asarray([c[1]], np.uint32))
        assert _prng._threefry_scalar(int(k[0]), int(k[1]), int(c[0]), int(c[1])) == (int(y0[0]), int(y1[0]))
        # randint through split() + the array block function == the scalar fast path
        k1, k2 = _prng.split(k)
        hi = int(_prng.threefry2x32(k1, np.zeros(1, np.uint32), np.zeros(1, np.uint32))[0][0])
        lo = int(_prng.threefry2x32(k2, np.zeros(1, np.uint32), np.zeros(1, np.uint32))[0][0])
        span = int(rng.integers(1, 19))
        mult = (2 ** 16) % span
        mult = (mult * mult) % span
        off = ((((hi % span) * mult) & 0xFFFFFFFF) + (lo % span) & 0xFFFFFFFF) % span
        assert _prng.randint(k, 0, span) == off


def test_prng_matches_the_values_the_jax_documentation_prints():
    """Pins of the derived draws against published JAX output (default threefry implementation; "Pseudorandom numbers"
    tutorial of the JAX documentation): PRNGKey(42) -> [0 42]; split(PRNGKey(42)) -> [2465931498 3679230171], [255383827
    267815257]; split(PRNGKey(0)) -> [4146024105 967050713], [2718843009 1272950319]; uniform(PRNGKey(0)) -> 0.41845703.
    jax is not installable here, so these literature values are the pin for the counter layout of split and the
    bits -> float mapping of uniform; randint is built from the same two primitives (jax._src.random._randint)."""
    from idqn_b200 import _prng
    np.testing.assert_array_equal(_prng.as_key(42), [0, 42])
    np.testing.assert_array_equal(_prng.split(42), [[2465931498, 3679230171], [255383827, 267815257]])
    np.testing.assert_array_equal(_prng.split(0), [[4146024105, 967050713], [2718843009, 1272950319]])
    assert _prng.uniform(0) == np.float32(0.41845703)


def test_c_draws_equal_the_python_restatement():
    """idqn_prng / idqn_select_action draw in C what _prng draws in Python (no GPU involved)."""
    import ctypes as C
    from idqn_b200 import _lib as L
    from idqn_b200 import _prng
    lib = L.lib()
    rng = np.random.default_rng(1)
    for _ in range(200):
        k = rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32)
        num = int(rng.integers(1, 5))
        out = np.zeros(2 * num, np.uint32)
        L.check(lib.idqn_prng(0, int(k[0]), int(k[1]), num, 0, L.ptr(out)))
        np.testing.assert_array_equal(out.reshape(num, 2), _prng.split(k, num))
        L.check(lib.idqn_prng(1, int(k[0]), int(k[1]), 0, 0, L.ptr(out)))
        assert out[:1].view(np.float32)[0] == _prng.uniform(k)
        lo, span = int(rng.integers(-5, 5)), int(rng.integers(1, 40))
        L.check(lib.idqn_prng(2, int(k[0]), int(k[1]), lo, lo + span, L.ptr(out)))
        assert int(out[:1].view(np.int32)[0]) == _prng.randint(k, lo, lo + span)
