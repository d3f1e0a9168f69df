"""The CPU oracle (oracle/sum_tree.py, samplers.py, replay_buffer.py) against
(1) golden vectors produced by executing the reference's own code (tests/golden/make_golden.py) and
(2) the known-answer tests of the reference (tests/test_sum_tree.py, test_samplers.py, test_replay_buffer.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle.replay_buffer import ReplayBufferOracle, Transition
from oracle.samplers import PrioritizedSamplerOracle, UniformSamplerOracle
from oracle.sum_tree import SumTreeOracle


@pytest.fixture(scope="module")
def g_tree(golden_dir):
    return np.load(os.path.join(golden_dir, "sum_tree.npz"))


@pytest.mark.parametrize("cap", [1, 2, 5, 8, 100, 1000])
def test_sum_tree_small_bit_exact(g_tree, cap):
    tree = SumTreeOracle(cap)
    idx, val = g_tree[f"small{cap}_idx"], g_tree[f"small{cap}_val"]
    for it in range(idx.shape[0]):
        n = int((idx[it] >= 0).sum())
        tree.set(idx[it, :n], val[it, :n])
        assert tree.nodes.tobytes() == g_tree[f"small{cap}_nodes"][it].tobytes()  # bit-exact f64
        if tree.root > 0:
            np.testing.assert_array_equal(tree.query(g_tree[f"small{cap}_targets"][it]), g_tree[f"small{cap}_query"][it])
    assert tree.max_recorded_priority == float(g_tree[f"small{cap}_maxprio"])


def test_sum_tree_1m_workload_bit_exact(g_tree):
    """SURVEY §8(d) sampler workload at 1M capacity, shortened to 200 rounds."""
    cap = 1_000_000
    rng = np.random.default_rng(0)
    tree = SumTreeOracle(cap)
    prio = rng.uniform(0.1, 1.0, cap)
    # bulk load: ascending distinct leaves -> same per-node add order as the chunked reference calls
    for lo in range(0, cap, 100_000):
        tree.set(np.arange(lo, lo + 100_000), prio[lo:lo + 100_000])
    assert tree.root == g_tree["big_roots"][0]
    for r in range(200):
        idx = rng.integers(0, cap, 32).astype(np.int32)
        val = rng.uniform(0.0, 2.0, 32)
        tree.set(idx, val)
        t = rng.uniform(0.0, tree.root, 32)
        np.testing.assert_array_equal(tree.query(t), g_tree["big_query"][r])
        assert tree.root == g_tree["big_roots"][r + 1]
    assert hashlib.sha256(tree.nodes.tobytes()).hexdigest() == str(g_tree["big_nodes_sha256"])


# ---- reference KATs (tests/test_sum_tree.py) -------------------------------------------------

def test_kat_capacity_and_negative():
    with pytest.raises(AssertionError):
        SumTreeOracle(-1)
    with pytest.raises(AssertionError):
        SumTreeOracle(100).set(0, -1)


def test_kat_small_and_dups():
    t = SumTreeOracle(1)
    t.set(0, 1.5)
    assert t.root == 1.5
    t = SumTreeOracle(100)
    t.set(np.array([1, 1, 1, 2, 2], np.int32), np.array([3.0, 3.0, 3.0, 4.0, 4.0], np.float32))
    assert t.get(1) == 3.0 and t.get(2) == 4.0 and t.root == 7.0
    with pytest.raises(ValueError):
        SumTreeOracle(100).query(1.0)


def test_kat_queries():
    t = SumTreeOracle(4)
    t.set(np.arange(4), np.array([0.5, 1.0, 0.5, 0.5], np.float32))
    assert t.root == 2.5 and t.depth == 3 and t.nodes.size == 7
    np.testing.assert_array_equal(t.query(np.array([1.5, 1.0])), [2, 1])
    t.set(0, 0.25)
    assert t.root == 2.25
    assert t.query(0.249)[0] == 0 and t.query(0.5)[0] == 1 and t.query(1.25)[0] == 2
    t = SumTreeOracle(8)
    t.set(np.arange(8), np.ones(8, np.float32))
    assert t.root == 8.0 and t.depth == 4 and t.nodes.size == 15
    np.testing.assert_array_equal(t.query(np.arange(8)), np.arange(8))
    t = SumTreeOracle(100)
    t.set(0, 0)
    assert t.max_recorded_priority == 1
    for i in range(1, 32):
        t.set(i, i)
        assert t.max_recorded_priority == i


# ---- samplers ---------------------------------------------------------------------------------

def host_pow_matches_golden(g):
    x = g["pow_probe_in"]
    return ((x ** 0.6).tobytes() == g["pow_probe_array"].tobytes()
            and np.asarray([float(v) ** 0.6 for v in x]).tobytes() == g["pow_probe_scalar"].tobytes())


@pytest.mark.parametrize("tag,exponent", [("prio1_", 1.0), ("prio_", 0.6)])
def test_prioritized_sampler_golden(golden_dir, tag, exponent):
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    if exponent != 1.0 and not host_pow_matches_golden(g):
        pytest.skip("numpy pow on this host differs (SIMD dispatch) from the host that generated the golden file")
    s = PrioritizedSamplerOracle(seed=3, max_capacity=50, priority_exponent=exponent)
    for it, (op, key, p) in enumerate(g[tag + "ops"]):
        op, key, p = int(op), int(key), float(p)
        if op == 0:
            s.add(key, priority=p)
        elif op == 1:
            s.update(np.asarray([key]), np.asarray([p]))
        elif op == 2:
            s.remove(key)
        if s.tree.root > 0:
            np.testing.assert_array_equal(s.sample(8), g[tag + "samples"][it])
    assert s.tree.nodes.tobytes() == g[tag + "nodes"].tobytes()
    np.testing.assert_array_equal(np.asarray(s.index_to_key), g[tag + "index_to_key"])


def test_uniform_sampler_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    u = UniformSamplerOracle(seed=11)
    for key in range(200):
        u.add(key)
        if key >= 37:
            u.remove(key - 37)
        np.testing.assert_array_equal(u.sample(6), g["uni_samples"][key])
    np.testing.assert_array_equal(np.asarray(u.index_to_key), g["uni_index_to_key"])


def test_kat_prioritized_sampler():
    """reference tests/test_samplers.py:16-35"""
    s = PrioritizedSamplerOracle(seed=0, max_capacity=10)
    for key, p in zip([0, 1, 2, 3, 4], [1.0, 2.0, 3.0, 4.0, 0.0]):
        s.add(key, priority=p)
    assert (s.sample(5) < 4).all()
    s.update(keys=np.array([2, 3]), priorities=np.array([0.0, 0.0]))
    assert (s.sample(5) < 2).all()
    s.remove(0)
    np.testing.assert_array_equal(s.sample(5), 1)


# ---- replay buffer ------------------------------------------------------------------------------

@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_replay_buffer_golden(golden_dir, ci):
    g = np.load(os.path.join(golden_dir, "replay_buffer.npz"))
    p = f"cfg{ci}_"
    stack, n, gamma, cap = g[p + "cfg"]
    rb = ReplayBufferOracle(UniformSamplerOracle(seed=ci), 8, int(cap), int(stack), int(n), float(gamma))
    bi = 0
    for t in range(g[p + "obs"].shape[0]):
        rb.add(Transition(g[p + "obs"][t], int(g[p + "act"][t]), float(g[p + "rew"][t]), bool(g[p + "term"][t]),
                          bool(g[p + "trunc"][t])))
        if rb.add_count and t % 10 == 9:
            b, _ = rb.sample()
            np.testing.assert_array_equal(b["state"], g[p + "b_state"][bi])
            np.testing.assert_array_equal(b["next_state"], g[p + "b_next_state"][bi])
            np.testing.assert_array_equal(b["action"], g[p + "b_action"][bi])
            assert b["reward"].tobytes() == g[p + "b_reward"][bi].tobytes()
            np.testing.assert_array_equal(b["is_terminal"], g[p + "b_is_terminal"][bi])
            bi += 1
    keys = list(rb.memory.keys())
    np.testing.assert_array_equal(keys, g[p + "keys"])
    assert rb.add_count == int(g[p + "add_count"])
    np.testing.assert_array_equal(np.stack([rb.memory[k]["state"] for k in keys]), g[p + "state"])
    np.testing.assert_array_equal(np.stack([rb.memory[k]["next_state"] for k in keys]), g[p + "next_state"])
    assert np.asarray([rb.memory[k]["reward"] for k in keys]).tobytes() == g[p + "reward"].tobytes()
