"""GPU SumTree / samplers / replay buffer: bit-exact against the golden vectors produced by the reference's own
code, plus the reference's known-answer tests ported to this package's drop-in classes."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OBSERVATION_SHAPE = (84, 84)
STACK_SIZE = 4
BATCH_SIZE = 32


@pytest.fixture(scope="module")
def g_tree(golden_dir):
    return np.load(os.path.join(golden_dir, "sum_tree.npz"))


@pytest.mark.parametrize("cap", [1, 2, 5, 8, 100, 1000])
def test_sum_tree_small_bit_exact(g_tree, cap):
    from idqn_b200.sample_collection.sum_tree import SumTree
    tree = SumTree(cap)
    idx, val = g_tree[f"small{cap}_idx"], g_tree[f"small{cap}_val"]
    for it in range(idx.shape[0]):
        n = int((idx[it] >= 0).sum())
        tree.set(idx[it, :n], val[it, :n])
        assert tree._nodes.tobytes() == g_tree[f"small{cap}_nodes"][it].tobytes()
        if tree.root > 0:
            np.testing.assert_array_equal(tree.query(g_tree[f"small{cap}_targets"][it]), g_tree[f"small{cap}_query"][it])
    assert tree.max_recorded_priority == float(g_tree[f"small{cap}_maxprio"])


def test_sum_tree_1m_workload_bit_exact(g_tree):
    from idqn_b200.sample_collection.sum_tree import SumTree
    cap = 1_000_000
    rng = np.random.default_rng(0)
    tree = SumTree(cap)
    assert tree._depth == 21 and tree._nodes.size == 2_097_151
    prio = rng.uniform(0.1, 1.0, cap)
    for lo in range(0, cap, 100_000):
        tree.set(np.arange(lo, lo + 100_000, dtype=np.int32), prio[lo:lo + 100_000])
    assert tree.root == g_tree["big_roots"][0]
    for r in range(200):
        idx = rng.integers(0, cap, 32).astype(np.int32)
        val = rng.uniform(0.0, 2.0, 32)
        tree.set(idx, val)
        t = rng.uniform(0.0, tree.root, 32)
        np.testing.assert_array_equal(tree.query(t), g_tree["big_query"][r])
        assert tree.root == g_tree["big_roots"][r + 1]
    assert hashlib.sha256(tree._nodes.tobytes()).hexdigest() == str(g_tree["big_nodes_sha256"])


def test_sum_tree_bulk_with_duplicates_matches_oracle():
    """bulk path (n > 1024) with duplicate leaves: first occurrence wins, ordered adds."""
    from idqn_b200.sample_collection.sum_tree import SumTree
    from oracle.sum_tree import SumTreeOracle
    rng = np.random.default_rng(5)
    cap = 5000
    a, b = SumTree(cap), SumTreeOracle(cap)
    for _ in range(3):
        idx = rng.integers(0, cap, 7000).astype(np.int32)
        val = rng.uniform(0, 1, 7000)
        a.set(idx, val), b.set(idx, val)
        assert a._nodes.tobytes() == b.nodes.tobytes()
    u = rng.random(4096)
    np.testing.assert_array_equal(a.sample_unit(u), b.query(b.root * u))


# ---- reference tests/test_sum_tree.py, ported -------------------------------------------------------

def test_ref_sum_tree_kats():
    from idqn_b200.sample_collection import sum_tree
    with pytest.raises(AssertionError):
        sum_tree.SumTree(capacity=-1)
    tree = sum_tree.SumTree(capacity=100)
    with pytest.raises(AssertionError):
        tree.set(0, -1)
    t1 = sum_tree.SumTree(capacity=1)
    t1.set(0, 1.5)
    assert t1.root == 1.5
    tree.set(0, 1.0)
    assert tree.get(0) == 1.0
    leaf_index, nodes = tree._first_leaf_offset, tree._nodes
    while leaf_index > 0:
        leaf_index = leaf_index // 2
        assert nodes[leaf_index] == 1.0
    tree = sum_tree.SumTree(capacity=100)
    tree.set(np.array([1, 2], dtype=np.int32), np.array([3.0, 4.0], dtype=np.float32))
    assert tree.get(1) == 3.0 and tree.get(2) == 4.0 and tree.root == 7.0
    tree = sum_tree.SumTree(capacity=100)
    tree.set(np.array([1, 1, 1, 2, 2], dtype=np.int32), np.array([3.0, 3.0, 3.0, 4.0, 4.0], dtype=np.float32))
    assert tree.get(1) == 3.0 and tree.get(2) == 4.0 and tree.root == 7.0
    assert tree._nodes.size >= 100
    with pytest.raises(ValueError):
        sum_tree.SumTree(capacity=100).query(1.0)
    tree = sum_tree.SumTree(capacity=100)
    tree.set(5, 1.0)
    assert tree.query(0.99) == 5


def test_ref_sum_tree_query_kats():
    from idqn_b200.sample_collection import sum_tree
    tree = sum_tree.SumTree(capacity=4)
    tree.set(np.array([0, 1, 2, 3], dtype=np.int32), np.array([0.5, 1.0, 0.5, 0.5], dtype=np.float32))
    assert tree.root == 2.5 and tree._depth == 3 and tree._nodes.size == 7
    np.testing.assert_array_equal(tree.query(np.array([1.5, 1.0])), np.array([2, 1], np.int32))
    tree.set(0, 0.25)
    assert tree.root == 2.25
    assert tree.query(0.249) == 0 and tree.query(0.5) == 1 and tree.query(1.25) == 2
    tree = sum_tree.SumTree(capacity=8)
    tree.set(np.arange(8, dtype=np.int32), np.ones((8,), dtype=np.float32))
    assert tree.root == 8.0 and tree._depth == 4 and tree._nodes.size == 15
    np.testing.assert_array_equal(tree.query(np.arange(8, dtype=np.int32)), np.arange(8, dtype=np.int32))
    tree = sum_tree.SumTree(capacity=100)
    tree.set(0, 0)
    assert tree.max_recorded_priority == 1
    for i in range(1, 32):
        tree.set(i, i)
        assert tree.max_recorded_priority == i


# ---- samplers ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("tag,exponent", [("prio1_", 1.0), ("prio_", 0.6)])
def test_prioritized_sampler_golden_and_oracle(golden_dir, tag, exponent):
    """Bit-exact against the reference-generated golden (exponent 1.0 always; 0.6 where this host's numpy pow
    reproduces the generating host's) and, on every host, against the CPU oracle fed the same operations."""
    from idqn_b200.sample_collection import samplers
    from oracle.samplers import PrioritizedSamplerOracle
    g = np.load(os.path.join(golden_dir, "samplers.npz"))
    x = g["pow_probe_in"]
    golden_ok = exponent == 1.0 or ((x ** 0.6).tobytes() == g["pow_probe_array"].tobytes() and
                                    np.asarray([float(v) ** 0.6 for v in x]).tobytes() == g["pow_probe_scalar"].tobytes())
    s = samplers.PrioritizedSamplingDistribution(seed=3, max_capacity=50, priority_exponent=exponent)
    o = PrioritizedSamplerOracle(seed=3, max_capacity=50, priority_exponent=exponent)
    for it, (op, key, p) in enumerate(g[tag + "ops"]):
        op, key, p = int(op), int(key), float(p)
        for t in (s, o):
            if op == 0:
                t.add(key, priority=p)
            elif op == 1:
                t.update(np.asarray([key]), np.asarray([p]))
            elif op == 2:
                t.remove(key)
        if o.tree.root > 0:
            got = s.sample(8)
            np.testing.assert_array_equal(got, o.sample(8))
            if golden_ok:
                np.testing.assert_array_equal(got, g[tag + "samples"][it])
        assert s._sum_tree._nodes.tobytes() == o.tree.nodes.tobytes(), f"op {it}"
    if golden_ok:
        assert s._sum_tree._nodes.tobytes() == g[tag + "nodes"].tobytes()
        np.testing.assert_array_equal(np.asarray(s._index_to_key), g[tag + "index_to_key"])


def test_ref_prioritized_sampler_kat():
    from idqn_b200.sample_collection import samplers
    sampler = samplers.PrioritizedSamplingDistribution(seed=0, max_capacity=10)
    for key, priority in zip([0, 1, 2, 3, 4], [1.0, 2.0, 3.0, 4.0, 0.0]):
        sampler.add(key, priority=priority)
    np.testing.assert_array_less(sampler.sample(5), 4)
    sampler.update(keys=np.array([2, 3]), priorities=np.array([0.0, 0.0]))
    np.testing.assert_array_less(sampler.sample(5), 2)
    sampler.remove(0)
    np.testing.assert_array_almost_equal(sampler.sample(5), 1)


# ---- replay buffer ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("ci", [0, 1, 2, 3])
def test_replay_buffer_golden(golden_dir, ci):
    from idqn_b200.sample_collection import replay_buffer, samplers
    g = np.load(os.path.join(golden_dir, "replay_buffer.npz"))
    p = f"cfg{ci}_"
    stack, n, gamma, cap = g[p + "cfg"]
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=ci), batch_size=8, max_capacity=int(cap),
                                    stack_size=int(stack), update_horizon=int(n), gamma=float(gamma), compress=False)
    bi = 0
    for t in range(g[p + "obs"].shape[0]):
        rb.add(replay_buffer.TransitionElement(g[p + "obs"][t], int(g[p + "act"][t]), float(g[p + "rew"][t]),
                                               bool(g[p + "term"][t]), bool(g[p + "trunc"][t])))
        if rb.add_count and t % 10 == 9:
            b = rb.sample()
            np.testing.assert_array_equal(b.state, g[p + "b_state"][bi])
            np.testing.assert_array_equal(b.next_state, g[p + "b_next_state"][bi])
            np.testing.assert_array_equal(b.action, g[p + "b_action"][bi])
            assert b.reward.tobytes() == g[p + "b_reward"][bi].tobytes()
            np.testing.assert_array_equal(b.is_terminal, g[p + "b_is_terminal"][bi])
            bi += 1
    keys = list(rb._memory.keys())
    np.testing.assert_array_equal(keys, g[p + "keys"])
    assert rb.add_count == int(g[p + "add_count"])
    np.testing.assert_array_equal(np.stack([rb._memory[k].state for k in keys]), g[p + "state"])
    np.testing.assert_array_equal(np.stack([rb._memory[k].next_state for k in keys]), g[p + "next_state"])
    assert np.asarray([rb._memory[k].reward for k in keys]).tobytes() == g[p + "reward"].tobytes()


def test_ref_add_up_to_capacity():
    """reference tests/test_replay_buffer.py:51-88"""
    from idqn_b200.sample_collection import replay_buffer, samplers
    from idqn_b200.sample_collection.replay_buffer import TransitionElement
    capacity = 10
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), batch_size=BATCH_SIZE,
                                    max_capacity=capacity, stack_size=STACK_SIZE, update_horizon=1, gamma=1.0,
                                    compress=False)
    transitions = []
    for i in range(16):
        transitions.append(TransitionElement(np.full(OBSERVATION_SHAPE, i), i, i, False, False))
        rb.add(transitions[-1])
    assert len(rb._memory) == capacity
    expected_keys = list(range(5, 5 + capacity))
    assert list(rb._memory.keys()) == expected_keys
    for i in expected_keys:
        el = rb._memory[i]
        np.testing.assert_array_equal(
            el.state, np.array([t.observation for t in transitions[i - STACK_SIZE + 1: i + 1]]).transpose(1, 2, 0))
        np.testing.assert_array_equal(
            el.next_state, np.array([t.observation for t in transitions[i - STACK_SIZE + 2: i + 2]]).transpose(1, 2, 0))
        assert el.action == transitions[i].action and el.reward == transitions[i].reward
        assert el.is_terminal == int(transitions[i].is_terminal) and el.episode_end == int(transitions[i].episode_end)


def test_ref_nstep_and_stack_and_key_mappings():
    """reference tests/test_replay_buffer.py:90-136 and :230-299"""
    from idqn_b200.sample_collection import replay_buffer, samplers
    from idqn_b200.sample_collection.replay_buffer import TransitionElement
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), batch_size=BATCH_SIZE, max_capacity=10,
                                    stack_size=STACK_SIZE, update_horizon=5, gamma=1.0, compress=False)
    for i in range(50):
        rb.add(TransitionElement(np.full(OBSERVATION_SHAPE, i), 0, 2.0, False))
    for _ in range(10):
        np.testing.assert_array_equal(rb.sample().reward, np.ones(BATCH_SIZE) * 10.0)
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), batch_size=BATCH_SIZE, max_capacity=50,
                                    stack_size=STACK_SIZE, update_horizon=5, gamma=1.0, compress=False)
    for i in range(11):
        rb.add(TransitionElement(np.full(OBSERVATION_SHAPE, i), 0, 0, False))
    for i in rb._memory:
        assert rb._memory[i].state.shape == OBSERVATION_SHAPE + (4,)
    np.testing.assert_array_equal(np.zeros(OBSERVATION_SHAPE + (3,)), rb._memory[0].state[:, :, :3])
    state = rb._memory[3].state
    for i in range(4):
        np.testing.assert_array_equal(np.full(OBSERVATION_SHAPE, i), state[:, :, i])
    # key mappings after eviction
    capacity = 10
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), batch_size=BATCH_SIZE,
                                    max_capacity=capacity, stack_size=1, update_horizon=1, gamma=0.99, compress=False)
    sampler = rb._sampling_distribution
    for i in range(capacity + 1):
        rb.add(TransitionElement(np.full(OBSERVATION_SHAPE, i), i, i, False, False))
    for i in range(capacity):
        assert sampler._key_to_index[i] == i and sampler._index_to_key[i] == i
    next_key = capacity
    rb.add(TransitionElement(np.full(OBSERVATION_SHAPE, next_key + 1), next_key + 1, next_key + 1, False, False))
    assert 0 not in sampler._key_to_index and sampler._index_to_key[0] != 0 and next_key in sampler._key_to_index
    assert sampler._index_to_key[sampler._key_to_index[next_key]] == next_key
    indices = np.random.default_rng(seed=0).integers(len(sampler._index_to_key), size=BATCH_SIZE)
    keys = [sampler._index_to_key[i] for i in indices]
    samples = rb.sample()
    for i, key in enumerate(keys):
        np.testing.assert_array_equal(samples.state[i, ...], np.full(OBSERVATION_SHAPE, key)[..., None])
        np.testing.assert_array_equal(samples.next_state[i, ...], np.full(OBSERVATION_SHAPE, key + 1)[..., None])
        assert samples.action[i] == key and samples.reward[i] == key
        assert samples.is_terminal[i] == 0 and samples.episode_end[i] == 0


def test_ref_sampling_with_terminal_in_trajectory():
    """reference tests/test_replay_buffer.py:185-228"""
    from idqn_b200.sample_collection import replay_buffer, samplers
    from idqn_b200.sample_collection.replay_buffer import TransitionElement
    rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=0), batch_size=2, max_capacity=10,
                                    stack_size=1, update_horizon=3, gamma=1.0, compress=False)
    for i in range(rb._max_capacity):
        rb.add(TransitionElement(np.full(OBSERVATION_SHAPE, i), action=i * 2, reward=i, is_terminal=i == 3,
                                 episode_end=False))
    indices = np.random.default_rng(seed=0).integers(rb.add_count, size=5)
    batch = rb.sample(size=5)
    expected_states = np.array([np.full(OBSERVATION_SHAPE + (1,), i) if i < 3 else np.full(OBSERVATION_SHAPE + (1,), i + 1)
                                for i in indices])
    expected_actions = np.array([i * 2 if i < 3 else (i + 1) * 2 for i in indices])
    expected_rewards = np.array([3, 6, 5, 15, 18, 21, 24])
    expected_terminals = np.array([1, 1, 1, 0, 0, 0, 0])
    np.testing.assert_array_equal(batch.state, expected_states)
    np.testing.assert_array_equal(batch.action, expected_actions)
    np.testing.assert_array_equal(batch.reward, expected_rewards[indices])
    np.testing.assert_array_equal(batch.is_terminal, expected_terminals[indices])


def test_device_replay_feeds_learner_like_host_path():
    """update_online_params through the device gather == the same batch through the host path (bit-identical)."""
    from idqn_b200.networks.idqn import iDQN
    from idqn_b200.sample_collection import replay_buffer, samplers
    from idqn_b200.sample_collection.replay_buffer import TransitionElement
    rng = np.random.default_rng(0)
    obs = (84, 84)

    def fill(seed):
        rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=seed), batch_size=32, max_capacity=300,
                                        stack_size=4, update_horizon=1, gamma=0.99, clipping=lambda x: np.clip(x, -1, 1))
        r = np.random.default_rng(1)
        for t in range(400):
            term = r.random() < 0.02
            rb.add(TransitionElement(r.integers(0, 256, obs).astype(np.uint8), int(r.integers(0, 6)),
                                     float(r.integers(-1, 2)), bool(term), bool(term)))
        return rb

    rb_a, rb_b = fill(5), fill(5)
    mk = lambda: iDQN(3, (84, 84, 4), 6, 2, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4)
    a, b = mk(), mk()
    class HostOnly:  # hides learn_step_on: update_online_params falls back to rb.sample() + the host-buffer call
        def sample(self):
            return rb_b.sample()

    for step in range(1, 4):
        a.update_online_params(step, rb_a)  # device gather path
        b.update_online_params(step, HostOnly())  # host path with the same sampler stream
    pa, pb = a.params.to_host(), b.params.to_host()
    for m in pa["params"]:
        np.testing.assert_array_equal(pa["params"][m]["kernel"], pb["params"][m]["kernel"])
    np.testing.assert_array_equal(a.cumulated_losses, b.cumulated_losses)


def test_prioritized_sampler_with_an_all_zero_tree_falls_back_to_uniform():
    """samplers.py:105-108: with every priority zero (add(priority=None)) the prioritised sampler draws like the uniform
    one -- same PCG64 stream, same keys -- instead of raising; once a positive priority exists the tree decides."""
    from idqn_b200.sample_collection.samplers import PrioritizedSamplingDistribution, UniformSamplingDistribution
    p = PrioritizedSamplingDistribution(seed=3, max_capacity=64)
    u = UniformSamplingDistribution(seed=3)
    for key in range(20):
        p.add(key, priority=None)
        u.add(key)
    for _ in range(3):
        np.testing.assert_array_equal(p.sample(8), u.sample(8))
    p.update(np.asarray([7], np.int32), np.asarray([2.5]))
    assert (p.sample(16) == 7).all()


def test_prioritized_replay_wired_to_the_learner_end_to_end():
    """§8f N3: a prioritised buffer feeding the learner with the priorities written back after every step
    (replay_buffer.py:232-237 + samplers.py:75-87 + sum_tree.py:18,32, which the reference ships but never connects):
    new elements enter at max_recorded_priority, the step's per-sample |TD| (mean over the heads) goes into the tree on
    the device, the next batch is sampled from the updated tree.  The reference-side oracle classes driven with the
    SAME priorities (read back from the device) must produce the same keys every step and, at the end, bit-identical
    tree nodes and max_recorded_priority."""
    from idqn_b200.networks.idqn import iDQN
    from idqn_b200.sample_collection import replay_buffer, samplers
    from idqn_b200.sample_collection.replay_buffer import TransitionElement
    from oracle.samplers import PrioritizedSamplerOracle
    cap, K = 200, 2
    rb = replay_buffer.ReplayBuffer(samplers.PrioritizedSamplingDistribution(seed=11, max_capacity=cap), batch_size=32,
                                    max_capacity=cap, stack_size=4, update_horizon=1, gamma=0.99,
                                    clipping=lambda x: np.clip(x, -1, 1))
    oracle = PrioritizedSamplerOracle(seed=11, max_capacity=cap)
    agent = iDQN(3, (84, 84, 4), 6, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4)
    r = np.random.default_rng(2)

    def add(n):
        for _ in range(n):
            before = rb.add_count
            term = r.random() < 0.02
            rb.add(TransitionElement(r.integers(0, 256, (84, 84)).astype(np.uint8), int(r.integers(0, 6)),
                                     float(r.integers(-1, 2)), bool(term), bool(term)))  # no priority: enters at the maximum
            for key in range(before, rb.add_count):  # mirror ReplayBuffer.add (:207-213) on the oracle
                oracle.add(key, priority=oracle.tree.max_recorded_priority)
                if key + 1 > cap:
                    oracle.remove(key - cap)

    add(150)
    for step in range(1, 13):
        assert agent.update_online_params(step, rb) is None
        keys = np.asarray(rb.last_keys)
        np.testing.assert_array_equal(keys, oracle.sample(32), err_msg=f"sampled keys, step {step}")
        rb.update_from_learner(agent._engine)
        td = agent._engine.td_abs().astype(np.float64)
        assert td.shape == (K, 32) and (td >= 0).all() and td.max() > 0
        oracle.update(keys, (td[0] + td[1]) / K)
        agent.update_target_params(step)
        add(10)  # the buffer wraps (evictions + re-insertions at the running maximum) while the learner runs
    s = rb._sampling_distribution
    assert s._sum_tree._nodes.tobytes() == oracle.tree.nodes.tobytes()
    assert s._sum_tree.max_recorded_priority == oracle.tree.max_recorded_priority
    np.testing.assert_array_equal(np.asarray(s._index_to_key), np.asarray(oracle.index_to_key))
