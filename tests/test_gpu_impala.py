"""§8f N4 -- the IMPALA architecture (architectures/dqn.py:7-29, 54-60) through the same engine and C ABI: 3x3 SAME
convolutions, 3x3 / 2 SAME max-pool, pre-activation residual blocks, on the generic tcgen05 implicit-GEMM kernels (fp32
CUDA-core kernels where a layer's channel counts do not allow 16-byte operand vectors) and the max-pool kernels.

Tolerances: fp32 storage / accumulation with bf16x3 split products against the fp32 oracle -> Q-values / losses rtol 1e-4,
parameters 1e-4 relative L2, gradients / first moments 3e-4, second moments 6e-4, with the GPU's relu and max-pool decisions
handed to the oracle (both are discontinuities of the gradient; every decision that differs from the float64 oracle's is
asserted to sit at a numerical tie)."""
import numpy as np
import pytest
import torch

from oracle import networks as O

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den > 0 else np.linalg.norm(a)


def assert_trees_close(got, want, tol, what, path=""):
    if isinstance(want, dict):
        assert set(got) == set(want), f"{what} {path}: modules {sorted(got)} != {sorted(want)}"
        for k in want:
            assert_trees_close(got[k], want[k], tol, what, f"{path}/{k}")
        return
    e = rel_l2(got, want)
    assert e <= tol, f"{what} {path}: rel-L2 {e:.3e} > {tol}"


def batch_of(rng, B, obs, A, u8):
    if u8:
        s, s2 = (rng.integers(0, 256, (B,) + obs).astype(np.uint8) for _ in range(2))
    else:
        s, s2 = (rng.uniform(0, 255, (B,) + obs).astype(np.float32) for _ in range(2))
    return dict(state=s, next_state=s2, action=rng.integers(0, A, B).astype(np.int32),
                reward=rng.uniform(-1, 1, B).astype(np.float32), is_terminal=(rng.random(B) < 0.2))


def test_impala_tree_layout_and_apply():
    """Parameter tree = flax's nested auto-naming (Stack_i/Conv_j, Dense_k) in creation order; Q-values of every head on
    an odd-sized observation (SAME padding (0,1) and (1,1) cases of the pool) against the oracle."""
    from idqn_b200 import _lib as L
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K = (37, 29, 3), [4, 6, 3, 10], 5, 2
    agent = iDQN(3, obs, A, K, feats, "impala", 1e-3, 0.94, 1, 1, 1, 1, batch_size=8)
    tree = agent.params.to_host()["params"]
    assert sorted(tree) == ["Dense_0", "Dense_1", "Stack_0", "Stack_1", "Stack_2"]
    assert sorted(tree["Stack_1"]) == [f"Conv_{j}" for j in range(5)]
    assert tree["Stack_0"]["Conv_0"]["kernel"].shape == (K, 3, 3, 3, 4) and tree["Stack_1"]["Conv_0"]["kernel"].shape == (K, 3, 3, 4, 6)
    assert tree["Stack_2"]["Conv_3"]["kernel"].shape == (K, 3, 3, 3, 3) and tree["Dense_0"]["kernel"].shape == (K, 5 * 4 * 3, 10)
    assert [l["module"] for l in agent._engine.leaves[::2]] == [n for n, _, _ in O.layer_shapes(obs, feats, "impala", A)]
    assert all(np.all(np.asarray(tree[f"Stack_{i}"][f"Conv_{j}"]["bias"]) == 0) for i in range(3) for j in range(5))
    rng = np.random.default_rng(5)
    params = O.init_params(rng, obs, feats, "impala", A, n_networks=K, bias_scale=0.1)
    agent.params = params
    for x in (rng.uniform(0, 255, (5,) + obs).astype(np.float32), rng.integers(0, 256, (3,) + obs).astype(np.uint8)):
        for k in range(K):
            got = agent._engine.apply(L.ONLINE, k, x)
            want = O.apply(O.tree_index(params, k), x, "impala")
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # the model pickle (experiments/base/utils.py:123-134: pickle.dump(agent.get_model())) round-trips through a new agent
    import pickle
    model = pickle.loads(pickle.dumps(agent.get_model()))
    other = iDQN(99, obs, A, K, feats, "impala", 1e-3, 0.94, 1, 1, 1, 1, batch_size=8)
    other.params = model["params"]
    np.testing.assert_array_equal(other._engine.apply(L.ONLINE, 1, x), agent._engine.apply(L.ONLINE, 1, x))


def gpu_gates(agent, k, z64):
    """The relu decisions the GPU took for online head k, in the order the oracle applies its relus (per Stack and block: the
    block input, the first conv's output; then the Stack_2 output; then the dense trunk), read from the stored activations.
    Where they differ from sign(z) of the float64 oracle the unit must be numerically zero."""
    eng = agent._engine
    layers = []  # engine layer index behind every relu
    for st in range(3):
        base = 6 * st
        layers += [base + 1, base + 2, base + 3, base + 4]  # block 1: pool output, Conv_1; block 2: Conv_2 output, Conv_3
    layers += [17, 18]  # relu(Stack_2 output) = Conv_4 of the last stack; Dense_0
    gates, flips = [], 0
    for li, z in zip(layers, z64):
        g = eng.download_activation(k, li).reshape(z.shape) > 0
        diff = g != (z > 0)
        if diff.any():
            assert np.abs(z[diff]).max() <= 2e-5 * max(1.0, np.abs(z).max()), f"layer {li}: gate differs at a clearly non-zero unit"
            flips += int(diff.sum())
        gates.append(g)
    return gates, flips


def gpu_pools(agent, k, pool_in64):
    """The window element every max-pool output of online head k took (first maximum of the GPU's own Conv_0 output, the rule
    of maxpool_fwd / maxpool_bwd_kernel).  Where it differs from the float64 oracle's choice, the two elements must be
    numerically equal: like a relu at zero, the pool is discontinuous in its gradient at a tie."""
    args, flips = [], 0
    for st, z in enumerate(pool_in64):
        zg = agent._engine.download_activation(k, 6 * st).reshape(z.shape)
        a_g, w64 = O.pool_argmax_same(zg), O.pool_windows_same(z)
        a_o = np.argmax(w64, axis=4)
        diff = a_g != a_o
        if diff.any():
            gap = np.take_along_axis(w64, a_o[..., None], 4)[..., 0] - np.take_along_axis(w64, a_g[..., None], 4)[..., 0]
            assert gap[diff].max() <= 2e-5 * max(1.0, np.abs(z).max()), f"Stack_{st}: pool choice differs at a clear maximum"
            flips += int(diff.sum())
        args.append(a_g)
    return args, flips


def check_learning_steps(obs, feats, A, K, B, steps, u8, seed, lr=1e-3, eps=1e-5, T=2, D=1):
    """`steps` learning steps (+ the schedule's target events) of a K-head impala agent: losses, every gradient, parameters and
    both Adam moments against the fp32 oracle restarted from the GPU's state each step and handed the GPU's relu and max-pool
    decisions (they may differ from the oracle's only at numerically-zero units / numerically-equal window elements: checked
    against the float64 oracle's pre-activations and pool inputs)."""
    from idqn_b200 import _lib as L
    from idqn_b200.networks.idqn import iDQN
    rng = np.random.default_rng(seed)
    agent = iDQN(0, obs, A, K, feats, "impala", lr, 0.94, 1, 1, T, D, eps, batch_size=B, flags=L.F_KEEP_GRADS)
    agent.params = O.init_params(rng, obs, feats, "impala", A, n_networks=K, bias_scale=0.05)
    agent.target_params = O.init_params(rng, obs, feats, "impala", A, n_networks=K, bias_scale=0.05)
    for step in range(1, steps + 1):
        batch = batch_of(rng, B, obs, A, u8)
        st = agent.optimizer_state[0]
        s_p, s_t = agent.params.to_host(), agent.target_params.to_host()
        s_o = {"count": np.asarray(st.count).copy(), "mu": st.mu.to_host(), "nu": st.nu.to_host()}
        _, _, g_l = agent.learn_on_batch(agent.params, agent.target_params, agent.optimizer_state, batch)
        gates, pools = [], []
        for k in range(K):
            _, _, z64, pin64 = O.loss_and_grad(O.tree_index(s_p, k), O.tree_index(s_t, k), batch, "impala", 0.94, 1, torch.float64,
                                               preacts=True, pool_inputs=True)
            gates.append(gpu_gates(agent, k, z64)[0])
            pools.append(gpu_pools(agent, k, pin64)[0])
        o_p, o_s, o_l, o_g = O.learn_on_batch(s_p, s_t, s_o, batch, "impala", 0.94, 1, lr, eps, torch.float32, return_grads=True,
                                              gates=gates, pools=pools)
        np.testing.assert_allclose(g_l, o_l, rtol=1e-4, err_msg=f"losses step {step}")
        assert_trees_close(agent.gradients(), o_g, 3e-4, f"grad step {step}")
        assert_trees_close(agent.params.to_host(), o_p, 1e-4, f"params step {step}")
        st = agent.optimizer_state[0]
        assert_trees_close(st.mu.to_host(), o_s["mu"], 3e-4, f"mu step {step}")
        assert_trees_close(st.nu.to_host(), o_s["nu"], 6e-4, f"nu step {step}")
        np.testing.assert_array_equal(np.asarray(st.count), o_s["count"])
        s_p, s_t = agent.params.to_host(), agent.target_params.to_host()
        updated, _ = agent.update_target_params(step)
        ev = O.ScheduleOracle(1, T, D).events(step)
        assert updated == ("T" in ev)
        if "T" in ev:
            s_t, s_p = O.tree_map(np.copy, s_p), O.shift_params(s_p)
        elif "D" in ev:
            s_t = O.sync_target_params(s_p, s_t)
        assert_trees_close(agent.target_params.to_host(), s_t, 0.0, f"target after schedule step {step}")
        assert_trees_close(agent.params.to_host(), s_p, 0.0, f"params after schedule step {step}")


@pytest.mark.parametrize("u8", [False, True])
def test_impala_learning_step_matches_oracle(u8):
    """Three steps (T = 2: one target update + window shift) of a K = 2 agent on a small net whose channel counts (8, 6, 8)
    take the tcgen05, the CUDA-core, the vectorised and the generic pool kernels."""
    check_learning_steps((22, 20, 4), [8, 6, 8, 16], 4, 2, 8, 3, u8, 11 + int(u8))


def test_impala_atari_size_learning_step_matches_oracle():
    """The same comparison at the sizes the bench times (84 x 84 x 4 uint8 frames, features 32 / 64 / 64 / 512, batch 32,
    A = 6): one step of a K = 2 agent with a D-sync after it -- every gradient of the full-size tcgen05 tiles / split-K
    reductions, not only the losses."""
    check_learning_steps((84, 84, 4), [32, 64, 64, 512], 6, 2, 32, 1, True, 31, lr=3e-4, eps=1.5e-4, T=200, D=1)


def test_reference_unit_tests_of_idqn_on_impala():
    """tests/test_idqn.py:44-84 of the reference, restated call for call: a random impala net on (84, 84, 4) with 1..9
    features per layer, K in 1..9, 2..9 actions, gamma 0.94, float32 states in [0, 1) -- compute_target == the inline
    formula, loss == (target - prediction)^2, best_action == argmax of the head drawn from the key.  Exact equality as in
    the reference (both sides of each assertion run the same deterministic kernels); the oracle is checked on top."""
    from idqn_b200 import _prng
    from idqn_b200.networks.idqn import iDQN
    for seed in (7, 481):
        rng = np.random.default_rng(seed)
        A, K = int(rng.integers(2, 10)), int(rng.integers(1, 10))
        feats = [int(rng.integers(1, 10)) for _ in range(4)]
        obs = (84, 84, 4)
        q = iDQN(seed, obs, A, K, feats, "impala", 0.001, 0.94, 1, 1, 1, 1)
        sample = dict(state=rng.random(obs, np.float32), action=int(rng.integers(0, A)), reward=np.float32(rng.random()),
                      next_state=rng.random(obs, np.float32), is_terminal=int(rng.integers(0, 2)))
        idx = int(rng.integers(0, K))
        host = q.params.to_host()
        params_k = O.tree_index(host, idx)
        # test_compute_target
        computed_target = q.compute_target(params_k, sample)
        next_q = q.network.apply(params_k, sample["next_state"])
        target = sample["reward"] + (1 - sample["is_terminal"]) * np.float32(q.gamma) * np.max(next_q)
        assert next_q.shape == (A,)
        assert np.float32(target) == computed_target
        np.testing.assert_allclose(next_q, O.apply(params_k, sample["next_state"][None], "impala")[0], rtol=1e-4, atol=1e-6)
        # test_loss
        computed_loss = q.loss(params_k, params_k, sample)
        prediction = q.network.apply(params_k, sample["state"])[sample["action"]]
        assert np.square(q.compute_target(params_k, sample) - prediction) == computed_loss
        # test_best_action
        key = _prng.PRNGKey(seed) if hasattr(_prng, "PRNGKey") else seed
        state = rng.random(obs, np.float32)
        computed = q.best_action(q.params, state, key)
        head = _prng.randint(key, 0, K)
        q_values = q.network.apply(O.tree_index(host, head), state)
        assert q_values.shape == (A,) and int(np.argmax(q_values)) == computed


def test_impala_at_atari_size_takes_learning_steps():
    """The architecture at the sizes the reference's Atari launch scripts would give it (84 x 84 x 4 uint8 frames,
    features 32 / 64 / 64 / 512, batch 32): two steps of a K = 1 agent, losses against the oracle, loss decreasing on a
    repeated batch."""
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, B = (84, 84, 4), [32, 64, 64, 512], 6, 32
    rng = np.random.default_rng(3)
    agent = iDQN(1, obs, A, 1, feats, "impala", 1e-4, 0.99, 1, 1, 100, 100, 1.5e-4)
    batch = batch_of(rng, B, obs, A, True)
    p0, t0 = agent.params.to_host(), agent.target_params.to_host()
    want = O.loss_on_batch(O.tree_index(p0, 0), O.tree_index(t0, 0), batch, "impala", 0.99, 1)
    losses = [float(agent._engine.learn_host(batch, want_losses=True)[0]) for _ in range(3)]
    np.testing.assert_allclose(losses[0], want, rtol=1e-4)
    assert np.isfinite(losses).all() and losses[2] < losses[0]
