"""GPU parity of the learning step (through the C ABI via the iDQN/DQN classes) against the CPU oracle.

Tolerance (north_star: "within 1e-4 relative (fp32)"): per-head losses rtol 1e-4; every gradient / parameter /
Adam-moment tensor within 1e-4 relative L2 of the fp32 oracle (the two fp32 implementations differ only by
summation order; the fp64 oracle is used to show both sit at the same distance from the truth)."""
import numpy as np
import pytest
import torch

from oracle import networks as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den > 0 else np.linalg.norm(a)


def assert_tree_close(got, want, tol, what):
    for mod in want["params"]:
        for leaf in want["params"][mod]:
            e = rel_l2(got["params"][mod][leaf], want["params"][mod][leaf])
            assert e <= tol, f"{what} {mod}/{leaf}: rel-L2 {e:.3e} > {tol}"


def make_batch(rng, B, obs, A, u8):
    if u8:
        s = rng.integers(0, 256, (B,) + obs).astype(np.uint8)
        s2 = rng.integers(0, 256, (B,) + obs).astype(np.uint8)
        r = rng.integers(-1, 2, B).astype(np.float32)
    else:
        s = rng.standard_normal((B,) + obs).astype(np.float32)
        s2 = rng.standard_normal((B,) + obs).astype(np.float32)
        r = rng.uniform(-1, 1, B).astype(np.float32)
    return dict(state=s, next_state=s2, action=rng.integers(0, A, B).astype(np.int32), reward=r,
                is_terminal=(rng.random(B) < 0.1))


def gpu_gates(agent, k, z64_layers):
    """Gate decisions (relu output > 0) the GPU took for online head k, checked against the fp64 pre-activations:
    wherever they disagree with sign(z64) the unit must be numerically zero (the relu derivative is discontinuous
    there and any fp32 implementation — the reference's included — may fall on either side)."""
    gates, flips = [], 0
    for li, z in enumerate(z64_layers):
        g = agent._engine.download_activation(k, li).reshape(z.shape) > 0
        diff = g != (z > 0)
        if diff.any():
            assert np.abs(z[diff]).max() <= 1e-5 * max(1.0, np.abs(z).max()), f"layer {li}: gate differs at a clearly non-zero unit"
            flips += int(diff.sum())
        gates.append(g)
    return gates, flips


def run_parity(arch, obs, feats, A, K, steps, T, D, lr, eps, u8, seed=0, flags=0, B=32, check_grads=True):
    """Teacher-forced parity: before every step the oracle is loaded with the GPU's own state, both take one
    step on the same batch, and loss / gradients / params / mu / nu / count are compared; then the T/D target
    events are applied on both sides and compared exactly.  The oracle is given the GPU's relu gates (see
    gpu_gates) so that units with a numerically zero pre-activation cannot break the 1e-4 comparison."""
    from idqn_b200 import _lib
    from idqn_b200.networks.idqn import iDQN
    rng = np.random.default_rng(seed)
    params = O.init_params(rng, obs, feats, arch, A, n_networks=K, bias_scale=0.01)
    target = O.init_params(np.random.default_rng(seed + 1000), obs, feats, arch, A, n_networks=K, bias_scale=0.01)
    if check_grads:  # materialise every gradient (the production path fuses Adam into the Dense_0 wgrad epilogue)
        flags |= _lib.F_KEEP_GRADS
    agent = iDQN(0, obs if arch == "cnn" else obs[0], A, K, feats, arch, lr, 0.99, 1, 1, T, D, eps, batch_size=B,
                 flags=flags)
    agent.params = params
    agent.target_params = target
    total_flips = 0
    for step in range(1, steps + 1):
        batch = make_batch(rng, B, obs if arch == "cnn" else obs + (1,), A, u8)
        st = agent.optimizer_state[0]
        s_p, s_t = agent.params.to_host(), agent.target_params.to_host()
        s_o = {"count": np.asarray(st.count).copy(), "mu": st.mu.to_host(), "nu": st.nu.to_host()}
        _, _, g_l = agent.learn_on_batch(agent.params, agent.target_params, agent.optimizer_state, batch)
        gates = []
        for k in range(K):
            _, _, z64 = O.loss_and_grad(O.tree_index(s_p, k), O.tree_index(s_t, k), batch, arch, 0.99, 1,
                                        torch.float64, preacts=True)
            gk, flips = gpu_gates(agent, k, z64)
            gates.append(gk)
            total_flips += flips
        o_p, o_s, o_l, o_g = O.learn_on_batch(s_p, s_t, s_o, batch, arch, 0.99, 1, lr, eps, torch.float32,
                                              return_grads=True, gates=gates)
        np.testing.assert_allclose(g_l, o_l, rtol=RTOL, err_msg=f"losses step {step}")
        if check_grads:
            assert_tree_close(agent.gradients(), o_g, RTOL, f"grad step {step}")
        assert_tree_close(agent.params.to_host(), o_p, RTOL, f"params step {step}")
        st = agent.optimizer_state[0]
        assert_tree_close(st.mu.to_host(), o_s["mu"], RTOL, f"mu step {step}")
        assert_tree_close(st.nu.to_host(), o_s["nu"], 2 * RTOL, f"nu step {step}")
        np.testing.assert_array_equal(np.asarray(st.count), o_s["count"])
        # schedule (idqn.py:74-94): exact copies on both sides
        s_p, s_t = agent.params.to_host(), agent.target_params.to_host()
        updated, logs = agent.update_target_params(step)
        ev = O.ScheduleOracle(1, T, D).events(step)
        assert updated == ("T" in ev)
        if "T" in ev:
            s_t = O.tree_map(np.copy, s_p)
            s_p = O.shift_params(s_p)
        elif "D" in ev:
            s_t = O.sync_target_params(s_p, s_t)
        assert_tree_close(agent.target_params.to_host(), s_t, 0.0, f"target after schedule step {step}")
        assert_tree_close(agent.params.to_host(), s_p, 0.0, f"params after schedule step {step}")
    print(f"[parity {arch} K={K}] {steps} steps, {total_flips} numerically-zero gate disagreements with fp64")
    return agent


def test_mlp_k3_ten_steps_with_schedule():
    """configs[0]: Lunar Lander i-DQN K=3 MLP batch 32; T=8, D=4 so the 10 steps include one D-sync and one T-shift."""
    run_parity("fc", (8,), [100, 100], 4, 3, 10, 8, 4, 3e-4, 1e-8, u8=False)


def test_mlp_no_graph_matches():
    from idqn_b200 import _lib
    run_parity("fc", (8,), [100, 100], 4, 2, 3, 8, 4, 3e-4, 1e-8, u8=False, flags=_lib.F_NO_GRAPH)


def test_nature_cnn_dqn_k1():
    """configs[1]: Atari NatureCNN DQN (K=1) batch 32, uint8 84x84x4 frames."""
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 1, 2, 8, 4, 3e-4, 1.5e-4, u8=True)


def test_nature_cnn_idqn_k3():
    """configs[2]: Atari NatureCNN i-DQN K=3; 4 steps with T=4, D=2 (one D-sync, one T-shift)."""
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 3, 4, 4, 2, 3e-4, 1.5e-4, u8=True)


def test_nature_cnn_production_path_without_materialised_grads():
    """default flags: Adam fused into the Dense_0 wgrad epilogue, gradient of that kernel never written."""
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 2, 3, 2, 3, 3e-4, 1.5e-4, u8=True, check_grads=False)


def test_nature_cnn_generic_tensor_path():
    """IDQN_F_NO_IMG: the generic tcgen05 implicit-GEMM kernels (gemm_tc.cuh) instead of the image-resident TMA
    conv kernels and the weight-streaming Dense_0 kernels stay a valid implementation of the same step."""
    from idqn_b200 import _lib
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 2, 2, 2, 3, 3e-4, 1.5e-4, u8=True, flags=_lib.F_NO_IMG)


def test_nature_cnn_without_programmatic_dependent_launch():
    """IDQN_F_NO_PDL: the step's kernels launched without programmatic stream serialization (griddepcontrol is the
    default) give the same step."""
    from idqn_b200 import _lib
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 3, 3, 2, 3, 3e-4, 1.5e-4, u8=True, flags=_lib.F_NO_PDL,
               check_grads=False)


def test_nature_cnn_backward_variants():
    """The Dense_0 wgrad+Adam runs as a TMA pipeline (dense_wgrad_tma.cuh) by default; the generic tcgen05 kernel with
    Adam in its epilogue (IDQN_F_OLD_WGRAD), the un-graphed order, the single-branch graph (IDQN_F_NO_FORK), and the backward pass split over two SM partitions
    (green contexts, IDQN_F_PARTITION, graphed and un-graphed) are the same step."""
    from idqn_b200 import _lib
    for flags in (_lib.F_NO_GRAPH, _lib.F_OLD_WGRAD, _lib.F_NO_FORK, _lib.F_NO_DEFER, _lib.F_PARTITION, _lib.F_PARTITION | _lib.F_NO_GRAPH,
                  _lib.F_PARTITION | _lib.F_OLD_WGRAD):
        run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 2, 2, 2, 3, 3e-4, 1.5e-4, u8=True, flags=flags)


def test_nature_cnn_k5_production_path():
    """K = 5 (the benchmark configuration: two head groups on the first layer, split-K Dense_0 summed in the head kernel)."""
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 5, 2, 8, 2, 3e-4, 1.5e-4, u8=True, check_grads=False)


def test_nature_cnn_k8_production_path():
    """configs[3] on one GPU: K = 8 heads (three head groups on the first layer, 16 Dense_0 nets)."""
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 8, 1, 8, 2, 3e-4, 1.5e-4, u8=True, check_grads=False)


def test_nature_cnn_simt_cross_check():
    """the exact-fp32 CUDA-core path (IDQN_F_SIMT_ONLY) stays a valid implementation of the same step."""
    from idqn_b200 import _lib
    run_parity("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 2, 2, 2, 3, 3e-4, 1.5e-4, u8=True, flags=_lib.F_SIMT_ONLY)


def test_nature_cnn_k5_a18_float_states():
    """K=5 (N = 160 on the concatenated first layer), 18 actions, float32 frames (two bf16 planes for conv0)."""
    from idqn_b200.networks.idqn import iDQN  # noqa: F401
    import oracle.networks  # noqa: F401
    rng = np.random.default_rng(11)
    obs, feats, A, K = (84, 84, 4), [32, 64, 64, 512], 18, 5
    from idqn_b200 import _lib
    agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4, flags=_lib.F_KEEP_GRADS)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    target = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    agent.params, agent.target_params = params, target
    batch = make_batch(rng, 32, obs, A, True)
    batch["state"] = batch["state"].astype(np.float32)
    batch["next_state"] = batch["next_state"].astype(np.float32)
    _, _, g_l = agent.learn_on_batch(agent.params, agent.target_params, agent.optimizer_state, batch)
    gates = []
    for k in range(K):
        _, _, z64 = O.loss_and_grad(O.tree_index(params, k), O.tree_index(target, k), batch, "cnn", 0.99, 1,
                                    torch.float64, preacts=True)
        gates.append(gpu_gates(agent, k, z64)[0])
    o_p, o_s, o_l, o_g = O.learn_on_batch(params, target, O.init_optimizer_state(params), batch, "cnn", 0.99, 1, 3e-4,
                                          1.5e-4, torch.float32, return_grads=True, gates=gates)
    np.testing.assert_allclose(g_l, o_l, rtol=RTOL)
    assert_tree_close(agent.gradients(), o_g, RTOL, "grad")
    assert_tree_close(agent.params.to_host(), o_p, RTOL, "params")


def test_small_cnn_odd_shapes():
    """ragged geometry: non-square frames, channel counts that are not multiples of the tile sizes, A=18, B=5."""
    run_parity("cnn", (37, 50, 3), [5, 7, 9, 33], 18, 2, 3, 2, 1, 1e-3, 1e-8, u8=True, B=5)


def test_cnn_float_inputs():
    """float32 states holding 0..255 (atari.py:43-45 as seen by best_action) go through the same /255 path."""
    from idqn_b200.networks.idqn import iDQN
    rng = np.random.default_rng(3)
    obs, feats, A, K = (84, 84, 4), [32, 64, 64, 512], 6, 2
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4)
    agent.params = params
    for k in range(K):
        state = rng.integers(0, 256, obs).astype(np.float32)
        q = O.apply(O.tree_index(params, k), state[None], "cnn")[0]
        assert agent.best_action(agent.params, state, idx_params=k) == int(np.argmax(q))
        got = agent._engine.apply(0, k, state)[0]
        np.testing.assert_allclose(got, q, rtol=1e-4, atol=1e-5)
        got_u8 = agent._engine.apply(0, k, state.astype(np.uint8))[0]
        np.testing.assert_allclose(got_u8, q, rtol=1e-4, atol=1e-5)


def test_functional_learn_on_batch_is_pure_and_dqn_api():
    from idqn_b200.networks.dqn import DQN
    rng = np.random.default_rng(5)
    obs, feats, A = (8,), [32, 32], 4
    agent = DQN(7, 8, A, feats, "fc", 1e-3, 0.99, 1, 1, 4, 1e-8)
    before = agent.params.to_host()
    p = O.init_params(rng, obs, feats, "fc", A)
    t = O.init_params(rng, obs, feats, "fc", A)
    st = {"count": np.int32(0), "mu": O.tree_map(np.zeros_like, p), "nu": O.tree_map(np.zeros_like, p)}
    batch = make_batch(rng, 32, obs + (1,), A, False)
    new_p, new_s, loss = agent.learn_on_batch(p, t, st, batch)
    o_p, o_s, o_l = O.learn_on_batch_dqn(p, t, st, batch, "fc", 0.99, 1, 1e-3, 1e-8)
    np.testing.assert_allclose(loss, o_l, rtol=RTOL)
    assert_tree_close(new_p, o_p, RTOL, "functional params")
    assert int(new_s[0].count) == 1
    after = agent.params.to_host()
    for m in before["params"]:
        np.testing.assert_array_equal(before["params"][m]["kernel"], after["params"][m]["kernel"])
    # reference-style self-consistency checks (tests/test_dqn.py:39-73)
    sample = {k: v[0] for k, v in batch.items()}
    tgt = agent.compute_target(p, sample)
    q_next = agent.network.apply(p, sample["next_state"])
    assert q_next.shape == (A,)
    assert tgt == np.float32(sample["reward"]) + np.float32(1 - int(sample["is_terminal"])) * np.float32(0.99) * np.max(q_next)
    pred = agent.network.apply(p, sample["state"])[sample["action"]]
    assert agent.loss(p, p, sample) == np.square(pred - agent.compute_target(p, sample))
    assert agent.best_action(p, sample["state"]) == int(np.argmax(agent.network.apply(p, sample["state"])))
    upd, logs = agent.update_target_params(4)
    assert upd and "loss" in logs
    assert agent.update_target_params(3) == (False, {})


class _ListBuffer:
    """Stand-in replay buffer for update_online_params (idqn.py:65-72): hands out prepared batches in order."""

    def __init__(self, batches):
        self._batches, self._i = batches, 0

    def sample(self):
        b = self._batches[self._i]
        self._i += 1
        return b


def test_loss_logging_and_cumulated_losses():
    """idqn.py:72,82-87: update_online_params feeds cumulated_losses, update_target_params logs and resets them; a
    direct learn_on_batch call (idqn.py:96-109) returns the losses and leaves the running sums alone."""
    from idqn_b200.networks.idqn import iDQN
    rng = np.random.default_rng(9)
    agent = iDQN(1, 8, 4, 3, [16], "fc", 1e-3, 0.99, 1, 1, 4, 2, 1e-8)
    twin = iDQN(1, 8, 4, 3, [16], "fc", 1e-3, 0.99, 1, 1, 4, 2, 1e-8)  # same key: same parameters
    batches = [make_batch(rng, 32, (8, 1), 4, False) for _ in range(4)]
    rb = _ListBuffer(batches)
    tot = np.zeros(3)
    for step in range(1, 5):
        agent.update_online_params(step, rb)
        _, _, l = twin.learn_on_batch(twin.params, twin.target_params, twin.optimizer_state, batches[step - 1])
        tot += l
        if step < 4:
            assert agent.update_target_params(step)[0] is False
            twin.update_target_params(step)
    assert (twin.cumulated_losses == 0).all(), "learn_on_batch must not feed cumulated_losses"
    np.testing.assert_allclose(agent.cumulated_losses, tot, rtol=1e-6)
    upd, logs = agent.update_target_params(4)
    assert upd
    np.testing.assert_allclose(logs["loss"], tot.mean() / 4, rtol=1e-6)
    np.testing.assert_allclose(logs["networks/2_loss"], tot[2] / 4, rtol=1e-6)
    assert (agent.cumulated_losses == 0).all()


def test_pipelined_host_submission_equals_blocking_calls():
    """idqn_submit_batch_host / idqn_wait_losses (H2D of batch t+1 on the copy stream while step t computes, losses
    read one step behind) give bit-identical losses and parameters to the blocking idqn_learn_on_batch_host."""
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B = (84, 84, 4), [32, 64, 64, 512], 6, 2, 32
    rng = np.random.default_rng(3)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    target = O.init_params(np.random.default_rng(1003), obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    batches = [make_batch(rng, B, obs, A, True) for _ in range(5)]
    out = []
    for pipelined in (False, True):
        agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4)
        agent.params, agent.target_params = params, target
        eng, losses, pending = agent._engine, [], None
        for b in batches:
            if pipelined:
                t = eng.submit_host(b)
                if pending is not None:
                    losses.append(eng.wait_losses(pending))
                pending = t
            else:
                losses.append(eng.learn_host(b, want_losses=True))
        if pipelined:
            losses.append(eng.wait_losses(pending))
            with pytest.raises(Exception):
                eng.wait_losses(0)  # only the two most recent tickets can be waited on
        out.append((np.stack(losses), agent.params.to_host()))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    assert_tree_close(out[0][1], out[1][1], 0.0, "params after pipelined steps")


def test_k5_steps_are_bit_reproducible():
    """Race detector for the step graph (two branches, programmatic dependent launch, split-K partials summed in a
    fixed order, no floating-point atomics): 60 graph-replayed K=5 steps with T/D target events, run twice from the
    same state on the same batches, give bit-identical losses, parameters and Adam moments."""
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B = (84, 84, 4), [32, 64, 64, 512], 6, 5, 32
    rng = np.random.default_rng(11)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    target = O.init_params(np.random.default_rng(1011), obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    batches = [make_batch(rng, B, obs, A, True) for _ in range(6)]
    runs = []
    for _ in range(2):
        agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 20, 5, 1.5e-4)
        agent.params, agent.target_params = params, target
        eng, losses = agent._engine, []
        for step in range(1, 61):
            losses.append(eng.learn_host(batches[step % len(batches)], want_losses=True))
            agent.update_target_params(step)
        st = agent.optimizer_state[0]
        runs.append((np.stack(losses), agent.params.to_host(), st.mu.to_host(), st.nu.to_host()))
    assert np.isfinite(runs[0][0]).all()
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    for i, what in ((1, "params"), (2, "mu"), (3, "nu")):
        assert_tree_close(runs[0][i], runs[1][i], 0.0, f"{what} after 60 steps")


def test_freshly_constructed_agent_bootstraps_from_its_own_init():
    """idqn.py:56 ``target_params = params``: an agent used straight from its constructor (no explicit assignment of
    target_params) must bootstrap every head from its own initial parameters -- the bf16 planes of the target arena have
    to follow the constructor's upload(ONLINE) + copy_online_to_target(), and stay right across a D-sync."""
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B = (84, 84, 4), [32, 64, 64, 512], 6, 3, 32
    rng = np.random.default_rng(21)
    agent = iDQN(7, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 100, 2, 1.5e-4)
    for step in (1, 2, 3):
        p0, t0 = agent.params.to_host(), agent.target_params.to_host()
        if step == 1:
            assert_tree_close(t0, p0, 0.0, "target == online after construction")
        st = agent.optimizer_state[0]
        s_o = {"count": np.asarray(st.count).copy(), "mu": st.mu.to_host(), "nu": st.nu.to_host()}
        batch = make_batch(rng, B, obs, A, True)
        _, _, g_l = agent.learn_on_batch(agent.params, agent.target_params, agent.optimizer_state, batch)
        _, _, o_l = O.learn_on_batch(p0, t0, s_o, batch, "cnn", 0.99, 1, 3e-4, 1.5e-4, torch.float32)
        np.testing.assert_allclose(g_l, o_l, rtol=RTOL, err_msg=f"losses step {step}")
        agent.update_target_params(step)  # D-sync at step 2


def test_best_action_fast_path_equals_generic_path():
    """best_action on a uint8 (or integral float32) Atari state runs through the learning step's own kernels; the
    generic batch-1 kernels (IDQN_F_SLOW_APPLY) and the oracle's argmax give the same action for every head, also
    right after a learning step (the fast path shares the step's activation buffers)."""
    from idqn_b200 import _lib
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B = (84, 84, 4), [32, 64, 64, 512], 6, 5, 32
    rng = np.random.default_rng(31)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    agents = []
    for flags in (0, _lib.F_SLOW_APPLY):
        ag = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4, flags=flags)
        ag.params, ag.target_params = params, params
        agents.append(ag)
    batch = make_batch(rng, B, obs, A, True)
    for rnd in range(2):
        for trial in range(6):
            state = rng.integers(0, 256, obs).astype(np.uint8)
            for head in range(K):
                fast = agents[0].best_action(agents[0].params, state, idx_params=head)
                slow = agents[1].best_action(agents[1].params, state, idx_params=head)
                as_float = agents[0].best_action(agents[0].params, state.astype(np.float32), idx_params=head)
                q = O.apply(O.tree_index(agents[0].params.to_host(), head), state[None], "cnn")
                top2 = np.sort(np.asarray(q).ravel())[-2:]
                if top2[1] - top2[0] > 1e-5 * max(1.0, abs(top2[1])):  # a clear maximum: all paths must agree
                    assert fast == slow == as_float == int(np.argmax(q)), (rnd, trial, head, fast, slow, q)
        for ag in agents:  # one learning step, then again: the step's buffers were reused in between
            ag._engine.learn_host(batch, want_losses=False)


def test_head_results_do_not_depend_on_how_many_heads_share_the_gpu():
    """Every reduction of the step (conv weight-gradient partial sums, Dense_0 split-K, batch sums) is grouped by the
    batch / layer shape only: head k of a K=3 agent and the same head alone in a K=1 agent (its target events fed by
    hand from its neighbours, as parallel.py does across GPUs) stay BIT-IDENTICAL over 6 steps with one D-sync and one
    T-shift.  This is the single-GPU proof that the head-sharded chain equals the unsharded one."""
    from idqn_b200 import _lib as L
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B, T, D = (84, 84, 4), [32, 64, 64, 512], 6, 3, 32, 4, 2
    rng = np.random.default_rng(41)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    target = O.init_params(np.random.default_rng(1041), obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    full = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4)
    full.params, full.target_params = params, target
    solo = []
    for k in range(K):
        ag = iDQN(0, obs, A, 1, feats, "cnn", 3e-4, 0.99, 1, 1, T, D, 1.5e-4)
        ag.params = O.tree_map(lambda a: a[k:k + 1], params)
        ag.target_params = O.tree_map(lambda a: a[k:k + 1], target)
        solo.append(ag)
    for step in range(1, 7):
        batch = make_batch(rng, B, obs, A, True)
        lf = full._engine.learn_host(batch, want_losses=True)
        ls = np.concatenate([ag._engine.learn_host(batch, want_losses=True) for ag in solo])
        np.testing.assert_array_equal(lf, ls, err_msg=f"losses step {step}")
        full.update_target_params(step)
        # the same events on the one-head agents, boundary heads moved by hand (idqn.py:13-24,74-94)
        on = [ag.params.to_host() for ag in solo]
        if step % T == 0:
            for k, ag in enumerate(solo):
                ag.target_params = on[k]
                ag.params = on[min(k + 1, K - 1)]
        elif step % D == 0:
            for k, ag in enumerate(solo):
                if k > 0:
                    ag.target_params = on[k - 1]
    for which, name in ((L.ONLINE, "online"), (L.TARGET, "target"), (L.MU, "mu"), (L.NU, "nu")):
        f = full._engine.download_arena(which)
        for k, ag in enumerate(solo):
            np.testing.assert_array_equal(f[k], ag._engine.download_arena(which)[0], err_msg=f"{name} head {k}")


def test_more_actions_than_the_head_kernels_hold_is_refused():
    from idqn_b200.networks.idqn import iDQN
    with pytest.raises(ValueError):
        iDQN(0, 8, 33, 2, [16], "fc", 1e-3, 0.99, 1, 1, 4, 2, 1e-8)


def test_select_action_in_one_call_equals_the_host_mirror():
    """utils.py:8-15 through idqn_select_action (C draws + graph-launched best_action of the drawn head) against the
    Python mirror (split / uniform / randint of _prng, then best_action with the explicit head): same action, same
    explore decision, same head, for greedy and exploring steps."""
    from idqn_b200 import _prng
    from idqn_b200.networks.idqn import iDQN
    from idqn_b200.sample_collection.utils import select_action
    obs, feats, A, K = (84, 84, 4), [32, 64, 64, 512], 6, 5
    rng = np.random.default_rng(51)
    agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4)
    agent.params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    explored = 0
    for trial in range(40):
        key = _prng.split(1000 + trial, 1)[0]
        state = rng.integers(0, 256, obs).astype(np.float32 if trial % 2 else np.uint8)
        eps = 0.5
        u_key, a_key, kw_key = _prng.split(key, 3)
        want_explore = bool(_prng.uniform(u_key) <= eps)
        want_head = _prng.randint(kw_key, 0, K)
        want = _prng.randint(a_key, 0, A) if want_explore else agent.best_action(agent.params, state, idx_params=want_head)
        got, got_explore, got_head = agent._engine.select_action(state, key, A, eps)
        assert (got, got_explore, got_head) == (want, want_explore, want_head), trial
        assert select_action(agent.best_action, agent.params, state, key, A, lambda n: eps, 0) == want
        explored += want_explore
    assert 5 < explored < 35


def test_training_loop_mirror_runs_the_reference_schedule():
    """§8f N2: experiments/base/dqn.py:12-69 over the device agent -- acting (select_action in one call), replay add,
    asynchronous learning steps with update_to_data = 2, D-syncs, T-updates with the device-side loss sums.  The learner
    must have taken exactly the steps the reference's schedule prescribes and the loop's bookkeeping must add up."""
    from idqn_b200.experiments.dqn import SyntheticAtari, linear_schedule, train
    from idqn_b200.networks.idqn import iDQN
    from idqn_b200.sample_collection.replay_buffer import ReplayBuffer
    from idqn_b200.sample_collection.samplers import UniformSamplingDistribution
    K, T, D, utd = 2, 16, 4, 2
    agent = iDQN(5, (84, 84, 4), 6, K, [32, 64, 64, 512], "cnn", 3e-4, 0.99, 1, utd, T, D, 1.5e-4)
    rb = ReplayBuffer(UniformSamplingDistribution(seed=1), batch_size=32, max_capacity=500, stack_size=4,
                      clipping=lambda r: np.clip(r, -1, 1))
    env = SyntheticAtari(episode_length=41)
    logs = []

    class Log:
        def log(self, d):
            logs.append(d)

    p = dict(n_epochs=2, n_training_steps_per_epoch=100, n_initial_samples=60, epsilon_end=0.1, epsilon_duration=150,
             horizon=1000, wandb=Log())
    returns, lengths = train(7, p, agent, env, rb)
    n_steps = sum(sum(l) for l in lengths)
    assert n_steps == len(env.actions) >= 200 and all(0 <= a < 6 for a in env.actions)
    want_learn = sum(1 for s in range(61, n_steps + 1) if s % utd == 0)
    assert int(np.asarray(agent.optimizer_state[0].count)[0]) == want_learn
    t_logs = [d for d in logs if "loss" in d]
    assert len(t_logs) == sum(1 for s in range(61, n_steps + 1) if s % T == 0)
    assert all(np.isfinite(d["loss"]) and d["loss"] > 0 for d in t_logs)
    assert [d["epoch"] for d in logs if "epoch" in d] == [0, 1]
    sched = linear_schedule(1.0, 0.1, 150)
    assert sched(0) == 1.0 and abs(sched(75) - 0.55) < 1e-12 and sched(150) == sched(10 ** 6) == 0.1


def test_forward_chain_is_bit_identical_to_the_two_kernel_path():
    """conv1 -> conv2 in one kernel (IDQN_F_CHAIN: the intermediate image in shared memory) issues the same MMAs in the same
    order as the two conv_taps launches: losses and every arena stay bit-identical over 5 steps with a D-sync."""
    from idqn_b200 import _lib as L
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B = (84, 84, 4), [32, 64, 64, 512], 6, 3, 32
    rng = np.random.default_rng(61)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    target = O.init_params(np.random.default_rng(1061), obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    batches = [make_batch(rng, B, obs, A, True) for _ in range(5)]
    runs = []
    for flags in (0, L.F_CHAIN):
        agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4, flags=flags)
        agent.params, agent.target_params = params, target
        losses = []
        for step, b in enumerate(batches, start=1):
            losses.append(agent._engine.learn_host(b, want_losses=True))
            agent.update_target_params(step)
        runs.append((np.stack(losses), [agent._engine.download_arena(w) for w in (L.ONLINE, L.TARGET, L.MU, L.NU)]))
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        np.testing.assert_array_equal(a, b)


def test_backward_schedule_does_not_change_the_numbers():
    """The Dense_0 update runs either after the conv backward chain on every SM (idqn_set_dense_update_ctas 0), next to it on
    the automatic 58 + 6.5 K CTAs, or on an arbitrary cap: three graphs with different branches and grids, the same kernels on
    the same operands -- losses and every arena bit-identical over 5 steps with a D-sync.  The per-CTA timeline of the
    instrumented build reports the capped grid."""
    import ctypes as C
    from idqn_b200 import _lib as L
    from idqn_b200.networks.idqn import iDQN
    obs, feats, A, K, B = (84, 84, 4), [32, 64, 64, 512], 6, 3, 32
    rng = np.random.default_rng(67)
    params = O.init_params(rng, obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    target = O.init_params(np.random.default_rng(1067), obs, feats, "cnn", A, n_networks=K, bias_scale=0.01)
    batches = [make_batch(rng, B, obs, A, True) for _ in range(5)]
    runs = []
    for ctas in (0, -1, 40):
        agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4)
        eng = agent._engine
        L.check(eng.lib.idqn_set_dense_update_ctas(eng.h, ctas))
        got = int(eng.lib.idqn_dense_update_ctas(eng.h))
        assert got == ctas if ctas >= 0 else 0 < got < torch.cuda.get_device_properties(0).multi_processor_count
        agent.params, agent.target_params = params, target
        losses = []
        for step, b in enumerate(batches, start=1):
            losses.append(eng.learn_host(b, want_losses=True))
            agent.update_target_params(step)
        runs.append((np.stack(losses), [eng.download_arena(w) for w in (L.ONLINE, L.TARGET, L.MU, L.NU)]))
    for other in runs[1:]:
        np.testing.assert_array_equal(runs[0][0], other[0])
        for a, b in zip(runs[0][1], other[1]):
            np.testing.assert_array_equal(a, b)
    # instrumented build: the update kernel's launch (slot 8 of the step) ran on the capped grid
    agent = iDQN(0, obs, A, K, feats, "cnn", 3e-4, 0.99, 1, 1, 8, 4, 1.5e-4, flags=L.F_TIMELINE)
    eng = agent._engine
    kt, names, n = np.zeros(128, np.uint64), np.zeros(64 * 32, np.uint8), C.c_int(0)
    for b in batches[:3]:
        eng.learn_host(b, want_losses=True)
        L.check(eng.lib.idqn_kernel_timeline(eng.h, L.ptr(kt), L.ptr(names), 64, C.byref(n)))  # reads and clears: the last step's stamps
    got = [bytes(names[32 * i:32 * i + 32]).split(b"\0")[0].decode() for i in range(n.value)]
    assert n.value == 15 and "dense_wgrad_adam_L3" in got and got[0] == "s2d_input_L0"
    t = kt[:2 * n.value].astype(np.int64).reshape(-1, 2)
    assert (t[:, 1] > t[:, 0]).all()
    upd, chain_end = got.index("dense_wgrad_adam_L3"), t[got.index("img_wgrad_L0"), 1]
    assert t[upd, 0] < chain_end, "the update must start before the conv backward chain has finished"
