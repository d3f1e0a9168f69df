"""Self-checks of the (unpinned) network oracle: SAME-padding conv vs a naive loop restatement of
lax.conv_general_dilated, autograd gradients vs fp64 finite differences, fp32 vs fp64 agreement,
optax-Adam arithmetic, shift/sync semantics (idqn.py:13-24) and the schedule trace of SURVEY App. B."""
import numpy as np
import torch

from oracle import networks as O


def small_batch(rng, B, obs, A, u8=True):
    if u8:
        s = rng.integers(0, 256, (B,) + obs).astype(np.uint8)
        s2 = rng.integers(0, 256, (B,) + obs).astype(np.uint8)
    else:
        s = rng.standard_normal((B,) + obs).astype(np.float32)
        s2 = rng.standard_normal((B,) + obs).astype(np.float32)
    return dict(state=s, next_state=s2, action=rng.integers(0, A, B).astype(np.int32),
                reward=rng.integers(-1, 2, B).astype(np.float32), is_terminal=(rng.random(B) < 0.2))


def test_same_pad_values():
    # conv0 (2,2), conv1 (1,2), conv2 (1,1); 84 -> 21 -> 11 -> 11
    assert O.same_pad(84, 8, 4) == (21, 2, 2)
    assert O.same_pad(21, 4, 2) == (11, 1, 2)
    assert O.same_pad(11, 3, 1) == (11, 1, 1)
    shapes = O.layer_shapes((84, 84, 4), [32, 64, 64, 512], "cnn", 6)
    assert shapes[3][1] == (7744, 512)
    assert sum(int(np.prod(k)) + int(np.prod(b)) for _, k, b in shapes) == 4_046_502


def test_conv_same_vs_naive():
    rng = np.random.default_rng(0)
    for (H, W, C, O_, k, s) in [(21, 21, 3, 5, 4, 2), (13, 10, 2, 3, 8, 4), (11, 11, 4, 4, 3, 1)]:
        x = rng.standard_normal((2, H, W, C))
        w = rng.standard_normal((k, k, C, O_))
        b = rng.standard_normal(O_)
        y = O._conv_same(torch.as_tensor(x), torch.as_tensor(w), torch.as_tensor(b), s).numpy()
        np.testing.assert_allclose(y, O.conv_same_naive(x, w, b, s), rtol=1e-12, atol=1e-12)


def test_grad_finite_differences_fp64():
    rng = np.random.default_rng(1)
    obs, feats, A = (20, 20, 2), [3, 4, 4, 6], 3
    p = O.init_params(rng, obs, feats, "cnn", A, bias_scale=0.1)
    pt = O.init_params(rng, obs, feats, "cnn", A, bias_scale=0.1)
    batch = small_batch(rng, 4, obs, A)
    loss, g = O.loss_and_grad(p, pt, batch, "cnn", 0.99, 1, torch.float64)
    eps = 1e-6
    for name in p["params"]:
        for leaf in ("kernel", "bias"):
            arr = p["params"][name][leaf].astype(np.float64)
            flat_idx = rng.integers(0, arr.size, 3)
            for fi in flat_idx:
                pp = {"params": {n: {l: v.astype(np.float64).copy() for l, v in d.items()} for n, d in p["params"].items()}}
                pm = {"params": {n: {l: v.astype(np.float64).copy() for l, v in d.items()} for n, d in p["params"].items()}}
                pp["params"][name][leaf].flat[fi] += eps
                pm["params"][name][leaf].flat[fi] -= eps
                fd = (O.loss_on_batch(pp, pt, batch, "cnn", 0.99, 1, torch.float64)
                      - O.loss_on_batch(pm, pt, batch, "cnn", 0.99, 1, torch.float64)) / (2 * eps)
                assert abs(fd - g["params"][name][leaf].flat[fi]) <= 1e-5 * max(1.0, abs(fd)), (name, leaf, fi)


def test_fp32_vs_fp64_and_adam():
    rng = np.random.default_rng(2)
    obs, feats, A, K = (8,), [16, 16], 4, 3
    params = O.init_params(rng, obs, feats, "fc", A, n_networks=K)
    target = O.init_params(rng, obs, feats, "fc", A, n_networks=K)
    batch = small_batch(rng, 32, obs + (1,), A, u8=False)
    st = O.init_optimizer_state(params)
    p32, s32, l32 = O.learn_on_batch(params, target, st, batch, "fc", 0.99, 1, 3e-4, 1e-8, torch.float32)
    p64, s64, l64 = O.learn_on_batch(params, target, st, batch, "fc", 0.99, 1, 3e-4, 1e-8, torch.float64)
    np.testing.assert_allclose(l32, l64, rtol=1e-5)
    assert (s32["count"] == 1).all()
    # first Adam step moves every weight with a non-negligible gradient by ~lr (|m_hat|/sqrt(v_hat) == 1)
    _, g = O.loss_and_grad(O.tree_index(params, 0), O.tree_index(target, 0), batch, "fc", 0.99, 1, torch.float64)
    gk = g["params"]["Dense_0"]["kernel"]
    d = p64["params"]["Dense_0"]["kernel"][0] - params["params"]["Dense_0"]["kernel"][0]
    mask = np.abs(gk) > 1e-4
    np.testing.assert_allclose(d[mask], -3e-4 * np.sign(gk[mask]), rtol=1e-3)


def test_shift_and_sync():
    K = 4
    p = {"params": {"Dense_0": {"kernel": np.arange(K * 6, dtype=np.float32).reshape(K, 2, 3), "bias": np.arange(K, dtype=np.float32)[:, None]}}}
    t = O.tree_map(lambda a: -a, p)
    s = O.shift_params(p)
    np.testing.assert_array_equal(s["params"]["Dense_0"]["bias"][:, 0], [1, 2, 3, 3])
    y = O.sync_target_params(p, t)
    np.testing.assert_array_equal(y["params"]["Dense_0"]["bias"][:, 0], [-0.0, 0, 1, 2])


def test_schedule_trace_appendix_b():
    sch = O.ScheduleOracle(1, 8, 4)
    got = {s: sch.events(s) for s in range(1, 17)}
    assert got[3] == ["grad"] and got[4] == ["grad", "D"] and got[8] == ["grad", "T"]
    assert got[12] == ["grad", "D"] and got[16] == ["grad", "T"]
    sch = O.ScheduleOracle(2.0, 8, 4)  # update_to_data is a float flag in the reference
    assert sch.events(3) == [] and sch.events(4) == ["grad", "D"]


# ---- the second, independent oracle (NumPy float64, hand-derived backward) and the committed fixtures -----------------
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_network as GN  # noqa: E402
from golden_network import G  # noqa: E402
from oracle import networks_np as N  # noqa: E402


@pytest.mark.parametrize("arch,obs,feats,A,K,u8", [("fc", (8,), [100, 100], 4, 3, False),
                                                   ("cnn", (84, 84, 4), [32, 64, 64, 512], 6, 1, True),
                                                   ("cnn", (37, 50, 3), [5, 7, 9, 33], 18, 2, True)])
def test_two_independent_oracles_agree_in_float64(arch, obs, feats, A, K, u8):
    """torch autograd + F.conv2d (oracle/networks.py) vs explicit patch gathers + hand-derived chain rule
    (oracle/networks_np.py): losses, every gradient, updated parameter and Adam moment agree to float64 rounding."""
    rng = np.random.default_rng(7)
    params = O.init_params(rng, obs, feats, arch, A, n_networks=K, bias_scale=0.01)
    target = O.init_params(rng, obs, feats, arch, A, n_networks=K, bias_scale=0.01)
    batch = small_batch(rng, 32, obs if arch == "cnn" else obs + (1,), A, u8=u8)
    st = O.init_optimizer_state(params)
    n_p, n_s, n_l, n_g = N.learn_on_batch(params, target, st, batch, arch, 0.99, 1, 3e-4, 1.5e-4)
    o_p, o_s, o_l, o_g = O.learn_on_batch(params, target, st, batch, arch, 0.99, 1, 3e-4, 1.5e-4, torch.float64, return_grads=True)
    np.testing.assert_allclose(n_l, o_l, rtol=1e-12)
    for mod in n_g["params"]:
        for leaf in n_g["params"][mod]:
            for a, b in ((n_g, o_g), (n_p, o_p), (n_s["mu"], o_s["mu"]), (n_s["nu"], o_s["nu"])):
                x, y = np.asarray(a["params"][mod][leaf], np.float64), np.asarray(b["params"][mod][leaf], np.float64)
                assert np.linalg.norm(x - y) <= 1e-12 * max(np.linalg.norm(y), 1e-30), (mod, leaf)
    assert (n_s["count"] == o_s["count"]).all()
    assert N.schedule_events(8, 1, 8, 4) == ["learn", "T"] and N.schedule_events(4, 1, 8, 4) == ["learn", "D"]


class NumpyChain:
    """The fixture's own generator, re-run: the committed numbers must reproduce."""

    def __init__(self, name):
        self.c = G.CONFIGS[name]

    def start(self, params, target):
        f64 = lambda t: O.tree_map(lambda a: np.asarray(a, np.float64), t)
        self.p, self.t = f64(params), f64(target)
        K = self.c["K"]
        self.st = {"count": np.zeros(K, np.int32), "mu": O.tree_map(np.zeros_like, self.p), "nu": O.tree_map(np.zeros_like, self.p)}

    def step(self, batch):
        c = self.c
        self.p, self.st, losses, _ = N.learn_on_batch(self.p, self.t, self.st, batch, c["arch"], G.GAMMA, G.NSTEP, c["lr"], c["eps"])
        return losses

    def state(self):
        return self.p, self.st["mu"], self.st["nu"], self.st["count"]

    def events(self, step):
        ev = N.schedule_events(step, 1, G.T, G.D)
        if "T" in ev:
            self.t = O.tree_map(np.copy, self.p)
            self.p = N.shift_params(self.p)
        elif "D" in ev:
            self.t = N.sync_target_params(self.p, self.t)

    def final(self):
        return self.t, self.p


class TorchChain(NumpyChain):
    """oracle/networks.py free-running in float32: the drift a TRUE fp32 implementation shows against the float64
    fixture -- the calibration of the envelope the CUDA path is held to (tests/golden_network.py)."""

    def start(self, params, target):
        params, target = N.nest_modules(params), N.nest_modules(target)  # flax's nesting (impala: Stack_i / Conv_j)
        self.p, self.t, self.st = params, target, O.init_optimizer_state(params)

    def step(self, batch):
        c = self.c
        self.p, self.st, losses = O.learn_on_batch(self.p, self.t, self.st, batch, c["arch"], G.GAMMA, G.NSTEP, c["lr"], c["eps"], torch.float32)[:3]
        return losses

    def events(self, step):
        ev = O.ScheduleOracle(1, G.T, G.D).events(step)
        if "T" in ev:
            self.t = O.tree_map(np.copy, self.p)
            self.p = O.shift_params(self.p)
        elif "D" in ev:
            self.t = O.sync_target_params(self.p, self.t)


@pytest.mark.parametrize("name,steps", [("mlp_k3", None), ("cnn_k1", 2), ("impala_k2", 2)])
def test_numpy_oracle_reproduces_the_committed_fixture(name, steps):
    for row in GN.run_chain(name, NumpyChain(name), steps=steps):
        for what in ("loss", "param", "mu", "nu", "final_target", "final_param"):
            assert row.get(what, 0.0) <= 1e-12, (row["step"], what, row[what])


@pytest.mark.parametrize("name", ["mlp_k3", "cnn_k1", "impala_k2"])
def test_torch_fp32_oracle_free_running_stays_inside_the_stated_envelope(name):
    GN.check(GN.run_chain(name, TorchChain(name)))


# ---- impala (§8f N4): the two independent restatements against each other, and the pool / residual rules by hand ----------

def test_impala_layer_table_and_naive_maxpool():
    """Layer table in flax creation order; the SAME max-pool (architectures/dqn.py:20) against a window-by-window scan with
    -inf padding, on the three padding cases the 84 -> 42 -> 21 -> 11 chain meets ((0,1), (0,1), (1,1)) and an odd width."""
    names = [n for n, _, _ in O.layer_shapes((84, 84, 4), [32, 64, 64, 512], "impala", 6)]
    assert names == [f"Stack_{i}/Conv_{j}" for i in range(3) for j in range(5)] + ["Dense_0", "Dense_1"]
    shapes = {n: k for n, k, _ in O.layer_shapes((84, 84, 4), [32, 64, 64, 512], "impala", 6)}
    assert shapes["Stack_0/Conv_0"] == (3, 3, 4, 32) and shapes["Stack_1/Conv_0"] == (3, 3, 32, 64)
    assert shapes["Stack_2/Conv_4"] == (3, 3, 64, 64) and shapes["Dense_0"] == (11 * 11 * 64, 512)
    rng = np.random.default_rng(2)
    for H, W in ((84, 84), (42, 42), (21, 21), (7, 10)):
        x = rng.standard_normal((2, H, W, 3))
        got = O._max_pool_same(torch.as_tensor(x)).numpy()
        OH, lo_h, _ = O.same_pad(H, 3, 2)
        OW, lo_w, _ = O.same_pad(W, 3, 2)
        want = np.full((2, OH, OW, 3), -np.inf)
        for oy in range(OH):
            for ox in range(OW):
                ys = [y for y in range(oy * 2 - lo_h, oy * 2 - lo_h + 3) if 0 <= y < H]
                xs = [v for v in range(ox * 2 - lo_w, ox * 2 - lo_w + 3) if 0 <= v < W]
                want[:, oy, ox, :] = x[:, ys][:, :, xs].max(axis=(1, 2))
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(N.maxpool_forward(x)[0], want)


@pytest.mark.parametrize("obs,feats,A", [((21, 19, 4), [3, 5, 2, 7], 4), ((30, 30, 2), [1, 9, 4, 3], 9)])
def test_impala_two_independent_oracles_agree_in_float64(obs, feats, A):
    """torch autograd + F.conv2d / F.max_pool2d vs NumPy patch gathers with a hand-derived backward (pool routing to the first
    maximum, residual gradient fan-in, pre-activation gates): Q-values, loss and every gradient to 1e-12."""
    rng = np.random.default_rng(17)
    p = O.init_params(rng, obs, feats, "impala", A, bias_scale=0.1)
    t = O.init_params(rng, obs, feats, "impala", A, bias_scale=0.1)
    B = 5
    batch = dict(state=rng.uniform(0, 255, (B,) + obs), next_state=rng.uniform(0, 255, (B,) + obs),
                 action=rng.integers(0, A, B), reward=rng.uniform(-1, 1, B), is_terminal=rng.random(B) < 0.3)
    np.testing.assert_allclose(O.apply(p, batch["state"], "impala", dtype=torch.float64), N.impala_forward(p, batch["state"]),
                               rtol=1e-12, atol=1e-13)
    l1, g1 = O.loss_and_grad(p, t, batch, "impala", 0.94, 1, dtype=torch.float64)
    l2, g2 = N.impala_loss_and_grad(p, t, batch, 0.94, 1)
    assert abs(l1 - l2) <= 1e-12 * max(1.0, abs(l1))

    def walk(a, b, path=""):
        if isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                walk(a[k], b[k], f"{path}/{k}")
        else:
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * max(1.0, float(np.abs(a).max())), err_msg=path)
    walk(g1, g2)
    # one step of the K-vmapped update on the nested tree (tree helpers must recurse through the Stack level)
    pk, tk = O.tree_stack([p, t]), O.tree_stack([t, p])
    opt = O.init_optimizer_state(pk)
    p2, o2, losses = O.learn_on_batch(pk, tk, opt, batch, "impala", 0.94, 1, 1e-3, 1e-5)
    assert losses.shape == (2,) and o2["count"].tolist() == [1, 1]
    assert p2["params"]["Stack_2"]["Conv_3"]["kernel"].shape == (2, 3, 3, feats[2], feats[2])
    assert not np.array_equal(p2["params"]["Stack_0"]["Conv_0"]["kernel"], pk["params"]["Stack_0"]["Conv_0"]["kernel"])


def test_impala_forced_pool_choices():
    """The max-pool decisions the GPU parity test hands to the oracle (`pools=`): with the oracle's own choices nothing changes;
    `pool_argmax_same` is the first maximum in row-major window order (the NumPy oracle's rule, = the CUDA kernels'); and moving
    ONE choice to another window element moves gradients of the Conv_0 that feeds that pool."""
    rng = np.random.default_rng(23)
    obs, feats, A, B = (13, 12, 4), [4, 6, 5, 8], 3, 3
    p = O.init_params(rng, obs, feats, "impala", A, bias_scale=0.1)
    t = O.init_params(rng, obs, feats, "impala", A, bias_scale=0.1)
    batch = small_batch(rng, B, obs, A)
    l0, g0, z, pin = O.loss_and_grad(p, t, batch, "impala", 0.9, 1, torch.float64, preacts=True, pool_inputs=True)
    assert [x.shape[1:3] for x in pin] == [(13, 12), (7, 6), (4, 3)] and len(z) == 14
    pools = [O.pool_argmax_same(x) for x in pin]
    for x, a in zip(pin, pools):
        np.testing.assert_array_equal(a, N.maxpool_forward(x)[1][0])
        w = O.pool_windows_same(x)
        np.testing.assert_array_equal(np.take_along_axis(w, a[..., None], 4)[..., 0], N.maxpool_forward(x)[0])
    # ties: the first of equal elements wins
    tie = np.zeros((1, 4, 4, 1))
    assert O.pool_argmax_same(tie)[0, 0, 0, 0] == 0 and O.pool_argmax_same(tie)[0, 1, 1, 0] == 0
    l1, g1 = O.loss_and_grad(p, t, batch, "impala", 0.9, 1, torch.float64, pools=pools)
    assert abs(l1 - l0) <= 1e-14
    for a, b in zip(O.tree_leaves(g0), O.tree_leaves(g1)):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-14)
    moved = [a.copy() for a in pools]
    w = O.pool_windows_same(pin[1])[1, 2, 2, 3]
    moved[1][1, 2, 2, 3] = int(np.argsort(w)[-2])  # the runner-up of an interior window (finite: -inf only pads the border)
    l2, g2 = O.loss_and_grad(p, t, batch, "impala", 0.9, 1, torch.float64, pools=moved)
    assert abs(l2 - l0) > 1e-9
    assert not np.array_equal(g2["params"]["Stack_1"]["Conv_0"]["kernel"], g0["params"]["Stack_1"]["Conv_0"]["kernel"])


def test_module_name_flattening_round_trip():
    """The fixtures and the NumPy chain functions use two-level trees with dotted module names; flax's nesting comes back
    unchanged, cnn / fc trees are untouched, K-stacked leaves keep their axis."""
    rng = np.random.default_rng(1)
    nested = O.init_params(rng, (20, 18, 4), [4, 3, 5, 6], "impala", 3, n_networks=2)
    flat = N.flatten_modules(nested)
    assert sorted(flat["params"])[:3] == ["Dense_0", "Dense_1", "Stack_0.Conv_0"] and len(flat["params"]) == 17
    assert flat["params"]["Stack_2.Conv_4"]["kernel"].shape == (2, 3, 3, 5, 5)
    back = N.nest_modules(flat)
    assert sorted(back["params"]) == sorted(nested["params"])
    for a, b in zip(O.tree_leaves(back), O.tree_leaves(nested)):
        assert a is b
    cnn = O.init_params(rng, (84, 84, 4), [32, 64, 64, 512], "cnn", 6)
    assert N.flatten_modules(cnn)["params"].keys() == cnn["params"].keys() == N.nest_modules(cnn)["params"].keys()
    assert N.flatten_modules(flat)["params"].keys() == flat["params"].keys()  # idempotent
