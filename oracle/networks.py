"""CPU restatement (torch, fp32 or fp64) of the reference's network path.  TEST ORACLE ONLY.

**Parity unpinned** – jax/flax/optax cannot be imported in this environment and the
reference ships no golden vectors for ``learn_on_batch``; this file restates the
documented semantics of the pinned third-party versions (flax 0.10.2, optax 0.2.4,
``setup.cfg:18,27`` of the reference) at the reference's own call sites.  It is
cross-checked against a second restatement that shares nothing with it
(``oracle/networks_np.py``: NumPy float64, hand-derived backward) and against the
committed float64 fixtures ``tests/golden/network_*.npz`` — which pins the two
restatements and the CUDA path to each other, not to JAX:

* ``slimdqn/networks/architectures/dqn.py:37-70``  -> :func:`apply` ('cnn', 'fc' and -- ``:7-29,54-60`` -- 'impala')
* ``slimdqn/networks/idqn.py:13-24``               -> :func:`shift_params`, :func:`sync_target_params`
* ``slimdqn/networks/idqn.py:96-124``              -> :func:`learn_on_batch`, :func:`loss_on_batch`,
                                                      :func:`loss`, :func:`compute_target`
* ``slimdqn/networks/idqn.py:65-94``               -> :class:`ScheduleOracle`
* ``slimdqn/networks/dqn.py:41-92``                -> the same functions with ``K`` axis absent

Parameters are nested dicts of numpy arrays in the Flax layout
``{"params": {"Conv_0": {"kernel": [kh,kw,in,out], "bias": [out]}, ...}}`` (impala: one more level,
``{"Stack_0": {"Conv_0": ..., ..., "Conv_4": ...}, ...}``); the i-DQN functions take leaves with a leading K axis.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------
# architecture description (flax auto-naming; architectures/dqn.py:39-70)
# ---------------------------------------------------------------------------------------

CNN_SPECS = ((8, 4), (4, 2), (3, 1))  # (kernel, stride) of Conv_0..2, architectures/dqn.py:43,48,51


def same_pad(in_size: int, k: int, s: int) -> Tuple[int, int, int]:
    """XLA 'SAME' padding: out=ceil(in/s), total=max((out-1)s+k-in,0), lo=total//2, hi=total-lo."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    lo = total // 2
    return out, lo, total - lo


def layer_shapes(observation_dim, features: Sequence[int], architecture_type: str, n_actions: int):
    """[(module name, kernel shape, bias shape)] in flax creation order."""
    features = [int(f) for f in features]
    layers = []
    if architecture_type == "cnn":
        h, w, c = observation_dim
        for i, (k, s) in enumerate(CNN_SPECS):
            layers.append((f"Conv_{i}", (k, k, c, features[i]), (features[i],)))
            h, w, c = same_pad(h, k, s)[0], same_pad(w, k, s)[0], features[i]
        in_dim = h * w * c
        start = 3
    elif architecture_type == "impala":
        # architectures/dqn.py:54-60: Stack(features[0..2]); a Stack (:7-29) is Conv_0, max-pool, then two residual blocks of
        # two convs each (flax auto-names them Conv_1..Conv_4 in creation order); every conv 3x3, stride 1, SAME
        h, w, c = observation_dim
        for i in range(3):
            f = features[i]
            layers.append((f"Stack_{i}/Conv_0", (3, 3, c, f), (f,)))
            for j in range(1, 5):
                layers.append((f"Stack_{i}/Conv_{j}", (3, 3, f, f), (f,)))
            h, w, c = same_pad(h, 3, 2)[0], same_pad(w, 3, 2)[0], f  # the 3x3 / 2 SAME max-pool (:20)
        in_dim = h * w * c
        start = 3
    elif architecture_type == "fc":
        in_dim = int(np.prod(observation_dim))
        start = 0
    else:
        raise ValueError(architecture_type)
    d = 0
    for f in features[start:]:
        layers.append((f"Dense_{d}", (in_dim, f), (f,)))
        in_dim = f
        d += 1
    layers.append((f"Dense_{d}", (in_dim, n_actions), (n_actions,)))
    return layers


def init_params(rng: np.random.Generator, observation_dim, features, architecture_type, n_actions,
                n_networks: int | None = None, bias_scale: float = 0.0):
    """Synthetic parameters with the reference's initialiser *distributions*
    (xavier-uniform for cnn incl. its dense trunk, lecun-normal for fc; biases zero unless
    ``bias_scale``).  The reference draws them from jax's threefry PRNG (idqn.py:48-50), which is
    not reproducible here, so parity tests always inject parameters."""
    def one():
        tree = {}
        for name, kshape, bshape in layer_shapes(observation_dim, features, architecture_type, n_actions):
            rf = int(np.prod(kshape[:-2]))
            fan_in, fan_out = rf * kshape[-2], rf * kshape[-1]
            # impala: xavier-uniform for the first conv of a Stack (:14-19) and the dense trunk (:55,68,70), flax's default
            # lecun-normal for the convs of the residual blocks (:25-26 pass no kernel_init)
            xavier = architecture_type == "cnn" or (architecture_type == "impala" and
                                                    (name.endswith("/Conv_0") or name.startswith("Dense_")))
            if xavier:
                a = math.sqrt(6.0 / (fan_in + fan_out))
                kern = rng.uniform(-a, a, kshape)
            else:
                std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
                kern = np.clip(rng.standard_normal(kshape), -2, 2) * std
            bias = rng.standard_normal(bshape) * bias_scale
            tree[name] = {"kernel": kern.astype(np.float32), "bias": bias.astype(np.float32)}
        return {"params": nest_modules(tree)}
    if n_networks is None:
        return one()
    heads = [one() for _ in range(n_networks)]
    return tree_stack(heads)


# ---------------------------------------------------------------------------------------
# tiny pytree helpers (nested dicts only)
# ---------------------------------------------------------------------------------------

def nest_modules(flat: dict) -> dict:
    """{"Stack_0/Conv_0": m, "Dense_0": n} -> {"Stack_0": {"Conv_0": m}, "Dense_0": n} (flax sub-module nesting)."""
    out: dict = {}
    for name, mod in flat.items():
        node = out
        parts = name.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = mod
    return out


def tree_map(fn, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: tree_map(fn, *[t[k] for t in trees]) for k in t0}
    return fn(*trees)


def tree_leaves(tree):
    if isinstance(tree, dict):
        out = []
        for k in tree:
            out.extend(tree_leaves(tree[k]))
        return out
    return [tree]


def tree_stack(trees):
    return tree_map(lambda *xs: np.stack(xs), *trees)


def tree_index(tree, k):
    return tree_map(lambda x: x[k], tree)


# ---------------------------------------------------------------------------------------
# DQNNet.__call__  (architectures/dqn.py:37-70), batched over a leading sample axis
# ---------------------------------------------------------------------------------------

def _conv_same(x_nhwc: torch.Tensor, kernel_hwio: torch.Tensor, bias: torch.Tensor, stride: int) -> torch.Tensor:
    """flax.linen.Conv with padding='SAME' (the default; none is passed at architectures/dqn.py:43,48,51)."""
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    _, hlo, hhi = same_pad(x_nhwc.shape[1], kh, stride)
    _, wlo, whi = same_pad(x_nhwc.shape[2], kw, stride)
    x = x_nhwc.permute(0, 3, 1, 2)
    x = F.pad(x, (wlo, whi, hlo, hhi))
    y = F.conv2d(x, kernel_hwio.permute(3, 2, 0, 1).contiguous(), bias, stride=stride)  # (contiguous: torch's 1-channel CPU path insists)
    return y.permute(0, 2, 3, 1)


def _max_pool_same(x_nhwc: torch.Tensor, k: int = 3, stride: int = 2, arg=None) -> torch.Tensor:
    """flax.linen.max_pool(x, (3, 3), strides=(2, 2), padding="SAME") (architectures/dqn.py:20): -inf padding."""
    _, hlo, hhi = same_pad(x_nhwc.shape[1], k, stride)
    _, wlo, whi = same_pad(x_nhwc.shape[2], k, stride)
    x = F.pad(x_nhwc.permute(0, 3, 1, 2), (wlo, whi, hlo, hhi), value=float("-inf"))
    if arg is None:
        return F.max_pool2d(x, k, stride).permute(0, 2, 3, 1)
    # the window element to take is given (ky * k + kx per output, [N, OH, OW, C]): parity tests hand over the decisions of
    # the implementation under test where two window elements are numerically equal
    patches = x.unfold(2, k, stride).unfold(3, k, stride)  # [N, C, OH, OW, k, k]
    patches = patches.reshape(*patches.shape[:4], k * k)
    a = torch.as_tensor(np.asarray(arg)).to(torch.int64).permute(0, 3, 1, 2).unsqueeze(-1)
    return torch.gather(patches, 4, a)[..., 0].permute(0, 2, 3, 1)


def pool_windows_same(x_nhwc: np.ndarray, k: int = 3, stride: int = 2) -> np.ndarray:
    """The 3x3 / 2 SAME max-pool windows of x (-inf padding), float64 [N, OH, OW, C, k * k], window elements row-major."""
    x = torch.as_tensor(np.asarray(x_nhwc, dtype=np.float64))
    _, hlo, hhi = same_pad(x.shape[1], k, stride)
    _, wlo, whi = same_pad(x.shape[2], k, stride)
    xp = F.pad(x.permute(0, 3, 1, 2), (wlo, whi, hlo, hhi), value=float("-inf"))
    patches = xp.unfold(2, k, stride).unfold(3, k, stride)
    return patches.reshape(*patches.shape[:4], k * k).permute(0, 2, 3, 1, 4).numpy()


def pool_argmax_same(x_nhwc: np.ndarray, k: int = 3, stride: int = 2) -> np.ndarray:
    """First maximum (row-major over the window) of every max-pool window: [N, OH, OW, C] of ky * k + kx."""
    return np.argmax(pool_windows_same(x_nhwc, k, stride), axis=4)  # np.argmax returns the first maximum


def _to_torch(tree, dtype):
    return tree_map(lambda a: torch.as_tensor(np.asarray(a)).to(dtype), tree)


def apply_t(p: Dict[str, Dict[str, torch.Tensor]], x: torch.Tensor, architecture_type: str,
            collect: list | None = None, gates: list | None = None, pools: list | None = None,
            collect_pool: list | None = None) -> torch.Tensor:
    """Torch forward of one head on a batch ``x`` ([N,H,W,C] for cnn, [N,obs] for fc).
    ``p`` is the inner ``params`` dict.  ``collect`` receives the pre-activations (for gate margins).
    ``gates`` (one 0/1 tensor per hidden layer) replaces relu(z) by z*gate: used by the parity tests to give the
    oracle the gate decisions of the implementation under test where a pre-activation is numerically zero.
    ``pools`` (impala: one index array per Stack) does the same for the max-pool's choice inside a window;
    ``collect_pool`` receives the pool inputs."""
    d = 0
    gi = 0

    def act(z):
        nonlocal gi
        if collect is not None:
            collect.append(z)
        if gates is None:
            return torch.relu(z)
        g = gates[gi].reshape(z.shape).to(z.dtype)
        gi += 1
        return z * g

    if architecture_type == "cnn":
        x = x / 255.0  # architectures/dqn.py:44
        for i, (_, s) in enumerate(CNN_SPECS):
            x = act(_conv_same(x, p[f"Conv_{i}"]["kernel"], p[f"Conv_{i}"]["bias"], s))
        x = x.reshape(x.shape[0], -1)  # (h,w,c) order, architectures/dqn.py:53
    elif architecture_type == "impala":
        x = x / 255.0  # architectures/dqn.py:57
        for i in range(3):
            s = p[f"Stack_{i}"]
            x = _conv_same(x, s["Conv_0"]["kernel"], s["Conv_0"]["bias"], 1)  # :14-19 (no activation)
            if collect_pool is not None:
                collect_pool.append(x)
            x = _max_pool_same(x, arg=None if pools is None else pools[i])  # :20
            for b in range(2):  # :22-27
                block_input = x
                x = act(x)
                x = act(_conv_same(x, s[f"Conv_{1 + 2 * b}"]["kernel"], s[f"Conv_{1 + 2 * b}"]["bias"], 1))
                x = _conv_same(x, s[f"Conv_{2 + 2 * b}"]["kernel"], s[f"Conv_{2 + 2 * b}"]["bias"], 1)
                x = x + block_input
        x = act(x)  # :59
        x = x.reshape(x.shape[0], -1)  # :60
    elif architecture_type == "fc":
        x = x.reshape(x.shape[0], -1)  # jnp.squeeze of the trailing stack axis, :65
    else:
        raise ValueError(architecture_type)
    n_dense = sum(1 for k in p if k.startswith("Dense_"))
    for d in range(n_dense - 1):
        x = act(x @ p[f"Dense_{d}"]["kernel"] + p[f"Dense_{d}"]["bias"])
    last = p[f"Dense_{n_dense - 1}"]
    return x @ last["kernel"] + last["bias"]


def apply(params, x, architecture_type: str, dtype=torch.float32) -> np.ndarray:
    """``network.apply(params, x)`` for a batch of inputs; returns Q-values [N, A]."""
    p = _to_torch(params["params"], dtype)
    xt = torch.as_tensor(np.asarray(x)).to(dtype)
    with torch.no_grad():
        return apply_t(p, xt, architecture_type).numpy()


# ---------------------------------------------------------------------------------------
# loss / target (idqn.py:111-124) and the Adam step (optax 0.2.4 adam, idqn.py:52,106-107)
# ---------------------------------------------------------------------------------------

def _batch_t(batch, dtype):
    return dict(
        state=torch.as_tensor(np.asarray(batch["state"])).to(dtype),
        next_state=torch.as_tensor(np.asarray(batch["next_state"])).to(dtype),
        action=torch.as_tensor(np.asarray(batch["action"]).astype(np.int64)),
        reward=torch.as_tensor(np.asarray(batch["reward"])).to(dtype),
        is_terminal=torch.as_tensor(np.asarray(batch["is_terminal"]).astype(np.int64)),
    )


def compute_target_t(pt, b, arch, gamma, n):
    # idqn.py:120-124:  r + (1 - done) * gamma**n * max_a' Q(target, s')
    q_next = apply_t(pt, b["next_state"], arch)
    coef = (1 - b["is_terminal"]).to(q_next.dtype) * torch.tensor(gamma ** n, dtype=q_next.dtype)
    return b["reward"] + coef * q_next.max(dim=1).values


def loss_on_batch_t(p, pt, b, arch, gamma, n, collect=None, gates=None, pools=None, collect_pool=None):
    with torch.no_grad():
        y = compute_target_t(pt, b, arch, gamma, n)  # value_and_grad is w.r.t. arg 0 only (idqn.py:105)
    q = apply_t(p, b["state"], arch, collect, gates, pools, collect_pool)
    q_sa = q.gather(1, b["action"][:, None])[:, 0]  # idqn.py:117
    return torch.square(q_sa - y).mean()  # idqn.py:118,112


def compute_target(params, sample_or_batch, arch, gamma, n, dtype=torch.float32):
    b = _batch_t(sample_or_batch, dtype)
    with torch.no_grad():
        return compute_target_t(_to_torch(params["params"], dtype), b, arch, gamma, n).numpy()


def loss_on_batch(params, params_target, batch, arch, gamma, n, dtype=torch.float32):
    b = _batch_t(batch, dtype)
    with torch.no_grad():
        return float(loss_on_batch_t(_to_torch(params["params"], dtype), _to_torch(params_target["params"], dtype),
                                     b, arch, gamma, n))


def loss_and_grad(params, params_target, batch, arch, gamma, n, dtype=torch.float32, margins=False, gates=None,
                  preacts=False, pools=None, pool_inputs=False):
    """(loss, grads pytree[, min |pre-activation| per relu layer]) for ONE head (no K axis)."""
    p = _to_torch(params["params"], dtype)
    for leaf in tree_leaves(p):
        leaf.requires_grad_(True)
    pt = _to_torch(params_target["params"], dtype)
    b = _batch_t(batch, dtype)
    collect = [] if (margins or preacts) else None
    gates_t = None if gates is None else [torch.as_tensor(np.asarray(g, dtype=np.float32)) for g in gates]
    collect_pool = [] if pool_inputs else None
    loss = loss_on_batch_t(p, pt, b, arch, gamma, n, collect, gates_t, pools, collect_pool)
    loss.backward()
    grads = {"params": tree_map(lambda t: t.grad.detach().numpy().copy(), p)}
    if pool_inputs:  # (loss, grads, relu pre-activations, max-pool inputs)
        return float(loss.detach()), grads, [z.detach().numpy() for z in collect], [z.detach().numpy() for z in collect_pool]
    if preacts:
        return float(loss.detach()), grads, [z.detach().numpy() for z in collect]
    if margins:
        return float(loss.detach()), grads, [float(z.detach().abs().min()) for z in collect]
    return float(loss.detach()), grads


def init_optimizer_state(params):
    """vmapped ``optax.adam(...).init`` (idqn.py:53): (count int32[K] or scalar, mu, nu)."""
    zeros = lambda a: np.zeros_like(np.asarray(a))
    first = tree_leaves(params)[0]
    return {"count": np.zeros(first.shape[0], np.int32), "mu": tree_map(zeros, params), "nu": tree_map(zeros, params)}


def adam_step(params, grads, mu, nu, count: int, lr, eps, dtype=torch.float32, b1=0.9, b2=0.999):
    """optax.scale_by_adam + scale(-lr) + apply_updates, all in ``dtype`` (f32 in the reference)."""
    T = lambda a: torch.as_tensor(np.asarray(a)).to(dtype)
    c = count + 1
    tb1, tb2 = torch.tensor(b1, dtype=dtype), torch.tensor(b2, dtype=dtype)
    bc1 = 1 - tb1 ** c
    bc2 = 1 - tb2 ** c

    def one(p, g, m, v):
        p, g, m, v = T(p), T(g), T(m), T(v)
        m = (1 - tb1) * g + tb1 * m
        v = (1 - tb2) * (g * g) + tb2 * v
        u = (m / bc1) / (torch.sqrt(v / bc2) + torch.tensor(eps, dtype=dtype))
        p = p + torch.tensor(-lr, dtype=dtype) * u
        return p.numpy(), m.numpy(), v.numpy()

    out = tree_map(one, params, grads, mu, nu)
    pick = lambda i: tree_map_tuple(out, i)
    return pick(0), pick(1), pick(2), c


def tree_map_tuple(tree, i):
    if isinstance(tree, dict):
        return {k: tree_map_tuple(v, i) for k, v in tree.items()}
    return tree[i]


def learn_on_batch(params, params_target, opt_state, batch, arch, gamma, n, lr, eps, dtype=torch.float32,
                   return_grads=False, gates=None, pools=None):
    """idqn.py:96-109 — vmap over K heads of value_and_grad + adam + apply_updates on a shared batch.

    ``opt_state = {"count": int32[K], "mu": tree[K,...], "nu": tree[K,...]}``.
    Returns (params', opt_state', losses[K]) (+ grads[K] if requested)."""
    K = tree_leaves(params)[0].shape[0]
    new_p, new_m, new_v, losses, grads_all = [], [], [], [], []
    counts = np.asarray(opt_state["count"]).copy()
    for k in range(K):
        pk, tk = tree_index(params, k), tree_index(params_target, k)
        loss, g = loss_and_grad(pk, tk, batch, arch, gamma, n, dtype, gates=None if gates is None else gates[k],
                                pools=None if pools is None else pools[k])
        p2, m2, v2, c2 = adam_step(pk, g, tree_index(opt_state["mu"], k), tree_index(opt_state["nu"], k),
                                   int(counts[k]), lr, eps, dtype)
        counts[k] = c2
        new_p.append(p2), new_m.append(m2), new_v.append(v2), losses.append(loss), grads_all.append(g)
    out = (tree_stack(new_p), {"count": counts, "mu": tree_stack(new_m), "nu": tree_stack(new_v)},
           np.asarray(losses))
    if return_grads:
        return out + (tree_stack(grads_all),)
    return out


def learn_on_batch_dqn(params, params_target, opt_state, batch, arch, gamma, n, lr, eps, dtype=torch.float32):
    """dqn.py:60-72 — the K-less variant."""
    add = lambda t: tree_map(lambda a: np.asarray(a)[None], t)
    p, o, l = learn_on_batch(add(params), add(params_target),
                             {"count": np.asarray([opt_state["count"]]), "mu": add(opt_state["mu"]),
                              "nu": add(opt_state["nu"])}, batch, arch, gamma, n, lr, eps, dtype)
    sq = lambda t: tree_index(t, 0)
    return sq(p), {"count": o["count"][0], "mu": sq(o["mu"]), "nu": sq(o["nu"])}, l[0]


# ---------------------------------------------------------------------------------------
# window shift / target sync (idqn.py:13-24) and the schedule (idqn.py:65-94)
# ---------------------------------------------------------------------------------------

def shift_params(params):
    """idqn.py:13-17: params[k] <- params[k+1] for k < K-1; params[K-1] unchanged."""
    def f(a):
        a = np.array(a, copy=True)
        a[:-1] = a[1:].copy()
        return a
    return tree_map(f, params)


def sync_target_params(params, target_params):
    """idqn.py:20-24: target[k] <- params[k-1] for k >= 1; target[0] unchanged."""
    def f(p, t):
        t = np.array(t, copy=True)
        t[1:] = np.asarray(p)[:-1]
        return t
    return tree_map(f, params, target_params)


class ScheduleOracle:
    """Symbolic replay of update_online_params / update_target_params (idqn.py:65-94) as driven by
    experiments/base/dqn.py:45-47.  Emits the event list for a step: subset of
    ["grad", "T", "D"] in execution order."""

    def __init__(self, update_to_data, target_update_frequency, target_sync_frequency):
        self.utd, self.T, self.D = update_to_data, target_update_frequency, target_sync_frequency

    def events(self, step: int) -> List[str]:
        ev = []
        if step % self.utd == 0:  # idqn.py:66 (float modulo when utd is a float flag)
            ev.append("grad")
        if step % self.T == 0:  # idqn.py:75 — returns before the D check (:89)
            ev.append("T")
        elif step % self.D == 0:  # idqn.py:91
            ev.append("D")
        return ev


# ---------------------------------------------------------------------------------------
# naive loop convolution used to cross-check the SAME-padding restatement
# ---------------------------------------------------------------------------------------

def conv_same_naive(x_nhwc: np.ndarray, kernel_hwio: np.ndarray, bias: np.ndarray, stride: int) -> np.ndarray:
    """Direct evaluation of lax.conv_general_dilated(NHWC, HWIO, 'SAME') + bias in float64 loops."""
    N, H, W, C = x_nhwc.shape
    kh, kw, _, O = kernel_hwio.shape
    oh, hlo, _ = same_pad(H, kh, stride)
    ow, wlo, _ = same_pad(W, kw, stride)
    out = np.zeros((N, oh, ow, O), np.float64)
    for oy in range(oh):
        for ox in range(ow):
            acc = np.zeros((N, O), np.float64)
            for ky in range(kh):
                iy = oy * stride + ky - hlo
                if iy < 0 or iy >= H:
                    continue
                for kx in range(kw):
                    ix = ox * stride + kx - wlo
                    if ix < 0 or ix >= W:
                        continue
                    acc += x_nhwc[:, iy, ix, :].astype(np.float64) @ kernel_hwio[ky, kx].astype(np.float64)
            out[:, oy, ox, :] = acc + bias
    return out


# ---------------------------------------------------------------------------------------
# tensor-resident CPU learner: same arithmetic as learn_on_batch above, without the per-step
# numpy<->torch conversions — this is the variant bench.py times as the CPU baseline.
# ---------------------------------------------------------------------------------------

class CpuLearner:
    def __init__(self, params, params_target, arch, gamma, n, lr, eps, dtype=torch.float32):
        self.arch, self.gamma, self.n, self.lr, self.eps, self.dtype = arch, gamma, n, lr, eps, dtype
        self.K = tree_leaves(params)[0].shape[0]
        clone = lambda tree: tree_map(lambda x: x.clone(), tree)  # never alias the caller's numpy arrays
        self.p = [clone(_to_torch(tree_index(params, k)["params"], dtype)) for k in range(self.K)]
        self.t = [clone(_to_torch(tree_index(params_target, k)["params"], dtype)) for k in range(self.K)]
        for p in self.p:
            for leaf in tree_leaves(p):
                leaf.requires_grad_(True)
        self.mu = [[torch.zeros_like(l) for l in tree_leaves(p)] for p in self.p]
        self.nu = [[torch.zeros_like(l) for l in tree_leaves(p)] for p in self.p]
        self.count = 0

    def step(self, batch):
        b = _batch_t(batch, self.dtype)
        self.count += 1
        bc1 = 1 - 0.9 ** self.count
        bc2 = 1 - 0.999 ** self.count
        losses = []
        for k in range(self.K):
            leaves = tree_leaves(self.p[k])
            loss = loss_on_batch_t(self.p[k], self.t[k], b, self.arch, self.gamma, self.n)
            grads = torch.autograd.grad(loss, leaves)
            with torch.no_grad():
                torch._foreach_mul_(self.mu[k], 0.9)
                torch._foreach_add_(self.mu[k], grads, alpha=0.1)
                torch._foreach_mul_(self.nu[k], 0.999)
                torch._foreach_addcmul_(self.nu[k], grads, grads, value=0.001)
                denom = torch._foreach_div(self.nu[k], bc2)
                torch._foreach_sqrt_(denom)
                torch._foreach_add_(denom, self.eps)
                torch._foreach_addcdiv_(leaves, self.mu[k], denom, value=-self.lr / bc1)
            losses.append(float(loss.detach()))
        return np.asarray(losses)

    def params_host(self):
        return tree_stack([{"params": tree_map(lambda t: t.detach().numpy().copy(), p)} for p in self.p])
