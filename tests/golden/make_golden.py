"""Generate golden vectors by EXECUTING THE REFERENCE's own code in the builder container.

Run once here (``python tests/golden/make_golden.py``); the outputs (``*.npz``) are committed
because ``/root/reference`` does not exist on the GPU box.

What runs unmodified from /root/reference:
  * ``slimdqn/sample_collection/sum_tree.py``      (imports as is)
  * ``slimdqn/sample_collection/samplers.py``      (its ``import jax`` is unused -> empty stub module)
  * ``slimdqn/sample_collection/replay_buffer.py`` (``compress=False``; the third-party CONTAINER
    plumbing it imports — ``flax.struct.PyTreeNode``, ``jax.tree_util.tree_map``, ``snappy`` — is
    stubbed below with dataclass equivalents; all replay logic is the reference's)

The network path (jax/flax/optax) cannot run here -> no golden vectors for it ("parity unpinned").
"""
import dataclasses
import hashlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    jax = types.ModuleType("jax")
    tree_util = types.ModuleType("jax.tree_util")

    def tree_map(fn, *trees):
        t0 = trees[0]
        if dataclasses.is_dataclass(t0):
            return type(t0)(**{f.name: fn(*[getattr(t, f.name) for t in trees]) for f in dataclasses.fields(t0)})
        return fn(*trees)

    tree_util.tree_map = tree_map
    jax.tree_util = tree_util
    sys.modules["jax"] = jax
    sys.modules["jax.tree_util"] = tree_util

    flax = types.ModuleType("flax")
    struct = types.ModuleType("flax.struct")

    class PyTreeNode:
        def __init_subclass__(cls, **kw):
            super().__init_subclass__(**kw)
            dataclasses.dataclass(frozen=True)(cls)

        def replace(self, **kw):
            return dataclasses.replace(self, **kw)

    struct.PyTreeNode = PyTreeNode
    flax.struct = struct
    sys.modules["flax"] = flax
    sys.modules["flax.struct"] = struct
    sys.modules["snappy"] = types.ModuleType("snappy")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_sum_tree(sum_tree):
    out = {}
    # --- small trees: full node arrays after every op --------------------------------------
    rng = np.random.default_rng(1234)
    for cap in (1, 2, 5, 8, 100, 1000):
        tree = sum_tree.SumTree(cap)
        nodes_log, q_log, t_log, idx_log, val_log = [], [], [], [], []
        for it in range(24):
            n = int(rng.integers(1, min(2 * cap, 40) + 1))
            idx = rng.integers(0, cap, n).astype(np.int32)  # duplicates on purpose
            val = rng.uniform(0.0, 3.0, n)
            if it % 5 == 4:
                val[rng.integers(0, n)] = 0.0
            tree.set(idx, val)
            idx_log.append(np.pad(idx, (0, 80 - n), constant_values=-1))
            val_log.append(np.pad(val, (0, 80 - n)))
            nodes_log.append(tree._nodes.copy())
            t = rng.uniform(0.0, tree.root, 16)
            t_log.append(t)
            q_log.append(tree.query(t) if tree.root > 0 else np.full(16, -1, np.int32))
        out[f"small{cap}_idx"] = np.stack(idx_log)
        out[f"small{cap}_val"] = np.stack(val_log)
        out[f"small{cap}_nodes"] = np.stack(nodes_log)
        out[f"small{cap}_targets"] = np.stack(t_log)
        out[f"small{cap}_query"] = np.stack(q_log)
        out[f"small{cap}_maxprio"] = np.float64(tree.max_recorded_priority)
    # --- SURVEY §8(d) sampler workload, 1M capacity (depth 21): inputs regenerate from the seed ---
    cap = 1_000_000
    rng = np.random.default_rng(0)
    tree = sum_tree.SumTree(cap)
    prio = rng.uniform(0.1, 1.0, cap)
    for lo in range(0, cap, 100_000):
        tree.set(np.arange(lo, lo + 100_000, dtype=np.int32), prio[lo:lo + 100_000])
    roots, queries = [tree.root], []
    for _ in range(200):
        idx = rng.integers(0, cap, 32).astype(np.int32)
        val = rng.uniform(0.0, 2.0, 32)
        tree.set(idx, val)
        t = rng.uniform(0.0, tree.root, 32)
        queries.append(tree.query(t))
        roots.append(tree.root)
    out["big_roots"] = np.asarray(roots)
    out["big_query"] = np.stack(queries)
    out["big_nodes_sha256"] = np.asarray(sha(tree._nodes))
    out["big_level_sums"] = np.asarray([tree._nodes[2 ** d - 1: 2 ** (d + 1) - 1].sum() for d in range(tree._depth)])
    np.savez_compressed(os.path.join(OUT, "sum_tree.npz"), **out)


def golden_samplers(samplers):
    out = {}
    # prioritised: adds, updates, removes (incl. last-index remove) interleaved with sampling.
    # exponent 1.0 -> p**1.0 == p exactly: portable across hosts.  exponent 0.6 -> numpy's vectorised pow is
    # CPU-dispatch dependent (AVX512 vs AVX2 vs libm differ by 1 ulp), so a probe of this host's pow is stored and
    # the bit-exact check of that sequence only runs where the probe reproduces.
    for tag, exponent in (("prio1_", 1.0), ("prio_", 0.6)):
        rng = np.random.default_rng(7)
        s = samplers.PrioritizedSamplingDistribution(seed=3, max_capacity=50, priority_exponent=exponent)
        ops, samples = [], []
        next_key, live = 0, []
        for it in range(300):
            r = rng.random()
            if not live or (r < 0.45 and len(live) < 50):
                p = float(rng.uniform(0, 2)) if rng.random() > 0.1 else 0.0
                s.add(next_key, priority=p)
                ops.append((0, next_key, p))
                live.append(next_key)
                next_key += 1
            elif r < 0.65:
                k = live[int(rng.integers(len(live)))]
                p = float(rng.uniform(0, 2))
                s.update(np.asarray([k]), np.asarray([p]))
                ops.append((1, k, p))
            elif r < 0.8 and len(live) > 1:
                k = live.pop(int(rng.integers(len(live))))
                s.remove(k)
                ops.append((2, k, 0.0))
            else:
                ops.append((3, -1, 0.0))
            if s._sum_tree.root > 0:
                samples.append(s.sample(8))
            else:
                samples.append(np.full(8, -1, np.int32))
        out[tag + "ops"] = np.asarray(ops, dtype=np.float64)
        out[tag + "samples"] = np.stack(samples)
        out[tag + "nodes"] = s._sum_tree._nodes.copy()
        out[tag + "index_to_key"] = np.asarray(s._index_to_key, np.int64)
    probe = np.random.default_rng(99).uniform(0, 2, 4096)
    out["pow_probe_in"] = probe
    out["pow_probe_array"] = probe ** 0.6
    out["pow_probe_scalar"] = np.asarray([float(v) ** 0.6 for v in probe])
    # uniform: FIFO evictions through swap-remove
    u = samplers.UniformSamplingDistribution(seed=11)
    usamples = []
    for key in range(200):
        u.add(key)
        if key >= 37:
            u.remove(key - 37)
        usamples.append(u.sample(6))
    out["uni_samples"] = np.stack(usamples)
    out["uni_index_to_key"] = np.asarray(u._index_to_key, np.int64)
    np.savez_compressed(os.path.join(OUT, "samplers.npz"), **out)


def golden_replay(replay_buffer, samplers):
    out = {}
    cfgs = [(4, 1, 0.99, 10), (4, 5, 0.9, 25), (1, 3, 1.0, 10), (2, 2, 0.5, 7)]
    for ci, (stack, n, gamma, cap) in enumerate(cfgs):
        rng = np.random.default_rng(100 + ci)
        rb = replay_buffer.ReplayBuffer(samplers.UniformSamplingDistribution(seed=ci), batch_size=8,
                                        max_capacity=cap, stack_size=stack, update_horizon=n, gamma=gamma,
                                        compress=False)
        T = 120
        obs = rng.integers(0, 256, (T, 3, 2)).astype(np.uint8)
        act = rng.integers(0, 5, T)
        rew = rng.uniform(-1, 1, T)
        term = rng.random(T) < 0.08
        trunc = (rng.random(T) < 0.05) | term
        batches = []
        for t in range(T):
            rb.add(replay_buffer.TransitionElement(obs[t], int(act[t]), float(rew[t]), bool(term[t]), bool(trunc[t])))
            if rb.add_count and t % 10 == 9:
                b = rb.sample()
                batches.append((b.state, b.next_state, b.action, b.reward, b.is_terminal))
        keys = np.asarray(list(rb._memory.keys()), np.int64)
        p = f"cfg{ci}_"
        out[p + "cfg"] = np.asarray([stack, n, gamma, cap], np.float64)
        out[p + "obs"], out[p + "act"], out[p + "rew"], out[p + "term"], out[p + "trunc"] = obs, act, rew, term, trunc
        out[p + "keys"] = keys
        out[p + "add_count"] = np.int64(rb.add_count)
        out[p + "state"] = np.stack([rb._memory[k].state for k in keys])
        out[p + "next_state"] = np.stack([rb._memory[k].next_state for k in keys])
        out[p + "action"] = np.asarray([rb._memory[k].action for k in keys], np.int64)
        out[p + "reward"] = np.asarray([rb._memory[k].reward for k in keys], np.float64)
        out[p + "is_terminal"] = np.asarray([rb._memory[k].is_terminal for k in keys], np.bool_)
        out[p + "b_state"] = np.stack([b[0] for b in batches])
        out[p + "b_next_state"] = np.stack([b[1] for b in batches])
        out[p + "b_action"] = np.stack([b[2] for b in batches])
        out[p + "b_reward"] = np.stack([b[3] for b in batches])
        out[p + "b_is_terminal"] = np.stack([b[4] for b in batches])
    np.savez_compressed(os.path.join(OUT, "replay_buffer.npz"), **out)


def main():
    install_stubs()
    sys.path.insert(0, REF)
    from slimdqn.sample_collection import replay_buffer, samplers, sum_tree

    golden_sum_tree(sum_tree)
    golden_samplers(samplers)
    golden_replay(replay_buffer, samplers)
    for f in ("sum_tree.npz", "samplers.npz", "replay_buffer.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
