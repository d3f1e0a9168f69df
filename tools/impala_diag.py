"""Per-tensor gradient error of one impala learning step against the fp32 oracle (and the float64 oracle's distance to it)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import networks as O
from idqn_b200 import _lib as L
from idqn_b200.networks.idqn import iDQN
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_impala import batch_of, rel_l2, gpu_gates

u8 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
obs, feats, A, K, B = (22, 20, 4), [8, 6, 8, 16], 4, 2, 8
rng = np.random.default_rng(11 + u8)
agent = iDQN(0, obs, A, K, feats, "impala", 1e-3, 0.94, 1, 1, 2, 1, 1e-5, batch_size=B, flags=L.F_KEEP_GRADS)
agent.params = O.init_params(rng, obs, feats, "impala", A, n_networks=K, bias_scale=0.05)
agent.target_params = O.init_params(rng, obs, feats, "impala", A, n_networks=K, bias_scale=0.05)
batch = batch_of(rng, B, obs, A, bool(u8))
s_p, s_t = agent.params.to_host(), agent.target_params.to_host()
_, _, g_l = agent.learn_on_batch(agent.params, agent.target_params, agent.optimizer_state, batch)
got = agent.gradients()
for k in range(K):
    l32, g32 = O.loss_and_grad(O.tree_index(s_p, k), O.tree_index(s_t, k), batch, "impala", 0.94, 1, torch.float32)
    l64, g64, z64 = O.loss_and_grad(O.tree_index(s_p, k), O.tree_index(s_t, k), batch, "impala", 0.94, 1, torch.float64, preacts=True)
    gates, flips = gpu_gates(agent, k, z64)
    print(f"head {k}: loss gpu {g_l[k]:.7f} fp32 {l32:.7f} fp64 {l64:.7f}; gate flips vs fp64: {flips}")

    def walk(a, b, c, path=""):
        if isinstance(b, dict):
            for kk in b:
                walk(a[kk], b[kk], c[kk], path + "/" + kk)
        else:
            print(f"   {path:32s} gpu-vs-fp64 {rel_l2(np.asarray(a)[k], c):.2e}   fp32-vs-fp64 {rel_l2(b, c):.2e}   |g| {np.linalg.norm(c):.3e}")
    walk(got["params"], g32["params"], g64["params"])

# ---- where inside Stack_1/Conv_0's kernel gradient (head 1) the error sits, and whether the stored activations still match
k = 1
l64, g64, z64 = O.loss_and_grad(O.tree_index(s_p, k), O.tree_index(s_t, k), batch, "impala", 0.94, 1, torch.float64, preacts=True)
gk = np.asarray(got["params"]["Stack_1"]["Conv_0"]["kernel"])[k]
wk = g64["params"]["Stack_1"]["Conv_0"]["kernel"]
err = np.abs(gk - wk).reshape(9, gk.shape[2], gk.shape[3])
print("S1/C0 kernel |err| per tap (max over c, o):", np.round(err.max(axis=(1, 2)), 5), " |g| max", np.abs(wk).max())
print("S1/C0 kernel |err| per input channel:", np.round(err.max(axis=(0, 2)), 5))
print("S1/C0 kernel |err| per output channel:", np.round(err.max(axis=(0, 1)), 5))
eng = agent._engine
# ---- max-pool argmax decisions: the GPU's (recomputed from its stored Conv_0 outputs) against the float64 oracle's
from oracle import networks_np as N
for k in range(K):
    p64 = O.tree_index(s_p, k)
    for st in range(3):
        k0, b0 = N._impala_leaf(p64["params"], st, 0)
        # input of this Stack's Conv_0 in float64: replay the oracle forward up to it
        hh = np.asarray(batch["state"], np.float64) / 255.0
        for i in range(st + 1):
            kk, bb = N._impala_leaf(p64["params"], i, 0)
            z, _ = N.conv_forward(hh, kk, bb, 1)
            if i == st:
                break
            hh, _ = N.maxpool_forward(z)
            for b in range(2):
                ka, ba = N._impala_leaf(p64["params"], i, 1 + 2 * b)
                kb, bb2 = N._impala_leaf(p64["params"], i, 2 + 2 * b)
                za, _ = N.conv_forward(np.maximum(hh, 0.0), ka, ba, 1)
                zb, _ = N.conv_forward(np.maximum(za, 0.0), kb, bb2, 1)
                hh = zb + hh
        zg = eng.download_activation(k, 6 * st).reshape(z.shape).astype(np.float64)
        y64, c64 = N.maxpool_forward(z)
        yg, cg = N.maxpool_forward(zg)
        diff = c64[0] != cg[0]
        print(f"head {k} Stack_{st}: conv0 rel-L2 {rel_l2(zg, z):.2e}; pool argmax differs in {int(diff.sum())} of {diff.size} windows", end="")
        if diff.any():
            idx = np.argwhere(diff)
            print("; channels", sorted(set(idx[:, 3].tolist())), end="")
        print()
