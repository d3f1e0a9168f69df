"""FREE-RUNNING parity of the CUDA learning step against the committed float64 fixtures (tests/golden/network_*.npz).

Unlike tests/test_gpu_learn.py nothing is teacher-forced: the agent starts from the fixture's seeded parameters and runs
10 consecutive steps (T=8, D=4: one D-sync, one T-shift) on its own state; the float64 trajectory was computed once by
the hand-derived NumPy oracle (oracle/networks_np.py) and is never fed the GPU's state or relu gates.  Tolerances and
their growth per step: tests/golden_network.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_network as GN  # noqa: E402
from golden_network import G  # noqa: E402

pytestmark = pytest.mark.gpu


class GpuChain:
    """The product path through the reference's own class API (iDQN over the C ABI)."""

    def __init__(self, name, flags=0):
        from idqn_b200.networks.idqn import iDQN
        c = G.CONFIGS[name]
        obs = c["obs"] if c["arch"] in ("cnn", "impala") else c["obs"][0]
        self.agent = iDQN(0, obs, c["A"], c["K"], c["feats"], c["arch"], c["lr"], G.GAMMA, G.NSTEP, 1, G.T, G.D, c["eps"],
                          batch_size=G.B, flags=flags)

    def start(self, params, target):
        self.agent.params, self.agent.target_params = GN.N.nest_modules(params), GN.N.nest_modules(target)

    def step(self, batch):
        a = self.agent
        return a.learn_on_batch(a.params, a.target_params, a.optimizer_state, batch)[2]

    def state(self):
        st = self.agent.optimizer_state[0]
        return self.agent.params.to_host(), st.mu.to_host(), st.nu.to_host(), np.asarray(st.count)

    def events(self, step):
        self.agent.update_target_params(step)

    def final(self):
        return self.agent.target_params.to_host(), self.agent.params.to_host()


def first_step_gradients(name):
    """Step 1 with every gradient materialised (IDQN_F_KEEP_GRADS) against the fixture's step-1 gradients."""
    from idqn_b200 import _lib
    z, params, target, batches = GN.load(name)
    chain = GpuChain(name, flags=_lib.F_KEEP_GRADS)
    chain.start(params, target)
    losses = chain.step(batches[0])
    e_loss = float(np.max(np.abs(losses - z["losses"][0]) / np.abs(z["losses"][0])))
    return (e_loss,) + GN.tensor_errors(name, z, "s1/grad", chain.agent.gradients(), G.CONFIGS[name]["K"])


@pytest.mark.parametrize("name", ["mlp_k3", "cnn_k1", "cnn_k3", "cnn_k5", "cnn_k8", "impala_k2"])
def test_free_running_ten_steps_against_the_float64_fixture(name):
    """BASELINE configs [0], [1], [2], the K=5 benchmark configuration, K=8 (configs[3] on one GPU) and the impala
    architecture: 10 un-synchronised steps incl. one D-sync and one T-shift; losses every step, parameters / mu / nu / count at the checkpoint steps,
    target and online parameters after the last events."""
    GN.check(GN.run_chain(name, GpuChain(name)))


@pytest.mark.parametrize("name", ["mlp_k3", "cnn_k1", "cnn_k5", "cnn_k8", "impala_k2"])
def test_first_step_gradients_against_the_float64_fixture(name):
    e_loss, e_grad, where = first_step_gradients(name)
    assert e_loss <= 1e-4, f"loss off by {e_loss:.2e}"
    assert e_grad <= GN.tol(1, "grad"), f"gradient off by {e_grad:.2e} at {where}"


if __name__ == "__main__":  # drift report: python tests/test_gpu_golden.py [config ...]
    for nm in (sys.argv[1:] or ["mlp_k3", "cnn_k1", "cnn_k5", "cnn_k8"]):
        print(f"== {nm}: first-step (loss, gradient) errors", first_step_gradients(nm), flush=True)
        GN.run_chain(nm, GpuChain(nm), report=lambda row: print(nm, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in row.items()}, flush=True))
