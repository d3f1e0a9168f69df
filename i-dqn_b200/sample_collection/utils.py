"""Acting helpers — host-side mirror of slimdqn/sample_collection/utils.py (``select_action`` :8-15,
``collect_single_sample`` :18-40).  SURVEY §8(f) N1: outside the accelerated path (env-bound); kept so the
reference training loop runs unchanged against this package.  PRNG semantics: ``idqn_b200._prng``."""
from __future__ import annotations

from .. import _prng
from .replay_buffer import ReplayBuffer, TransitionElement


def select_action(best_action_fn, params, state, key, n_actions, epsilon_fn, n_training_steps):
    agent = getattr(best_action_fn, "__self__", None)
    engine = getattr(agent, "_engine", None)
    if engine is not None and getattr(params, "_engine_ref", lambda: None)() is engine and getattr(params, "_which", None) == 0 \
            and getattr(agent, "n_networks", 0) == engine.K and not getattr(agent, "_squeeze", False):
        # the agent's own live parameters: draws + (on a greedy step) the forward pass in one call into the library
        return engine.select_action(state, key, n_actions, epsilon_fn(n_training_steps))[0]
    uniform_key, action_key, kwargs_key = _prng.split(key, 3)
    if _prng.uniform(uniform_key) <= epsilon_fn(n_training_steps):  # utils.py:12
        return _prng.randint(action_key, 0, n_actions)
    return best_action_fn(params, state, key=kwargs_key)  # only the taken branch is evaluated


def collect_single_sample(key, env, agent, rb: ReplayBuffer, p, epsilon_schedule, n_training_steps: int):
    action = int(select_action(agent.best_action, agent.params, env.state, key, env.n_actions, epsilon_schedule,
                               n_training_steps))
    obs = env.observation
    reward, absorbing = env.step(action)
    episode_end = absorbing or env.n_steps >= p["horizon"]
    rb.add(TransitionElement(observation=obs, action=action,
                             reward=reward if rb._clipping is None else rb._clipping(reward),
                             is_terminal=absorbing, episode_end=episode_end))
    if episode_end:
        env.reset()
    return reward, episode_end
