from .dqn import DQN  # noqa: F401
from .idqn import iDQN, shift_params, sync_target_params  # noqa: F401
