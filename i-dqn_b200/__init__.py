"""idqn_b200 — B200-native (sm_100a) i-DQN learning step behind the slimdqn API of theovincent/i-DQN.

Module map (mirrors the reference package layout):

    idqn_b200.networks.idqn                -> slimdqn/networks/idqn.py          (iDQN, shift_params, sync_target_params)
    idqn_b200.networks.dqn                 -> slimdqn/networks/dqn.py           (DQN)
    idqn_b200.networks.architectures.dqn   -> slimdqn/networks/architectures/dqn.py (DQNNet)
    idqn_b200.sample_collection.*          -> slimdqn/sample_collection/*       (ReplayBuffer, samplers, SumTree, utils)
    idqn_b200.parallel                     -> head sharding over several B200s (no reference counterpart)

All arithmetic runs in hand-written CUDA inside ``libidqn_b200.so`` (C ABI: ``include/idqn_b200.h``), bound with
ctypes.  There is no CPU fallback: importing works anywhere, *using* an agent/tree/buffer without the library
or without a Blackwell GPU raises.
"""
__version__ = "0.1.0"

from ._lib import LibraryError, lib, library_path  # noqa: F401
