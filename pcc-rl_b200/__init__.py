"""pcc_rl_b200 -- B200-native batched congestion-control simulator: the gym hot path of
PCCproject/PCC-RL (network_sim.py + sender_obs.py) as one hand-written sm_100a CUDA kernel
behind the reference's gym.Env surface.  Import name: `pcc_rl_b200` (this directory is
`pcc-rl_b200/`; the alias module pcc_rl_b200.py at the repo root maps the import name to it)."""
from . import build as build_mod
from . import _lib, sender_obs, params
from .params import LinkRanges, sample_link_params
from .batch_env import PccBatchEnv
from .multi_env import PccMultiSenderEnv, grid_sweep_params
from .flow_monitor import PccFlowMonitor
from . import flow_monitor
from . import event_log
from .event_log import EventRecorder


def build(force=False, verbose=False):
    """Compiles libpcc_b200.so for sm_100a with nvcc (no GPU needed)."""
    return build_mod.build(force=force, verbose=verbose)


def load_library():
    return _lib.load()


def __getattr__(name):  # lazy: importing network_sim touches gym registration
    if name in ("SimulatedNetworkEnv", "network_sim", "distributed"):
        import importlib
        mod = importlib.import_module("." + ("distributed" if name == "distributed" else "network_sim"), __name__)
        return mod.SimulatedNetworkEnv if name == "SimulatedNetworkEnv" else mod
    if name in ("ShimNetworkEnv", "shim_env"):   # opens a TCP socket when constructed (gym/online/shim_env.py)
        import importlib
        mod = importlib.import_module(".shim_env", __name__)
        return mod.ShimNetworkEnv if name == "ShimNetworkEnv" else mod
    raise AttributeError(name)
