"""Test infrastructure: import the UNMODIFIED reference (PCCproject/PCC-RL) in this container.

This is NOT product code.  It exists only where /root/reference exists (the build
container); nothing under tests -m gpu, smoke() or bench.py may depend on it.

What it does:
  * injects a ~20-line in-memory `gym` stub into sys.modules (gym.Env, gym.spaces.Box,
    gym.utils.seeding.np_random, gym.envs.registration.register) — the reference imports
    those four names at network_sim.py:15-18 and uses nothing else from gym;
  * puts /root/reference/src/gym on sys.path and imports `network_sim` with a clean argv
    (the reference scans sys.argv at import, simple_arg_parse.py:17-23);
  * silences the reference's prints and keeps its JSON dumps out of the repo (cwd → tmp);
  * offers a per-env RNG shim so that several lock-stepped reference envs can each consume
    their own stream (`network_sim.random = shim`); no reference file is modified.
"""
import contextlib
import io
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("PCC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "gym", "network_sim.py"))


def _install_gym_stub():
    if "gym" in sys.modules and not getattr(sys.modules["gym"], "_pcc_stub", False):
        return  # a real gym is importable: use it
    import numpy as np

    gym = types.ModuleType("gym")
    gym._pcc_stub = True

    class Env(object):
        pass

    class Box(object):
        def __init__(self, low, high, dtype=None):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.dtype = np.dtype(dtype)
            self.shape = self.low.shape

    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")

    def np_random(seed=None):
        return np.random.RandomState(seed), seed

    seeding.np_random = np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.registry = {}

    def register(id, entry_point=None, **kw):
        registration.registry[id] = entry_point

    registration.register = register
    envs.registration = registration
    gym.Env = Env
    gym.spaces = spaces
    gym.utils = utils
    gym.envs = envs
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils),
                      ("gym.utils.seeding", seeding), ("gym.envs", envs),
                      ("gym.envs.registration", registration)]:
        sys.modules[name] = mod


_ref_mod = None


def load_reference():
    """Returns the reference's `network_sim` module (imported once)."""
    global _ref_mod
    if _ref_mod is not None:
        return _ref_mod
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_gym_stub()
    gym_dir = os.path.join(REFERENCE_ROOT, "src", "gym")
    saved_argv = sys.argv
    sys.argv = ["refharness"]
    sys.path.insert(0, gym_dir)
    # a product module is also called network_sim: make sure we import the reference's
    saved = sys.modules.pop("network_sim", None)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import importlib
            mod = importlib.import_module("network_sim")
    finally:
        sys.argv = saved_argv
        sys.path.remove(gym_dir)
    sys.modules["pcc_reference_network_sim"] = mod
    sys.modules.pop("network_sim", None)
    if saved is not None:
        sys.modules["network_sim"] = saved
    _ref_mod = mod
    return mod


@contextlib.contextmanager
def quiet_tmp_cwd():
    """The reference prints on every reset and dumps JSON into cwd every 100 episodes."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                yield
        finally:
            os.chdir(old)


class StreamShim(object):
    """Replacement for the `random` module inside the reference's network_sim namespace
    (`network_sim.random = shim`; no reference file is modified).

    `streams` is a list of objects with .random(); `select(i)` picks which one feeds the
    reference's random.random()/random.uniform() calls (network_sim.py:73, 455-466).
    uniform(a, b) follows CPython: a + (b - a) * random() -- unless `script` holds values,
    which are then handed back in order instead (used to make create_new_links_and_senders,
    network_sim.py:454-467, build exactly the link we want: its five uniform() calls are, in
    order, bw, lat, queue exponent, loss, start-rate factor).
    """

    def __init__(self, streams):
        self.streams = streams
        self.cur = 0
        self.script = []

    def select(self, i):
        self.cur = i

    def random(self):
        return self.streams[self.cur].random()

    def uniform(self, a, b):
        if self.script:
            return self.script.pop(0)
        return a + (b - a) * self.streams[self.cur].random()


# ---------------------------------------------------------------------------------------------
# The MI-sample ingestion path: common/sender_obs.py, the loaded-model
# client udt-plugins/testing/loaded_client.py and the online shim (gym/online/shim_env.py +
# udt-plugins/training/shim.py), all imported unmodified.
# ---------------------------------------------------------------------------------------------
class StubAgent(object):
    """Stands in for loaded_agent.LoadedModelAgent (TensorFlow saved model; TF is absent here):
    a fixed, deterministic function of the observation.  `StubAgent.fn` may be replaced."""

    fn = None

    def __init__(self, model_path=None):
        self.n_reset = 0

    def reset(self):
        self.n_reset += 1

    def act(self, ob):
        return StubAgent.fn(ob)


_ref_flows = None


def load_reference_flows():
    """Returns (sender_obs, loaded_client) of the reference.  loaded_client imports `loaded_agent`
    (TensorFlow); a stub module with the same class name is injected for the import."""
    global _ref_flows
    if _ref_flows is not None:
        return _ref_flows
    ns = load_reference()                      # makes `common.sender_obs` the reference's module
    sender_obs = sys.modules["common.sender_obs"]
    assert ns.sender_obs is sender_obs
    stub = types.ModuleType("loaded_agent")
    stub.LoadedModelAgent = StubAgent
    testing_dir = os.path.join(REFERENCE_ROOT, "src", "udt-plugins", "testing")
    saved_argv, saved_agent = sys.argv, sys.modules.get("loaded_agent")
    sys.argv = ["refharness"]
    sys.modules["loaded_agent"] = stub
    sys.path.insert(0, testing_dir)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import importlib
            lc = importlib.import_module("loaded_client")
    finally:
        sys.argv = saved_argv
        sys.path.remove(testing_dir)
        if saved_agent is None:
            sys.modules.pop("loaded_agent", None)
        else:
            sys.modules["loaded_agent"] = saved_agent
    sys.modules["pcc_reference_loaded_client"] = sys.modules.pop("loaded_client")
    assert lc.sender_obs is sender_obs
    _ref_flows = (sender_obs, lc)
    return _ref_flows


_ref_shim = None


def load_reference_shim():
    """Returns (shim_env, shim): the online training shim, server and client side.  Both use a real TCP
    socket on localhost:9787 (shim_env.py:62-64, shim.py:24-25); the caller runs them in two threads."""
    global _ref_shim
    if _ref_shim is not None:
        return _ref_shim
    load_reference()
    _install_gym_stub()
    saved_argv = sys.argv
    sys.argv = ["refharness"]
    dirs = [os.path.join(REFERENCE_ROOT, "src", "gym", "online"),
            os.path.join(REFERENCE_ROOT, "src", "udt-plugins", "training")]
    for d in dirs:
        sys.path.insert(0, d)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import importlib
            shim_env = importlib.import_module("shim_env")
            shim = importlib.import_module("shim")
    finally:
        sys.argv = saved_argv
        for d in dirs:
            sys.path.remove(d)
    sys.modules["pcc_reference_shim_env"] = sys.modules.pop("shim_env")
    sys.modules["pcc_reference_shim"] = sys.modules.pop("shim")
    _ref_shim = (shim_env, shim)
    return _ref_shim
