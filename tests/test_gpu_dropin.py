"""-m gpu: the drop-in SimulatedNetworkEnv (GPU-backed, MT19937 shared with Python's global
`random`) replays the UNMODIFIED reference's own trajectories bit for bit: random.seed(S);
SimulatedNetworkEnv(); reset(); 400 x step()."""
import random

import numpy as np
import pytest

from golden_util import assert_step_equal, golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("name", golden_names("mt_"))
def test_dropin_env_replays_reference(name, strict):
    """strict=False (default): the MT19937 state stays on the device between steps and meets Python's `random` at every
    reset; strict=True: exchanged around every step.  Same trajectories either way."""
    import pcc_rl_b200
    g = load_golden(name)
    random.seed(g["seed"])
    env = pcc_rl_b200.SimulatedNetworkEnv(strict_rng=strict)
    assert env.observation_space.shape == (30,) and env.action_space.shape == (1,)
    k = 0
    for ep in range(len(g["ep_params"])):
        obs0 = env.reset()
        p = env.link_params
        assert (p["bw"], p["lat"], float(p["queue"]), p["loss"], p["start_rate"]) == tuple(g["ep_params"][ep])
        assert np.array_equal(obs0, g["ep_obs0"][ep])
        for _ in range(g["steps_per_episode"]):
            obs, r, d, info = env.step([g["action"][k]])
            assert isinstance(info, dict)
            assert_step_equal(g, k, obs, r, d, env.last_counts, env.cur_time, env.run_dur, env.rate, what=name)
            k += 1
    env.close()


def test_dropin_survey_kat():
    import pcc_rl_b200
    random.seed(1234)
    env = pcc_rl_b200.SimulatedNetworkEnv()
    env.reset()
    want = [((206, 200, 5), 1.129879565565137, [0.0, 1.0, 1.035175879396985]),
            ((68, 64, 5), 0.9394036045958437, [0.0, 1.0, 1.0793650793650793]),
            ((69, 62, 6), 0.8506430480655987, [0.0, 1.0000000000000002, 1.1311475409836067])]
    for counts, reward, newest in want:
        obs, r, d, _ = env.step([0.0])
        assert env.last_counts == counts and float(r) == reward and obs[-3:].tolist() == newest and not d
    # float32 actions (what PPO passes) are promoted to float64 before use
    obs, r, d, _ = env.step(np.array([0.5], dtype=np.float32))
    assert np.isfinite(r)


def _gym_stub():
    """The four names the reference's module takes from gym (network_sim.py:15-18), as in oracle/refharness.py."""
    import sys
    import types
    if "gym" in sys.modules:
        return sys.modules["gym"]
    gym = types.ModuleType("gym")
    gym._pcc_stub = True

    class Env(object):
        pass

    class Box(object):
        def __init__(self, low, high, dtype=None):
            self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.dtype, self.shape = np.dtype(dtype), self.low.shape

    spaces, utils = types.ModuleType("gym.spaces"), types.ModuleType("gym.utils")
    seeding, envs = types.ModuleType("gym.utils.seeding"), types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.registry = {}
    registration.register = lambda id, entry_point=None, **kw: registration.registry.__setitem__(id, entry_point)
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    spaces.Box, utils.seeding, envs.registration = Box, seeding, registration
    gym.Env, gym.spaces, gym.utils, gym.envs = Env, spaces, utils, envs
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("gym.envs", envs), ("gym.envs.registration", registration)]:
        sys.modules[name] = mod
    return gym


def test_dropin_route_of_stable_solve(monkeypatch):
    """The actual drop-in route: the package directory first on sys.path, `import network_sim` as a TOP-LEVEL module
    (what stable_solve.py:21 does), the 'PccNs-v0' registration, construction through the registered entry point with
    the command-line defaults (--history-len / --input-features, network_sim.py:347-351), then stable_solve.py's call
    pattern -- float32 actions of shape (1,), reset() after done -- replaying the reference's own trajectory."""
    import importlib
    import os
    import sys
    gym = _gym_stub()
    pkg_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pcc-rl_b200")
    g = load_golden("mt_seed2019")
    monkeypatch.setattr(sys, "argv", ["stable_solve.py", "--history-len=%d" % g["history_len"]])
    monkeypatch.syspath_prepend(pkg_dir)
    sys.modules.pop("network_sim", None)
    ns = importlib.import_module("network_sim")
    try:
        assert os.path.dirname(os.path.abspath(ns.__file__)) == pkg_dir
        reg = gym.envs.registration
        if hasattr(reg, "registry") and isinstance(reg.registry, dict) and "PccNs-v0" in reg.registry:
            mod_name, cls_name = reg.registry["PccNs-v0"].split(":")
            cls = getattr(importlib.import_module(mod_name), cls_name)
        else:
            cls = ns.SimulatedNetworkEnv
        assert cls is ns.SimulatedNetworkEnv
        random.seed(g["seed"])
        env = cls()
        assert env.history_len == g["history_len"] and env.observation_space.shape == (3 * g["history_len"],)
        k = 0
        obs = env.reset()
        assert np.array_equal(obs, g["ep_obs0"][0])
        for ep in range(len(g["ep_params"])):
            for _ in range(g["steps_per_episode"]):
                a32 = np.array([g["action"][k]], dtype=np.float32)        # PPO1 hands float32 (1,) actions
                if float(a32[0]) != g["action"][k]:
                    pytest.skip("golden actions are not float32-representable")
                obs, r, d, info = env.step(a32)
                assert_step_equal(g, k, obs, r, d, env.last_counts, env.cur_time, env.run_dur, env.rate, what="route")
                k += 1
            if d and ep + 1 < len(g["ep_params"]):
                obs = env.reset()
                assert np.array_equal(obs, g["ep_obs0"][ep + 1])
        env.close()
        # a different command line changes the default constructor, like the reference's
        monkeypatch.setattr(sys, "argv", ["stable_solve.py", "--history-len=4", "--input-features=send rate,loss ratio"])
        sys.modules.pop("network_sim", None)
        ns2 = importlib.import_module("network_sim")
        e2 = ns2.SimulatedNetworkEnv()
        assert e2.history_len == 4 and e2.features == ["send rate", "loss ratio"] and e2.observation_space.shape == (8,)
        e2.reset()
        e2.step(np.zeros(1, dtype=np.float32))
        e2.close()
    finally:
        sys.modules.pop("network_sim", None)


def test_dropin_module_switches_route_to_the_variant_engine(monkeypatch):
    """USE_CWND / USE_LATENCY_NOISE set on the module (as one would edit network_sim.py:51-54): 2-dim action space
    (:376-379) and the variant engine; with the Philox seed and the link parameters of a golden file injected, the env
    reproduces the reference's own trajectory with the switch on."""
    import os
    from golden_util import GOLDEN_DIR
    import pcc_rl_b200
    from pcc_rl_b200 import network_sim as ns
    z = np.load(os.path.join(GOLDEN_DIR, "variant_cwnd.npz"))
    monkeypatch.setattr(ns, "USE_CWND", True)
    monkeypatch.setattr(random, "getrandbits", lambda k: int(z["seed"]))
    env = ns.SimulatedNetworkEnv(features=str(z["features"]))
    assert env.action_space.shape == (2,)
    eps = {"i": 0}

    def fixed_link():
        bw, lat, q, loss, rate = z["ep_params"][min(eps["i"], len(z["ep_params"]) - 1)]
        env.link_params = dict(bw=bw, lat=lat, queue=int(q), loss=loss, start_rate=rate)
        env.run_dur = 3 * lat
    env.create_new_links_and_senders = fixed_link
    obs = env.reset()
    assert np.array_equal(obs, z["ep_obs0"][0])
    for k in range(int(z["steps_per_episode"])):
        obs, r, d, _ = env.step(np.array([z["action"][k], z["cwnd_action"][k]]))
        assert env.last_counts == tuple(z["counts"][k]) and r == z["reward"][k] and np.array_equal(obs, z["obs"][k]), k
    env.close()


def test_event_recorder_from_live_batch_env(tmp_path):
    """EventRecorder fed by a live PccBatchEnv(want_info=True): the dumped record of env 3 equals, key for key, the
    record the drop-in single env keeps for the same link and the same loss stream."""
    import json
    import pcc_rl_b200
    n, steps = 16, 30
    env = pcc_rl_b200.PccBatchEnv(n_envs=n, seed=55, want_info=True, auto_reset=False)
    env.reset()
    rec = pcc_rl_b200.EventRecorder([3, 9])
    import oracle
    o = oracle.OracleEnv(10, oracle.DEFAULT_FEATURES)
    o.seed_philox(55 + 3)
    o.reset(*[env.params[k][3] for k in ("bw", "lat", "queue", "loss", "start_rate")])
    g = np.random.default_rng(2)
    want = []
    for t in range(steps):
        a = g.normal(0, 1, n)
        obs, r, d, info = env.step(a)
        rec.record(r, info["metrics"], d)
        _o, o_r, _d, _c, o_info = o.step(a[3])
        want.append((t + 1, o_r, o_info[:7]))
    path = tmp_path / "pcc_env_log_run_100.json"
    rec.dump(3, str(path))
    ev = json.load(open(path))["Events"]
    assert len(ev) == steps and list(ev[0].keys()) == list(pcc_rl_b200.event_log.EVENT_KEYS)
    for e, (tm, rw, inf) in zip(ev, want):
        assert e["Time"] == tm and e["Reward"] == rw
        assert [e[k] for k in pcc_rl_b200.event_log._INFO_COLUMNS] == list(inf)
