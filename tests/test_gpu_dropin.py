"""-m gpu: the drop-in SimulatedNetworkEnv (GPU-backed, MT19937 shared with Python's global
`random`) replays the UNMODIFIED reference's own trajectories bit for bit: random.seed(S);
SimulatedNetworkEnv(); reset(); 400 x step()."""
import random

import numpy as np
import pytest

from golden_util import assert_step_equal, golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("mt_"))
def test_dropin_env_replays_reference(name):
    import pcc_rl_b200
    g = load_golden(name)
    random.seed(g["seed"])
    env = pcc_rl_b200.SimulatedNetworkEnv()
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
