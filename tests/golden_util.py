"""Helpers shared by the oracle and CUDA parity tests: load tests/golden/*.npz (outputs of the
unmodified reference, see oracle/gen_golden.py) and replay them through a backend."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["features"] = str(g["features"])
    g["rng"] = str(g["rng"])
    g["seed"] = int(g["seed"])
    g["history_len"] = int(g["history_len"])
    g["steps_per_episode"] = int(g["steps_per_episode"])
    return g


def assert_step_equal(g, k, obs, reward, done, counts, cur_time=None, run_dur=None, rate=None,
                      info=None, what=""):
    """Bit-exact comparison of step k of golden file g with a backend's outputs."""
    tag = "%s step %d" % (what, k)
    assert tuple(int(c) for c in counts) == tuple(int(c) for c in g["counts"][k]), tag + " counts"
    assert np.array_equal(np.asarray(obs, dtype=np.float64), g["obs"][k]), tag + " obs"
    assert float(reward) == g["reward"][k], tag + " reward"
    assert bool(done) == bool(g["done"][k]), tag + " done"
    if cur_time is not None:
        assert cur_time == g["cur_time"][k], tag + " cur_time"
    if run_dur is not None:
        assert run_dur == g["run_dur"][k], tag + " run_dur"
    if rate is not None:
        assert rate == g["rate"][k], tag + " rate"
    if info is not None:
        assert np.array_equal(np.asarray(info)[:7], g["info"][k]), tag + " info"
