"""Experiment helper (needs a library built with -DPCC_PROFILE, PCC_B200_LIB=...): per-phase SM cycles of the
group kernel for the heaviest envs of a step.   python tools/phase_profile.py [n_envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcc_rl_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
env = pcc_rl_b200.PccBatchEnv(n_envs=n, seed=100, want_info=True, auto_reset=False)
env.reset()
g = torch.Generator(device=env.device); g.manual_seed(101)
names = ["send", "hop1", "bnd1", "hop2", "bnd2", "cross", "means", "sent", "acked", "total"]
if os.environ.get("PCC_B200_MODE", "") == "warp" or (os.environ.get("PCC_B200_MODE") is None and n > 16384):
    names = ["warp_send", "consume", "means", "warp_phaseB", "envs_in_warp", "-", "-", "sent", "acked", "total"]
for t in range(steps):
    a = torch.randn(n, generator=g, device=env.device, dtype=torch.float64)
    obs, r, d, info = env.step(a)
m = info["metrics"].cpu().numpy()
c = info["counts"].cpu().numpy()
print("counts of the last step: mean sent/acked/lost =", c.mean(0), " max sent =", c[:, 0].max())
print("cur_time col mean", env.column("cur_time").mean().item(), "run_dur mean", env.column("run_dur").mean().item(),
      "rate mean", env.column("rate").mean().item())
order = np.argsort(-m[:, 9])
print("last step: per-env cycles; columns:", names)
for i in order[:8]:
    print("env %5d " % i + " ".join("%s=%d" % (nm, m[i, k]) for k, nm in enumerate(names)))
print("median env:", " ".join("%s=%d" % (nm, np.median(m[:, k])) for k, nm in enumerate(names)))
print("mean   env:", " ".join("%s=%d" % (nm, np.mean(m[:, k])) for k, nm in enumerate(names)))
tot = m[:, 9]
print("max total %d cycles (%.1f us at 1.965 GHz); p99 %d; sum/148/4 = %d" % (tot.max(), tot.max() / 1965, np.percentile(tot, 99), tot.sum() / 592))
