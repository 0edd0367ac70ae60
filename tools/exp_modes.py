"""Experiment driver: device time per step of the env-step kernels under several execution modes (environment
variables read by pcc_create).  python tools/exp_modes.py N 'K=V,K=V' 'K=V' ...   (each argument = one setting)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pcc_rl_b200


def run(n, setting, K=60, W=30, seed=100):
    keys = []
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    dev = torch.device("cuda", 0)
    env = pcc_rl_b200.PccBatchEnv(n_envs=n, device=dev, seed=seed, auto_reset=False)
    for k in keys:
        del os.environ[k]
    env.reset()
    g = torch.Generator(device=dev); g.manual_seed(seed + 1)
    acts = torch.randn((W + K, n), generator=g, device=dev, dtype=torch.float64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for t in range(W):
        env.step_device(acts[t])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    tot = torch.zeros(3, dtype=torch.int64, device=dev)
    for t in range(K):
        flush.fill_(t & 255)
        ev[t][0].record()
        env.step_device(acts[W + t])
        ev[t][1].record()
        tot += env.counts.sum(0)
    torch.cuda.synchronize()
    env.check()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    chk = float(env.reward.sum().item())
    print(json.dumps(dict(n=n, setting=setting, ms_median=round(ms[len(ms) // 2], 4), ms_min=round(ms[0], 4),
                          ms_mean=round(sum(ms) / len(ms), 4), M_env_steps_s=round(n / ms[len(ms) // 2] / 1e3, 1),
                          sent=int(tot[0]), reward_sum=chk)), flush=True)
    env.close()
    del env, flush, acts
    torch.cuda.empty_cache()


if __name__ == "__main__":
    n = int(sys.argv[1])
    W = int(os.environ.get("EXP_WARMUP", "30"))      # the timed window starts at this step of the episode
    for s in sys.argv[2:]:
        run(n, s, W=W)
