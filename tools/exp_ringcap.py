"""Experiment: how much of the env-step time is the 1 MiB-per-env ring stride?
Runs the device pass of config 3 with several global ring capacities (overflowing envs are ignored:
their results are wrong, but the timing of the rest is representative) and prints the sent-per-step
distribution of the batch."""
import json, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import pcc_rl_b200

def run(n, cap, K=60, W=30, seed=100):
    dev = torch.device("cuda", 0)
    env = pcc_rl_b200.PccBatchEnv(n_envs=n, device=dev, seed=seed, auto_reset=False, ring_capacity=cap)
    env.reset()
    g = torch.Generator(device=dev); g.manual_seed(seed + 1)
    acts = torch.randn((W + K, n), generator=g, device=dev, dtype=torch.float64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for t in range(W):
        env.step_device(acts[t])
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sent = []
    for t in range(K):
        flush.fill_(t & 255)
        ev[t][0].record()
        env.step_device(acts[W + t])
        ev[t][1].record()
        if t % 20 == 0:
            sent.append(env.counts[:, 0].cpu().numpy().copy())
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    s = np.concatenate(sent)
    try:
        env.check(); ovf = None
    except Exception as e:
        ovf = str(e)[:100]
    out = dict(n=n, cap=cap, ms_median=ms[len(ms) // 2], ms_min=ms[0], ms_max=ms[-1],
               sent_mean=float(s.mean()), sent_pct={p: float(np.percentile(s, p)) for p in (50, 90, 99, 99.9)},
               sent_max=int(s.max()), frac_pk_gt256=float(s[s > 256].sum() / s.sum()), frac_pk_gt1024=float(s[s > 1024].sum() / s.sum()),
               frac_pk_gt2048=float(s[s > 2048].sum() / s.sum()), n_gt2048=int((s > 2048).sum() / len(sent)), overflow=ovf)
    print(json.dumps(out), flush=True)
    env.close()
    del env, flush, acts
    torch.cuda.empty_cache()

if __name__ == "__main__":
    for n, cap in ((65536, 65536), (65536, 16384), (65536, 4096), (4096, 65536), (4096, 4096)):
        run(n, cap)
