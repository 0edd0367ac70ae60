"""Experiment helper: the episode-boundary step of the batched env (step 400: step kernel + reset of every env with
fresh link parameters) as the GPU sees it, with the host queue running ahead as in bench.py's timed loop.
python tools/time_boundary.py [n_envs] [episodes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcc_rl_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
env = pcc_rl_b200.PccBatchEnv(n_envs=n, device=dev, seed=100)
env.reset()
g = torch.Generator(device=dev); g.manual_seed(5)
acts = torch.randn((400, n), generator=g, device=dev, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if os.environ.get("FLUSH", "1") == "1" else None
def ev(): return torch.cuda.Event(enable_timing=True)
for ep in range(eps):
    evs = []
    t_host = []
    for t in range(400):
        if flush is not None:
            flush.fill_(t & 255)
        s, e = ev(), ev()
        h0 = time.perf_counter()
        s.record(); env.step(acts[t]); e.record()
        t_host.append(time.perf_counter() - h0)
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    print("episode %d: median step %.3f ms, steps 396-399: %s ms; host time of the boundary call %.2f ms (median call %.3f ms)"
          % (ep, float(np.median(ms)), " ".join("%.2f" % x for x in ms[396:]), 1e3 * t_host[399], 1e3 * float(np.median(t_host))))
# the pieces, synchronised: reset kernel + uploads alone
for k in range(2):
    for t in range(399):
        env.step_device(acts[t])
    torch.cuda.synchronize()
    s, e = ev(), ev()
    h0 = time.perf_counter(); s.record(); env.reset(); e.record(); torch.cuda.synchronize()
    print("explicit env.reset() after a drained queue: device %.2f ms, wall %.2f ms" % (s.elapsed_time(e), 1e3 * (time.perf_counter() - h0)))
