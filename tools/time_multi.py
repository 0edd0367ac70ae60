"""Throughput of the multi-sender path (BASELINE config 5: bw x delay grid, 2 senders per link) for its engines
(warp = streaming MI, one link per warp; thread = the same, one link per thread; heap = per-env event heap).
python tools/time_multi.py [grid_side] [steps] [modes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for mode in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("warp", "thread", "heap")):
    os.environ["PCC_MULTI_MODE"] = mode
    import pcc_rl_b200
    p = pcc_rl_b200.grid_sweep_params(n_bw=side, n_lat=side, queue=40, loss=0.01)
    n, S = side * side, 2
    g = np.random.default_rng(7)
    env = pcc_rl_b200.PccMultiSenderEnv(n, n_senders=S, seed=500, ring_capacity=1 << 13)
    r0 = g.uniform(40, 1000, (n, S))
    env.reset(p, r0)
    torch.cuda.synchronize()
    rs, re_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rs.record(); env.reset(p, r0); re_.record()
    torch.cuda.synchronize()
    reset_ms = rs.elapsed_time(re_)
    acts = torch.randn((steps + 5, n, S), dtype=torch.float64, device=env.device) * 2.0
    for t in range(5):
        env.step(acts[t])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(steps):
        env.step(acts[5 + t])
    e.record()
    torch.cuda.synchronize()
    env.check()
    ms = s.elapsed_time(e) / steps
    print("%-6s grid %dx%d, %d senders: %.3f ms/step = %.3f M env-steps/s (%.3f M sender-steps/s); reset of every link %.2f ms"
          % (mode, side, side, S, ms, n / ms / 1e3, n * S / ms / 1e3, reset_ms))
