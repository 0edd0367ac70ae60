"""Experiment helper (needs a library built with -DPCC_PROFILE, PCC_B200_LIB=...): where the cycles of the warp
kernel go at a large batch.   python tools/phase_profile_warp.py [n_envs] [steps]
Per env the profiling build reports: cycles of its warp's phase A (send chains), its own consume and means
cycles inside phase B, its warp's whole phase B, envs in its warp, packets sent / acked, warp total."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcc_rl_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
env = pcc_rl_b200.PccBatchEnv(n_envs=n, seed=100, want_info=True, auto_reset=False)
env.reset()
g = torch.Generator(device=env.device); g.manual_seed(101)
acc = []
for t in range(steps):
    a = torch.randn(n, generator=g, device=env.device, dtype=torch.float64)
    obs, r, d, info = env.step(a)
    if t >= steps - 8:
        acc.append(info["metrics"].cpu().numpy().copy())
for m in acc[-3:]:
    sendA, cons, means, phB, cnt, sent, acked, tot = m[:, 0], m[:, 1], m[:, 2], m[:, 3], m[:, 4], m[:, 7], m[:, 8], m[:, 9]
    # one representative per warp: envs of a warp share sendA / phB / tot; weight 1/cnt
    w = 1.0 / np.maximum(cnt, 1)
    n_warps = w.sum()
    print("warps %.0f  envs/warp mean %.1f | per-warp cycles: phaseA mean %.0f  phaseB mean %.0f  total mean %.0f  max %.0f  p99 %.0f"
          % (n_warps, n / n_warps, (sendA * w).sum() / n_warps, (phB * w).sum() / n_warps, (tot * w).sum() / n_warps,
             tot.max(), np.percentile(tot, 99)))
    print("   sum of warp totals = %.3e cycles -> / (148 SMs x 16 warps) = %.0f cycles = %.1f us at 1.965 GHz"
          % ((tot * w).sum(), (tot * w).sum() / (148 * 16), (tot * w).sum() / (148 * 16) / 1965))
    print("   per env: consume mean %.0f  means mean %.0f cycles; packets sent mean %.1f acked %.1f"
          % (cons.mean(), means.mean(), sent.mean(), acked.mean()))
    for lo, hi in ((0, 16), (16, 64), (64, 256), (256, 1024), (1024, 1e9)):
        sel = (sent >= lo) & (sent < hi)
        if sel.any():
            print("   sent in [%d,%g): %6d envs  consume %.0f  means %.0f cycles/env   (%.1f / %.1f per packet)"
                  % (lo, hi, sel.sum(), cons[sel].mean(), means[sel].mean(), cons[sel].mean() / max(sent[sel].mean(), 1),
                     means[sel].mean() / max(acked[sel].mean(), 1)))
    single = cnt == 1
    multi = cnt >= 8
    if multi.any():
        print("   multi-env warps: phaseA cycles per (max packet of warp) ~ %.1f ; envs %d" % (
            (sendA[multi] / np.maximum(sent[multi], 1)).mean(), multi.sum()))
    if single.any():
        print("   single-env warps: %d, phaseA %.0f cycles for %.0f packets (%.1f/packet)" % (
            single.sum(), sendA[single].mean(), sent[single].mean(), sendA[single].mean() / max(sent[single].mean(), 1)))
