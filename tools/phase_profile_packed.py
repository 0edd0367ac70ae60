"""Experiment helper (needs the -DPCC_PROFILE library: tools/build_profile_lib.sh, PCC_B200_LIB=gpurun_exp_prof/libpcc_b200_prof.so):
where the cycles of the packed env-step kernel go.   python tools/phase_profile_packed.py [n_envs] [steps]
Per env the profiling build reports the cycles of ITS WARP's phases (A sends, B1 hop-1, B2 hop-2, crossing, B3 means, emit),
envs in the warp (-1: a solo warp), packets sent / acked, the warp's total cycles and its start / end on the global timer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcc_rl_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60      # the profile is taken at the last 3 of these steps of the episode
env = pcc_rl_b200.PccBatchEnv(n_envs=n, seed=100, want_info=True, auto_reset=False)
env.reset()
g = torch.Generator(device=env.device); g.manual_seed(101)
acc = []
for t in range(steps):
    a = torch.randn(n, generator=g, device=env.device, dtype=torch.float64)
    obs, r, d, info = env.step(a)
    if t >= steps - 3:
        acc.append(info["metrics"].cpu().numpy().copy())
names = ["A sends", "B1 hop-1", "B2 hop-2", "crossing", "B3 means", "emit"]
for m in acc:
    cnt, sent, acked, tot, g0, g1 = m[:, 6], m[:, 7], m[:, 8], m[:, 9], m[:, 10], m[:, 11]
    t0 = g0.min()
    print("step: kernel span %.1f us (global timer, first warp start to last warp end)" % ((g1.max() - t0) / 1e3))
    solo = cnt == -1
    pk = cnt > 0
    quad = cnt == -4
    if solo.any():
        print("  solo warps %d: packets mean %.0f max %.0f | cycles mean %.0f max %.0f (%.1f per packet) | last end at %.1f us"
              % (solo.sum(), sent[solo].mean(), sent[solo].max(), tot[solo].mean(), tot[solo].max(),
                 (tot[solo] / np.maximum(sent[solo], 1)).mean(), (g1[solo].max() - t0) / 1e3))
        ms = m[solo]
        print("     solo per packet: send %.1f  consume %.1f  means %.1f cycles"
              % tuple((ms[:, i] / np.maximum(ms[:, 7], 1)).mean() for i in (0, 1, 2)))
    if quad.any():
        q = m[quad]
        print("  quad envs %d (%.0f warps): packets mean %.0f max %.0f | warp cycles mean %.0f max %.0f (%.1f per packet of the env) | "
              "send %.1f  hop1 %.1f  hop2+cross %.1f  means %.1f  emit %.1f cycles per packet | last end at %.1f us"
              % (quad.sum(), quad.sum() / 4, q[:, 7].mean(), q[:, 7].max(), q[:, 9].mean(), q[:, 9].max(),
                 (q[:, 9] / np.maximum(q[:, 7], 1)).mean(), *[(q[:, i] / np.maximum(q[:, 7], 1)).mean() for i in (0, 1, 2, 4, 5)],
                 (q[:, 11].max() - t0) / 1e3))
    # one representative per packed warp: the lane with the most packets
    order = np.argsort(-tot[pk])
    mp = m[pk][order]
    # group lanes of a warp: same (start time, total)
    key = np.stack([mp[:, 10], mp[:, 9]], 1)
    _, idx, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    nw = len(idx)
    wmax = np.zeros(nw); wsum = np.zeros(nw); wcnt = np.zeros(nw)
    np.maximum.at(wmax, inv, mp[:, 7]); np.add.at(wsum, inv, mp[:, 7]); np.add.at(wcnt, inv, 1)
    rep = mp[idx]
    print("  packed warps %d: rows (max packets of the warp) mean %.0f max %.0f, lane balance (mean/max packets) %.2f"
          % (nw, wmax.mean(), wmax.max(), (wsum / wcnt / np.maximum(wmax, 1)).mean()))
    print("  packed warp cycles: mean %.0f  p99 %.0f  max %.0f | sum %.3e -> / (148 x 16) = %.1f us"
          % (rep[:, 9].mean(), np.percentile(rep[:, 9], 99), rep[:, 9].max(), rep[:, 9].sum(), rep[:, 9].sum() / (148 * 16) / 1965))
    for lo, hi in ((0, 32), (32, 128), (128, 512), (512, 1e9)):
        sel = (wmax >= lo) & (wmax < hi)
        if sel.any():
            ph = rep[sel][:, :6]
            print("   warps with %4d <= rows < %-6g: %5d | cycles per row: " % (lo, hi, sel.sum())
                  + "  ".join("%s %.0f" % (nm, (ph[:, i] / np.maximum(wmax[sel], 1)).mean()) for i, nm in enumerate(names))
                  + " | total %.0f/row, %.0f cycles" % ((rep[sel][:, 9] / np.maximum(wmax[sel], 1)).mean(), rep[sel][:, 9].mean()))
    hv = np.argmax(rep[:, 9])
    print("   slowest packed warp: rows %.0f, cycles %s total %.0f, ends at %.1f us"
          % (wmax[hv], " ".join("%s %.0f" % (nm, rep[hv, i]) for i, nm in enumerate(names)), rep[hv, 9], (rep[hv, 11] - t0) / 1e3))
