#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched MI simulator on N B200s, next to the CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config2|config4|config5|flows]

One "step" = one monitor interval for every env of the batch (one pcc_step call), auto-reset
included at its natural frequency when K >= 400 (the window then covers a whole 400-step episode).
For a shorter K the window is centred on the middle of an episode and the boundary step (reset kernel +
fresh link parameters for every env) is timed separately and reported as config.episode_boundary.
Workloads (BASELINE.json `configs`):
    config3: 65 536 envs per GPU (524 288 on 8), default (ICML'19) ranges, fresh parameters at every reset   [default]
    config2: 4 096 envs per GPU, same ranges, history_len 10 (also reported as a sub-key of the default line)
    config4: the closed-loop rollout (policy + value head on the device, 8 192 steps by default; also a sub-key)
Weak scaling: every rank owns --envs envs (global ids are contiguous, parameters and RNG streams
are functions of the global id); no data-path collective; one all-gather of episode returns.

Timing: W untimed warm-up steps, then K steps, each bracketed by CUDA events on the launching
stream, with a 256 MiB L2-flush write between steps (outside the brackets); the timed region as
a whole is bracketed by barrier + torch.cuda.synchronize(); time = sum of the K brackets, max
over ranks.  A second pass runs K steps back to back without flushes (`back_to_back`).  `e2e`
is the same K steps through pcc_step_host: pinned HOST actions in, obs/reward/done out, copies
and synchronisation inside the wall-clock timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    "config2": dict(envs=4096, desc="4 096 envs on 1xB200, link params sampled from ICML'19 ranges, history_len=10"),
    "config3": dict(envs=65536, desc="65 536 envs on 1xB200, per-reset randomized bw/lat/queue/loss, 1 sender per env"),
    "config4": dict(envs=65536, desc="65 536 envs per GPU (524 288 on 8), PPO-style rollout end to end: on-device MLP policy "
                                     "30-32-16-1 + value head + env step + auto-reset (pcc_rollout), NCCL gather of episode returns"),
}
WORKLOADS["flows"] = dict(envs=1 << 20, desc="MI-sample ingestion (SURVEY 8f rank 4): 1 Mi live flows per GPU, one MI record per flow "
                                             "per step (~150 RTT samples each), history_len=10, 3 features")
WORKLOADS["config5"] = dict(envs=64 * 64, desc="link-parameter grid sweep (bw 1-1000 Mbit/s x delay 1-500 ms, 64 x 64 log grid), "
                                                "2 senders per link, 1xB200")
ACTION_SIGMA = 1.0   # a ~ N(0,1), BASELINE.md §3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 400 (config4: 8192)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (overrides the workload's)")
    ap.add_argument("--seed", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--ring-capacity", type=int, default=None, help="experiment: override the safe ring capacity")
    ap.add_argument("--only-device-pass", action="store_true", help="experiment: skip the b2b / e2e / cpu passes")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 8192 if a.workload == "config4" else 400
    return a


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(n_env_steps, sent, acked, hist_len=10, n_feat=3):
    """(221 + 8*H*F + 8*(H-1)*F) + 48*sent + 8*acked bytes per env-step."""
    fixed = 221 + 8 * hist_len * n_feat + 8 * (hist_len - 1) * n_feat
    return fixed * n_env_steps + 48 * sent + 8 * acked


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


REF_CHUNK = 100     # env-steps per process and "step" of the reference arm


def time_python_reference(steps, warmup, seed, n_procs=None):
    """The UNMODIFIED Python reference on this box's host cores (oracle/ref_timing.py): one SimulatedNetworkEnv per
    process, `steps` x REF_CHUNK timed env-steps each (bounded), all cores and one core.  None if no reference tree."""
    import ref_timing
    root = ref_timing.find_reference_root()
    if root is None:
        return None
    cores = n_procs or os.cpu_count() or 1
    timed = int(min(max(steps, 1) * REF_CHUNK, 30000))
    warm = int(min(max(warmup, 0) * REF_CHUNK, 2000))
    allc = ref_timing.time_reference(cores, warm, timed, seed=seed, root=root)
    one = ref_timing.time_reference(1, warm, min(timed, 6000), seed=seed, root=root)
    return dict(value=float(sum(allc["per_process"])), cores=cores, single_core=one["value"],
                bound_by_slowest_process=allc["value"], seconds=allc["seconds"], env_steps_per_process=timed,
                root=root)


def time_c_port(args, n, steps, warmup):
    """The C restatement of the reference (oracle/) on all host threads over the arm's env batch."""
    import oracle
    from pcc_rl_b200 import sample_link_params
    cores = os.cpu_count() or 1
    seeds = (np.uint64(args.seed) + np.arange(n, dtype=np.uint64))
    ob = oracle.OracleBatch(seeds, n_threads=cores)
    episode = 1
    ob.reset(sample_link_params(args.seed, episode, np.arange(n), n))
    g = np.random.default_rng(args.seed + 1)
    steps_in_ep = 0

    def one():
        nonlocal steps_in_ep, episode
        ob.step(g.normal(0, ACTION_SIGMA, n))
        steps_in_ep += 1
        if steps_in_ep >= 400:
            episode += 1
            ob.reset(sample_link_params(args.seed, episode, np.arange(n), n))
            steps_in_ep = 0
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return dict(value=n * steps / dt, cores=cores, seconds=dt, envs=n, steps=steps)


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same metric.
    If a reference tree is present (PCC_REFERENCE_ROOT, baseline/_ref staged by __graft_entry__.build(), /root/reference)
    the line's value is the UNMODIFIED Python SimulatedNetworkEnv, one process per core, a bounded sample per step
    (REF_CHUNK env-steps per process); the C restatement (oracle/) over the arm's env batch is reported beside it.
    Without a tree the C restatement is the value and the line says reference_absent.  Rank 0 only."""
    if rank != 0:
        return
    import oracle
    if args.workload == "flows":
        return run_reference_arm_flows(args, oracle)
    wl = args.workload if args.workload in ("config2", "config3", "config4") else "config3"
    n = (args.envs or WORKLOADS[wl]["envs"]) * max(1, args.gpus)   # the whole job's env batch
    ks = min(args.steps, 40)                                      # bounded sample of the C restatement
    port = time_c_port(args, n, ks, min(args.warmup, 5))
    ref = time_python_reference(args.steps, args.warmup, args.seed)
    port_desc = ("%d envs x %d steps on %d host threads, C restatement of the Python reference (oracle/)"
                 % (n, ks, port["cores"]))
    if ref is not None:
        v, secs = ref["value"], ref["seconds"]
        cpu = {"value": v, "unit": "env-steps/s", "cores": ref["cores"], "kind": "reference",
               "sample": "unmodified SimulatedNetworkEnv (%s), one process per host core, %d env-steps per process incl. "
                         "resets, default link-parameter ranges, actions N(0,1); sum of the per-process rates"
                         % (ref["root"], ref["env_steps_per_process"]),
               "single_core": ref["single_core"], "bound_by_slowest_process": ref["bound_by_slowest_process"],
               "port": {"value": port["value"], "cores": port["cores"], "sample": port_desc}}
    else:
        v, secs = port["value"], port["seconds"]
        cpu = {"value": v, "unit": "env-steps/s", "cores": port["cores"], "kind": "port", "reference_absent": True,
               "sample": port_desc + "; no reference tree on this box (the unmodified Python reference measured "
                                     "1.3-2.1k env-steps/s per core in the build container, BASELINE.md section 2)"}
    line = {"impl": "reference", "metric": "env-steps/sec (batched MI sim)", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl + ": " + WORKLOADS[wl]["desc"], "envs_per_gpu": n // max(1, args.gpus),
                       "global_envs": n, "actions": "N(0,1)", "auto_reset": True,
                       "reference_step": "%d env-steps of every host process" % REF_CHUNK if ref is not None else
                                         "one MI of every env of the batch"},
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_reference_arm_flows(args, oracle):
    """--impl reference --workload flows: the C restatement of sender_obs.py / loaded_client.give_sample on all host
    threads (flows partitioned over threads), a bounded sample of the same batches per step."""
    n = args.envs or WORKLOADS["flows"]["envs"]
    cores = os.cpu_count() or 1
    ns = min(n, 1 << 18)                          # records per step of the CPU arm (bounded sample)
    rng = np.random.default_rng(args.seed)
    fl = oracle.OracleFlows(ns, 10, oracle.DEFAULT_FEATURES)
    batches = [flows_batch(rng, ns) for _ in range(2)]
    for t in range(args.warmup):
        fl.give_batch(batches[t % 2], n_threads=cores)
    secs = 0.0
    for t in range(args.steps):
        secs += fl.give_batch(batches[t % 2], n_threads=cores)
    v = ns * args.steps / secs
    print(json.dumps({
        "impl": "reference", "metric": "MI records/sec (flow-monitor ingestion)", "value": v, "unit": "records/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "flows: " + WORKLOADS["flows"]["desc"], "records_per_step": ns},
        "cpu_baseline": {"value": v, "unit": "records/s", "cores": cores, "kind": "port",
                         "sample": "%d records per step (one per flow) on %d host threads; C restatement of "
                                   "sender_obs.py / loaded_client.give_sample (oracle/pcc_oracle_flows.c)" % (ns, cores)},
        "e2e": {"value": v, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_config5(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global):
    """BASELINE config 5: every rank sweeps the same 64 x 64 (bw, delay) grid with 2 senders per link (its own seeds).
    One step = one MI of every link; an env-step here carries two senders' MIs."""
    K, W, S = args.steps, args.warmup, 2
    side = int(round(n ** 0.5))
    p = pcc_rl_b200.grid_sweep_params(n_bw=side, n_lat=side, queue=40, loss=0.01)
    n = side * side
    g = np.random.default_rng(args.seed + rank)
    rates = g.uniform(40, 1000, (n, S))
    env = pcc_rl_b200.PccMultiSenderEnv(n, n_senders=S, seed=args.seed + 1000 * rank, ring_capacity=1 << 13, device=dev)
    env.reset(p, rates)
    acts = torch.randn((W + K, n, S), dtype=torch.float64, device=dev) * 2.0
    for t in range(W):
        env.step(acts[t])
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else dev.index)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    sampler.start()
    tot = torch.zeros(3, dtype=torch.int64, device=dev)
    for t in range(K):
        if flush is not None:
            flush.fill_(t & 0xFF)
        ev_s[t].record()
        obs, rew, done, info = env.step(acts[W + t])
        ev_e[t].record()
        tot += info["counts"].sum((0, 1))
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    env.check()
    dev_ms = D.max_over_ranks(sum(a.elapsed_time(b) for a, b in zip(ev_s, ev_e)), dev)
    sent, acked, _ = [D.sum_over_ranks(int(x), dev) for x in tot.cpu().tolist()]
    # end to end: host actions in, obs / reward / done back
    ke = min(K, 50)
    h_act = acts[W:W + ke].cpu().pin_memory()
    t0 = time.perf_counter()
    acc = 0.0
    for t in range(ke):
        o_, r_, d_, _i = env.step(h_act[t].to(dev, non_blocking=True))
        acc += float(r_.cpu()[0, 0]) + float(o_.cpu()[0, 0, 0])
        d_.cpu()
    torch.cuda.synchronize(dev)
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    launches = env.launch_count()
    if rank != 0:
        return
    peak, peak_kind = measured_peak_gbs()
    hf = env.obs_dim
    engine = os.environ.get("PCC_MULTI_MODE", "warp")
    kernel_name = {"thread": "pcc_mfast_step_kernel", "heap": "pcc_multi_step_kernel"}.get(engine, "pcc_mwarp_step_kernel<2>")
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh).get("config5")
        if tj and engine not in ("thread", "heap") and n == 4096:
            traffic, traffic_src, kernel_name = tj["dram_bytes_per_launch"], "profiles/" + tj["source"], tj["kernel"]
    except (OSError, ValueError, KeyError):
        pass
    # per sender the single-sender figure (677 + 48 sent + 8 acked) without a second copy of the link parameters
    bytes_total = (677 * S - 32 * (S - 1)) * n * world * K + 48 * sent + 8 * acked
    achieved = bytes_total / (dev_ms * 1e-3) / 1e9 / world
    line = {
        "metric": "env-steps/sec (batched MI sim)", "value": n * world * K / (dev_ms * 1e-3), "unit": "env-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config5: " + WORKLOADS["config5"]["desc"], "links_per_gpu": n, "senders_per_link": S,
                   "engine": {"thread": "streaming MI, one link per thread", "heap": "per-env event heap"}.get(
                       engine, "streaming MI (heap-free), one link per warp, links re-sorted by predicted packets every 8 steps"),
                   "actions": "N(0,2) per sender",
                   "l2": "no flush" if args.no_flush else "flushed between steps (256 MiB write outside the event brackets)"},
        "sender_steps_per_s": n * world * S * K / (dev_ms * 1e-3),
        "e2e": {"value": n * world * ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": 8 * n * S,
                "d2h_bytes_per_step": n * S * (8 * hf + 8) + n, "ms_per_step": 1e3 * e2e_s / ke, "steps": ke,
                "api": "PccMultiSenderEnv.step with pinned host actions, obs / reward / done copied back"},
        "gpu_launches": launches,
        "gpu_launches_note": "this library's kernels over reset + warm-up + timed + e2e steps (step kernels + the sort's key kernel; cub's radix sort not counted)",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                     "algorithmic_bytes_per_launch": bytes_total / (K * world)},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        import oracle
        ns, ks = 64, 40                       # a bounded sample: 64 grid points x (reset + 40 steps), one host thread
        idx = np.linspace(0, n - 1, ns).astype(int)
        t0 = time.perf_counter()
        for i in idx:
            o = oracle.OracleEnv()
            o.seed_philox(args.seed + int(i))
            o.reset_multi(p["bw"][i], p["lat"][i], int(p["queue"][i]), p["loss"][i], rates[i])
            for k in range(ks):
                o.step_multi(g.normal(0, 2.0, S))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns * ks / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                "sample": "%d grid points spread over the grid x (reset + %d steps), one host thread; C "
                                          "restatement of the reference's heap loop with 2 senders (oracle/)" % (ns, ks)}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_config4(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global):
    """BASELINE config 4: K-step rollout (default 8 192 = PPO1's timesteps_per_actorbatch, stable_solve.py:52) with the
    policy and the value head on the device; `value` = env-steps/s of the whole rollout incl. policy, auto-resets and the
    return gather."""
    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)
    roll = run_rollout(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global, args.steps, barrier, lambda k, w: 0)
    if rank == 0:
        print(json.dumps({
            "metric": "env-steps/sec (batched MI sim)", "value": roll["value"], "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": roll["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config4: " + WORKLOADS["config4"]["desc"], "envs_per_gpu": n, "global_envs": n_global,
                       "policy": roll["policy"]},
            "device_ms_per_step": roll["device_ms_per_step"], "gpu_launches": roll["gpu_launches"],
            "e2e": {"value": roll["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "closed loop on the device: nothing crosses PCIe per step"},
            "episode_returns": roll["episode_returns"], "note": roll["note"]}))
    if world > 1:
        torch.distributed.destroy_process_group()


def flows_batch(rng, n_flows, mean_samples=150):
    """One synthetic MI record per flow (flow order shuffled), SoA + CSR -- numpy, host."""
    n = rng.poisson(mean_samples, n_flows).astype(np.int64)
    n[rng.random(n_flows) < 0.01] = 0
    ps = rng.choice(np.array([1500, 1400, 1000], dtype=np.int64), n_flows)
    lost = rng.integers(0, 6, n_flows)
    dur = rng.uniform(0.01, 0.5, n_flows)
    base = rng.uniform(0.01, 0.4, n_flows)
    off = np.zeros(n_flows + 1, dtype=np.int64)
    off[1:] = np.cumsum(n)
    rtt = np.repeat(base, n) * (1.0 + 0.5 * rng.random(int(off[-1])))
    return dict(flow=rng.permutation(n_flows).astype(np.int32), bytes_sent=(n + lost + 1) * ps, bytes_acked=n * ps,
                bytes_lost=lost * ps, send_start=np.zeros(n_flows), send_end=dur, recv_start=base, recv_end=dur + base,
                packet_size=ps, rtt_off=off, rtt=rtt)


def flows_algorithmic_bytes(n_records, n_samples, H=10, F=3):
    """Per record: 76 B of fields (flow 4, 4 x i64, 4 x f64, CSR offset 8) + 8 B per RTT sample (read once) + per-flow
    state read+write (conn-min entry 16, head/flags 8, record counter 8, batch stamp 8 = 40) + the new history row
    (8F write) + the observation (8(H-1)F history read + 8HF write)."""
    return n_records * (76 + 40 + 8 * F + 8 * (H - 1) * F + 8 * H * F) + 8 * n_samples


def run_flows(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global):
    """--workload flows: steps of one record per flow through pcc_flows_give_samples (unique batch, one launch)."""
    K, W = args.steps, args.warmup
    NB = 3                                               # distinct pre-generated batches, cycled (each >> L2)
    rng = np.random.default_rng(args.seed + 17 * rank)
    mon = pcc_rl_b200.PccFlowMonitor(n, device=dev)
    host = [flows_batch(rng, n) for _ in range(NB)]
    batches = [mon.make_batch(**b) for b in host]
    n_samples = [int(b["rtt_off"][-1]) for b in host]
    obs = torch.empty((n, mon.obs_dim), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for t in range(W):
        mon.give_samples(batches[t % NB], unique_flows=True, obs_out=obs)
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else dev.index)
    barrier()
    launches0 = mon.launches
    sampler.start()
    samples_total = 0
    for t in range(K):
        ev_s[t].record()
        mon.give_samples(batches[(W + t) % NB], unique_flows=True, obs_out=obs)
        ev_e[t].record()
        samples_total += n_samples[(W + t) % NB]
    barrier()
    clocks = sampler.stop()
    launches = mon.launches - launches0
    mon.check()
    dev_ms = D.max_over_ranks(sum(s.elapsed_time(e) for s, e in zip(ev_s, ev_e)), dev)
    g_samples = D.sum_over_ranks(samples_total, dev)
    value = n_global * K / (dev_ms * 1e-3)

    # general (non-unique) batches: ingest + cub sort + per-flow apply
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kg = max(1, K // 4)
    mon.give_samples(batches[0], unique_flows=False, obs_out=obs)
    barrier()
    s.record()
    for t in range(kg):
        mon.give_samples(batches[t % NB], unique_flows=False, obs_out=obs)
    e.record()
    barrier()
    gen_ms = D.max_over_ranks(s.elapsed_time(e), dev)

    # end to end: pinned HOST batch -> device, ingest, observation back to the host
    ke = max(1, min(K, 8))
    pin = [{k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()} for b in host[:2]]
    h_obs = torch.empty((n, mon.obs_dim), dtype=torch.float64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in pin[0].values())

    def e2e_step(t):
        b = {k: v.to(dev, non_blocking=True) for k, v in pin[t % 2].items()}
        mon.give_samples(b, unique_flows=True, obs_out=obs)
        h_obs.copy_(obs, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return float(h_obs[0, -1])
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    acc = 0.0
    for t in range(ke):
        acc += e2e_step(t)
    barrier()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    mon.check()
    if rank != 0:
        return
    peak, peak_kind = measured_peak_gbs()
    bytes_total = flows_algorithmic_bytes(n_global * K, g_samples)
    achieved = bytes_total / (dev_ms * 1e-3) / 1e9 / world
    traffic, traffic_src, kernel_name = None, None, "pcc_flows_ingest_tma_kernel<true>"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)["flows"]
        if n == WORKLOADS["flows"]["envs"]:
            traffic, traffic_src, kernel_name = tj["dram_bytes_per_launch"], "profiles/" + tj["source"], tj["kernel"]
    except Exception:
        pass
    line = {
        "metric": "MI records/sec (flow-monitor ingestion)", "value": value, "unit": "records/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "flows: " + WORKLOADS["flows"]["desc"], "flows_per_gpu": n, "global_flows": n_global,
                   "history_len": 10, "features": 3, "samples_per_record": g_samples / (n_global * K),
                   "l2": "inputs larger than L2 (%.2f GB per batch, %d batches cycled)" % (h2d / 1e9, NB),
                   "parallelism": "flow sharding x%d, no collective" % world},
        "general_batches": {"value": n_global * kg / (gen_ms * 1e-3), "ms_per_step": gen_ms / kg,
                            "note": "unique_flows=0: ingest + cub radix sort by flow + per-flow apply in batch order"},
        "e2e": {"value": n_global * ke / e2e_s, "unit": "records/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(h_obs.numel() * 8), "ms_per_step": 1e3 * e2e_s / ke, "steps": ke,
                "api": "PccFlowMonitor.give_samples on pinned host tensors (H2D batch, D2H observations, synchronous)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                     "algorithmic_bytes_per_launch": bytes_total / (K * world)},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        fl = oracle.OracleFlows(n, 10, oracle.DEFAULT_FEATURES)
        secs = fl.give_batch(host[0], n_threads=cores) + fl.give_batch(host[1], n_threads=cores)
        line["cpu_baseline"] = {"value": 2 * n / secs, "unit": "records/s", "cores": cores, "kind": "port",
                                "sample": "2 of the same batches (%d records each) on %d host threads, flows partitioned "
                                          "over threads; C restatement of sender_obs.py / loaded_client.give_sample "
                                          "(oracle/pcc_oracle_flows.c)" % (n, cores)}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    import torch
    import pcc_rl_b200
    from pcc_rl_b200 import distributed as D
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    rank, local_rank, world = D.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.envs or WORKLOADS[args.workload]["envs"]
    n_global = n * world
    K, W = args.steps, args.warmup

    if args.workload == "config4":
        run_config4(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global)
        return
    if args.workload == "flows":
        run_flows(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global)
        return
    if args.workload == "config5":
        run_config5(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global)
        return

    MAX_STEPS = 400

    def make_env(ne=n):
        return pcc_rl_b200.PccBatchEnv(n_envs=ne, device=dev, seed=args.seed, global_offset=rank * ne,
                                       n_global=ne * world, auto_reset=True, ring_capacity=args.ring_capacity)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def preroll(k, w):
        """Untimed steps before the warm-up.  K >= 400: none -- the window covers a whole episode and its boundary at
        the natural frequency.  K < 400: the window is centred on the middle of the episode (early steps are lighter
        than late ones: the rates random-walk apart); the boundary step is then timed separately (episode_boundary)."""
        return 0 if k >= MAX_STEPS else max(0, MAX_STEPS // 2 - w - k // 2)

    def make_actions(ne, rows, seed):
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        return torch.randn((rows, ne), generator=gen, device=dev, dtype=torch.float64) * ACTION_SIGMA

    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def device_pass(ne, k, w, sample_clocks):
        """Device-resident inputs, per-step CUDA-event brackets, L2 flushed between steps (outside the brackets)."""
        env = make_env(ne)
        env.reset()
        pre = preroll(k, w)
        acts = make_actions(ne, pre + w + k, args.seed + 1 + rank)
        for t in range(pre + w):
            env.step(acts[t])
        ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
        sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
        barrier()
        launches0 = env.launches
        if sample_clocks:
            sampler.start()
        wall0 = time.perf_counter()
        tot = torch.zeros(3, dtype=torch.int64, device=dev)
        finished = []
        for t in range(k):
            if flush is not None:
                flush.fill_(t & 0xFF)
            ev_s[t].record()
            obs, rew, done, info = env.step(acts[pre + w + t])
            ev_e[t].record()
            tot += info["counts"].sum(0)
            if int(env._steps.max()) == 0:     # host bookkeeping: a synchronized auto-reset just happened
                finished.append(env.column("last_episode_return"))
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if sample_clocks else None
        launches = env.launches - launches0
        env.check()
        boundary_ms = None
        if not finished and sample_clocks:
            # the episode boundary is outside the window: walk to it untimed and time that one step (reset kernel +
            # parameter upload for every env + the step itself)
            ag = make_actions(ne, 1, args.seed + 99)[0]
            while int(env._steps.max()) < MAX_STEPS - 1:
                env.step(ag)
            torch.cuda.synchronize(dev)
            bs, be = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tb0 = time.perf_counter()
            bs.record()
            env.step(ag)
            be.record()
            torch.cuda.synchronize(dev)
            if os.environ.get("PCC_BENCH_DEBUG"):
                sys.stderr.write("boundary: wall %.2f ms, events %.2f ms\n" % (1e3 * (time.perf_counter() - tb0), bs.elapsed_time(be)))
            boundary_ms = D.max_over_ranks(bs.elapsed_time(be), dev)
            finished.append(env.column("last_episode_return"))
            env.check()
        per_step = [a.elapsed_time(b) for a, b in zip(ev_s, ev_e)]
        if os.environ.get("PCC_BENCH_DEBUG"):
            top = sorted(range(len(per_step)), key=lambda i: -per_step[i])[:4]
            sys.stderr.write("slowest steps (index in the window: ms): %s; preroll + warm-up = %d steps\n"
                             % (", ".join("%d: %.2f" % (i, per_step[i]) for i in top), pre + w))
        dev_ms = D.max_over_ranks(sum(per_step), dev)
        sent, acked, _lost = [int(x) for x in tot.cpu().tolist()]
        rets = torch.cat(finished) if finished else torch.zeros(0, dtype=torch.float64, device=dev)
        out = dict(dev_ms=dev_ms, wall=wall, clocks=clocks, launches=int(launches), g_sent=D.sum_over_ranks(sent, dev),
                   g_acked=D.sum_over_ranks(acked, dev), rets=rets, reset_step_ms=max(per_step),
                   median_step_ms=float(np.median(per_step)), boundary_inside=boundary_ms is None and bool(finished),
                   boundary_ms=boundary_ms)
        env.close()
        del env, acts
        torch.cuda.empty_cache()
        return out

    # ---------------- pass 1: the headline device pass ----------------
    p1 = device_pass(n, K, W, True)
    ret_stats = D.gather_episode_returns(p1["rets"])   # the path's only collective (NCCL all-gather, a few bytes)
    dev_ms, g_sent, g_acked, clocks = p1["dev_ms"], p1["g_sent"], p1["g_acked"], p1["clocks"]
    value = n_global * K / (dev_ms * 1e-3)
    if args.only_device_pass:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": dev_ms / K, "sent_per_step": g_sent / (n_global * K),
                              "clocks": clocks}))
        return

    # ---------------- pass 2: back to back, no flush, one bracket around all K steps ----------------
    env = make_env()
    env.reset()
    acts = make_actions(n, W + K, args.seed + 1 + rank)
    for t in range(W):
        env.step(acts[t])
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(K):
        env.step(acts[W + t])
    e.record()
    barrier()
    b2b_ms = D.max_over_ranks(s.elapsed_time(e), dev)
    env.check()
    env.close()
    del env
    torch.cuda.empty_cache()

    # ---------------- pass 3: end to end through host buffers ----------------
    # pinned HOST actions in, obs / reward / done out, every step; (a) two steps in flight through
    # pcc_step_host_submit / _wait (the download of step k overlaps step k + 1), (b) the synchronous pcc_step_host
    env = make_env()
    env.reset()
    hf = env.obs_dim
    pre3 = preroll(K, W)
    for t in range(pre3):
        env.step(acts[t % (W + K)])
    h_act = torch.empty((W + K, n), dtype=torch.float64).pin_memory()
    h_act.copy_(acts.cpu())
    bufs = [dict(o=torch.empty((n, hf), dtype=torch.float64).pin_memory(), r=torch.empty(n, dtype=torch.float64).pin_memory(),
                 d=torch.empty(n, dtype=torch.uint8).pin_memory()) for _ in range(2)]
    a_np = h_act.numpy()
    nb = [{k: v.numpy() for k, v in b.items()} for b in bufs]

    e2e_trace = [] if os.environ.get("PCC_BENCH_DEBUG") else None

    def e2e_run(first, count, pipelined):
        acc, tk = 0.0, [None, None]
        for t in range(first, first + count):
            sl = t & 1
            if pipelined:
                if tk[sl] is not None:
                    env.step_host_wait(tk[sl])
                    acc += float(nb[sl]["r"][0])           # the result is really read on the host
                if e2e_trace is not None:
                    e2e_trace.append(time.perf_counter())
                tk[sl] = env.step_host_submit(a_np[t], nb[sl]["o"], nb[sl]["r"], nb[sl]["d"])
                if e2e_trace is not None:
                    e2e_trace.append(time.perf_counter())
            else:
                env.step_host(a_np[t], nb[sl]["o"], nb[sl]["r"], nb[sl]["d"])
                acc += float(nb[sl]["r"][0])
            if env._steps[0] >= env.max_steps:             # the caller resets finished envs, as PPO does
                for q in (0, 1):
                    if tk[q] is not None:
                        env.step_host_wait(tk[q]); acc += float(nb[q]["r"][0]); tk[q] = None
                env.reset()
        for q in (0, 1):
            if tk[q] is not None:
                env.step_host_wait(tk[q]); acc += float(nb[q]["r"][0])
        return acc

    e2e_run(0, W, True)
    barrier()
    t0 = time.perf_counter()
    e2e_run(W, K, True)
    barrier()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    if e2e_trace is not None:
        tr = e2e_trace[-2 * K:]
        sys.stderr.write("e2e pipelined, per step [wait-done -> submit-returned, submit-returned -> next wait-done] ms: %s\n"
                         % " ".join("%.2f/%.2f" % (1e3 * (tr[2 * i + 1] - tr[2 * i]), 1e3 * ((tr[2 * i + 2] if 2 * i + 2 < len(tr) else tr[-1]) - tr[2 * i + 1]))
                                    for i in range(min(K, 24))))
        e2e_trace = None
    ks = min(K, 50)
    barrier()
    t0 = time.perf_counter()
    e2e_run(W, ks, False)
    barrier()
    e2e_sync_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = n_global * K / e2e_s
    env.check()
    env.close()
    del env, acts
    torch.cuda.empty_cache()

    # ---------------- sub-key: the closed-loop rollout of config 4 (policy + value head on the device) ----------------
    roll = run_rollout(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global, max(64, min(K, 512)), barrier, preroll)

    # ---------------- sub-key: BASELINE config 2 (4 096 envs per GPU) ----------------
    c2 = None
    if args.workload == "config3" and n > 4096:
        k2 = min(K, 100)
        q = device_pass(4096, k2, 10, False)
        b2 = algorithmic_bytes(4096 * world * k2, q["g_sent"], q["g_acked"])
        c2 = {"workload": "config2: " + WORKLOADS["config2"]["desc"], "value": 4096 * world * k2 / (q["dev_ms"] * 1e-3),
              "unit": "env-steps/s", "ms_per_step": q["dev_ms"] / k2, "steps": k2,
              "roofline_frac": b2 / (q["dev_ms"] * 1e-3) / 1e9 / world / measured_peak_gbs()[0],
              "boundary_inside": q["boundary_inside"]}
        r2 = run_rollout(args, pcc_rl_b200, D, torch, dev, rank, world, 4096, 4096 * world, max(64, min(K, 256)), barrier, preroll)
        c2["rollout"] = {k: r2[k] for k in ("value", "ms_per_step", "device_ms_per_step", "steps", "gpu_launches")}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peak, peak_kind = measured_peak_gbs()
    traffic, traffic_src, kernel_name = None, None, "pcc_step_packed_kernel" if n > 16384 else "pcc_step_warp_kernel"
    try:   # dram bytes per launch of the dominant kernel, from the committed ncu capture of this workload
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)[args.workload]
        if n == WORKLOADS[args.workload]["envs"]:
            traffic, traffic_src, kernel_name = tj["dram_bytes_per_launch"], "profiles/" + tj["source"], tj["kernel"]
    except Exception:
        pass
    bytes_total = algorithmic_bytes(n_global * K, g_sent, g_acked)
    achieved = bytes_total / (dev_ms * 1e-3) / 1e9 / world   # per GPU
    line = {
        "metric": "env-steps/sec (batched MI sim)", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + WORKLOADS[args.workload]["desc"], "envs_per_gpu": n,
                   "global_envs": n_global, "history_len": 10, "features": 3, "rng": "philox4x32-10",
                   "actions": "N(0,1) pre-generated on device", "auto_reset": True,
                   "episode_boundary": ("inside the timed window (that step took %.3f ms, the median step %.3f ms)"
                                        % (p1["reset_step_ms"], p1["median_step_ms"])) if p1["boundary_inside"] else
                                       ("outside the %d-step window (steps %d.. of a 400-step episode, median step %.3f ms); the "
                                        "boundary step, walked to and timed separately incl. the host-side reset preparation, "
                                        "took %.1f ms; a run with --steps 400 covers a whole episode with the boundary inside"
                                        % (K, preroll(K, W) + W, p1["median_step_ms"], p1["boundary_ms"] or 0.0)),
                   "l2": "no flush" if args.no_flush else "flushed between steps (256 MiB write outside the event brackets)",
                   "parallelism": "env-batch sharding x%d, no data-path collective" % world},
        "back_to_back": {"value": n_global * K / (b2b_ms * 1e-3), "ms_per_step": b2b_ms / K,
                         "note": "K steps enqueued back to back, one event bracket, warm L2"},
        "rollout": roll,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": 8 * n,
                "d2h_bytes_per_step": n * (8 * hf + 8 + 1), "ms_per_step": 1e3 * e2e_s / K,
                "api": "PccBatchEnv.step_host_submit / step_host_wait -> pcc_step_host_submit / _wait: pinned host buffers, "
                       "two steps in flight (the download of step k overlaps the upload and kernel of step k + 1)",
                "sync": {"value": n_global * ks / e2e_sync_s, "ms_per_step": 1e3 * e2e_sync_s / ks, "steps": ks,
                         "api": "PccBatchEnv.step_host -> pcc_step_host (one step in flight, synchronous)"}},
        "gpu_launches": p1["launches"],
        "wall_s_timed_region": p1["wall"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                     "algorithmic_bytes_per_launch": bytes_total / (K * world),
                     "algorithmic_bytes_per_env_step": bytes_total / (n_global * K),
                     "packets_sent_per_env_step": g_sent / (n_global * K)},
        "clocks": clocks,
        "episode_returns": {"count": ret_stats["count"], "mean": ret_stats["mean"]},
    }
    if c2 is not None:
        line["config2"] = c2
    if not args.no_cpu_baseline:
        ref = time_python_reference(min(K, 60), min(W, 5), args.seed)
        ns, ks2 = min(n, 4096), 100
        import oracle
        from pcc_rl_b200 import sample_link_params
        cores = os.cpu_count() or 1
        pp = sample_link_params(args.seed, 1, np.arange(ns), n_global)
        pacts = np.random.default_rng(args.seed + 1).normal(0, ACTION_SIGMA, (ks2, ns))
        r = oracle.batch_run(pp["bw"], pp["lat"], pp["queue"], pp["loss"], pp["start_rate"],
                             np.uint64(args.seed) + np.arange(ns, dtype=np.uint64), ks2, actions=pacts, n_threads=cores)
        port = {"value": ns * ks2 / r["seconds"], "cores": cores,
                "sample": "%d envs x (reset + %d steps) of the same workload on %d host threads; C restatement of the "
                          "reference (oracle/)" % (ns, ks2, cores)}
        if ref is not None:
            line["cpu_baseline"] = {"value": ref["value"], "unit": "env-steps/s", "cores": ref["cores"], "kind": "reference",
                                    "sample": "unmodified Python SimulatedNetworkEnv (%s), one process per host core, %d "
                                              "env-steps per process incl. resets, default ranges, actions N(0,1); sum of "
                                              "the per-process rates" % (ref["root"], ref["env_steps_per_process"]),
                                    "single_core": ref["single_core"], "port": port}
        else:
            line["cpu_baseline"] = dict(port, unit="env-steps/s", kind="port", reference_absent=True)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_rollout(args, pcc_rl_b200, D, torch, dev, rank, world, n, n_global, K, barrier, preroll):
    """The closed loop of BASELINE config 4 on every rank: K monitor intervals with the MLP policy 30-32-16-1 and the
    value network of the same shape on the device (PPO1's ob / ac / vpred / rew / new segment, stable_solve.py:30-58),
    in-sequence auto-reset, nothing crossing PCIe per step, and the NCCL gather of the finished episodes' returns."""
    RK = 256
    env = pcc_rl_b200.PccBatchEnv(n_envs=n, device=dev, seed=args.seed, global_offset=rank * n, n_global=n_global)
    env.reset()
    g = torch.Generator(device=dev)
    g.manual_seed(args.seed + 7)    # same random-init networks on every rank
    r = lambda *sh, sc: torch.randn(*sh, generator=g, device=dev, dtype=torch.float64) * sc
    pol = dict(w1=r(32, 30, sc=0.2), b1=r(32, sc=0.05), w2=r(16, 32, sc=0.2), b2=r(16, sc=0.05), w3=r(1, 16, sc=0.5),
               b3=r(1, sc=0.05), vw1=r(32, 30, sc=0.2), vb1=r(32, sc=0.05), vw2=r(16, 32, sc=0.2), vb2=r(16, sc=0.05),
               vw3=r(1, 16, sc=0.5), vb3=r(1, sc=0.05), stochastic=True, log_std=-0.7, noise_seed=args.seed + rank)
    pre = preroll(K, 16) + 16                       # includes the warm-up launches of every kernel of the sequence
    done_pre = 0
    while done_pre < pre:
        k = min(RK, pre - done_pre)
        env.rollout(k, policy=pol, want_obs=False, want_counts=False)
        done_pre += k
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = env.launches
    t0 = time.perf_counter()
    s.record()
    finished, steps_done = [], 0
    while steps_done < K:
        k = min(RK, K - steps_done)
        crosses = bool(((env._steps + k) // env.max_steps).max() > 0)       # host bookkeeping, no device sync
        env.rollout(k, policy=pol, want_obs=False, want_counts=False)
        steps_done += k
        if crosses:
            finished.append(env.column("last_episode_return"))
    e.record()
    rets = torch.cat(finished) if finished else torch.zeros(0, dtype=torch.float64, device=dev)
    stats = D.gather_episode_returns(rets)          # NCCL all-gather of [count, sum, sum^2]
    barrier()
    wall = D.max_over_ranks(time.perf_counter() - t0, dev)
    dev_ms = D.max_over_ranks(s.elapsed_time(e), dev)
    launches = int(env.launches - launches0)
    env.check()
    env.close()
    del env
    torch.cuda.empty_cache()
    return {"value": n_global * K / wall, "unit": "env-steps/s", "ms_per_step": 1e3 * wall / K, "device_ms_per_step": dev_ms / K,
            "steps": K, "gpu_launches": launches, "policy": "random-init MLP 30-32-16-1 (tanh), stochastic, + value head",
            "episode_returns": {"count": stats["count"], "mean": stats["mean"]},
            "note": "pcc_rollout: per step policy / value kernel + step kernel + bank gather + masked reset, enqueued on the "
                    "device; wall clock incl. the NCCL gather of episode returns; nothing crosses PCIe per step"}


if __name__ == "__main__":
    main()
