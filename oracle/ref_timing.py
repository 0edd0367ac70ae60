"""Test / measurement infrastructure: time the UNMODIFIED Python reference (SimulatedNetworkEnv.step,
gym/network_sim.py:406-444) on this box's host cores, one independent process per core.

Only bench.py's `cpu_baseline` leg and `--impl reference` arm use it.  The reference tree is looked for in
PCC_REFERENCE_ROOT, then <repo>/baseline/_ref (staged, unmodified and git-ignored, by __graft_entry__.build() when the
build container has /root/reference), then /root/reference; nothing is timed if none exists."""
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def find_reference_root():
    for r in (os.environ.get("PCC_REFERENCE_ROOT"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if r and os.path.isfile(os.path.join(r, "src", "gym", "network_sim.py")):
            return r
    return None


def _worker(args):
    root, rank, warm_steps, timed_steps, seed = args
    os.environ["PCC_REFERENCE_ROOT"] = root
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import random
    import refharness
    refharness.REFERENCE_ROOT = root
    ns = refharness.load_reference()
    with refharness.quiet_tmp_cwd():
        random.seed(seed + rank)                      # the reference's only RNG
        act = random.Random(seed + 1000 + rank)       # actions ~ N(0, 1), BASELINE.md section 3
        env = ns.SimulatedNetworkEnv()
        env.reset()

        def run(k):
            for _ in range(k):
                _o, _r, done, _i = env.step([act.gauss(0.0, 1.0)])
                if done:
                    env.reset()
        run(warm_steps)
        t0 = time.perf_counter()
        run(timed_steps)
        dt = time.perf_counter() - t0
    return timed_steps, dt


def time_reference(n_procs, warm_steps, timed_steps, seed=100, root=None):
    """Every process: one SimulatedNetworkEnv (default link-parameter ranges, history 10, 3 features, 400-step
    episodes, resets included), warm_steps untimed then timed_steps timed env-steps.  Returns dict(value = aggregate
    env-steps/s = total steps / slowest process, per_process, cores, root) or None if there is no reference tree."""
    root = root or find_reference_root()
    if root is None:
        return None
    ctx = mp.get_context("spawn")
    jobs = [(root, r, int(warm_steps), int(timed_steps), int(seed)) for r in range(int(n_procs))]
    if n_procs == 1:
        res = [_worker(jobs[0])]
    else:
        with ctx.Pool(int(n_procs)) as pool:
            res = pool.map(_worker, jobs)
    total = sum(s for s, _ in res)
    slowest = max(dt for _, dt in res)
    return dict(value=total / slowest, seconds=slowest, env_steps=total, cores=int(n_procs), root=root,
                per_process=[s / dt for s, dt in res])


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    print(time_reference(1, 200, 2000))
    print(time_reference(n, 200, 2000))
