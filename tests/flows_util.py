"""Helpers for the MI-sample ingestion ("flows") tests: load tests/golden/flows_*.npz (outputs of the
unmodified reference modules, oracle/gen_golden_flows.py) and synthesise record batches."""
import os

import numpy as np

from golden_util import GOLDEN_DIR

# rate-control constants of the two callers (loaded_client.py:33-35; shim_env.py:38-44)
CLIENT_RATE = dict(delta_scale=0.05, min_rate=0.5, max_rate=300.0, style=0, start=6.0)
SHIM_RATE = dict(delta_scale=0.025, min_rate=0.25, max_rate=1000.0, style=1, start=2.0)


def load_flows_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["features"] = str(g["features"])
    for k in ("n_flows", "history_len", "reset_mode"):
        g[k] = int(g[k])
    return g


def record_of(g, k):
    lo, hi = int(g["rtt_off"][k]), int(g["rtt_off"][k + 1])
    return dict(flow=int(g["flow"][k]), bytes_sent=int(g["bytes_sent"][k]), bytes_acked=int(g["bytes_acked"][k]),
                bytes_lost=int(g["bytes_lost"][k]), send_start=float(g["send_start"][k]),
                send_end=float(g["send_end"][k]), recv_start=float(g["recv_start"][k]),
                recv_end=float(g["recv_end"][k]), packet_size=int(g["packet_size"][k]), rtt=g["rtt"][lo:hi])


def synth_batch(rng, n_records, n_flows, mean_samples=150, unique=True, t0=0.0):
    """A batch of synthetic MI records in the SoA + CSR layout of pcc_flows_give_samples."""
    if unique:
        assert n_records <= n_flows
        flow = rng.permutation(n_flows)[:n_records].astype(np.int32)
    else:
        flow = rng.integers(0, n_flows, n_records).astype(np.int32)
    n = rng.poisson(mean_samples, n_records).astype(np.int64)
    n[rng.random(n_records) < 0.02] = 0
    ps = rng.choice(np.array([1500, 1400, 1000], dtype=np.int64), n_records)
    lost = rng.integers(0, 6, n_records)
    dur = rng.uniform(0.01, 0.5, n_records)
    base = rng.uniform(0.01, 0.4, n_records)
    off = np.zeros(n_records + 1, dtype=np.int64)
    off[1:] = np.cumsum(n)
    rtt = np.repeat(base, n) * (1.0 + 0.5 * rng.random(int(off[-1])))
    return dict(flow=flow, bytes_sent=(n + lost + 1) * ps, bytes_acked=n * ps, bytes_lost=lost * ps,
                send_start=np.full(n_records, t0), send_end=t0 + dur, recv_start=t0 + base,
                recv_end=t0 + dur + base, packet_size=ps, rtt_off=off, rtt=rtt)
