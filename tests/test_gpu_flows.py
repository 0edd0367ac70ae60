"""GPU parity of the MI-sample ingestion path (pcc_flows_* through the C ABI) against the reference's own
outputs (tests/golden/flows_*.npz) and against the CPU oracle on large synthetic batches.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from flows_util import CLIENT_RATE, SHIM_RATE, load_flows_golden, record_of, synth_batch


def _mon(g, **kw):
    import pcc_rl_b200
    return pcc_rl_b200.PccFlowMonitor(g["n_flows"], g["history_len"], g["features"], **kw)


def _batch_of(mon, g, ks):
    lens = [int(g["rtt_off"][k + 1] - g["rtt_off"][k]) for k in ks]
    off = np.zeros(len(ks) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    rtt = np.concatenate([g["rtt"][g["rtt_off"][k]:g["rtt_off"][k + 1]] for k in ks]) if ks else np.zeros(0)
    return mon.make_batch(g["flow"][ks], g["bytes_sent"][ks], g["bytes_acked"][ks], g["bytes_lost"][ks],
                          g["send_start"][ks], g["send_end"][ks], g["recv_start"][ks], g["recv_end"][ks],
                          g["packet_size"][ks], off, rtt)


@pytest.mark.parametrize("name", ["flows_client_default", "flows_allfeatures", "flows_shim"])
@pytest.mark.parametrize("unique", [True, False])
def test_flows_golden_record_by_record(name, unique):
    """Every op of the reference's stream as a batch of one; obs (and the 12 metrics) after every op."""
    g = load_flows_golden(name)
    mon = _mon(g)
    for k in range(len(g["op"])):
        i = int(g["flow"][k])
        if g["op"][k] == 1:
            m = np.zeros(g["n_flows"], dtype=np.uint8)
            m[i] = 1
            mon.reset(mask=m, mode=g["reset_mode"])
            obs = mon.obs()[i].cpu().numpy()
        else:
            o, met = mon.give_samples(_batch_of(mon, g, [k]), unique_flows=unique, want_metrics=True)
            obs = o[0].cpu().numpy()
            if not np.isnan(g["metrics"][k]).any():
                assert np.array_equal(met[0].cpu().numpy(), g["metrics"][k]), "metrics of op %d" % k
            assert np.array_equal(mon.obs()[i].cpu().numpy(), obs)
        assert np.array_equal(obs, g["obs"][k]), "obs after op %d" % k
    mon.check()


@pytest.mark.parametrize("name", ["flows_client_default", "flows_allfeatures"])
def test_flows_golden_in_general_batches(name):
    """The same stream cut into batches at the resets: many records per flow per batch, applied in batch order."""
    g = load_flows_golden(name)
    mon = _mon(g)
    ks = []

    def flush():
        if not ks:
            return
        o, _ = mon.give_samples(_batch_of(mon, g, ks), unique_flows=False)
        assert np.array_equal(o.cpu().numpy(), g["obs"][ks])
        del ks[:]

    for k in range(len(g["op"])):
        if g["op"][k] == 1:
            flush()
            m = np.zeros(g["n_flows"], dtype=np.uint8)
            m[int(g["flow"][k])] = 1
            mon.reset(mask=m, mode=g["reset_mode"])
        else:
            ks.append(k)
            if len(ks) == 97:
                flush()
    flush()
    mon.check()


@pytest.mark.parametrize("name,cfg", [("flows_client_default", CLIENT_RATE), ("flows_shim", SHIM_RATE)])
def test_flows_rates_golden(name, cfg):
    """Rate control of both callers (loaded_client.apply_rate_delta, ShimNetworkEnv.apply_action) on the device."""
    g = load_flows_golden(name)
    mon = _mon(g, delta_scale=cfg["delta_scale"], min_rate=cfg["min_rate"], max_rate=cfg["max_rate"],
               rate_style=cfg["style"], start_rate=cfg["start"])
    n = g["n_flows"]
    for k in range(len(g["op"])):
        i = int(g["flow"][k])
        sel = np.zeros(n, dtype=np.uint8)
        sel[i] = 1
        acts = np.zeros(n)
        acts[i] = g["action"][k] if g["op"][k] == 0 else 0.0
        if g["op"][k] == 1:
            mon.reset(mask=sel, mode=g["reset_mode"])
            if cfg["style"] == 1:
                mon.set_rates(rate=cfg["start"], mask=sel)          # ShimNetworkEnv.reset: set_rate(STARTING_RATE)
            rate = mon.get_rates()[i].item()
            assert rate == g["rate"][k]
            continue
        if cfg["style"] == 1:                                       # shim: action first, then the record
            rate = mon.get_rates(actions=acts, mask=sel)[i].item()
            mon.give_samples(_batch_of(mon, g, [k]), unique_flows=True, want_obs=False)
            assert rate == g["rate"][k], "op %d" % k
        else:                                                       # client: record, then get_rate
            mon.give_samples(_batch_of(mon, g, [k]), unique_flows=True, want_obs=False)
            rate = mon.get_rates(actions=acts, mask=sel)[i].item()
            assert rate * 1e6 == g["rate"][k], "op %d" % k
    mon.check()


def test_flows_module_api_replays_loaded_client():
    """The reference module's own API (init / give_sample / get_rate / reset) on the GPU path, replaying the
    stream the unmodified loaded_client.py produced."""
    from pcc_rl_b200 import flow_monitor as fm
    g = load_flows_golden("flows_client_default")

    class Agent(object):
        action = 0.0

        def act(self, ob):
            Agent.seen = np.asarray(ob).copy()
            return Agent.action

        def reset(self):
            pass

    fm.configure(agent_factory=Agent, history_len=g["history_len"], features=g["features"], max_flows=64)
    ids = [1000 + 7 * i for i in range(g["n_flows"])]
    for fid in ids:
        fm.init(fid)
    for k in range(len(g["op"])):
        fid = ids[int(g["flow"][k])]
        if g["op"][k] == 1:
            fm.reset(fid)
            assert fm.PccGymDriver.get_by_flow_id(fid).rate == g["rate"][k]
            continue
        r = record_of(g, k)
        fm.give_sample(fid, r["bytes_sent"], r["bytes_acked"], r["bytes_lost"], r["send_start"], r["send_end"],
                       r["recv_start"], r["recv_end"], list(r["rtt"]), r["packet_size"], 0.0)
        Agent.action = float(g["action"][k])
        assert fm.get_rate(fid) == g["rate"][k], "op %d" % k
        assert np.array_equal(Agent.seen, g["obs"][k]), "obs given to the agent at op %d" % k
    fm.configure()


def _oracle_replay(b, n_flows, H, features, want):
    import oracle
    fl = oracle.OracleFlows(n_flows, H, features)
    out = np.zeros((len(want), H * len(features.split(","))))
    pos = {int(r): q for q, r in enumerate(want)}
    for r in range(len(b["flow"])):
        lo, hi = int(b["rtt_off"][r]), int(b["rtt_off"][r + 1])
        i = int(b["flow"][r])
        fl.give_sample(i, b["bytes_sent"][r], b["bytes_acked"][r], b["bytes_lost"][r], b["send_start"][r],
                       b["send_end"][r], b["recv_start"][r], b["recv_end"][r], b["rtt"][lo:hi], b["packet_size"][r])
        if r in pos:
            out[pos[r]] = fl.obs(i)
    return out, fl


@pytest.mark.parametrize("unique", [True, False])
def test_flows_large_batches_vs_oracle(unique):
    """Three consecutive batches over 50 000 flows (unique: one record per flow; general: repeats), every
    observation of the last batch and every flow's final history against the CPU oracle."""
    import pcc_rl_b200
    feats = "sent latency inflation,latency ratio,send ratio,conn min latency,recv rate"
    n_flows, H = 50000, 10
    mon = pcc_rl_b200.PccFlowMonitor(n_flows, H, feats)
    rng = np.random.default_rng(5 + unique)
    import oracle
    fl = oracle.OracleFlows(n_flows, H, feats)
    for it in range(3):
        b = synth_batch(rng, 40000, n_flows, mean_samples=60 if it < 2 else 150, unique=unique, t0=float(it))
        if it == 1:   # a few very long sample lists: numpy's recursion at depth > 1, the whole-warp path
            n = np.diff(b["rtt_off"])
            n[:8] = [129, 1000, 5000, 300, 257, 20000, 131, 4096]
            b["rtt_off"] = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
            b["rtt"] = 0.05 * (1.0 + rng.random(int(b["rtt_off"][-1])))
            b["bytes_acked"] = n * b["packet_size"]
        o, _ = mon.give_samples(mon.make_batch(**{k: b[k] for k in b}), unique_flows=unique)
        o = o.cpu().numpy()
        for r in range(len(b["flow"])):
            lo, hi = int(b["rtt_off"][r]), int(b["rtt_off"][r + 1])
            i = int(b["flow"][r])
            fl.give_sample(i, b["bytes_sent"][r], b["bytes_acked"][r], b["bytes_lost"][r], b["send_start"][r],
                           b["send_end"][r], b["recv_start"][r], b["recv_end"][r], b["rtt"][lo:hi],
                           b["packet_size"][r])
            if r % 7 == 0 or r < 16:
                assert np.array_equal(o[r], fl.obs(i)), "batch %d record %d" % (it, r)
    final = mon.obs().cpu().numpy()
    for i in range(0, n_flows, 11):
        assert np.array_equal(final[i], fl.obs(i)), "flow %d" % i
    cm = mon.column("conn_min").cpu().numpy()
    assert all(cm[i] == fl.conn_min(i) for i in range(0, n_flows, 101))
    mon.check()


def test_flows_unique_equals_general_and_order_independent():
    """Size-independent properties at a large size: the fused (unique) kernel and the sorted general path agree,
    and the record order inside a unique batch does not matter."""
    import pcc_rl_b200
    import torch
    n_flows = 1 << 20                   # the bench workload's size
    rng = np.random.default_rng(77)
    b = synth_batch(rng, n_flows, n_flows, mean_samples=100, unique=True)
    mons = [pcc_rl_b200.PccFlowMonitor(n_flows) for _ in range(3)]
    mons[0].give_samples(mons[0].make_batch(**b), unique_flows=True, want_obs=False)
    mons[1].give_samples(mons[1].make_batch(**b), unique_flows=False, want_obs=False)
    perm = rng.permutation(n_flows)
    n = np.diff(b["rtt_off"])[perm]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    idx = np.concatenate([np.arange(b["rtt_off"][p], b["rtt_off"][p + 1]) for p in perm[:2000]])  # spot-check gather
    rtt = np.concatenate([b["rtt"][b["rtt_off"][p]:b["rtt_off"][p + 1]] for p in perm])
    assert np.array_equal(rtt[:len(idx)], b["rtt"][idx])
    bp = {k: (b[k][perm] if k not in ("rtt", "rtt_off") else None) for k in b}
    bp["rtt"], bp["rtt_off"] = rtt, off
    mons[2].give_samples(mons[2].make_batch(**bp), unique_flows=True, want_obs=False)
    o = [m.obs() for m in mons]
    assert torch.equal(o[0], o[1]) and torch.equal(o[0], o[2])
    for m in mons:
        m.check()


def test_flows_errors_are_reported():
    import pcc_rl_b200
    from pcc_rl_b200 import _lib
    rng = np.random.default_rng(3)
    mon = pcc_rl_b200.PccFlowMonitor(100)
    b = synth_batch(rng, 50, 100, unique=True)
    b["flow"][7] = b["flow"][3]                           # duplicate in a batch declared unique
    mon.give_samples(mon.make_batch(**b), unique_flows=True)
    with pytest.raises(_lib.PccError):
        mon.check()
    mon2 = pcc_rl_b200.PccFlowMonitor(100)
    b = synth_batch(rng, 50, 100, unique=True)
    b["flow"][5] = 100                                    # out of range
    mon2.give_samples(mon2.make_batch(**b), unique_flows=True)
    with pytest.raises(_lib.PccError):
        mon2.check()


def test_flows_device_policy_matches_torch():
    """pcc_flows_act: the saved-model agent as an on-device MLP; binary64, tanh from CUDA's libm (1e-12)."""
    import pcc_rl_b200
    import torch
    n_flows = 4096
    rng = np.random.default_rng(11)
    mon = pcc_rl_b200.PccFlowMonitor(n_flows)
    for it in range(4):
        mon.give_samples(mon.make_batch(**synth_batch(rng, n_flows, n_flows, mean_samples=40, t0=float(it))),
                         unique_flows=True, want_obs=False)
    w = [rng.normal(0, 0.3, s) for s in ((32, 30), (32,), (16, 32), (16,), (1, 16), (1,))]
    mon.set_policy(*w)
    a = mon.act().cpu()
    obs = mon.obs().cpu()
    t = [torch.from_numpy(x) for x in w]
    ref = (torch.tanh(torch.tanh(obs @ t[0].T + t[1]) @ t[2].T + t[3]) @ t[4].T + t[5]).reshape(-1)
    assert torch.allclose(a, ref, rtol=1e-12, atol=1e-13)
    r0 = mon.get_rates().clone()
    r1 = mon.get_rates(actions=a.numpy())
    assert (r0 == 0).all() and r1.shape == (n_flows,)


@pytest.mark.parametrize("kernel", ["tma", "ldg"])
def test_flows_both_ingest_kernels_and_misaligned_samples(kernel, monkeypatch):
    """The TMA-staged kernel (default) and the read-only-path kernel give identical histories; a sample array that
    is not 16-byte aligned (cp.async.bulk cannot take it) silently uses the latter."""
    import pcc_rl_b200
    import torch
    monkeypatch.setenv("PCC_FLOWS_KERNEL", kernel)
    n_flows = 20000
    rng = np.random.default_rng(123)
    feats = "send rate,avg latency,latency increase,ack latency inflation,latency ratio"
    import oracle
    fl = oracle.OracleFlows(n_flows, 10, feats)
    mon = pcc_rl_b200.PccFlowMonitor(n_flows, 10, feats)
    for it in range(3):
        b = synth_batch(rng, n_flows - 3, n_flows, mean_samples=[3, 140, 400][it], unique=True, t0=float(it))
        dev = mon.make_batch(**b)
        if it == 1:     # shift the sample array by one element: 8-byte aligned only
            pad = torch.empty(dev["rtt"].numel() + 1, dtype=torch.float64, device=dev["rtt"].device)
            pad[1:] = dev["rtt"]
            dev["rtt"] = pad[1:]
            assert dev["rtt"].data_ptr() % 16 == 8
        o, _ = mon.give_samples(dev, unique_flows=True)
        fl.give_batch(b)
        o = o.cpu().numpy()
        for r in list(range(0, len(b["flow"]), 37)) + [len(b["flow"]) - 1, len(b["flow"]) - 2]:
            assert np.array_equal(o[r], fl.obs(int(b["flow"][r]))), "batch %d record %d" % (it, r)   # unique batch
    final = mon.obs().cpu().numpy()
    for i in range(0, n_flows, 7):
        assert np.array_equal(final[i], fl.obs(i)), "flow %d" % i
    mon.check()


def test_flows_checkpoint_attach():
    """The workspace is the whole state: a monitor attached to a copy continues bit-identically."""
    import pcc_rl_b200
    import torch
    n_flows = 5000
    rng = np.random.default_rng(9)
    feats = "conn min latency,latency ratio,send ratio,loss ratio"
    a = pcc_rl_b200.PccFlowMonitor(n_flows, 7, feats, start_rate=6.0)
    for it in range(3):
        a.give_samples(a.make_batch(**synth_batch(rng, n_flows, n_flows, mean_samples=50, t0=float(it))), unique_flows=True,
                       want_obs=False)
    a.get_rates(actions=rng.normal(0, 1, n_flows))
    torch.cuda.synchronize()
    b = pcc_rl_b200.PccFlowMonitor(n_flows, 7, feats, workspace=a.workspace.clone())
    assert torch.equal(a.obs(), b.obs()) and torch.equal(a.get_rates(), b.get_rates())
    nxt = synth_batch(rng, n_flows, n_flows, mean_samples=50, t0=9.0)
    oa, _ = a.give_samples(a.make_batch(**nxt), unique_flows=True)
    ob, _ = b.give_samples(b.make_batch(**nxt), unique_flows=False)
    assert torch.equal(oa, ob) and torch.equal(a.column("conn_min"), b.column("conn_min"))
    a.check(); b.check()
