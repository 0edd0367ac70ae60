"""Executable specification of the chunked send phase of the cooperative kernels (pcc_coop.cuh: group_send_chunks,
stages S3 / S4 / S5; pcc_multi_warp.cuh: the same on merged timers) against the direct per-packet statement of the
reference's link model (gym/network_sim.py:66-84, 170-175), in the same IEEE binary64 arithmetic (Python floats).

The kernels do not run the queue update packet by packet.  Per chunk they (S3) form x_k = t_k - t_(last packet before k
that reached the queue) over the packets that were not randomly dropped, compacted; (S4) run y = q - x, q = f(y) over
that array -- the only serial part; (S5) rebuild every record in parallel, a randomly dropped packet recomputing the
queue its predecessor left.  This test states that transformation lane by lane and checks, on random chunks incl.
one- and zero-packet queues, exact timer ties, loss 0 and 1, that records and carried state are bit-identical."""
import ctypes as C

import numpy as np
import pytest

import twin_util


def _w_full(d_bw, max_qd):
    L = twin_util.lib()
    L.twin_tail_drop_threshold.restype = C.c_double
    L.twin_tail_drop_threshold.argtypes = [C.c_double, C.c_double]
    return L.twin_tail_drop_threshold(d_bw, max_qd)


def direct(ts, lost_bits, qd, t_upd, dl, d_bw, max_qd):
    """network_sim.py:170-175 -> :66-84 per packet; returns records (a, ll, dropped) and the link state after."""
    recs = []
    for t, rl in zip(ts, lost_bits):
        w = max(0.0, qd - (t - t_upd))                       # :66-70
        ll = dl + w
        if rl:                                               # :73
            dropped = True
        else:
            qd, t_upd = w, t                                 # :75-76
            if d_bw + qd > max_qd:                           # :77-79, the reference's own test
                dropped = True
            else:
                qd += d_bw                                   # :82
                dropped = False
        recs.append((t + ll, ll, dropped))
    return recs, qd, t_upd


def chunked(ts, lost_bits, qd, t_upd, dl, d_bw, max_qd):
    """The kernels' formulation (one chunk of <= 64 packets)."""
    w_full = _w_full(d_bw, max_qd)
    k0 = 0.0 if 0.0 > w_full else d_bw                        # q' when the queue has drained
    full0 = 0.0 > w_full
    cnt = len(ts)
    reach = [k for k in range(cnt) if not lost_bits[k]]       # packets that reach the queue (ndm)
    # S3: every "lane" k in parallel
    xs = []
    for j, k in enumerate(reach):
        tuk = ts[reach[j - 1]] if j > 0 else t_upd
        xs.append(ts[k] - tuk)
    # S4: the serial recurrence, in place
    ys, state = [], qd
    for x in xs:
        y = state - x
        ys.append(y)
        cpos = d_bw + y
        state = ((y if y > w_full else cpos) if y > 0.0 else k0)
    # S5: every lane rebuilds its record
    recs = []
    for k in range(cnt):
        before = [r for r in reach if r < k]
        rank = len(before)
        if not lost_bits[k]:
            y = ys[rank]
        else:
            qp = qd
            if rank > 0:
                yp = ys[rank - 1]
                qp = ((yp if yp > w_full else d_bw + yp) if yp > 0.0 else k0)
            tuk = ts[before[-1]] if before else t_upd
            y = qp - (ts[k] - tuk)
        pos = y > 0.0
        w = y if pos else 0.0
        full = (y > w_full) if pos else full0
        ll = dl + w
        recs.append((ts[k] + ll, ll, bool(lost_bits[k]) or full))
    t_upd_out = ts[reach[-1]] if reach else t_upd
    return recs, state, t_upd_out


@pytest.mark.parametrize("seed", range(8))
def test_chunked_send_phase_equals_direct_statement(seed):
    g = np.random.default_rng(seed)
    for case in range(400):
        bw = float(np.exp(g.uniform(np.log(40), np.log(90000))))
        queue = int(g.choice([0, 1, 1, 2, 3, 10, 1000]))
        d_bw, max_qd = 1.0 / bw, queue / bw
        dl = float(np.exp(g.uniform(np.log(0.001), np.log(0.5))))
        p_loss = float(g.choice([0.0, 0.02, 0.5, 1.0]))
        cnt = int(g.integers(1, 65))
        # merged send times of 1-3 senders: non-decreasing, with exact ties when two timers coincide
        gaps = g.exponential(1.0 / float(g.uniform(40, 3000)), cnt)
        gaps[g.random(cnt) < 0.1] = 0.0
        t0 = float(g.uniform(0, 50))
        ts = list(t0 + np.cumsum(gaps))
        lost = list(g.random(cnt) < p_loss)
        # link state on entry: anything the previous chunk can leave, incl. a drained and an over-full queue
        qd = float(g.choice([0.0, d_bw, g.uniform(0, 1.5 * max(max_qd, d_bw))]))
        t_upd = t0 - float(g.choice([0.0, g.uniform(0, 3 * d_bw), g.uniform(0, 1.0)]))
        a, qa, ta = direct(ts, lost, qd, t_upd, dl, d_bw, max_qd)
        b, qb, tb = chunked(ts, lost, qd, t_upd, dl, d_bw, max_qd)
        assert a == b, (seed, case, queue, p_loss)
        assert qa == qb and ta == tb, (seed, case, queue, p_loss, qa, qb)
