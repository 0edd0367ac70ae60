// pcc_multi_fast.cuh -- several senders on one bottleneck WITHOUT the event heap (BASELINE config 5).
//
// pcc_multi_core.cuh keeps the reference's heap because S pacing timers interleave.  But all senders share link 0's
// queue and the same propagation delays, so what made the single-sender case streamable (pcc_core.cuh) still holds:
// a packet is fully determined when it is sent -- hop-1 time a = fl(t + ll), hop-2 time b = fl(a + dl), rtt = fl(ll + dl),
// ll = dl + queue delay seen -- and in GLOBAL send order the hop times are non-decreasing except inside a drop cluster
// (a run of dropped packets plus the accepted packet that ends it: a drop does not add to the shared queue, so the
// members' arrival times coincide up to rounding).  Hence: one shared in-flight ring with a sender id per record, the
// S timers merged by (time, sender index), and the three cursors of pcc_core.cuh.  The reference's event order is the
// tuple order (time, sender, 'A' < 'S', next_hop, cur_latency, dropped) -- the sender index now sits in second place,
// so it precedes type, hop and latency in every tie -- and as before it only matters for which event crosses the end
// of the MI: SEND events of different senders at one instant mutate the queue in sender order (the merge does that),
// ACK-type events do not touch shared state.
//
// Shared with the host twin (tests/twin), which checks it against the oracle and the reference's goldens.  The heap
// path stays for the cwnd / latency-noise variants (their event times are not monotone streams).
#pragma once
#include "pcc_multi_core.cuh"

namespace pcc {

struct MFast {
    double next_send[PCC_MAX_SENDERS];   // pending pacing-timer event of each sender
    uint32_t tail, h1, h2, pad;          // shared ring cursors, as in EnvState
};

// tuple order of two events (time, sender, type ['A' = 0 < 'S' = 1], hop, latency, dropped)
PCC_HD bool mf_less(double t1, int s1, int ty1, int h1, double l1, bool d1,
                    double t2, int s2, int ty2, int h2, double l2, bool d2)
{
    if (t1 != t2) return t1 < t2;
    if (s1 != s2) return s1 < s2;
    if (ty1 != ty2) return ty1 < ty2;
    if (h1 != h2) return h1 < h2;
    if (l1 != l2) return l1 < l2;
    return !d1 && d2;
}

// Ring: Rec load(i); void store(i, Rec); void store_a(i, double); int sid(i); void set_sid(i, int); uint32_t capacity()
template <class Ring, class Rng>
PCC_HD bool mfast_run_for_dur(MNet &net, MSender *snd, int S, MFast &f, Ring &ring, double *samples, int cap_s, Rng &rng,
                              double dur)
{
    bool ok = true;
    const double end = net.cur_time + dur;                           // network_sim.py:124
    for (int i = 0; i < S; i++) {                                    // reset_obs :125-126, :319-324
        snd[i].sent = 0; snd[i].acked = 0; snd[i].lost = 0; snd[i].n_rtt = 0;
        snd[i].obs_start = net.cur_time;
    }
    const uint32_t cap = ring.capacity();
    uint32_t tail = f.tail, h1 = f.h1, h2 = f.h2;
    double qd = net.qd, t_upd = net.t_upd;

    // the pending timer that fires first: smallest (time, sender index)
#define PCC_MF_NEXT_TIMER(it)                                                                 \
    int it = 0;                                                                               \
    for (int i_ = 1; i_ < S; i_++) if (f.next_send[i_] < f.next_send[it]) it = i_;
    // the SEND event of sender i (:156-178 -> :66-84)
#define PCC_MF_SEND(i)                                                                        \
    {                                                                                         \
        const double t = f.next_send[i];                                                      \
        snd[i].sent++;                                               /* :159-160 */           \
        const double w = py_max0(qd - (t - t_upd));                  /* :170 -> :66-70 */     \
        const double ll = net.dl + w;                                                         \
        bool dropped;                                                                         \
        if (rng.next() < net.lr) dropped = true;                     /* :73 */                \
        else {                                                                                \
            qd = w; t_upd = t;                                       /* :75-76 */             \
            if (w > net.w_full) dropped = true;                      /* :79 */                \
            else { qd += net.d_bw; dropped = false; }                /* :82 */                \
        }                                                                                     \
        Rec r; r.a = t + ll; r.l = dropped ? negd(ll) : ll;          /* :173-175 */           \
        if ((uint32_t)(tail - h2) >= cap) { ok = false; tail--; }    /* ring overflow: fatal, reported */ \
        ring.store(tail, r); ring.set_sid(tail, i); tail++;                                   \
        f.next_send[i] = t + (1.0 / snd[i].rate);                    /* :161 */               \
    }
    // a consumed hop-2 event of sender sd_ (:140-145)
#define PCC_MF_HOP2(sd_, dr_, l2_)                                                            \
    {                                                                                         \
        MSender &x_ = snd[sd_];                                                               \
        if (dr_) x_.lost++;                                                                   \
        else {                                                                                \
            x_.acked++;                                                                       \
            if (x_.n_rtt < cap_s) samples[(size_t)(sd_) * cap_s + x_.n_rtt] = (l2_); else ok = false; \
            x_.n_rtt++;                                                                       \
        }                                                                                     \
    }

    // ---- (1) every send with t < end, timers merged by (time, sender) ------------------------------------
    for (;;) {
        PCC_MF_NEXT_TIMER(it);
        if (!(f.next_send[it] < end)) break;
        PCC_MF_SEND(it);
    }

    // ---- (2) hop-1 events with a < end ---------------------------------------------------------------
    while (h1 != tail) {
        const Rec r = ring.load(h1);
        if (!sgn(r.a) && !(r.a < end)) break;
        h1++;
    }
    bool has1 = false;
    uint32_t m1 = 0; double m1a = 0.0, m1l = 0.0; bool m1d = false; int m1s = 0;
    for (uint32_t k = h1; k != tail; k++) {                          // the boundary cluster
        const Rec r = ring.load(k);
        const bool dr = sgn(r.l);
        if (!sgn(r.a)) {
            if (r.a < end) ring.store_a(k, negd(r.a));               // straggler: consumed out of order
            else {
                const double l = absd(r.l);
                const int sd = ring.sid(k);
                if (!has1 || mf_less(r.a, sd, 0, 1, l, dr, m1a, m1s, 0, 1, m1l, m1d)) {
                    has1 = true; m1 = k; m1a = r.a; m1l = l; m1d = dr; m1s = sd;
                }
            }
        }
        if (!dr) break;                                              // an accepted packet closes the cluster
    }

    // ---- (3) hop-2 events with b < end ---------------------------------------------------------------
    bool at_live = false;
    while (h2 != tail) {
        const Rec r = ring.load(h2);
        if (!is_dead(r.a)) {
            const bool c1 = ((int32_t)(h2 - h1) < 0) || sgn(r.a);
            if (!c1) break;
            const double b = absd(r.a) + net.dl;                     // link 1: latency == dl exactly (N1)
            if (!(b < end)) { at_live = true; break; }
            const int sd = ring.sid(h2);
            PCC_MF_HOP2(sd, sgn(r.l), absd(r.l) + net.dl);
        }
        h2++;
    }
    bool has2 = false;
    uint32_t m2 = 0; double m2b = 0.0, m2l = 0.0; bool m2d = false; int m2s = 0;
    if (at_live) {
        for (uint32_t k = h2; k != tail; k++) {
            const Rec r = ring.load(k);
            const bool dr = sgn(r.l);
            if (!is_dead(r.a)) {
                const bool c1 = ((int32_t)(k - h1) < 0) || sgn(r.a);
                if (!c1) break;                                      // later hop-2 events are >= end + dl
                const double b = absd(r.a) + net.dl;
                const double l2 = absd(r.l) + net.dl;
                const int sd = ring.sid(k);
                if (b < end) {                                       // straggler
                    PCC_MF_HOP2(sd, dr, l2);
                    ring.store_a(k, u2d(PCC_NEG_INF));
                } else if (!has2 || mf_less(b, sd, 0, 2, l2, dr, m2b, m2s, 0, 2, m2l, m2d)) {
                    has2 = true; m2 = k; m2b = b; m2l = l2; m2d = dr; m2s = sd;
                }
            }
            if (!dr) break;
        }
    }

    // ---- (4) the event that crosses `end`: tuple-order minimum of (timer, hop-1, hop-2) -----------------
    PCC_MF_NEXT_TIMER(it);
    int which = 0;
    double bt = f.next_send[it]; int bs = it, bty = 1, bh = 0; double bl = 0.0; bool bd = false;
    if (has1 && mf_less(m1a, m1s, 0, 1, m1l, m1d, bt, bs, bty, bh, bl, bd)) {
        which = 1; bt = m1a; bs = m1s; bty = 0; bh = 1; bl = m1l; bd = m1d;
    }
    if (has2 && mf_less(m2b, m2s, 0, 2, m2l, m2d, bt, bs, bty, bh, bl, bd)) which = 2;
    if (which == 0) {
        net.cur_time = f.next_send[it];
        PCC_MF_SEND(it);
    } else if (which == 1) {
        net.cur_time = m1a;
        if (m1 == h1) h1++; else ring.store_a(m1, negd(m1a));
    } else {
        net.cur_time = m2b;
        PCC_MF_HOP2(m2s, m2d, m2l);
        if (m2 == h2) h2++; else ring.store_a(m2, u2d(PCC_NEG_INF));
    }
#undef PCC_MF_NEXT_TIMER
#undef PCC_MF_SEND
#undef PCC_MF_HOP2
    net.qd = qd; net.t_upd = t_upd;
    f.tail = tail; f.h1 = h1; f.h2 = h2;
    return ok;
}

// reset: fresh links + S senders, two discarded warm-up MIs (network_sim.py:454-484)
template <class Ring, class Rng>
PCC_HD bool mfast_reset(MNet &net, MSender *snd, int S, MFast &f, Ring &ring, double *samples, int cap_s, Rng &rng,
                        double bw, double dl, int64_t queue, double lr, const double *rates)
{
    net.d_bw = 1.0 / bw; net.dl = dl; net.lr = lr; net.max_qd = (double)queue / bw;
    net.w_full = tail_drop_threshold(net.d_bw, net.max_qd);
    net.qd = 0.0; net.t_upd = 0.0; net.cur_time = 0.0; net.run_dur = 3 * dl; net.steps = 0; net.heap_n = 0;
    f.h1 = f.tail; f.h2 = f.tail;                                   // drop everything in flight; positions keep counting
    for (int i = 0; i < S; i++) {
        snd[i].rate = rates[i]; snd[i].conn_min = 0.0;
        snd[i].sent = snd[i].acked = snd[i].lost = snd[i].n_rtt = 0; snd[i].obs_start = 0.0;
        snd[i].cwnd = 0; snd[i].inflight = 0;
        f.next_send[i] = 1.0 / rates[i];                             // queue_initial_packets :107-111
    }
    bool ok = mfast_run_for_dur(net, snd, S, f, ring, samples, cap_s, rng, net.run_dur);   // :478
    ok = mfast_run_for_dur(net, snd, S, f, ring, samples, cap_s, rng, net.run_dur) && ok;  // :479
    return ok;
}

// step(actions[S]) for one env: multi_step of pcc_multi_core.cuh on the streaming MI
template <class Ring, class Rng>
PCC_HD bool mfast_step(MNet &net, MSender *snd, int S, MFast &f, Ring &ring, double *samples, int cap_s, Rng &rng,
                       const double *actions, const Consts &c, const int *ids, int F, bool need_increase,
                       double *rows, double *rewards, int32_t *counts, bool &done)
{
    for (int i = 0; i < S; i++) snd[i].rate = apply_rate_delta(snd[i].rate, actions[i], c);   // :409-412
    const bool ok = mfast_run_for_dur(net, snd, S, f, ring, samples, cap_s, rng, net.run_dur);   // :416
    double avg0 = 0.0;
    for (int i = 0; i < S; i++) {
        MiStats st;
        multi_sender_stats(net, snd[i], samples + (size_t)i * cap_s, c, need_increase, st);
        rewards[i] = st.reward;
        for (int k = 0; k < F; k++) rows[i * F + k] = metric_value(st, ids[k]);
        if (counts) { counts[3 * i] = snd[i].sent; counts[3 * i + 1] = snd[i].acked; counts[3 * i + 2] = snd[i].lost; }
        if (i == 0) avg0 = st.avg_lat;
    }
    net.steps += 1;                                                  // :419
    if (avg0 > 0.0) net.run_dur = 0.5 * avg0;                        // :437-438 (sender 0, as written)
    done = net.steps >= c.max_steps;                                 // :444
    return ok;
}

}  // namespace pcc
