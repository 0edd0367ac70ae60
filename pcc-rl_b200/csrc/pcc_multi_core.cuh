// pcc_multi_core.cuh -- several senders sharing one bottleneck (BASELINE config 5).
//
// The three-cursor algorithm of pcc_core.cuh needs a single pacing timer; with S senders on the
// same links the event order interleaves S timers with ties across senders, so this path keeps
// the reference's structure -- a binary heap of events per env (network_sim.py:100-178) -- and
// runs one env per thread.  It is the slow, general path: exact, not fast.  Shared with the host
// twin (tests/twin) so it can be checked against the oracle without a GPU.
//
// Semantics (the reference's env never creates more than one sender; pinned against the
// reference's Network/Link/Sender classes with the patch Sender.__lt__ = id order):
//   - all senders use path [l0, l1], dest 0: one shared queue on l0, l1 adds its delay only;
//   - heap order = (time, sender index, 'A' < 'S', next_hop, cur_latency, dropped);
//   - step applies action i to sender i, every sender gets its own MI / obs / reward
//     (network_sim.py:194,205 on its own MI), run_dur follows sender 0 (:437-438);
//   - one uniform draw per sent packet from the env's single stream, in event order (:73).
//
// The same heap loop also carries the two variants the reference compiles out behind module switches
// (pinned against the reference with the switches turned on, tests/golden/variant_*):
//   - USE_CWND (network_sim.py:54, 243-255, 283-289, 413-414): a pacing-timer event only emits a packet while
//     bytes_in_flight / 1500 < cwnd -- the timer is re-armed and the link is still "entered" (queue update and
//     loss draw, :170-175 are outside the can_send branch) -- and step() takes a second action for the window;
//   - USE_LATENCY_NOISE (:51-52, 150-151, 171-172): every hop's latency is multiplied by
//     random.uniform(1.0, 1.1) = 1.0 + (1.1 - 1.0) * random(), one extra draw per hop, before the loss draw.
// With either switch on, event times are no longer three monotone streams, which is why these variants live
// here (n_senders = 1 gives the reference's own single-sender env) and not in the three-cursor kernels.
#pragma once
#include "pcc_core.cuh"

namespace pcc {

#define PCC_MAX_SENDERS 4

struct MEvent {
    double time;
    double lat;
    uint32_t meta;   // sender | type << 8 ('A' = 0, 'S' = 1) | next_hop << 9 | dropped << 11
    uint32_t pad;
};
PCC_HD uint32_t mev_meta(int sender, int type, int hop, int dropped)
{
    return (uint32_t)sender | ((uint32_t)type << 8) | ((uint32_t)hop << 9) | ((uint32_t)dropped << 11);
}
// tuple order of the reference's heap entries
PCC_HD bool mev_less(const MEvent &a, const MEvent &b)
{
    if (a.time != b.time) return a.time < b.time;
    const uint32_t sa = a.meta & 0xffu, sb = b.meta & 0xffu;
    if (sa != sb) return sa < sb;
    const uint32_t ta = (a.meta >> 8) & 1u, tb = (b.meta >> 8) & 1u;
    if (ta != tb) return ta < tb;
    const uint32_t ha = (a.meta >> 9) & 3u, hb = (b.meta >> 9) & 3u;
    if (ha != hb) return ha < hb;
    if (a.lat != b.lat) return a.lat < b.lat;
    return ((a.meta >> 11) & 1u) < ((b.meta >> 11) & 1u);
}

struct MSender {
    double rate, obs_start, conn_min;
    int32_t sent, acked, lost, n_rtt;
    int32_t cwnd;        // Sender.cwnd (packets; int after set_cwnd)
    int32_t inflight;    // Sender.bytes_in_flight / BYTES_PER_PACKET (always a whole number of packets)
};
// module switches and window constants of network_sim.py:33-34, 51-54, 209
struct Variant {
    int32_t use_cwnd, use_noise;
    double max_noise;                          // MAX_LATENCY_NOISE 1.1
    int32_t initial_cwnd, min_cwnd, max_cwnd;  // 25, MIN_CWND 4, MAX_CWND 5000
};
PCC_HD Variant default_variant()
{
    Variant v; v.use_cwnd = 0; v.use_noise = 0; v.max_noise = 1.1; v.initial_cwnd = 25; v.min_cwnd = 4; v.max_cwnd = 5000;
    return v;
}
// Sender.apply_cwnd_delta + set_cwnd (:243-249, 283-289): cwnd = int(new) (truncation), then clamped.  The clamp
// is applied before the conversion where the value is out of int range -- int() then clamp gives the same.
PCC_HD int32_t apply_cwnd_delta(int32_t cwnd, double action, const Consts &c, const Variant &v)
{
    const double delta = action * c.delta_scale;
    const double nw = (delta >= 0.0) ? (double)cwnd * (1.0 + delta) : (double)cwnd / (1.0 - delta);
    if (nw >= (double)v.max_cwnd + 1.0) return v.max_cwnd;
    if (nw < (double)v.min_cwnd) return v.min_cwnd;
    int32_t r = (int32_t)nw;
    if (r > v.max_cwnd) r = v.max_cwnd;
    if (r < v.min_cwnd) r = v.min_cwnd;
    return r;
}
struct MNet {
    double d_bw, dl, lr, max_qd, w_full, qd, t_upd;   // link 0 (link 1: same dl, never queues)
    double cur_time, run_dur;
    int32_t steps, heap_n;
};

template <class Heap>   // Heap: MEvent get(int), void set(int, MEvent), int capacity()
PCC_HD bool mheap_push(Heap &h, int32_t &n, const MEvent &ev)
{
    if (n >= h.capacity()) return false;
    int i = n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        const MEvent pe = h.get(p);
        if (!mev_less(ev, pe)) break;
        h.set(i, pe);
        i = p;
    }
    h.set(i, ev);
    return true;
}
template <class Heap>
PCC_HD MEvent mheap_pop(Heap &h, int32_t &n)
{
    const MEvent top = h.get(0);
    const MEvent last = h.get(--n);
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        MEvent ce = h.get(c);
        if (c + 1 < n) {
            const MEvent ce2 = h.get(c + 1);
            if (mev_less(ce2, ce)) { c++; ce = ce2; }
        }
        if (!mev_less(ce, last)) break;
        h.set(i, ce);
        i = c;
    }
    if (n > 0) h.set(i, last);
    return top;
}

// Network.run_for_dur (network_sim.py:123-178) for S senders.  `samples` = [S][cap_s] acked latencies of
// this MI.  Returns false if the heap or a sample array overflowed.
template <class Heap, class Rng>
PCC_HD bool multi_run_for_dur(MNet &net, MSender *snd, int S, Heap &heap, double *samples, int cap_s, Rng &rng,
                              double dur, const Variant &v)
{
    const double noise_span = v.max_noise - 1.0;                     // uniform(a, b) = a + (b - a) * random()
    bool ok = true;
    const double end = net.cur_time + dur;                           // :124
    for (int i = 0; i < S; i++) {                                    // reset_obs :125-126, :319-324
        snd[i].sent = 0; snd[i].acked = 0; snd[i].lost = 0; snd[i].n_rtt = 0;
        snd[i].obs_start = net.cur_time;
    }
    while (net.cur_time < end) {                                     // :128
        const MEvent ev = mheap_pop(heap, net.heap_n);               // :129
        const int sid = (int)(ev.meta & 0xffu), type = (int)((ev.meta >> 8) & 1u);
        const int hop = (int)((ev.meta >> 9) & 3u), dropped = (int)((ev.meta >> 11) & 1u);
        MSender &sd = snd[sid];
        net.cur_time = ev.time;                                      // :131
        if (type == 0) {                                             // ACK :139
            if (hop == 2) {                                          // :140
                if (dropped) sd.lost++;                              // :141-142
                else {                                               // :144-145
                    sd.acked++;
                    if (sd.n_rtt < cap_s) samples[(size_t)sid * cap_s + sd.n_rtt] = ev.lat; else ok = false;
                    sd.n_rtt++;
                }
                sd.inflight--;                                       // :269, :273
            } else {                                                 // :147-154, link 1: latency == dl (N1)
                double l1 = net.dl;
                if (v.use_noise) l1 *= 1.0 + noise_span * rng.next(); // :150-151
                MEvent nw; nw.time = ev.time + l1; nw.lat = ev.lat + l1;
                nw.meta = mev_meta(sid, 0, hop + 1, dropped); nw.pad = 0;
                ok = mheap_push(heap, net.heap_n, nw) && ok;
            }
        } else {                                                     // SEND at hop 0 :155-175
            const bool can_send = !v.use_cwnd || sd.inflight < sd.cwnd;   // :251-255 (whole packets: exact)
            if (can_send) { sd.sent++; sd.inflight++; }              // :158-160
            MEvent timer; timer.time = net.cur_time + (1.0 / sd.rate); timer.lat = 0.0;
            timer.meta = mev_meta(sid, 1, 0, 0); timer.pad = 0;
            ok = mheap_push(heap, net.heap_n, timer) && ok;          // :161
            const double t = net.cur_time;
            const double w = py_max0(net.qd - (t - net.t_upd));      // :170 -> :66-70
            double ll = net.dl + w;
            if (v.use_noise) ll *= 1.0 + noise_span * rng.next();    // :171-172, before the loss draw
            int drop;
            if (rng.next() < net.lr) drop = 1;                       // :73 (also when the window held the packet back)
            else {
                net.qd = w; net.t_upd = t;                           // :75-76
                if (w > net.w_full) drop = 1;                        // :79 (tail_drop_threshold)
                else { net.qd += net.d_bw; drop = 0; }               // :82
            }
            MEvent nw; nw.time = t + ll; nw.lat = 0.0 + ll;          // :173-174
            nw.meta = mev_meta(sid, 0, 1, drop); nw.pad = 0;         // type flips to ACK: next_hop == dest (:166-168)
            if (can_send) ok = mheap_push(heap, net.heap_n, nw) && ok;   // :177 push_new_event
        }
    }
    return ok;
}

// np.mean over a plain array with numpy's pairwise summation
struct ArrayReader {
    const double *a; int i;
    PCC_HD double next() { return a[i++]; }
};

// MI metrics + reward of sender `sd` after an MI
PCC_HD void multi_sender_stats(const MNet &net, MSender &sd, const double *smp, const Consts &c, bool need_increase,
                               MiStats &st)
{
    MiOut o;
    o.sent = sd.sent; o.acked = sd.acked; o.lost = sd.lost; o.start = sd.obs_start; o.end = net.cur_time;
    const int n = sd.n_rtt;
    double avg = 0.0, inc = 0.0;
    if (n > 0) { ArrayReader rd{smp, 0}; avg = np_mean_stream(rd, n); }
    if (need_increase && n / 2 >= 1) {
        ArrayReader rd{smp, 0};
        const double first = np_mean_stream(rd, n / 2);
        const double second = np_mean_stream(rd, n - n / 2);
        inc = second - first;
    }
    mi_stats_finish(o, c, avg, inc, sd.conn_min, true, st);
}

// reset: fresh links + S senders, two discarded warm-up MIs (network_sim.py:454-484)
template <class Heap, class Rng>
PCC_HD bool multi_reset(MNet &net, MSender *snd, int S, Heap &heap, double *samples, int cap_s, Rng &rng, double bw,
                        double dl, int64_t queue, double lr, const double *rates, const Variant &v)
{
    net.d_bw = 1.0 / bw; net.dl = dl; net.lr = lr; net.max_qd = (double)queue / bw;
    net.w_full = tail_drop_threshold(net.d_bw, net.max_qd);
    net.qd = 0.0; net.t_upd = 0.0; net.cur_time = 0.0; net.run_dur = 3 * dl; net.steps = 0; net.heap_n = 0;
    bool ok = true;
    for (int i = 0; i < S; i++) {
        snd[i].rate = rates[i]; snd[i].conn_min = 0.0;
        snd[i].sent = snd[i].acked = snd[i].lost = snd[i].n_rtt = 0; snd[i].obs_start = 0.0;
        snd[i].cwnd = v.initial_cwnd; snd[i].inflight = 0;          // a fresh Sender (:209-226)
        MEvent first; first.time = 1.0 / rates[i]; first.lat = 0.0; first.meta = mev_meta(i, 1, 0, 0); first.pad = 0;
        ok = mheap_push(heap, net.heap_n, first) && ok;             // queue_initial_packets :107-111
    }
    ok = multi_run_for_dur(net, snd, S, heap, samples, cap_s, rng, net.run_dur, v) && ok;   // :478
    ok = multi_run_for_dur(net, snd, S, heap, samples, cap_s, rng, net.run_dur, v) && ok;   // :479
    return ok;
}

}  // namespace pcc

namespace pcc {

// step(actions[S]) for one env (network_sim.py:406-444 generalised to S senders).  Writes, per sender, the
// new history row (F values, already divided by the metric scale), the reward and the packet counts.
template <class Heap, class Rng>
PCC_HD bool multi_step(MNet &net, MSender *snd, int S, Heap &heap, double *samples, int cap_s, Rng &rng,
                       const double *actions, const double *cwnd_actions, const Consts &c, const Variant &v,
                       const int *ids, int F, bool need_increase,
                       double *rows, double *rewards, int32_t *counts, bool &done)
{
    for (int i = 0; i < S; i++) {
        snd[i].rate = apply_rate_delta(snd[i].rate, actions[i], c);                           // :409-412
        if (v.use_cwnd && cwnd_actions) snd[i].cwnd = apply_cwnd_delta(snd[i].cwnd, cwnd_actions[i], c, v);   // :413-414
    }
    const bool ok = multi_run_for_dur(net, snd, S, heap, samples, cap_s, rng, net.run_dur, v);   // :416
    double avg0 = 0.0;
    for (int i = 0; i < S; i++) {
        MiStats st;
        multi_sender_stats(net, snd[i], samples + (size_t)i * cap_s, c, need_increase, st);
        rewards[i] = st.reward;
        for (int f = 0; f < F; f++) rows[i * F + f] = metric_value(st, ids[f]);
        if (counts) { counts[3 * i] = snd[i].sent; counts[3 * i + 1] = snd[i].acked; counts[3 * i + 2] = snd[i].lost; }
        if (i == 0) avg0 = st.avg_lat;
    }
    net.steps += 1;                                                  // :419
    if (avg0 > 0.0) net.run_dur = 0.5 * avg0;                        // :437-438 (sender 0, as written)
    done = net.steps >= c.max_steps;                                 // :444
    return ok;
}

}  // namespace pcc
