// pcc_warp.cuh -- "a warp owns E envs": the MI of pcc_core.cuh::run_mi with the work split the way
// the hardware likes it.
//
//   phase A (per lane, E lanes active): the serial part -- pacing timer, Philox loss draws and
//     the binary64 queue recurrence (network_sim.py:72-84, 156-178) -- runs as E independent
//     chains, one per lane, each appending 16-byte records to its env's in-flight ring.
//     Envs are visited in work-sorted order (pcc_b200.cu: rebalance), so the chains of a warp
//     have similar lengths.
//   phase B (whole warp, one env after the other): hop-1 / hop-2 cursor scans over the env's
//     ring with 32-record coalesced windows (512 B), W windows per round; MI-boundary cluster
//     analysis with ballots and shuffle min-reductions; acked latencies compacted into the
//     warp's shared-memory staging buffer; numpy-exact pairwise means from that buffer with 8
//     lanes as numpy's 8 accumulators and a 3-level xor-shuffle tree.  The next env's ring
//     lines are prefetched while the current one is processed.
//   phase C (per lane): the event that crosses the MI end if it is a send, MI metrics, reward,
//     history/obs row, state write-back.
//
// Arithmetic and event order are those of pcc_core.cuh (proved against the oracle by the host
// twin); the cooperative pieces are the G = 32 instances of pcc_coop.cuh.
#pragma once
#include "pcc_core.cuh"
#include "pcc_coop.cuh"

namespace pcc {

#define PCC_WBUF 1024   // samples staged per warp (8 KB); MIs with more acks re-read the ring

struct ConsumeIn {
    double end, dl, tnext;
    uint32_t tail, h1, h2;
};
struct ConsumeOut {
    uint32_t h1, h2, s_begin, s_end;
    int32_t acked, lost;
    double extra;
    bool has_extra;
    int which;         // 0 = the pacing timer crosses `end` (owner lane sends), 1 = hop-1, 2 = hop-2
    double cur_time;   // valid for which != 0
};

// Phases (2)-(4) of run_mi for ONE env by the whole warp.  All inputs and outputs warp-uniform.
template <class Ring>
__device__ __forceinline__ void consume_mi_warp(const Grp<32> &g, const ConsumeIn &in, Ring &ring, double *buf,
                                                ConsumeOut &out)
{
    constexpr int G = 32;
    const double end = in.end;
    const uint32_t tail = in.tail;
    uint32_t h1 = in.h1, h2 = in.h2;
    int32_t acked = 0, lost = 0;
    out.has_extra = false;
    out.extra = 0.0;
    out.s_begin = h2;

    // ---- hop-1 events with a < end ----------------------------------------------------------
    for (;;) {
        unsigned bm[PCC_SCAN_W];
#pragma unroll
        for (int w = 0; w < PCC_SCAN_W; w++) {
            bool valid;
            const Rec r = load_window(g, ring, h1 + (uint32_t)(w * G), tail, true, valid);
            bm[w] = __ballot_sync(PCC_FULL, valid && (sgn(r.a) || r.a < end));
        }
        int adv = 0;
        bool stop = false;
#pragma unroll
        for (int w = 0; w < PCC_SCAN_W; w++) {
            const int nl = Grp<G>::lead_ones(bm[w]);
            if (!stop) adv += nl;
            stop = stop || (nl < G);
        }
        h1 += (uint32_t)adv;
        if (stop) break;
    }
    bool has1 = false;
    uint32_t m1 = 0; double m1a = 0.0, m1l = 0.0; bool m1d = false;
    {
        uint32_t kk = h1;
        bool open = (kk != tail);
        while (open) {
            bool valid;
            const Rec r = load_window(g, ring, kk, tail, true, valid);
            const bool dr = sgn(r.l);
            const unsigned validm = __ballot_sync(PCC_FULL, valid);
            const unsigned ndm = __ballot_sync(PCC_FULL, valid && !dr);
            const int pend = ndm ? (__ffs(ndm) - 1) : G;   // accepted record closes the cluster
            const bool pending = valid && (int)g.gl <= pend && !sgn(r.a);
            const bool strag = pending && (r.a < end);
            if (strag) ring.store_a(kk + g.gl, negd(r.a));
            window_argmin(g, pending && !strag, r.a, absd(r.l), dr, kk, has1, m1, m1a, m1l, m1d);
            open = (pend == G) && (validm == 0xffffffffu) && ((uint32_t)(kk + G) != tail);
            kk += G;
        }
        __syncwarp();   // straggler flags are read by the hop-2 scan
    }

    // ---- hop-2 events with b < end; acked latencies staged for np.mean ---------------------------
    bool at_live = false;
    for (;;) {
        unsigned bm[PCC_SCAN_W], am[PCC_SCAN_W], lm[PCC_SCAN_W], lv[PCC_SCAN_W];
        double l2[PCC_SCAN_W];
#pragma unroll
        for (int w = 0; w < PCC_SCAN_W; w++) {
            bool valid;
            const uint32_t i = h2 + (uint32_t)(w * G);
            const Rec r = load_window(g, ring, i, tail, true, valid);
            const bool dead = is_dead(r.a);
            const bool c1 = ((int32_t)(i + g.gl - h1) < 0) || sgn(r.a);
            const bool early = (absd(r.a) + in.dl) < end;         // :149-154, link 1 latency == dl
            const bool cons = valid && (dead || (c1 && early));
            bm[w] = __ballot_sync(PCC_FULL, cons);
            am[w] = __ballot_sync(PCC_FULL, cons && !dead && !sgn(r.l));        // :144-145
            lm[w] = __ballot_sync(PCC_FULL, cons && !dead && sgn(r.l));         // :141-142
            lv[w] = __ballot_sync(PCC_FULL, valid && !dead && c1 && !early);
            l2[w] = r.l + in.dl;                                  // rtt = fl(ll + dl)
        }
        int adv = 0;
        bool stop = false;
#pragma unroll
        for (int w = 0; w < PCC_SCAN_W; w++) {
            const int nl = Grp<G>::lead_ones(bm[w]);
            const unsigned lead = Grp<G>::lowmask(nl);
            const unsigned a_w = stop ? 0u : (am[w] & lead);
            if ((a_w >> g.gl) & 1u) {
                const int pos = acked + __popc(a_w & Grp<G>::lowmask((int)g.gl));
                if (pos < PCC_WBUF) buf[pos] = l2[w];
            }
            if (!stop) {
                acked += __popc(a_w);
                lost += __popc(lm[w] & lead);
                adv += nl;
                if (nl < G) at_live = ((lv[w] >> nl) & 1u) != 0u;
            }
            stop = stop || (nl < G);
        }
        h2 += (uint32_t)adv;
        if (stop) break;
    }
    out.s_end = h2;
    bool has2 = false;
    uint32_t m2 = 0; double m2b = 0.0, m2l = 0.0; bool m2d = false;
    {
        uint32_t kk = h2;
        bool open = at_live;
        while (open) {
            bool valid;
            const Rec r = load_window(g, ring, kk, tail, true, valid);
            const bool dr = sgn(r.l);
            const bool dead = is_dead(r.a);
            const bool c1 = ((int32_t)(kk + g.gl - h1) < 0) || sgn(r.a);
            const unsigned x1 = __ballot_sync(PCC_FULL, !valid || (!dead && !c1));   // stop BEFORE this record
            const unsigned x2 = __ballot_sync(PCC_FULL, valid && !dr);               // stop AFTER this record
            const int p1 = x1 ? (__ffs(x1) - 1) : G;
            const int p2 = x2 ? (__ffs(x2) - 1) : G;
            const bool act = (int)g.gl < p1 && (int)g.gl <= p2 && !dead;
            const double b = absd(r.a) + in.dl;
            const double l2 = absd(r.l) + in.dl;
            const bool strag = act && (b < end);
            const unsigned sa = __ballot_sync(PCC_FULL, strag && !dr), sl = __ballot_sync(PCC_FULL, strag && dr);
            const double ex = __shfl_sync(PCC_FULL, l2, sa ? (__ffs(sa) - 1) : 0);
            if (sa) { out.extra = ex; out.has_extra = true; }   // at most one acked per cluster
            acked += __popc(sa);
            lost += __popc(sl);
            if (strag) ring.store_a(kk + g.gl, u2d(PCC_NEG_INF));
            window_argmin(g, act && !strag, b, l2, dr, kk, has2, m2, m2b, m2l, m2d);
            open = (p1 == G) && (p2 == G);
            kk += G;
        }
    }

    // ---- the event that crosses `end` --------------------------------------------------------------
    int which;
    if (has1 && (!has2 || m1a <= m2b)) which = (m1a <= in.tnext) ? 1 : 0;
    else if (has2) which = (m2b <= in.tnext) ? 2 : 0;
    else which = 0;
    out.cur_time = in.tnext;
    if (which == 1) {
        out.cur_time = m1a;
        if (m1 == h1) h1++;
        else if (g.gl == 0) ring.store_a(m1, negd(m1a));
    } else if (which == 2) {
        out.cur_time = m2b;
        if (m2d) lost++; else { acked++; out.extra = m2l; out.has_extra = true; }
        if (m2 == h2) h2++;
        else if (g.gl == 0) ring.store_a(m2, u2d(PCC_NEG_INF));
    }
    // the one possible out-of-order sample is the MI's last sample
    if (out.has_extra && acked <= PCC_WBUF && g.gl == 0) buf[acked - 1] = out.extra;
    __syncwarp();
    out.which = which;
    out.h1 = h1; out.h2 = h2;
    out.acked = acked; out.lost = lost;
}

// numpy's pairwise sum over a[0..n) held in shared memory (n <= PCC_WBUF): the recursion of
// DOUBLE_pairwise_sum with coop_leaf at the leaves.  Warp-uniform.
__device__ __forceinline__ double pw_smem(const Grp<32> &g, const double *a, int n)
{
    if (n <= PCC_LEAF) return coop_leaf(g, a, n);
    int right_n[8];
    const double *right_p[8];
    double left_sum[8];
    bool have_left[8];
    int sp = 0;
    int cur = n;
    const double *p = a;
    for (;;) {
        while (cur > PCC_LEAF) {
            int n2 = cur / 2;
            n2 -= n2 % 8;
            right_n[sp] = cur - n2; right_p[sp] = p + n2; have_left[sp] = false; sp++;
            cur = n2;
        }
        double res = coop_leaf(g, p, cur);
        for (;;) {
            if (sp == 0) return res;
            if (!have_left[sp - 1]) {
                left_sum[sp - 1] = res; have_left[sp - 1] = true;
                cur = right_n[sp - 1]; p = right_p[sp - 1];
                break;
            }
            res = left_sum[sp - 1] + res;
            sp--;
        }
    }
}

// avg latency (sender_obs.py:119-122) and latency increase (:138-142) of one env's MI, warp-wide
template <class Ring>
__device__ __forceinline__ void mi_means_warp(const Grp<32> &g, const ConsumeOut &co, Ring &ring, double dl,
                                              double *buf, bool need_increase, double &avg_lat, double &lat_increase)
{
    const int n = co.acked;
    avg_lat = 0.0;
    lat_increase = 0.0;
    if (n <= 0) return;
    const int half = n / 2;
    if (n <= PCC_WBUF) {
        double sum = 0.0;
        sum += pw_smem(g, buf, n);
        avg_lat = sum / (double)n;
        if (need_increase && half >= 1) {
            double s1 = 0.0, s2 = 0.0;
            s1 += pw_smem(g, buf, half);
            s2 += pw_smem(g, buf + half, n - half);
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
    } else {
        MiOut o;
        o.s_begin = co.s_begin; o.s_end = co.s_end; o.extra = co.extra; o.has_extra = co.has_extra;
        {
            CoopSamples<32, Ring> st(g, ring, buf, o, dl);
            double sum = 0.0;
            sum += coop_pw_sum(g, st, n);
            avg_lat = sum / (double)n;
        }
        if (need_increase) {
            CoopSamples<32, Ring> st(g, ring, buf, o, dl);
            double s1 = 0.0, s2 = 0.0;
            s1 += coop_pw_sum(g, st, half);
            s2 += coop_pw_sum(g, st, n - half);
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
    }
    __syncwarp();
}

// One packet of the send phase, branch-free (network_sim.py:156-178 -> :66-84).
struct LaneChain {
    double t, q, tu;       // next send time, Link.queue_delay, Link.queue_delay_update_time
    uint32_t tail;
    int32_t sent;
    bool ovf;
};
template <class Ring>
__device__ __forceinline__ void lane_send_one(LaneChain &c, const EnvState &s, Ring &ring, uint32_t h2, uint32_t cap,
                                              double inv_rate, double u)
{
    const bool rdrop = u < s.lr;                                        // :73
    const long long yb = __double_as_longlong(c.q - (c.t - c.tu));     // :66-67
    const double w = __longlong_as_double(yb & ~(yb >> 63));            // max(0.0, y)
    const double cc = s.d_bw + w;                                       // :77-79
    const bool full = cc > s.max_qd;
    const double ll = s.dl + w;                                         // :69-70
    c.q = rdrop ? c.q : (full ? w : cc);                                // :74-82
    c.tu = rdrop ? c.tu : c.t;
    const bool dropped = rdrop || full;
    Rec r;
    r.a = c.t + ll;                                                     // :173-174
    r.l = __longlong_as_double(__double_as_longlong(ll) | (dropped ? (long long)PCC_SIGN : 0ll));
    if ((uint32_t)(c.tail - h2) >= cap) c.ovf = true;                   // fatal, reported by the host
    else { ring.store(c.tail, r); c.tail++; }
    c.t = c.t + inv_rate;                                               // :161
    c.sent++;
}

// All sends with t < end of one env, by its owner lane.  One Philox block feeds two packets.
template <class Ring>
__device__ __forceinline__ void lane_send_phase(LaneChain &c, const EnvState &s, Ring &ring, PhiloxRng &rng,
                                                uint32_t h2, uint32_t cap, double end, double inv_rate)
{
    if (c.t < end && (rng.draws & 1ull)) {
        lane_send_one(c, s, ring, h2, cap, inv_rate, res53(rng.w2, rng.w3));
        rng.draws++;
    }
    while (c.t < end) {
        uint32_t a, b;
        rng.block(rng.draws >> 1, a, b, rng.w2, rng.w3);
        lane_send_one(c, s, ring, h2, cap, inv_rate, res53(a, b));
        rng.draws++;
        if (c.t < end) {
            lane_send_one(c, s, ring, h2, cap, inv_rate, res53(rng.w2, rng.w3));
            rng.draws++;
        }
    }
}

}  // namespace pcc
