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

// Per-warp shared memory (dynamic): [wbuf + 32] doubles of acked-latency staging, followed by the shared memory of the
// single-env send phase (SoloSendSmem).  wbuf is a launch parameter (512-1024 for big batches, 4096 when the batch is
// small and occupancy is not the limit); MIs with more acks than wbuf stage in the warp's global scratch, and beyond
// that re-read the ring (streaming path).
#define PCC_SEND_SMEM_BYTES ((sizeof(SoloSendSmem) + 15) & ~(size_t)15)
__host__ __device__ inline size_t warp_smem_bytes(int wbuf) { return (size_t)(wbuf + 32) * 8 + PCC_SEND_SMEM_BYTES; }

struct ConsumeIn {
    double end, dl, tnext;
    uint32_t tail, h1, h2;
    int wbuf;          // capacity of the staging buffer
    // resuming after consume_scan_warp (a helper warp ran the in-order scans over the records that existed before
    // this MI's sends): events already counted and staged, and the MI's first hop-2 position.  Fresh: 0, 0, h2.
    int32_t acked0, lost0;
    uint32_t s_begin;
};
struct ScanOut { uint32_t h1, h2; int32_t acked, lost; };
struct ConsumeOut {
    uint32_t h1, h2, s_begin, s_end;
    int32_t acked, lost;
    double extra;
    bool has_extra;
    int which;         // 0 = the pacing timer crosses `end` (owner lane sends), 1 = hop-1, 2 = hop-2
    double cur_time;   // valid for which != 0
};

// ---- rare paths, kept out of line so the common path stays small in the instruction cache ------

// General hop-1 boundary analysis (pcc_core.cuh::run_mi phase 2): cluster starting at h1 whose first
// record is a dropped packet -- stragglers + tuple-order minimum over the cluster remainder.
template <class Ring>
__device__ __noinline__ void boundary1_slow(Ring ring, uint32_t h1, uint32_t tail, double end, bool &has1,
                                            uint32_t &m1, double &m1a, double &m1l, bool &m1d)
{
    constexpr int G = 32;
    const Grp<32> g;
    has1 = false; m1 = 0; m1a = 0.0; m1l = 0.0; m1d = false;
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

// General hop-2 boundary analysis (phase 3): the live record at h2 is a dropped packet.
template <class Ring>
__device__ __noinline__ void boundary2_slow(Ring ring, uint32_t h1, uint32_t h2, uint32_t tail, double dl, double end,
                                            int32_t &acked, int32_t &lost, double &extra, bool &has_extra,
                                            bool &has2, uint32_t &m2, double &m2b, double &m2l, bool &m2d)
{
    constexpr int G = 32;
    const Grp<32> g;
    has2 = false; m2 = 0; m2b = 0.0; m2l = 0.0; m2d = false;
    uint32_t kk = h2;
    bool open = true;
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
        const double b = absd(r.a) + dl;
        const double l2 = absd(r.l) + dl;
        const bool strag = act && (b < end);
        const unsigned sa = __ballot_sync(PCC_FULL, strag && !dr), sl = __ballot_sync(PCC_FULL, strag && dr);
        const double ex = __shfl_sync(PCC_FULL, l2, sa ? (__ffs(sa) - 1) : 0);
        if (sa) { extra = ex; has_extra = true; }   // at most one acked per cluster
        acked += __popc(sa);
        lost += __popc(sl);
        if (strag) ring.store_a(kk + g.gl, u2d(PCC_NEG_INF));
        window_argmin(g, act && !strag, b, l2, dr, kk, has2, m2, m2b, m2l, m2d);
        open = (p1 == G) && (p2 == G);
        kk += G;
    }
}

// The in-order part of phases (2) and (3) alone -- the two cursor scans, with the acked latencies staged -- over the
// records below `tail`.  Both scans are prefix scans that stop at the first record they cannot consume, so running
// them first with an earlier tail (the records that existed before this MI's sends) and then handing the cursors and
// counts to consume_mi_warp with the final tail gives exactly what consume_mi_warp alone would: a helper warp does
// this while the env's own warp is still in its send phase.  No boundary analysis, nothing is written to the ring.
template <class Ring>
__device__ __forceinline__ void consume_scan_warp(const Grp<32> &g, double end, double dl, uint32_t h1, uint32_t h2,
                                                  uint32_t tail, Ring &ring, double *buf, int wbuf, ScanOut &out)
{
    constexpr int G = 32;
    int32_t acked = 0, lost = 0;
    for (;;) {
        unsigned bm[PCC_SCAN_W];
        scan_prefetch(g, ring, h1, tail, true);
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
    for (;;) {
        unsigned bm[PCC_SCAN_W], am[PCC_SCAN_W], lm[PCC_SCAN_W];
        double l2[PCC_SCAN_W];
        scan_prefetch(g, ring, h2, tail, true);
#pragma unroll
        for (int w = 0; w < PCC_SCAN_W; w++) {
            bool valid;
            const uint32_t i = h2 + (uint32_t)(w * G);
            const Rec r = load_window(g, ring, i, tail, true, valid);
            const bool dead = is_dead(r.a);
            const bool c1 = ((int32_t)(i + g.gl - h1) < 0) || sgn(r.a);
            const bool early = (absd(r.a) + dl) < end;
            const bool cons = valid && (dead || (c1 && early));
            bm[w] = __ballot_sync(PCC_FULL, cons);
            am[w] = __ballot_sync(PCC_FULL, cons && !dead && !sgn(r.l));
            lm[w] = __ballot_sync(PCC_FULL, cons && !dead && sgn(r.l));
            l2[w] = r.l + dl;
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
                if (pos < wbuf) buf[pos] = l2[w];
            }
            if (!stop) {
                acked += __popc(a_w);
                lost += __popc(lm[w] & lead);
                adv += nl;
            }
            stop = stop || (nl < G);
        }
        h2 += (uint32_t)adv;
        if (stop) break;
    }
    out.h1 = h1; out.h2 = h2; out.acked = acked; out.lost = lost;
}

// Phases (2)-(4) of run_mi for ONE env by the whole warp.  All inputs and outputs warp-uniform.
template <class Ring>
__device__ __forceinline__ void consume_mi_warp(const Grp<32> &g, const ConsumeIn &in, Ring &ring, double *buf,
                                                ConsumeOut &out)
{
    constexpr int G = 32;
    const double end = in.end;
    const uint32_t tail = in.tail;
    uint32_t h1 = in.h1, h2 = in.h2;
    int32_t acked = in.acked0, lost = in.lost0;
    out.has_extra = false;
    out.extra = 0.0;
    out.s_begin = in.s_begin;

    // ---- hop-1 events with a < end ----------------------------------------------------------
    for (;;) {
        unsigned bm[PCC_SCAN_W];
        scan_prefetch(g, ring, h1, tail, true);
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
    // boundary: the record at h1 (if any) is pending with a >= end.  If it is an ACCEPTED packet it
    // is a cluster of its own: no stragglers, and it is the smallest pending hop-1 key.
    bool has1 = false;
    uint32_t m1 = h1; double m1a = 0.0, m1l = 0.0; bool m1d = false;
    if (h1 != tail) {
        const Rec r = ring.load(h1);               // same address in every lane: one broadcast load
        if (!sgn(r.l)) { has1 = true; m1a = r.a; m1l = r.l; }
        else boundary1_slow(ring, h1, tail, end, has1, m1, m1a, m1l, m1d);
    }

    // ---- hop-2 events with b < end; acked latencies staged for np.mean ---------------------------
    for (;;) {
        unsigned bm[PCC_SCAN_W], am[PCC_SCAN_W], lm[PCC_SCAN_W];
        double l2[PCC_SCAN_W];
        scan_prefetch(g, ring, h2, tail, true);
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
                if (pos < in.wbuf) buf[pos] = l2[w];
            }
            if (!stop) {
                acked += __popc(a_w);
                lost += __popc(lm[w] & lead);
                adv += nl;
            }
            stop = stop || (nl < G);
        }
        h2 += (uint32_t)adv;
        if (stop) break;
    }
    out.s_end = h2;
    // boundary: the record at h2 (if any) is not consumable.  It carries a live hop-2 event iff its
    // hop-1 event is consumed (then b >= end).  An accepted packet is a cluster of its own.
    bool has2 = false;
    uint32_t m2 = h2; double m2b = 0.0, m2l = 0.0; bool m2d = false;
    if (h2 != tail) {
        const Rec r = ring.load(h2);
        const bool c1 = ((int32_t)(h2 - h1) < 0) || sgn(r.a);
        if (c1) {
            if (!sgn(r.l)) { has2 = true; m2b = absd(r.a) + in.dl; m2l = r.l + in.dl; }
            else boundary2_slow(ring, h1, h2, tail, in.dl, end, acked, lost, out.extra, out.has_extra,
                                has2, m2, m2b, m2l, m2d);
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
    if (out.has_extra && acked <= in.wbuf && g.gl == 0) buf[acked - 1] = out.extra;
    __syncwarp();
    out.which = which;
    out.h1 = h1; out.h2 = h2;
    out.acked = acked; out.lost = lost;
}

// (np.mean for 128 < n <= staging capacity: means_solo_from_buf in pcc_coop.cuh -- the three sums of the env side by
// side on three 8-lane subgroups, each walking the leaves of numpy's recursion)

// n > wbuf (very rare): streaming re-read of the ring
template <class Ring>
__device__ __noinline__ void mi_means_stream(ConsumeOut co, Ring ring, double dl, double *buf, bool need_increase,
                                             double &avg_lat, double &lat_increase)
{
    const Grp<32> g;
    const int n = co.acked;
    const int half = n / 2;
    lat_increase = 0.0;
    MiOut o;
    o.s_begin = co.s_begin; o.s_end = co.s_end; o.extra = co.extra; o.has_extra = co.has_extra;
    {
        CoopSamples<32, Ring, 4> st(g, ring, buf, o, dl);
        double sum = 0.0;
        sum += coop_pw_sum(g, st, n);
        avg_lat = sum / (double)n;
    }
    if (need_increase) {
        CoopSamples<32, Ring, 4> st(g, ring, buf, o, dl);
        double s1 = 0.0, s2 = 0.0;
        s1 += coop_pw_sum(g, st, half);
        s2 += coop_pw_sum(g, st, n - half);
        lat_increase = s2 / (double)(n - half) - s1 / (double)half;
    }
    __syncwarp();
}

// avg latency (sender_obs.py:119-122) and latency increase (:138-142) of one env's MI, warp-wide
template <class Ring>
__device__ __forceinline__ void mi_means_warp(const Grp<32> &g, const ConsumeOut &co, Ring &ring, double dl,
                                              double *buf, int wbuf, double *smem_buf,
                                              bool need_increase, double &avg_lat, double &lat_increase)
{
    const int n = co.acked;
    avg_lat = 0.0;
    lat_increase = 0.0;
    if (n <= 0) return;
    const int half = n / 2;
    if (n <= PCC_LEAF) {
        // the common case: total, first half and second half are single numpy leaves; evaluate the
        // three concurrently on three 8-lane subgroups (lanes 0-7, 8-15, 16-23)
        const int sub = (int)(g.gl >> 3), j = (int)(g.gl & 7u);
        const int off = (sub == 2) ? half : 0;
        const int cnt = (sub == 0) ? n : (sub == 1) ? half : (sub == 2) ? (n - half) : 0;
        const double *a = buf + off;
        const int nb = cnt - (cnt % 8);
        double r = 0.0;
        if (cnt >= 8) {
            r = a[j];
            for (int k = 8; k < nb; k += 8) r += a[k + j];
        }
        r += __shfl_xor_sync(PCC_FULL, r, 1);
        r += __shfl_xor_sync(PCC_FULL, r, 2);
        r += __shfl_xor_sync(PCC_FULL, r, 4);
        double res;
        if (cnt >= 8) { res = r; for (int k = nb; k < cnt; k++) res += a[k]; }
        else { res = 0.; for (int k = 0; k < cnt; k++) res += a[k]; }
        const double tot = __shfl_sync(PCC_FULL, res, 0);
        const double f1 = __shfl_sync(PCC_FULL, res, 8);
        const double f2 = __shfl_sync(PCC_FULL, res, 16);
        double sum = 0.0;
        sum += tot;
        avg_lat = sum / (double)n;
        if (need_increase && half >= 1) {
            double s1 = 0.0, s2 = 0.0;
            s1 += f1;
            s2 += f2;
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
        __syncwarp();
    } else if (n <= wbuf) {
        means_solo_from_buf(buf, n, need_increase, avg_lat, lat_increase);
    } else {
        mi_means_stream(co, ring, dl, smem_buf, need_increase, avg_lat, lat_increase);
    }
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
    const double y = c.q - (c.t - c.tu);                                // :66-67
    const double cpos = s.d_bw + y;                                     // :82 if 0 < y <= w_full
    const bool pos = y > 0.0;
    const bool fullp = y > s.w_full;                                    // :77-79 (tail_drop_threshold)
    const double w = pos ? y : 0.0;                                     // max(0.0, y)
    const double ll = s.dl + w;                                         // :69-70
    double qn = fullp ? y : cpos;
    qn = pos ? qn : ((0.0 > s.w_full) ? 0.0 : s.d_bw);
    const bool full = pos ? fullp : (0.0 > s.w_full);
    c.q = rdrop ? c.q : qn;                                             // :74-82
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
#ifdef PCC_NO_PHILOX_PIPELINE
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
#else
    // Software pipeline: the Philox block of the NEXT two packets is computed next to the queue recurrence of the
    // current two.  The two are independent dependency chains (about 100 and 2 x 42 cycles of latency), so issued
    // interleaved they overlap instead of adding up; the price is one unused block per env and MI.
    if (!(c.t < end)) return;
    uint32_t a, b, w2, w3;
    rng.block(rng.draws >> 1, a, b, w2, w3);
    for (;;) {      // here: c.t < end, draws even, (a, b, w2, w3) = block draws >> 1
        uint32_t na, nb, nw2, nw3;
        rng.block((rng.draws >> 1) + 1, na, nb, nw2, nw3);
        lane_send_one(c, s, ring, h2, cap, inv_rate, res53(a, b));
        rng.draws++;
        rng.w2 = w2; rng.w3 = w3;                       // the cached second half of the block `draws` now points into
        if (!(c.t < end)) break;
        lane_send_one(c, s, ring, h2, cap, inv_rate, res53(w2, w3));
        rng.draws++;
        if (!(c.t < end)) break;
        a = na; b = nb; w2 = nw2; w3 = nw3;
    }
#endif
}

// Same as lane_send_phase, but the records are staged in shared memory (8 per lane) and written to
// the rings by the whole warp: 8 lanes copy one env's 128-byte line, 4 envs per store instruction.
// (Per-lane 16-byte stores are 32 wavefronts each and hold their registers until the LSU takes
// them -- ncu showed 30 % of the send phase stalled on exactly that.)  Warp-uniform control flow.
#define PCC_STAGE_N 8
struct WarpStage { double2 rec[32][PCC_STAGE_N + 1]; };   // rows padded to 144 B: conflict-free both ways

__device__ __forceinline__ void lane_send_phase_staged(LaneChain &c, const EnvState &s, Rec *ring_base, uint32_t mask,
                                                       PhiloxRng &rng, bool owner, uint32_t h2, uint32_t cap,
                                                       double end, double inv_rate, WarpStage &st)
{
    const unsigned lane = threadIdx.x & 31u;
    bool active = owner && (c.t < end);
    while (__any_sync(PCC_FULL, active)) {
        int nst = 0;
        if (active) {
#pragma unroll 1
            while (nst < PCC_STAGE_N && c.t < end) {
                const double u = rng.next();
                const bool rdrop = u < s.lr;                                        // :73
                const long long yb = __double_as_longlong(c.q - (c.t - c.tu));     // :66-67
                const double w = __longlong_as_double(yb & ~(yb >> 63));            // max(0.0, y)
                const double cc = s.d_bw + w;                                       // :77-79
                const bool full = w > s.w_full;                                      // tail_drop_threshold
                const double ll = s.dl + w;                                         // :69-70
                c.q = rdrop ? c.q : (full ? w : cc);                                // :74-82
                c.tu = rdrop ? c.tu : c.t;
                const bool dropped = rdrop || full;
                const double lsigned = __longlong_as_double(__double_as_longlong(ll) | (dropped ? (long long)PCC_SIGN : 0ll));
                if ((uint32_t)(c.tail + (uint32_t)nst - h2) >= cap) c.ovf = true;   // fatal, reported by the host
                else { st.rec[lane][nst] = make_double2(c.t + ll, lsigned); nst++; }
                c.t = c.t + inv_rate;                                               // :161
                c.sent++;
            }
            active = c.t < end;
        }
        __syncwarp();
        const unsigned have = __ballot_sync(PCC_FULL, nst > 0);
#pragma unroll 1
        for (int r = 0; r < 8; r++) {
            if (((have >> (4 * r)) & 0xFu) == 0u) continue;                         // warp-uniform
            const int j = 4 * r + (int)(lane >> 3);                                 // owner lane of the env served
            const int k = (int)(lane & 7u);
            const int nj = __shfl_sync(PCC_FULL, nst, j);
            const uint32_t tj = __shfl_sync(PCC_FULL, c.tail, j);
            const unsigned long long bj = __shfl_sync(PCC_FULL, (unsigned long long)ring_base, j);
            if (k < nj)
                *reinterpret_cast<double2 *>(reinterpret_cast<Rec *>(bj) + ((tj + (uint32_t)k) & mask)) = st.rec[j][k];
        }
        c.tail += (uint32_t)nst;
        __syncwarp();
    }
}

// Send phase of a warp that owns few envs (cnt <= 8): the chains stay on the owner lanes, but the
// loss draws are produced by ALL lanes -- LPE = 32/cnt (power of two) lanes per env, one Philox
// block (2 draws) per lane and round -- and handed to the owner lane through two ballots.
template <class Ring>
__device__ __forceinline__ void coop_send_phase(LaneChain &c, const EnvState &s, Ring &ring, PhiloxRng &rng,
                                                bool owner, int cnt, uint32_t h2, uint32_t cap, double end,
                                                double inv_rate)
{
    const unsigned lane = threadIdx.x & 31u;
    const int lpe = (cnt <= 1) ? 32 : (cnt <= 2) ? 16 : (cnt <= 4) ? 8 : 4;   // lanes per env
    const int sh = (lpe == 32) ? 5 : (lpe == 16) ? 4 : (lpe == 8) ? 3 : 2;
    const int serve = (int)(lane >> sh);            // env (owner lane) this lane draws for
    const unsigned sub = lane & (unsigned)(lpe - 1);
    const unsigned grp_low = (lpe == 32) ? 0xffffffffu : ((1u << lpe) - 1u);
    bool more = owner && (c.t < end);
    while (__any_sync(PCC_FULL, more)) {
        const unsigned long long sd = __shfl_sync(PCC_FULL, (unsigned long long)rng.seed, serve);
        const unsigned long long dr = __shfl_sync(PCC_FULL, (unsigned long long)rng.draws, serve);
        const double lr = __shfl_sync(PCC_FULL, s.lr, serve);
        uint32_t c0, c1, c2, c3;
        philox_block(sd, (dr >> 1) + sub, c0, c1, c2, c3);
        const unsigned be_all = __ballot_sync(PCC_FULL, res53(c0, c1) < lr);
        const unsigned bo_all = __ballot_sync(PCC_FULL, res53(c2, c3) < lr);
        if (more) {
            const unsigned off = (unsigned)(rng.draws & 1ull);
            const unsigned shl = lane << sh;        // this owner's lanes start at lane * lpe
            uint64_t dm = interleave_bits((be_all >> shl) & grp_low, (bo_all >> shl) & grp_low) >> off;
            const int navail = 2 * lpe - (int)off;
            int k = 0;
#pragma unroll 2
            for (; k < navail; ++k) {
                if (!(c.t < end)) break;
                const bool rdrop = (dm & 1ull) != 0ull;                            // :73
                dm >>= 1;
                const long long yb = __double_as_longlong(c.q - (c.t - c.tu));     // :66-67
                const double w = __longlong_as_double(yb & ~(yb >> 63));            // max(0.0, y)
                const double cc = s.d_bw + w;                                       // :77-79
                const bool full = w > s.w_full;                                      // tail_drop_threshold
                const double ll = s.dl + w;                                         // :69-70
                c.q = rdrop ? c.q : (full ? w : cc);                                // :74-82
                c.tu = rdrop ? c.tu : c.t;
                const bool dropped = rdrop || full;
                Rec r;
                r.a = c.t + ll;
                r.l = __longlong_as_double(__double_as_longlong(ll) | (dropped ? (long long)PCC_SIGN : 0ll));
                if ((uint32_t)(c.tail - h2) >= cap) c.ovf = true;
                else { ring.store(c.tail, r); c.tail++; }
                c.t = c.t + inv_rate;                                               // :161
                c.sent++;
            }
            rng.draws += (uint64_t)k;
            more = (k == navail) && (c.t < end);
        }
    }
    // keep PhiloxRng's cached half consistent for rng.next() (the crossing send)
    if (owner && (rng.draws & 1ull)) { uint32_t a, b; rng.block(rng.draws >> 1, a, b, rng.w2, rng.w3); }
}

}  // namespace pcc
