// pcc_multi_warp.cuh -- the heap-free multi-sender MI of pcc_multi_fast.cuh with one LINK per WARP (BASELINE config 5).
//
// Same state, same ring layout and the same arithmetic as mfast_run_for_dur (which the host twin proves against the
// oracle and the reference's goldens); what changes is who does the work.  All 32 lanes carry the link's scalar state
// in registers and execute the serial parts together (a warp-uniform instruction stream costs what one thread costs);
// the three loops that made the thread-per-link kernel slow are spread over the lanes:
//   * loss draws (network_sim.py:73): one Philox4x32-10 block per lane = 64 draws per round, handed to the queue
//     recurrence as a bit mask, so the generator leaves the per-packet dependency chain;
//   * the queue recurrence itself (:66-84) stays serial in binary64 -- it runs on every lane, lane (k mod 32) keeps
//     the k-th record, and 32 records go to the shared in-flight ring with one coalesced 16-byte store per lane;
//   * hop-1 / hop-2 cursors (:140-154): 32 consecutive records per round, predicate + ballot + ffs instead of a
//     dependent-load loop; per-sender ack / loss counts are popcounts of the ballots, the acked latencies are
//     compacted into the sender's sample array by ballot prefix;
//   * np.mean (sender_obs.py:119-122, 138-142): the 3 * S sums (all / first half / second half per sender) run four
//     at a time on the four 8-lane subgroups (pw_sum_subgroups: numpy's 8 accumulators + xor tree, bit-identical).
// The MI-boundary clusters and the crossing event are a handful of records: warp-uniform scalar code, lane 0 stores.
#pragma once
#include "pcc_multi_fast.cuh"
#include "pcc_coop.cuh"

namespace pcc {

// the env's slice of the heap region reinterpreted as the shared in-flight ring: Rec[cap] then sender ids[cap]
struct DevSidRing {
    Rec *base; uint8_t *sids; uint32_t mask;
    __device__ __forceinline__ uint32_t capacity() const { return mask + 1u; }
    __device__ __forceinline__ Rec load(uint32_t i) const
    {
        const double2 v = *reinterpret_cast<const double2 *>(base + (i & mask));
        Rec r; r.a = v.x; r.l = v.y;
        return r;
    }
    __device__ __forceinline__ double load_a(uint32_t i) const { return base[i & mask].a; }
    __device__ __forceinline__ void store(uint32_t i, Rec r) { *reinterpret_cast<double2 *>(base + (i & mask)) = make_double2(r.a, r.l); }
    __device__ __forceinline__ void store_a(uint32_t i, double a) { base[i & mask].a = a; }
    __device__ __forceinline__ int sid(uint32_t i) const { return sids[i & mask]; }
    __device__ __forceinline__ void set_sid(uint32_t i, int sd) { sids[i & mask] = (uint8_t)sd; }
    __device__ __forceinline__ const void *addr(uint32_t i) const { return base + (i & mask); }
};

// One MI of the link owned by this warp.  Every argument and every local is warp-uniform unless it says "lane".
template <int S>
__device__ __forceinline__ bool mwarp_run_for_dur(MNet &net, MSender *snd, MFast &f, DevSidRing &ring, double *samples,
                                                  int cap_s, uint64_t seed, uint64_t &draws, double dur)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    bool ok = true;
    const double end = net.cur_time + dur;                           // network_sim.py:124
    double ns[S], inv[S];
    int sent[S], acked[S], lost[S], nrtt[S];
#pragma unroll
    for (int i = 0; i < S; i++) {                                    // reset_obs :125-126, :319-324
        ns[i] = f.next_send[i]; inv[i] = 1.0 / snd[i].rate;
        sent[i] = 0; acked[i] = 0; lost[i] = 0; nrtt[i] = 0;
        snd[i].obs_start = net.cur_time;
    }
    const uint32_t cap = ring.capacity();
    uint32_t tail = f.tail, h1 = f.h1, h2 = f.h2;
    double qd = net.qd, t_upd = net.t_upd;
    const double dl = net.dl, lr = net.lr, w_full = net.w_full, d_bw = net.d_bw;

    // the SEND event of the pending timer (it, t): queue update, record (:156-178 -> :66-84)
#define PCC_MW_SEND(it, t, loss, r)                                                           \
    {                                                                                         \
        const double w_ = py_max0(qd - ((t) - t_upd));               /* :170 -> :66-70 */     \
        const double ll_ = dl + w_;                                                           \
        bool dropped_;                                                                        \
        if (loss) dropped_ = true;                                   /* :73 */                \
        else {                                                                                \
            qd = w_; t_upd = (t);                                    /* :75-76 */             \
            if (w_ > w_full) dropped_ = true;                        /* :79 */                \
            else { qd += d_bw; dropped_ = false; }                   /* :82 */                \
        }                                                                                     \
        (r).a = (t) + ll_; (r).l = dropped_ ? negd(ll_) : ll_;       /* :173-175 */           \
        _Pragma("unroll")                                                                     \
        for (int i_ = 0; i_ < S; i_++)                                                        \
            if (i_ == (it)) { sent[i_]++; ns[i_] = (t) + inv[i_]; }  /* :159-161 */           \
    }
#define PCC_MW_NEXT_TIMER(it, t)                                                              \
    int it = 0; double t = ns[0];                                                             \
    _Pragma("unroll")                                                                         \
    for (int i_ = 1; i_ < S; i_++) if (ns[i_] < t) { t = ns[i_]; it = i_; }

    // ---- (1) every send with t < end, timers merged by (time, sender) ------------------------------------
    for (bool more = true; more;) {
        uint32_t c0, c1, c2, c3;
        philox_block(seed, (draws >> 1) + lane, c0, c1, c2, c3);     // lane: draws 2 * (B0 + lane), + 1
        const uint32_t ev = __ballot_sync(PCC_FULL, res53(c0, c1) < lr), od = __ballot_sync(PCC_FULL, res53(c2, c3) < lr);
        const int skip = (int)(draws & 1u);
        uint64_t bits = interleave_bits(ev, od) >> skip;
        const int avail = 64 - skip;
        int k = 0;
        Rec mine; mine.a = 0.0; mine.l = 0.0;                        // lane: record k with k mod 32 == lane
        int mysid = 0;
        while (k < avail) {
            PCC_MW_NEXT_TIMER(it, t);
            if (!(t < end)) { more = false; break; }
            const bool loss = (bits & 1ull) != 0ull;
            bits >>= 1;
            Rec r;
            PCC_MW_SEND(it, t, loss, r);
            if ((unsigned)(k & 31) == lane) { mine = r; mysid = it; }
            k++;
            if ((k & 31) == 0) {                                     // 32 records staged: coalesced copy-out
                const uint32_t p = tail + lane;
                if ((uint32_t)(p - h2) >= cap) ok = false;           // ring overflow: fatal, reported
                else { ring.store(p, mine); ring.set_sid(p, mysid); }
                tail += 32u;
            }
        }
        const unsigned rem = (unsigned)(k & 31);
        if (rem) {
            const uint32_t p = tail + lane;
            if (lane < rem) {
                if ((uint32_t)(p - h2) >= cap) ok = false;
                else { ring.store(p, mine); ring.set_sid(p, mysid); }
            }
            tail += rem;
        }
        draws += (uint64_t)k;
    }
    __syncwarp();

    // ---- (2) hop-1 events with a < end ---------------------------------------------------------------
    for (;;) {
        const uint32_t n = tail - h1;
        if (n == 0u) break;
        const bool valid = lane < n;
        const double a = valid ? ring.load_a(h1 + lane) : 0.0;
        if (lane * 8u + 32u < n && (lane & 3u) == 0u) prefetch_l2_line(ring.addr(h1 + 32u + lane * 8u));
        const unsigned stop = __ballot_sync(PCC_FULL, valid && !sgn(a) && !(a < end));
        if (stop) { h1 += (uint32_t)(__ffs(stop) - 1); break; }
        h1 += n < 32u ? n : 32u;
    }
    bool has1 = false;
    uint32_t m1 = 0; double m1a = 0.0, m1l = 0.0; bool m1d = false; int m1s = 0;
    for (uint32_t k = h1; k != tail; k++) {                          // the boundary cluster
        const Rec r = ring.load(k);
        const bool dr = sgn(r.l);
        if (!sgn(r.a)) {
            if (r.a < end) { if (lane == 0) ring.store_a(k, negd(r.a)); }   // straggler: consumed out of order
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
    __syncwarp();

    // a consumed hop-2 event of sender sd_ outside the cooperative scan (:140-145)
#define PCC_MW_HOP2(sd_, dr_, l2_)                                                            \
    _Pragma("unroll")                                                                         \
    for (int i_ = 0; i_ < S; i_++)                                                            \
        if (i_ == (sd_)) {                                                                    \
            if (dr_) lost[i_]++;                                                              \
            else {                                                                            \
                acked[i_]++;                                                                  \
                if (nrtt[i_] < cap_s) { if (lane == 0) samples[(size_t)i_ * cap_s + nrtt[i_]] = (l2_); } \
                else ok = false;                                                              \
                nrtt[i_]++;                                                                   \
            }                                                                                 \
        }

    // ---- (3) hop-2 events with b < end ---------------------------------------------------------------
    bool at_live = false;
    for (;;) {
        const uint32_t n = tail - h2;
        if (n == 0u) break;
        const uint32_t idx = h2 + lane;
        const bool valid = lane < n;
        Rec r; r.a = 0.0; r.l = 0.0;
        int sd = 0;
        if (valid) { r = ring.load(idx); sd = ring.sid(idx); }
        if (lane * 8u + 32u < n && (lane & 3u) == 0u) prefetch_l2_line(ring.addr(h2 + 32u + lane * 8u));
        const bool live = valid && !is_dead(r.a);
        const bool c1 = ((int32_t)(idx - h1) < 0) || sgn(r.a);
        const bool early = live && c1 && (absd(r.a) + dl < end);     // link 1: latency == dl exactly (N1)
        const unsigned stop = __ballot_sync(PCC_FULL, live && !early);
        const unsigned first = stop ? (unsigned)(__ffs(stop) - 1) : (n < 32u ? n : 32u);
        const bool take = early && lane < first;
        const bool dr = sgn(r.l);
#pragma unroll
        for (int i = 0; i < S; i++) {
            const unsigned ma = __ballot_sync(PCC_FULL, take && sd == i && !dr);
            const unsigned ml = __ballot_sync(PCC_FULL, take && sd == i && dr);
            if (take && sd == i && !dr) {
                const int pos = nrtt[i] + __popc(ma & lt_mask);
                if (pos < cap_s) samples[(size_t)i * cap_s + pos] = absd(r.l) + dl; else ok = false;
            }
            lost[i] += __popc(ml);
            acked[i] += __popc(ma); nrtt[i] += __popc(ma);
        }
        h2 += first;
        if (stop) {
            at_live = __shfl_sync(PCC_FULL, (int)c1, (int)first) != 0;   // the stop is a live event at or after `end`
            break;
        }
    }
    ok = !__any_sync(PCC_FULL, !ok);
    bool has2 = false;
    uint32_t m2 = 0; double m2b = 0.0, m2l = 0.0; bool m2d = false; int m2s = 0;
    if (at_live) {
        for (uint32_t k = h2; k != tail; k++) {
            const Rec r = ring.load(k);
            const bool dr = sgn(r.l);
            if (!is_dead(r.a)) {
                const bool c1 = ((int32_t)(k - h1) < 0) || sgn(r.a);
                if (!c1) break;                                      // later hop-2 events are >= end + dl
                const double b = absd(r.a) + dl;
                const double l2 = absd(r.l) + dl;
                const int sd = ring.sid(k);
                if (b < end) {                                       // straggler
                    PCC_MW_HOP2(sd, dr, l2);
                    if (lane == 0) ring.store_a(k, u2d(PCC_NEG_INF));
                } else if (!has2 || mf_less(b, sd, 0, 2, l2, dr, m2b, m2s, 0, 2, m2l, m2d)) {
                    has2 = true; m2 = k; m2b = b; m2l = l2; m2d = dr; m2s = sd;
                }
            }
            if (!dr) break;
        }
    }

    // ---- (4) the event that crosses `end`: tuple-order minimum of (timer, hop-1, hop-2) -----------------
    PCC_MW_NEXT_TIMER(it, bt0);
    int which = 0;
    double bt = bt0; int bs = it, bty = 1, bh = 0; double bl = 0.0; bool bd = false;
    if (has1 && mf_less(m1a, m1s, 0, 1, m1l, m1d, bt, bs, bty, bh, bl, bd)) {
        which = 1; bt = m1a; bs = m1s; bty = 0; bh = 1; bl = m1l; bd = m1d;
    }
    if (has2 && mf_less(m2b, m2s, 0, 2, m2l, m2d, bt, bs, bty, bh, bl, bd)) which = 2;
    if (which == 0) {
        net.cur_time = bt0;
        uint32_t c0, c1, c2, c3;
        philox_block(seed, draws >> 1, c0, c1, c2, c3);
        const double u = (draws & 1u) ? res53(c2, c3) : res53(c0, c1);
        draws++;
        Rec r;
        PCC_MW_SEND(it, bt0, u < lr, r);
        if ((uint32_t)(tail - h2) >= cap) ok = false;
        else if (lane == 0) { ring.store(tail, r); ring.set_sid(tail, it); }
        tail++;
    } else if (which == 1) {
        net.cur_time = m1a;
        if (m1 == h1) h1++; else if (lane == 0) ring.store_a(m1, negd(m1a));
    } else {
        net.cur_time = m2b;
        PCC_MW_HOP2(m2s, m2d, m2l);
        if (m2 == h2) h2++; else if (lane == 0) ring.store_a(m2, u2d(PCC_NEG_INF));
    }
#undef PCC_MW_SEND
#undef PCC_MW_NEXT_TIMER
#undef PCC_MW_HOP2
    net.qd = qd; net.t_upd = t_upd;
    f.tail = tail; f.h1 = h1; f.h2 = h2;
#pragma unroll
    for (int i = 0; i < S; i++) {
        f.next_send[i] = ns[i];
        snd[i].sent = sent[i]; snd[i].acked = acked[i]; snd[i].lost = lost[i]; snd[i].n_rtt = nrtt[i];
    }
    __syncwarp();
    return ok;
}

// np.mean-derived quantities of all S senders: sum q = 3 * i + {0: all, 1: first half, 2: second half} runs on 8-lane
// subgroup (q mod 4) of round q / 4.
template <int S>
__device__ __forceinline__ void mwarp_means(const MSender *snd, const double *samples, int cap_s, bool need_increase,
                                            double *avg, double *inc)
{
    const int sub = (int)((threadIdx.x & 31u) >> 3);
    double sums[3 * S];
#pragma unroll
    for (int rd = 0; rd < (3 * S + 3) / 4; rd++) {
        const int q = rd * 4 + sub, i = q / 3, part = q - 3 * i;
        int n = 0;
#pragma unroll
        for (int j = 0; j < S; j++) if (j == i) n = snd[j].n_rtt;
        const int half = n / 2;
        const bool on = q < 3 * S && n > 0 && (part == 0 || (need_increase && half >= 1));
        const double *pa = samples + (size_t)(i < S ? i : 0) * cap_s + (part == 2 ? half : 0);
        const int cnt = part == 0 ? n : part == 1 ? half : n - half;
        const double sres = pw_sum_subgroups(on, pa, on ? cnt : 0);
#pragma unroll
        for (int s = 0; s < 4; s++)
            if (rd * 4 + s < 3 * S) sums[rd * 4 + s] = __shfl_sync(PCC_FULL, sres, 8 * s);
    }
#pragma unroll
    for (int i = 0; i < S; i++) {
        const int n = snd[i].n_rtt, half = n / 2;
        avg[i] = 0.0; inc[i] = 0.0;
        if (n > 0) { double s = 0.0; s += sums[3 * i]; avg[i] = s / (double)n; }           // sender_obs.py:119-122
        if (need_increase && half >= 1) {                                                  // :138-142
            double s1 = 0.0, s2 = 0.0;
            s1 += sums[3 * i + 1];
            s2 += sums[3 * i + 2];
            inc[i] = s2 / (double)(n - half) - s1 / (double)half;
        }
    }
}

}  // namespace pcc
