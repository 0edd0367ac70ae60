// pcc_multi_warp.cuh -- the heap-free multi-sender MI of pcc_multi_fast.cuh with one LINK per WARP (BASELINE config 5).
//
// Same state, same ring layout and the same arithmetic as mfast_run_for_dur (which the host twin proves against the
// oracle and the reference's goldens); what changes is who does the work.  All 32 lanes carry the link's scalar state
// in registers and execute the serial parts together (a warp-uniform instruction stream costs what one thread costs);
// the three loops that made the thread-per-link kernel slow are spread over the lanes:
//   * send phase in chunks of 64 packets: the S pacing timers are merged serially (compare, select, one add per
//     packet: network_sim.py:156-161) into shared memory; the loss draws (:73) are one Philox4x32-10 block per lane =
//     64 draws, compared as 53-bit integers and handed on as a bit mask; the queue recurrence (:66-84) -- the only
//     other serial part -- runs in place over the compacted packets that reach the queue (one LDS, two DADD, two
//     DSETP and the selects per packet, as in the single-sender group_send_chunks); all lanes then rebuild the 64
//     records and store them to the shared in-flight ring, coalesced;
//   * hop-1 / hop-2 cursors (:140-154): 4 x 32 consecutive records per round, predicate + ballot + ffs instead of a
//     dependent-load loop; per-sender ack / loss counts are popcounts of the ballots, the acked latencies are
//     compacted into the sender's sample array by ballot prefix;
//   * np.mean (sender_obs.py:119-122, 138-142): the 3 * S sums (all / first half / second half per sender) run four
//     at a time on the four 8-lane subgroups (pw_sum_subgroups: numpy's 8 accumulators + xor tree, bit-identical).
// The MI-boundary clusters and the crossing event are a handful of records: warp-uniform scalar code, lane 0 stores.
#pragma once
#include "pcc_multi_fast.cuh"
#include "pcc_coop.cuh"

namespace pcc {

#define PCC_MW_W 4            // windows of 32 records per scan round (loads issued together)

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
    __device__ __forceinline__ const void *sid_addr(uint32_t i) const { return sids + (i & mask); }
};

// per-warp staging of a chunk of the send phase
struct MwSendSmem {
    double ts[2][64];     // merged send times of the packets of the current / next chunk
    double xs[64 + 8];    // x of the packets that reach the queue, compact (+ slack: the loop reads one ahead)
    double ys[64];        // y = q - x of the same packets
    uint8_t sid[2][64];   // sender of packet k
};

// One MI of the link owned by this warp.  Every argument and every local is warp-uniform unless it says "lane".
template <int S>
__device__ __forceinline__ bool mwarp_run_for_dur(MNet &net, MSender *snd, MFast &f, DevSidRing &ring, double *samples,
                                                  int cap_s, uint64_t seed, uint64_t &draws, double dur, MwSendSmem &sm)
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

    // the records in flight are read by the cursors after the send phase: ask L2 for their lines now (16-byte records,
    // 8 per line; the sender ids, 128 per line), so that DRAM latency runs under the send phase
    {
        const uint32_t inflight = tail - h2;
        for (uint32_t i = lane * 8u; i < inflight; i += 256u) prefetch_l2_line(ring.addr(h2 + i));
        for (uint32_t i = lane * 128u; i < inflight; i += 4096u) prefetch_l2_line(ring.sid_addr(h2 + i));
    }

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

    // ---- (1) every send with t < end, timers merged by (time, sender), in chunks of <= 64 packets -----------
    // The shared queue does not care who sent a packet, so once the chunk's merged send times are known the rest is the
    // single-sender send phase (group_send_chunks in pcc_coop.cuh, stages S1 / S3 / S4 / S5): the queue recurrence runs
    // in place over x_k = t_k - t_(last packet before k that reached the queue), compacted over the packets that were
    // not randomly dropped; records are rebuilt by all lanes afterwards.
    {
        const uint64_t thr = loss_threshold(lr);
        const double k0 = (0.0 > w_full) ? 0.0 : d_bw;               // q' when the queue has drained (w = 0)
        const bool full0 = 0.0 > w_full;
        // M1: one step of the timers' merge (:156-161) into chunk buffer b_, branch-free: `mon_` says the merge is still
        // running; it ends at the first timer >= end or when the chunk is full
#define PCC_MW_MERGE_STEP(b_, mcnt_, mon_, navail_)                                           \
        {                                                                                     \
            PCC_MW_NEXT_TIMER(it, t);                                                         \
            const bool fire = mon_ && (t < end);                                              \
            if (lane == 0 && fire) { sm.ts[b_][mcnt_] = t; sm.sid[b_][mcnt_] = (uint8_t)it; } \
            _Pragma("unroll")                                                                 \
            for (int i = 0; i < S; i++)                                                       \
                if (fire && i == it) ns[i] += inv[i];             /* t == ns[it]: t + 1/rate */ \
            mcnt_ += fire ? 1 : 0;                                                            \
            mon_ = fire && mcnt_ != (navail_);                                                \
        }
        int b = 0, cnt = 0;
        {
            const int navail = 64 - (int)(draws & 1ull);
            bool mon = true;
            while (mon) PCC_MW_MERGE_STEP(0, cnt, mon, navail);
        }
        while (cnt > 0) {
            const unsigned off = (unsigned)(draws & 1ull);
            const bool last = cnt < 64 - (int)off;                   // the merge stopped at a timer >= end
            __syncwarp();
            // S1: bit k of dm = packet k of the chunk is randomly dropped (:73)
            uint32_t c0, c1, c2, c3;
            philox_block(seed, (draws >> 1) + lane, c0, c1, c2, c3); // lane: draws 2 * (B0 + lane), + 1
            const unsigned be = __ballot_sync(PCC_FULL, u53(c0, c1) < thr), bo = __ballot_sync(PCC_FULL, u53(c2, c3) < thr);
            const uint64_t dm = interleave_bits(be, bo) >> off;
            const uint64_t sentm = (cnt >= 64) ? ~0ull : ((1ull << cnt) - 1ull);
            const uint64_t ndm = ~dm & sentm;                        // sent packets that reach the queue
            const int cnt_nd = __popcll(ndm);
            const bool fits = (uint32_t)(tail - h2) + (uint32_t)cnt <= cap;
            if (!fits) ok = false;                                   // ring overflow: fatal, reported
            // S3: x_k (:66-67 with :75-76)
            double tk[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int k = (int)lane + 32 * u;
                tk[u] = (k < cnt) ? sm.ts[b][k] : 0.0;
                const uint64_t before = ndm & ((1ull << k) - 1ull);
                if ((ndm >> k) & 1ull) {
                    const double tuk = before ? sm.ts[b][63 - __clzll((long long)before)] : t_upd;
                    sm.xs[__popcll(before)] = tk[u] - tuk;
                }
            }
            __syncwarp();
            // S4 + M1: y = q - x, q = f(y) over the compact array -- every lane runs it, lane 0 stores y -- and, in the same
            // branch-free loop, the merge of the NEXT chunk: two independent dependency chains in one instruction stream
            double state = qd;
            int mcnt = 0;
            {
                const int navail2 = 64 - (int)((draws + (uint64_t)cnt) & 1ull);
                bool mon = !last;
                double xn = sm.xs[0];
#pragma unroll 4
                for (int k = 0; k < cnt_nd; ++k) {
                    const double y = state - xn;                     // :66-67
                    xn = sm.xs[k + 1];
                    if (lane == 0) sm.ys[k] = y;
                    const double cpos = d_bw + y;                    // :82 if 0 < y <= w_full
                    const bool pos = y > 0.0;
                    const bool fullp = y > w_full;                   // :77-79 (tail_drop_threshold)
                    const double qn = fullp ? y : cpos;
                    state = pos ? qn : k0;
                    PCC_MW_MERGE_STEP(b ^ 1, mcnt, mon, navail2);
                }
                while (mon) PCC_MW_MERGE_STEP(b ^ 1, mcnt, mon, navail2);
            }
            __syncwarp();
            // S5: records (:173-175), sender ids, packets sent per sender (:159-160), the chunk's carry
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int k = (int)lane + 32 * u;
                const int sdk = (k < cnt) ? (int)sm.sid[b][k] : -1;
#pragma unroll
                for (int i = 0; i < S; i++) sent[i] += __popc(__ballot_sync(PCC_FULL, sdk == i));
                if (k < cnt && fits) {
                    const uint64_t before = ndm & ((1ull << k) - 1ull);
                    const int rank = __popcll(before);
                    const bool rdrop = ((dm >> k) & 1ull) != 0ull;
                    double y;
                    if (!rdrop) y = sm.ys[rank];
                    else {
                        // the queue as the last packet that reached it left it (:74: a random drop does not touch it)
                        double qp = qd;
                        if (rank > 0) {
                            const double yp = sm.ys[rank - 1];
                            qp = (yp > 0.0) ? ((yp > w_full) ? yp : d_bw + yp) : k0;
                        }
                        const double tuk = before ? sm.ts[b][63 - __clzll((long long)before)] : t_upd;
                        y = qp - (tk[u] - tuk);
                    }
                    const bool pos = y > 0.0;
                    const double w = pos ? y : 0.0;                  // max(0.0, y)
                    const bool full = pos ? (y > w_full) : full0;
                    const bool dropped = rdrop || full;
                    const double ll = dl + w;                        // :69-70
                    Rec r;
                    r.a = tk[u] + ll;
                    r.l = __longlong_as_double(__double_as_longlong(ll) | (dropped ? (long long)PCC_SIGN : 0ll));
                    ring.store(tail + (uint32_t)k, r);
                    ring.set_sid(tail + (uint32_t)k, sdk);
                }
            }
            if (ndm) t_upd = sm.ts[b][63 - __clzll((long long)ndm)];
            qd = state;
            if (fits) tail += (uint32_t)cnt;
            draws += (uint64_t)cnt;
            cnt = mcnt;
            b ^= 1;
        }
#undef PCC_MW_MERGE_STEP
    }
    __syncwarp();

    // ---- (2) hop-1 events with a < end ---------------------------------------------------------------
    for (bool scanning = true; scanning;) {
        const uint32_t n = tail - h1;
        if (n == 0u) break;
        double a[PCC_MW_W];
#pragma unroll
        for (int w = 0; w < PCC_MW_W; w++) a[w] = (w * 32u + lane < n) ? ring.load_a(h1 + w * 32u + lane) : 0.0;
        if (PCC_MW_W * 32u + lane * 8u < n) prefetch_l2_line(ring.addr(h1 + PCC_MW_W * 32u + lane * 8u));
        uint32_t adv = n < PCC_MW_W * 32u ? n : PCC_MW_W * 32u;
#pragma unroll
        for (int w = PCC_MW_W - 1; w >= 0; w--) {                    // the first stop of the round, if any
            const unsigned stop = __ballot_sync(PCC_FULL, (w * 32u + lane < n) && !sgn(a[w]) && !(a[w] < end));
            if (stop) { adv = w * 32u + (uint32_t)(__ffs(stop) - 1); scanning = false; }
        }
        h1 += adv;
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
    for (bool scanning = true; scanning;) {
        const uint32_t n = tail - h2;
        if (n == 0u) break;
        Rec rr[PCC_MW_W];
        int sdd[PCC_MW_W];
#pragma unroll
        for (int w = 0; w < PCC_MW_W; w++) {
            rr[w].a = 0.0; rr[w].l = 0.0; sdd[w] = 0;
            if (w * 32u + lane < n) { rr[w] = ring.load(h2 + w * 32u + lane); sdd[w] = ring.sid(h2 + w * 32u + lane); }
        }
        if (PCC_MW_W * 32u + lane * 8u < n) prefetch_l2_line(ring.addr(h2 + PCC_MW_W * 32u + lane * 8u));
#pragma unroll
        for (int w = 0; w < PCC_MW_W; w++) {
            if (!scanning || w * 32u >= n) break;
            const uint32_t nw = n - w * 32u, idx = h2 + lane;
            const Rec r = rr[w];
            const int sd = sdd[w];
            const bool live = lane < nw && !is_dead(r.a);
            const bool c1 = ((int32_t)(idx - h1) < 0) || sgn(r.a);
            const bool early = live && c1 && (absd(r.a) + dl < end); // link 1: latency == dl exactly (N1)
            const unsigned stop = __ballot_sync(PCC_FULL, live && !early);
            const unsigned first = stop ? (unsigned)(__ffs(stop) - 1) : (nw < 32u ? nw : 32u);
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
                scanning = false;
            }
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
