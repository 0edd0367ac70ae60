// pcc_core.cuh -- per-env monitor-interval (MI) simulation, shared by the CUDA kernels and by
// the host-compiled "twin" harness in tests/twin (the same functions compiled with g++ so the
// streaming algorithm can be checked against the heap-based oracle without a GPU).
//
// What this replaces (reference file:line, under /root/reference/src):
//   Network.run_for_dur event loop + reward      gym/network_sim.py:123-205
//   Link queue / loss model                      gym/network_sim.py:56-96
//   Sender rate control and MI accounting        gym/network_sim.py:235-281, 298-324
//   SimulatedNetworkEnv.step / reset glue        gym/network_sim.py:406-484
//   MI metrics, history                          common/sender_obs.py:44-73, 110-206
//
// It is NOT a translation of the reference's heap.  With one sender on the path [l0, l1] the
// heap is a merge of three streams: the pacing timer, hop-1 arrivals and hop-2 arrivals (ACK
// or loss notification), and every quantity of a packet is fixed when it is sent:
//     a   = fl(t_send + ll)      time of the hop-1 event       (ll = dl + queue delay seen)
//     b   = fl(a + dl)           time of the hop-2 event
//     rtt = fl(ll + dl)          latency reported when acked
// So an MI is: (1) emit all sends with t < end, appending one 16-byte record (a, +-ll) per
// packet to the env's in-flight ring; (2) advance the hop-1 cursor over records with a < end;
// (3) advance the hop-2 cursor over records with b < end, counting acked / lost; (4) process
// the single event that crosses `end` (the reference's loop tests cur_time BEFORE popping, so
// the first event with t >= end is still processed and becomes the MI's end time).
//
// Exactness.  All arithmetic is binary64 in the reference's operation order (compile with
// -fmad=false / -ffp-contract=off).  Event order only matters for which event crosses `end`.
// Streams are sorted in send order except inside a "cluster": a run of dropped packets plus
// the accepted packet that ends it.  A drop does not add to the queue, so all members of a
// cluster that see a non-empty queue have the mathematically identical arrival time, and
// rounding makes their order arbitrary at the ulp level (one third of the
// reference's heap pops tie exactly).  Packets after an accepted one arrive >= 1/bw later,
// so disorder never crosses a cluster end.  At an MI boundary we therefore look ahead inside
// the boundary cluster only, consume stragglers (members behind the cursor's record whose
// time is < end) and pick the crossing event by the reference's tuple order
//     (time, 'A' < 'S', next_hop, cur_latency, dropped False < True).
// Out-of-order consumption is recorded in the record's spare sign bits:
//     sign(l) = packet was dropped;  sign(a) = hop-1 consumed out of order;  a = -inf = dead
//     (hop-2 consumed out of order).  In-order consumption writes nothing.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define PCC_HD __host__ __device__ __forceinline__
#else
#define PCC_HD inline
#endif

namespace pcc {

// metric ids = position in the reference's SENDER_MI_METRICS (sender_obs.py:193-206)
enum Metric {
    M_SEND_RATE = 0, M_RECV_RATE, M_RECV_DUR, M_SEND_DUR, M_AVG_LATENCY, M_LOSS_RATIO,
    M_ACK_LAT_INFL, M_SENT_LAT_INFL, M_CONN_MIN_LAT, M_LAT_INCREASE, M_LAT_RATIO, M_SEND_RATIO,
    N_METRICS
};
#define PCC_MAX_FEATURES 12

// Simulation constants (network_sim.py:33-54, config.py:17); runtime parameters of a handle.
struct Consts {
    double max_rate;       // MAX_RATE 1000
    double min_rate;       // MIN_RATE 40
    double delta_scale;    // DELTA_SCALE 0.025
    double reward_scale;   // REWARD_SCALE 0.001
    int32_t max_steps;     // MAX_STEPS 400
    int32_t bytes_per_packet; // BYTES_PER_PACKET 1500
};

struct Rec {
    double a;  // hop-1 event time; sign bit / -inf are flags (see header)
    double l;  // latency accumulated at hop 1 (= link-0 latency); sign bit = dropped
};

PCC_HD uint64_t d2u(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PCC_HD double u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
#define PCC_SIGN 0x8000000000000000ull
#define PCC_NEG_INF 0xFFF0000000000000ull
PCC_HD bool sgn(double x) { return (d2u(x) & PCC_SIGN) != 0; }
PCC_HD double absd(double x) { return u2d(d2u(x) & ~PCC_SIGN); }
PCC_HD double negd(double x) { return u2d(d2u(x) | PCC_SIGN); }
PCC_HD bool is_dead(double a) { return d2u(a) == PCC_NEG_INF; }
PCC_HD double py_max0(double x) { return (x > 0.0) ? x : 0.0; }  // Python max(0.0, x)

// ---------------------------------------------------------------------------------------
// RNG streams.  Both give CPython's genrand_res53 double from two 32-bit words.
// ---------------------------------------------------------------------------------------
PCC_HD double res53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// The loss draw `random.random() < lr` (network_sim.py:73) without the int -> double conversions:
// u = res53(a, b) = k * 2^-53 with the integer k = (a >> 5) * 2^26 + (b >> 6), and k * 2^-53 < lr  <=>  k < ceil(lr * 2^53)
// (scaling by 2^53 is exact, k is an integer).  lr <= 0 or NaN: never; lr >= 1: always.
PCC_HD uint64_t loss_threshold(double lr)
{
    if (!(lr > 0.0)) return 0ull;
    if (lr >= 1.0) return 1ull << 53;
    return (uint64_t)ceil(lr * 9007199254740992.0);
}
PCC_HD uint64_t u53(uint32_t a, uint32_t b) { return ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6); }

#define PCC_PHILOX_DOMAIN 0x50434352u
#if defined(__CUDA_ARCH__) && defined(PCC_PHILOX_NOINLINE)
__device__ __noinline__
#else
PCC_HD
#endif
void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1)
{
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// Draw j of the env's stream: block j>>1 of Philox4x32-10 keyed by the env's 64-bit seed,
// words (0,1) for even j, (2,3) for odd j.  The second half of a block is cached.
struct PhiloxRng {
    uint64_t seed, draws;
    uint32_t w2, w3;  // cached second half (valid when draws is odd)
    PCC_HD void init(uint64_t seed_, uint64_t draws_)
    {
        seed = seed_; draws = draws_;
        if (draws & 1u) { uint32_t a, b; block(draws >> 1, a, b, w2, w3); }
    }
    PCC_HD void block(uint64_t blk, uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) const
    {
        uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = PCC_PHILOX_DOMAIN, c3 = 0u;
        philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
        o0 = c0; o1 = c1; o2 = c2; o3 = c3;
    }
    PCC_HD double next()
    {
        double u;
        if (draws & 1u) {
            u = res53(w2, w3);
        } else {
            uint32_t a, b;
            block(draws >> 1, a, b, w2, w3);
            u = res53(a, b);
        }
        draws++;
        return u;
    }
};

// CPython's MT19937 with the state laid out exactly like random.getstate()[1]:
// mt[0..623] = the current (already twisted) block, mt[624] = position of the next output.
struct Mt19937Rng {
    uint32_t *mt;
    PCC_HD void init(uint32_t *state) { mt = state; }
    PCC_HD void regenerate()
    {
        int kk;
        uint32_t y;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    PCC_HD uint32_t next32()
    {
        uint32_t i = mt[624];
        if (i >= 624u) { regenerate(); i = 0; }
        uint32_t y = mt[i];
        mt[624] = i + 1;
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    PCC_HD double next()
    {
        uint32_t a = next32();
        uint32_t b = next32();
        return res53(a, b);
    }
};

// network_sim.py:77-79 drops a packet when fl(extra_delay + queue_delay) > max_queue_delay.  Rounding
// is monotone, so {w >= 0 : fl(d_bw + w) <= max_qd} is an interval [0, w_full] (or empty): the test
// is EXACTLY  w > w_full.  Evaluating it on w instead of on the sum takes one binary64 operation
// off the loop-carried dependency chain of the queue recurrence.
PCC_HD double tail_drop_threshold(double d_bw, double max_qd)
{
    if (d_bw + 0.0 > max_qd) return -1.0;            // even an empty queue is "full"
    // Non-negative doubles are ordered like their bit patterns, so the largest admissible w is found by bisection on
    // the bits: at most 63 steps.  (Walking there one ulp at a time from max_qd - d_bw does not terminate in practice
    // when that difference is 0 or tiny next to d_bw -- a queue of exactly one packet.)
    uint64_t lo = 0ull, hi = 0x7FF0000000000000ull;  // 0.0 is admissible, +inf is not
    while (hi - lo > 1ull) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (d_bw + u2d(mid) <= max_qd) lo = mid; else hi = mid;
    }
    return u2d(lo);
}

// ---------------------------------------------------------------------------------------
// Per-env scalar state (register copy; the kernels keep it structure-of-arrays in HBM).
// ---------------------------------------------------------------------------------------
struct EnvState {
    // link 0 (link 1 has the same dl and never queues: SURVEY.md N1)
    double d_bw;      // 1.0 / bw              (network_sim.py:77, loop invariant)
    double dl;        // propagation delay
    double lr;        // loss rate
    double max_qd;    // queue_size / bw       (network_sim.py:64)
    double w_full;    // largest w with fl(d_bw + w) <= max_qd: tail drop  <=>  w > w_full  (see tail_drop_threshold)
    double qd;        // Link.queue_delay
    double t_upd;     // Link.queue_delay_update_time
    // sender / network
    double rate;
    double next_send; // time of the pending pacing-timer event
    double cur_time;
    double run_dur;
    double conn_min;  // _conn_min_latencies[sender id]; 0.0 = no entry yet
    uint32_t tail;    // ring: next slot to write
    uint32_t h1;      // ring: first record whose hop-1 event has not been consumed in order
    uint32_t h2;      // ring: first record whose hop-2 event has not been consumed in order
    int32_t steps;
};

// Results of one MI, before feature evaluation.
struct MiOut {
    int32_t sent, acked, lost;
    double start, end;        // obs_start_time, cur_time at MI end
    uint32_t s_begin, s_end;  // ring range [s_begin, s_end) consumed in order at hop 2
    double extra;             // the one possible out-of-order acked sample (last in the MI)
    bool has_extra;
    bool overflow;            // ring capacity exceeded (fatal, reported by the host)
};

// ---------------------------------------------------------------------------------------
// One monitor interval.  Ring access is through `Ring`, which provides
//   Rec  load(uint32_t idx);  void store(uint32_t idx, Rec r);  void store_a(uint32_t idx, double a);
//   uint32_t capacity();
// ---------------------------------------------------------------------------------------
// In-order advance of the hop-1 cursor over records with a < bound (flagged records were
// consumed out of order earlier and are skipped).
template <class Ring>
PCC_HD uint32_t scan_hop1(Ring &ring, uint32_t i, uint32_t tail, double bound)
{
    while (i != tail) {
        if ((i & 7u) == 0u) ring.prefetch(i + 8u);   // next 128-byte line of this env's ring
        Rec r = ring.load(i);
        if (!sgn(r.a) && !(r.a < bound)) break;
        i++;
    }
    return i;
}

// In-order advance of the hop-2 cursor over records whose hop-1 event is consumed and whose
// b = fl(a + dl) < bound, counting acked / lost (network_sim.py:140-154).  Returns the new
// cursor; `at_live` tells whether it stopped on such a record with b >= bound.
template <class Ring>
PCC_HD uint32_t scan_hop2(Ring &ring, uint32_t i, uint32_t tail, uint32_t h1, double dl, double bound,
                          int32_t &acked, int32_t &lost, bool &at_live)
{
    at_live = false;
    while (i != tail) {
        if ((i & 7u) == 0u) ring.prefetch(i + 8u);
        Rec r = ring.load(i);
        if (!is_dead(r.a)) {
            bool c1 = ((int32_t)(i - h1) < 0) || sgn(r.a);
            if (!c1) break;
            double b = absd(r.a) + dl;               // link 1: latency == dl exactly (N1)
            if (!(b < bound)) { at_live = true; break; }
            if (sgn(r.l)) lost++; else acked++;      // :141-145
        }
        i++;
    }
    return i;
}

// The smallest pending event of a stream at an MI boundary, in the reference's tuple order.
struct Pending {
    bool has;
    uint32_t idx;      // ring position of the record
    double t, l;       // event time; latency carried by the event (hop 1: ll, hop 2: ll + dl)
    bool dropped;
};

// MI boundary of the hop-1 stream (network_sim.py:129 pops in tuple order): the cluster that starts at the
// cursor -- a run of dropped packets plus the accepted packet that ends it -- is the only place where ring
// order and time order can differ.  Members with a < end are consumed out of order (flagged), the rest
// compete for the crossing event by (a, l, dropped).
template <class Ring>
PCC_HD void boundary_hop1(Ring &ring, uint32_t h1, uint32_t tail, double end, Pending &m)
{
    m.has = false; m.idx = 0; m.t = 0.0; m.l = 0.0; m.dropped = false;
    for (uint32_t k = h1; k != tail; k++) {
        Rec r = ring.load(k);
        bool dr = sgn(r.l);
        if (!sgn(r.a)) {
            if (r.a < end) {
                ring.store_a(k, negd(r.a));          // straggler: consume out of order
            } else {
                double l = absd(r.l);
                bool less = !m.has || r.a < m.t || (r.a == m.t && (l < m.l || (l == m.l && !dr && m.dropped)));
                if (less) { m.has = true; m.idx = k; m.t = r.a; m.l = l; m.dropped = dr; }
            }
        }
        if (!dr) break;                              // accepted packet closes the cluster
    }
}

// MI boundary of the hop-2 stream; call only when the record at h2 carries a live hop-2 event (at_live).
template <class Ring>
PCC_HD void boundary_hop2(Ring &ring, uint32_t h1, uint32_t h2, uint32_t tail, double dl, double end,
                          int32_t &acked, int32_t &lost, double &extra, bool &has_extra, Pending &m)
{
    m.has = false; m.idx = 0; m.t = 0.0; m.l = 0.0; m.dropped = false;
    for (uint32_t k = h2; k != tail; k++) {
        Rec r = ring.load(k);
        bool dr = sgn(r.l);
        if (!is_dead(r.a)) {
            bool c1 = ((int32_t)(k - h1) < 0) || sgn(r.a);
            if (!c1) break;                      // later hop-2 events are >= end + dl
            double b = absd(r.a) + dl;
            double l2 = absd(r.l) + dl;
            if (b < end) {                       // straggler
                if (dr) lost++; else { acked++; extra = l2; has_extra = true; }
                ring.store_a(k, u2d(PCC_NEG_INF));
            } else {
                bool less = !m.has || b < m.t || (b == m.t && (l2 < m.l || (l2 == m.l && !dr && m.dropped)));
                if (less) { m.has = true; m.idx = k; m.t = b; m.l = l2; m.dropped = dr; }
            }
        }
        if (!dr) break;
    }
}

// `prescan`: run the two in-order cursor scans once BEFORE the sends, over the records that already exist, and let
// phases (2) and (3) continue from there.  Both scans are prefix scans that stop at the first record they cannot
// consume and neither writes the ring, so the result is the same as without -- this is the scalar statement of what
// the helper warp of pcc_step_warp_kernel<.., 1> does concurrently with the send phase (consume_scan_warp); the host
// twin checks it against the heap oracle (tests/test_twin.py).
template <class Ring, class Rng>
PCC_HD void run_mi(EnvState &s, Ring &ring, Rng &rng, double dur, MiOut &out, bool prescan = false)
{
    const double end = s.cur_time + dur;            // network_sim.py:124
    const double inv_rate = 1.0 / s.rate;           // :161 (rate is constant within an MI)
    const uint32_t cap = ring.capacity();
    int32_t sent = 0, acked = 0, lost = 0;
    out.start = s.cur_time;                         // reset_obs :319-324
    out.overflow = false;
    out.has_extra = false;
    out.extra = 0.0;
    out.s_begin = s.h2;

    double t = s.next_send;
    double qd = s.qd, t_upd = s.t_upd;
    uint32_t tail = s.tail, h1 = s.h1, h2 = s.h2;
    bool at_live;
#define PCC_SEND_ONE()                                                                     \
    {                                                                                      \
        sent++;                                              /* :159-160 */                \
        double w = py_max0(qd - (t - t_upd));                /* :170 -> :66-70 */          \
        double ll = s.dl + w;                                                              \
        bool dropped;                                                                      \
        double u = rng.next();                               /* :73 */                     \
        if (u < s.lr) {                                                                    \
            dropped = true;                                                                \
        } else {                                                                           \
            qd = w; t_upd = t;                               /* :75-76 */                  \
            if (w > s.w_full) dropped = true;                /* :79, see tail_drop_threshold */ \
            else { qd += s.d_bw; dropped = false; }          /* :82 */                     \
        }                                                                                  \
        Rec r; r.a = t + ll; r.l = dropped ? negd(ll) : ll;  /* :173-175; 0.0 + ll == ll */ \
        ring.store(tail, r); tail++;                                                       \
        t = t + inv_rate;                                    /* :161 */                    \
    }

    if (prescan) {
        h1 = scan_hop1(ring, h1, tail, end);
        h2 = scan_hop2(ring, h2, tail, h1, s.dl, end, acked, lost, at_live);
    }
    // ---- (1) sends with t < end -------------------------------------------------------
    // Ring capacity: records stay in the ring until the END of the MI in which their hop-2
    // event is consumed (the RTT samples are re-read from them for the exact np.mean), so the
    // ring must hold  in-flight at MI start + packets sent in the MI.  The host sizes it from
    // the declared parameter ranges (1.5 * max_rate * max RTT); exceeding it is a fatal,
    // reported error -- the packet is counted but its record is lost.
    // (capacity is measured from the MI-START hop-2 cursor out.s_begin: with `prescan` the cursor has already moved,
    // but the records it passed are re-read for the means at the end of the MI and must not be overwritten)
    while (t < end) {
        if ((uint32_t)(tail - out.s_begin) >= cap) { out.overflow = true; tail--; }
        PCC_SEND_ONE();
    }

    // ---- (2) hop-1 events with a < end ------------------------------------------------
    h1 = scan_hop1(ring, h1, tail, end);
    // boundary cluster: stragglers + the smallest pending key (a, l, dropped)
    Pending m1;
    boundary_hop1(ring, h1, tail, end, m1);

    // ---- (3) hop-2 events with b < end ------------------------------------------------
    h2 = scan_hop2(ring, h2, tail, h1, s.dl, end, acked, lost, at_live);
    out.s_end = h2;
    Pending m2;
    m2.has = false; m2.idx = 0; m2.t = 0.0; m2.l = 0.0; m2.dropped = false;
    if (at_live) boundary_hop2(ring, h1, h2, tail, s.dl, end, acked, lost, out.extra, out.has_extra, m2);
    const bool has1 = m1.has, has2 = m2.has, m2d = m2.dropped;
    const uint32_t m1i = m1.idx, m2i = m2.idx;
    const double m1a = m1.t, m2b = m2.t, m2l = m2.l;

    // ---- (4) the event that crosses `end` ----------------------------------------------
    // candidates: pacing timer (t, 'S'), hop-1 (m1a, 'A', hop 1), hop-2 (m2b, 'A', hop 2)
    int which;  // 0 = send, 1 = hop-1, 2 = hop-2
    if (has1 && (!has2 || m1a <= m2b)) which = (m1a <= t) ? 1 : 0;
    else if (has2) which = (m2b <= t) ? 2 : 0;
    else which = 0;
    if (which == 0) {
        s.cur_time = t;
        if ((uint32_t)(tail - out.s_begin) >= cap) { out.overflow = true; tail--; }
        PCC_SEND_ONE();
    } else if (which == 1) {
        s.cur_time = m1a;
        if (m1i == h1) h1++; else ring.store_a(m1i, negd(m1a));
    } else {
        s.cur_time = m2b;
        if (m2d) lost++; else { acked++; out.extra = m2l; out.has_extra = true; }
        if (m2i == h2) h2++; else ring.store_a(m2i, u2d(PCC_NEG_INF));
    }
#undef PCC_SEND_ONE
    s.next_send = t;
    s.qd = qd; s.t_upd = t_upd;
    s.tail = tail; s.h1 = h1; s.h2 = h2;
    out.sent = sent; out.acked = acked; out.lost = lost;
    out.end = s.cur_time;
}

// ---------------------------------------------------------------------------------------
// numpy's pairwise summation over the MI's acked-latency samples, streamed from the ring.
// ---------------------------------------------------------------------------------------
// Yields the samples of an MI in order: acked records in [s_begin, s_end), then `extra`.
template <class Ring>
struct SampleReader {
    Ring &ring; uint32_t i, end; double dl; double extra;
    PCC_HD SampleReader(Ring &r, const MiOut &o, double dl_) : ring(r), i(o.s_begin), end(o.s_end), dl(dl_), extra(o.extra) {}
    PCC_HD double next()
    {
        while (i != end) {
            if ((i & 7u) == 0u) ring.prefetch(i + 8u);
            Rec r = ring.load(i);
            i++;
            if (!is_dead(r.a) && !sgn(r.l)) return r.l + dl;   // rtt = fl(ll + dl)
        }
        return extra;
    }
};

// One leaf of numpy's DOUBLE_pairwise_sum (n <= 128), consuming n samples from the reader.
template <class Reader>
PCC_HD double pw_leaf(Reader &rd, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int k = 0; k < n; k++) res += rd.next();
        return res;
    }
    double r0 = rd.next(), r1 = rd.next(), r2 = rd.next(), r3 = rd.next();
    double r4 = rd.next(), r5 = rd.next(), r6 = rd.next(), r7 = rd.next();
    int k;
    for (k = 8; k < n - (n % 8); k += 8) {
        r0 += rd.next(); r1 += rd.next(); r2 += rd.next(); r3 += rd.next();
        r4 += rd.next(); r5 += rd.next(); r6 += rd.next(); r7 += rd.next();
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; k < n; k++) res += rd.next();
    return res;
}

// Full pairwise sum of the next n samples: iterative post-order walk of numpy's recursion
// (n > 128: n2 = n/2 rounded down to a multiple of 8; sum(left n2) + sum(right n - n2)).
#define PCC_PW_STACK 24
template <class Reader>
PCC_HD double pw_sum(Reader &rd, int n)
{
    int right_n[PCC_PW_STACK];
    double left_sum[PCC_PW_STACK];
    bool have_left[PCC_PW_STACK];
    int sp = 0;
    int cur = n;
    for (;;) {
        while (cur > 128) {
            int n2 = cur / 2;
            n2 -= n2 % 8;
            right_n[sp] = cur - n2; have_left[sp] = false; sp++;
            cur = n2;
        }
        double res = pw_leaf(rd, cur);
        for (;;) {
            if (sp == 0) return res;
            if (!have_left[sp - 1]) {
                left_sum[sp - 1] = res; have_left[sp - 1] = true;
                cur = right_n[sp - 1];
                break;  // descend into the right child
            }
            res = left_sum[sp - 1] + res;
            sp--;
        }
    }
}

// The same pairwise sum in PUSH form: the number of samples n is known in advance, the samples arrive one at a
// time (the lane-per-env kernels read them tile by tile and cannot hand a pull-style reader to pw_sum).  The walk of
// numpy's recursion is pw_sum's, turned inside out: `descend` opens the leftmost leaf of a subtree, `push` feeds the
// current leaf (8 accumulators r[k % 8], the combination tree when the last full block of 8 is in, then the tail
// sequentially), `leaf_done` folds finished subtrees.  Storage is a policy so that the scalar state stays in registers
// on the GPU: `Acc` holds the 8 accumulators (double get(int), void set(int, double): a shared-memory column per
// lane), `Stk` the recursion stack (int &rn(int), double &ls(int): caller-owned arrays, touched once per leaf).
struct PwHostStack {
    int right_n[PCC_PW_STACK]; double left_sum[PCC_PW_STACK];
    PCC_HD int &rn(int i) { return right_n[i]; }
    PCC_HD double &ls(int i) { return left_sum[i]; }
};
template <class Acc, class Stk = PwHostStack>
struct PwStream {
    Acc acc;
    Stk stk;
    int cur, k, nb;          // current leaf: size, samples consumed, size rounded down to a multiple of 8 (0 if < 8)
    double res, total;
    int sp;
    uint32_t have_left;      // bit i: stack level i already holds its left sum
    bool done;

    PCC_HD void descend(int m)
    {
        while (m > 128) {
            int n2 = m / 2;
            n2 -= n2 % 8;
            stk.rn(sp) = m - n2; have_left &= ~(1u << sp); sp++;
            m = n2;
        }
        cur = m; k = 0; nb = (m >= 8) ? m - (m % 8) : 0; res = 0.;
    }
    PCC_HD void begin(int n)
    {
        sp = 0; have_left = 0u; total = 0.0; res = 0.; cur = 0; k = 0; nb = 0;
        done = n <= 0;
        if (n > 0) descend(n);
    }
    PCC_HD void leaf_done()
    {
        double r = res;
        for (;;) {
            if (sp == 0) { total = r; done = true; return; }
            if (!((have_left >> (sp - 1)) & 1u)) {
                stk.ls(sp - 1) = r; have_left |= 1u << (sp - 1);
                descend(stk.rn(sp - 1));             // the right sibling's leftmost leaf
                return;
            }
            r = stk.ls(sp - 1) + r;
            sp--;
        }
    }
    PCC_HD void push(double x)
    {
        if (k < nb) {
            const int j = k & 7;
            if (k < 8) acc.set(j, x); else acc.set(j, acc.get(j) + x);
            k++;
            if (k == nb)
                res = ((acc.get(0) + acc.get(1)) + (acc.get(2) + acc.get(3))) +
                      ((acc.get(4) + acc.get(5)) + (acc.get(6) + acc.get(7)));
        } else { res += x; k++; }
        if (k == cur) leaf_done();
    }
    // np.mean of the n samples pushed: (0.0 + pairwise) / n
    PCC_HD double mean(int n) const { double s = 0.0; s += total; return s / (double)n; }
};

// np.mean of the next n samples: (0.0 + pairwise) / n
template <class Reader>
PCC_HD double np_mean_stream(Reader &rd, int n)
{
    double s = 0.0;
    s += pw_sum(rd, n);
    return s / (double)n;
}

// ---------------------------------------------------------------------------------------
// MI metrics (sender_obs.py:110-191) and reward (network_sim.py:180-205)
// ---------------------------------------------------------------------------------------
struct MiStats {
    double dur, send_rate, recv_rate, avg_lat, loss_ratio, lat_increase, lat_infl;
    double conn_min, lat_ratio, send_ratio, reward;
};

// Everything of mi_stats that is scalar: takes the two np.mean-derived quantities as inputs.
PCC_HD void mi_stats_finish(const MiOut &o, const Consts &c, double avg_lat, double lat_increase,
                            double &conn_min_state, bool update_conn_min, MiStats &st)
{
    const double bytes_sent = (double)((int64_t)o.sent * c.bytes_per_packet);
    const double bytes_acked_m1 = (double)((int64_t)o.acked * c.bytes_per_packet - c.bytes_per_packet);
    st.dur = o.end - o.start;                                                     // :116-117
    st.send_rate = (st.dur > 0.0) ? 8.0 * bytes_sent / st.dur : 0.0;              // :124-128
    st.recv_rate = (st.dur > 0.0) ? 8.0 * bytes_acked_m1 / st.dur : 0.0;          // :110-114
    st.avg_lat = avg_lat;                                                         // :119-122
    // bytes_lost / (bytes_lost + bytes_acked): both are exact ints * 1500; int/int true
    // division in Python is correctly rounded, as is this double division of exact values.
    st.loss_ratio = (o.lost + o.acked > 0)
        ? (double)((int64_t)o.lost * c.bytes_per_packet) /
          (double)(((int64_t)o.lost + o.acked) * c.bytes_per_packet) : 0.0;       // :133-136
    st.lat_increase = lat_increase;                                               // :138-142
    st.lat_infl = (st.dur > 0.0) ? st.lat_increase / st.dur : 0.0;                // :144-156
    // conn min latency (:158-176); the dict entry exists iff conn_min_state > 0
    double cm;
    if (conn_min_state > 0.0) {
        if (st.avg_lat == 0.0) cm = conn_min_state;
        else if (st.avg_lat < conn_min_state) { cm = st.avg_lat; if (update_conn_min) conn_min_state = cm; }
        else cm = conn_min_state;
    } else {
        if (st.avg_lat > 0.0) { cm = st.avg_lat; if (update_conn_min) conn_min_state = cm; }
        else cm = 0.0;
    }
    st.conn_min = cm;
    st.lat_ratio = (cm > 0.0) ? st.avg_lat / cm : 1.0;                            // :186-191
    st.send_ratio = (st.recv_rate > 0.0 && st.send_rate < 1000.0 * st.recv_rate)
        ? st.send_rate / st.recv_rate : 1.0;                                      // :179-184
    // reward :194,205 -- ((10*thr)/12000 - 1e3*lat) - 2e3*loss, then * REWARD_SCALE
    double rw = 10.0 * st.recv_rate / (double)(8 * c.bytes_per_packet) - 1e3 * st.avg_lat - 2e3 * st.loss_ratio;
    st.reward = rw * c.reward_scale;
}

template <class Ring>
PCC_HD void mi_stats(const MiOut &o, Ring &ring, double dl, const Consts &c, bool need_increase,
                     double &conn_min_state, bool update_conn_min, MiStats &st)
{
    const int n = o.acked;
    double avg_lat = 0.0, lat_increase = 0.0;
    if (n > 0) { SampleReader<Ring> rd(ring, o, dl); avg_lat = np_mean_stream(rd, n); }
    if (need_increase) {
        int half = n / 2;                                                         // :138-142
        if (half >= 1) {
            SampleReader<Ring> rd(ring, o, dl);
            double first = np_mean_stream(rd, half);
            double second = np_mean_stream(rd, n - half);
            lat_increase = second - first;
        }
    }
    mi_stats_finish(o, c, avg_lat, lat_increase, conn_min_state, update_conn_min, st);
}

PCC_HD double metric_value(const MiStats &st, int id)
{
    switch (id) {
    case M_SEND_RATE: return st.send_rate / 1e7;      // scale 1e7, sender_obs.py:194-195
    case M_RECV_RATE: return st.recv_rate / 1e7;
    case M_RECV_DUR: case M_SEND_DUR: return st.dur;
    case M_AVG_LATENCY: return st.avg_lat;
    case M_LOSS_RATIO: return st.loss_ratio;
    case M_ACK_LAT_INFL: case M_SENT_LAT_INFL: return st.lat_infl;
    case M_CONN_MIN_LAT: return st.conn_min;
    case M_LAT_INCREASE: return st.lat_increase;
    case M_LAT_RATIO: return st.lat_ratio;
    default: return st.send_ratio;
    }
}

// History row of an empty MI (SenderHistory.__init__, sender_obs.py:57-62): every metric 0
// except latency ratio and send ratio, which are 1.
PCC_HD double metric_empty(int id) { return (id == M_LAT_RATIO || id == M_SEND_RATIO) ? 1.0 : 0.0; }

PCC_HD bool features_need_increase(const int *ids, int n)
{
    for (int k = 0; k < n; k++)
        if (ids[k] == M_ACK_LAT_INFL || ids[k] == M_SENT_LAT_INFL || ids[k] == M_LAT_INCREASE) return true;
    return false;
}

// Sender.apply_rate_delta + set_rate (network_sim.py:235-241, 275-281)
PCC_HD double apply_rate_delta(double rate, double action, const Consts &c)
{
    double delta = action * c.delta_scale;
    double nr = (delta >= 0.0) ? rate * (1.0 + delta) : rate / (1.0 - delta);
    if (nr > c.max_rate) nr = c.max_rate;
    if (nr < c.min_rate) nr = c.min_rate;
    return nr;
}

// reset(): create_new_links_and_senders + Network() + two discarded warm-up MIs
// (network_sim.py:454-484).  Link parameters and the start rate are inputs.
template <class Ring, class Rng>
PCC_HD bool reset_env(EnvState &s, Ring &ring, Rng &rng, double bw, double dl, double lr,
                      int64_t queue_size, double start_rate)
{
    s.d_bw = 1.0 / bw;
    s.dl = dl;
    s.lr = lr;
    s.max_qd = (double)queue_size / bw;
    s.w_full = tail_drop_threshold(s.d_bw, s.max_qd);
    s.qd = 0.0; s.t_upd = 0.0;
    s.rate = start_rate;
    s.cur_time = 0.0;
    s.next_send = 1.0 / start_rate;     // queue_initial_packets :107-111
    s.run_dur = 3 * dl;                 // :467
    s.conn_min = 0.0;                   // fresh sender id
    s.h1 = s.tail; s.h2 = s.tail;   // drop everything in flight; ring positions keep counting
    s.steps = 0;
    MiOut o;
    run_mi(s, ring, rng, s.run_dur, o);  // :478
    bool ovf = o.overflow;
    run_mi(s, ring, rng, s.run_dur, o);  // :479
    return ovf || o.overflow;
}

}  // namespace pcc

namespace pcc {

// step(): apply the action, run one MI, evaluate the metrics, update run_dur / steps
// (network_sim.py:406-444).  History and obs are handled by the caller (layout differs
// between the CUDA kernels and the host twin).
struct StepOut {
    MiOut mi;
    MiStats st;
    bool done;
};

template <class Ring, class Rng>
PCC_HD void step_env(EnvState &s, Ring &ring, Rng &rng, double action, const Consts &c,
                     bool need_increase, StepOut &o, bool prescan = false)
{
    s.rate = apply_rate_delta(s.rate, action, c);                    // :412
    run_mi(s, ring, rng, s.run_dur, o.mi, prescan);                  // :416
    mi_stats(o.mi, ring, s.dl, c, need_increase, s.conn_min, true, o.st);
    s.steps += 1;                                                    // :419
    if (o.st.avg_lat > 0.0) s.run_dur = 0.5 * o.st.avg_lat;          // :437-438
    o.done = s.steps >= c.max_steps;                                 // :444
}

}  // namespace pcc
