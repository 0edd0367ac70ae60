// pcc_coop.cuh -- group-cooperative monitor interval: G lanes of a warp work on ONE env.
//
// Same algorithm and the same arithmetic as pcc_core.cuh::run_mi / mi_stats (which the host
// twin proves against the oracle); what changes is who does the work:
//   * queue recurrence (network_sim.py:72-84, inherently serial in binary64): lane 0 of the
//     group, with the per-packet loss draws precomputed by all G lanes (one Philox4x32-10
//     block per lane = 2G draws per chunk) and handed over as two ballot masks;
//   * ring scans (hop-1 / hop-2 cursors): G consecutive 16-byte records per load = full
//     128-byte lines for G = 8, predicate + ballot + ffs instead of a dependent-load loop;
//   * MI-boundary cluster analysis (stragglers, tuple-order minimum): ballots and
//     shuffle min-reductions over the window;
//   * np.mean: samples are compacted (ballot + popc prefix) into a shared-memory staging
//     buffer, numpy's 8 accumulators live in 8 lanes, combined with a 3-level xor-shuffle
//     tree -- bit-identical to DOUBLE_pairwise_sum.
#pragma once
#include "pcc_core.cuh"

namespace pcc {

#define PCC_INF_BITS 0x7FF0000000000000ull

template <int G>
struct Grp {
    static constexpr unsigned LOW = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    unsigned gl, gbase, gmask;
    __device__ __forceinline__ Grp()
    {
        const unsigned lane = threadIdx.x & 31u;
        gl = lane & (unsigned)(G - 1);
        gbase = lane - gl;
        gmask = LOW << gbase;
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(gmask, p) >> gbase) & LOW; }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(gmask, v, (int)gbase + src); }
    __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(gmask, v, (int)gbase + src); }
    __device__ __forceinline__ unsigned bcast(unsigned v, int src) const { return __shfl_sync(gmask, v, (int)gbase + src); }
    __device__ __forceinline__ double min(double v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(gmask, v, o));
        return v;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
    static __device__ __forceinline__ int lead_ones(unsigned bm) { return (bm == LOW) ? G : (__ffs(~bm) - 1); }
};

__device__ __forceinline__ void philox_block(uint64_t seed, uint64_t blk, uint32_t &c0, uint32_t &c1,
                                             uint32_t &c2, uint32_t &c3)
{
    c0 = (uint32_t)blk; c1 = (uint32_t)(blk >> 32); c2 = PCC_PHILOX_DOMAIN; c3 = 0u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// window load: lane gl gets record i + gl if it lies before `lim`, else a neutral dummy
// (a = +inf: never consumable, not flagged; l = +1: not dropped)
template <int G, class Ring>
__device__ __forceinline__ Rec load_window(const Grp<G> &g, Ring &ring, uint32_t i, uint32_t lim, bool &valid)
{
    const uint32_t idx = i + g.gl;
    valid = (int32_t)(idx - lim) < 0;
    Rec r;
    if (valid) r = ring.load(idx);
    else { r.a = u2d(PCC_INF_BITS); r.l = 1.0; }
    return r;
}

// argmin of (key1, key2, dropped False<True) over candidate lanes, merged into the running
// minimum (has, m_idx, m_k1, m_k2, m_d).  All outputs are group-uniform.
template <int G>
__device__ __forceinline__ void window_argmin(const Grp<G> &g, bool cand, double k1, double k2, bool dr,
                                              uint32_t base_idx, bool &has, uint32_t &m_idx, double &m_k1,
                                              double &m_k2, bool &m_d)
{
    const double inf = u2d(PCC_INF_BITS);
    const double k1min = g.min(cand ? k1 : inf);
    const bool c2 = cand && k1 == k1min;
    const double k2min = g.min(c2 ? k2 : inf);
    const bool c3 = c2 && k2 == k2min;
    const unsigned nd3 = g.ballot(c3 && !dr), all3 = g.ballot(c3);
    if (all3) {
        const bool pd = (nd3 == 0u);
        const unsigned pick = pd ? all3 : nd3;
        const uint32_t pidx = base_idx + (uint32_t)(__ffs(pick) - 1);
        const bool less = !has || k1min < m_k1 || (k1min == m_k1 && (k2min < m_k2 || (k2min == m_k2 && !pd && m_d)));
        if (less) { has = true; m_idx = pidx; m_k1 = k1min; m_k2 = k2min; m_d = pd; }
    }
}

// One monitor interval, cooperatively.  Every lane of the group holds the same EnvState copy
// on entry and on exit.  `draws` is the env's Philox draw counter.
template <int G, class Ring>
__device__ __forceinline__ void run_mi_coop(const Grp<G> &g, EnvState &s, Ring &ring, uint64_t seed,
                                            uint64_t &draws, double dur, MiOut &out)
{
    const double end = s.cur_time + dur;            // network_sim.py:124
    const double inv_rate = 1.0 / s.rate;           // :161
    const uint32_t cap = ring.capacity();
    int32_t sent = 0, acked = 0, lost = 0;
    out.start = s.cur_time;
    out.overflow = false;
    out.has_extra = false;
    out.extra = 0.0;
    out.s_begin = s.h2;
    double t = s.next_send, qd = s.qd, t_upd = s.t_upd;
    uint32_t tail = s.tail, h1 = s.h1, h2 = s.h2;
    bool ovf = false;

    // ---- (1) sends with t < end: chain on lane 0, loss draws from all lanes ---------------
    bool more = t < end;
    while (more) {
        const unsigned off = (unsigned)(draws & 1ull);
        uint32_t c0, c1, c2, c3;
        philox_block(seed, (draws >> 1) + g.gl, c0, c1, c2, c3);
        const unsigned m_even = g.ballot(res53(c0, c1) < s.lr);   // draw 2*blk   of lane's block
        const unsigned m_odd = g.ballot(res53(c2, c3) < s.lr);    // draw 2*blk+1
        const int navail = 2 * G - (int)off;
        int k = 0;
        if (g.gl == 0) {
            for (; k < navail && t < end; ++k) {
                const unsigned d = off + (unsigned)k;
                const bool rdrop = ((((d & 1u) ? m_odd : m_even) >> (d >> 1)) & 1u) != 0u;   // :73
                const double w = py_max0(qd - (t - t_upd));          // :170 -> :66-70
                const double ll = s.dl + w;
                bool dropped = rdrop;
                if (!rdrop) {
                    qd = w; t_upd = t;                               // :75-76
                    if (s.d_bw + qd > s.max_qd) dropped = true;      // :79
                    else qd += s.d_bw;                               // :82
                }
                Rec r; r.a = t + ll; r.l = dropped ? negd(ll) : ll;
                if ((uint32_t)(tail - h2) >= cap) ovf = true;
                else { ring.store(tail, r); tail++; }
                t = t + inv_rate;                                    // :161
            }
        }
        const int packed = g.bcast((int)(k | ((k == navail && t < end) ? 0x100 : 0)), 0);
        k = packed & 0xff;
        more = (packed & 0x100) != 0;
        draws += (uint64_t)k;
        sent += k;
    }
    // make the chain lane's state the group's state
    t = g.bcast(t, 0); qd = g.bcast(qd, 0); t_upd = g.bcast(t_upd, 0);
    tail = g.bcast(tail, 0);
    ovf = g.bcast((int)ovf, 0) != 0;
    g.sync();   // the chain lane's record stores are read by every lane below

    // ---- (2) hop-1 events with a < end ------------------------------------------------------
    {
        uint32_t i = h1;
        for (;;) {
            bool valid;
            const Rec r = load_window(g, ring, i, tail, valid);
            const bool cons = valid && (sgn(r.a) || r.a < end);
            const int nlead = Grp<G>::lead_ones(g.ballot(cons));
            i += (uint32_t)nlead;
            if (nlead < G) break;
        }
        h1 = i;
    }
    bool has1 = false;
    uint32_t m1 = 0; double m1a = 0.0, m1l = 0.0; bool m1d = false;
    {
        uint32_t kk = h1;
        bool open = (kk != tail);
        while (open) {
            bool valid;
            const Rec r = load_window(g, ring, kk, tail, valid);
            const bool dr = sgn(r.l);
            const unsigned validm = g.ballot(valid);
            const unsigned ndm = g.ballot(valid && !dr);
            const int pend = ndm ? (__ffs(ndm) - 1) : G;   // accepted record closes the cluster
            const bool pending = valid && (int)g.gl <= pend && !sgn(r.a);
            const bool strag = pending && (r.a < end);
            if (strag) ring.store_a(kk + g.gl, negd(r.a));
            window_argmin(g, pending && !strag, r.a, absd(r.l), dr, kk, has1, m1, m1a, m1l, m1d);
            open = (pend == G) && (validm == Grp<G>::LOW) && ((uint32_t)(kk + G) != tail);
            kk += G;
        }
        g.sync();   // straggler flags are read by the hop-2 scan
    }

    // ---- (3) hop-2 events with b < end ------------------------------------------------------
    bool at_live = false;
    {
        uint32_t i = h2;
        for (;;) {
            bool valid;
            const Rec r = load_window(g, ring, i, tail, valid);
            const bool dead = is_dead(r.a);
            const bool c1 = ((int32_t)(i + g.gl - h1) < 0) || sgn(r.a);
            const double b = absd(r.a) + s.dl;               // :149-154, link 1 latency == dl
            const bool cons = valid && (dead || (c1 && b < end));
            const int nlead = Grp<G>::lead_ones(g.ballot(cons));
            const unsigned leadmask = (nlead >= 32) ? 0xffffffffu : ((1u << nlead) - 1u);
            acked += __popc(g.ballot(cons && !dead && !sgn(r.l)) & leadmask);   // :144-145
            lost += __popc(g.ballot(cons && !dead && sgn(r.l)) & leadmask);     // :141-142
            i += (uint32_t)nlead;
            if (nlead < G) {
                const unsigned lv = g.ballot(valid && !dead && c1 && !(b < end));
                at_live = ((lv >> nlead) & 1u) != 0u;
                break;
            }
        }
        h2 = i;
    }
    out.s_end = h2;
    bool has2 = false;
    uint32_t m2 = 0; double m2b = 0.0, m2l = 0.0; bool m2d = false;
    if (at_live) {
        uint32_t kk = h2;
        bool open = true;
        while (open) {
            bool valid;
            const Rec r = load_window(g, ring, kk, tail, valid);
            const bool dr = sgn(r.l);
            const bool dead = is_dead(r.a);
            const bool c1 = ((int32_t)(kk + g.gl - h1) < 0) || sgn(r.a);
            const unsigned x1 = g.ballot(!valid || (!dead && !c1));   // stop BEFORE this record
            const unsigned x2 = g.ballot(valid && !dr);               // stop AFTER this record
            const int p1 = x1 ? (__ffs(x1) - 1) : G;
            const int p2 = x2 ? (__ffs(x2) - 1) : G;
            const bool act = (int)g.gl < p1 && (int)g.gl <= p2 && !dead;
            const double b = absd(r.a) + s.dl;
            const double l2 = absd(r.l) + s.dl;
            const bool strag = act && (b < end);
            const unsigned sa = g.ballot(strag && !dr), sl = g.ballot(strag && dr);
            acked += __popc(sa);
            lost += __popc(sl);
            if (sa) { out.extra = g.bcast(l2, __ffs(sa) - 1); out.has_extra = true; }
            if (strag) ring.store_a(kk + g.gl, u2d(PCC_NEG_INF));
            window_argmin(g, act && !strag, b, l2, dr, kk, has2, m2, m2b, m2l, m2d);
            open = (p1 == G) && (p2 == G);
            kk += G;
        }
    }

    // ---- (4) the event that crosses `end` (group-uniform) -----------------------------------
    int which;
    if (has1 && (!has2 || m1a <= m2b)) which = (m1a <= t) ? 1 : 0;
    else if (has2) which = (m2b <= t) ? 2 : 0;
    else which = 0;
    if (which == 0) {
        s.cur_time = t;
        uint32_t c0, c1, c2, c3;
        philox_block(seed, draws >> 1, c0, c1, c2, c3);
        const double u = (draws & 1ull) ? res53(c2, c3) : res53(c0, c1);
        draws++;
        sent++;
        const double w = py_max0(qd - (t - t_upd));
        const double ll = s.dl + w;
        bool dropped;
        if (u < s.lr) dropped = true;
        else {
            qd = w; t_upd = t;
            if (s.d_bw + qd > s.max_qd) dropped = true;
            else { qd += s.d_bw; dropped = false; }
        }
        Rec r; r.a = t + ll; r.l = dropped ? negd(ll) : ll;
        if ((uint32_t)(tail - h2) >= cap) ovf = true;
        else { if (g.gl == 0) ring.store(tail, r); tail++; }
        t = t + inv_rate;
    } else if (which == 1) {
        s.cur_time = m1a;
        if (m1 == h1) h1++;
        else if (g.gl == 0) ring.store_a(m1, negd(m1a));
    } else {
        s.cur_time = m2b;
        if (m2d) lost++; else { acked++; out.extra = m2l; out.has_extra = true; }
        if (m2 == h2) h2++;
        else if (g.gl == 0) ring.store_a(m2, u2d(PCC_NEG_INF));
    }
    g.sync();   // order this MI's flag stores before any later window load by other lanes
    s.next_send = t;
    s.qd = qd; s.t_upd = t_upd;
    s.tail = tail; s.h1 = h1; s.h2 = h2;
    out.sent = sent; out.acked = acked; out.lost = lost;
    out.end = s.cur_time;
    out.overflow = ovf;
}

// ---------------------------------------------------------------------------------------------
// np.mean over the MI's samples, cooperatively.  Samples = acked, not-dead records of
// [s_begin, s_end) in ring order, then the one possible out-of-order sample `extra`.
// ---------------------------------------------------------------------------------------------
#define PCC_LEAF 128

template <int G, class Ring>
struct CoopSamples {
    const Grp<G> &g;
    Ring &ring;
    double *buf;          // shared memory, PCC_LEAF + G doubles, private to the group
    uint32_t i, end;
    double dl, extra;
    bool extra_pending;
    int fill;
    __device__ __forceinline__ CoopSamples(const Grp<G> &g_, Ring &r, double *b, const MiOut &o, double dl_)
        : g(g_), ring(r), buf(b), i(o.s_begin), end(o.s_end), dl(dl_), extra(o.extra),
          extra_pending(o.has_extra), fill(0) {}

    // make at least `need` (<= PCC_LEAF) samples available in buf[0..fill)
    __device__ __forceinline__ void fill_until(int need)
    {
        while (fill < need && (i != end || extra_pending)) {
            if (i != end) {
                bool valid;
                const Rec r = load_window(g, ring, i, end, valid);
                const bool keep = valid && !is_dead(r.a) && !sgn(r.l);
                const unsigned km = g.ballot(keep);
                if (keep) buf[fill + __popc(km & ((1u << g.gl) - 1u))] = r.l + dl;   // rtt = fl(ll + dl)
                fill += __popc(km);
                i = ((uint32_t)(end - i) < (uint32_t)G) ? end : i + (uint32_t)G;
            } else {
                if (g.gl == 0) buf[fill] = extra;
                fill++;
                extra_pending = false;
            }
        }
        g.sync();
    }
    // drop the first cnt samples
    __device__ __forceinline__ void consume(int cnt)
    {
        const int rem = fill - cnt;   // < G by construction of fill_until
        double v = 0.0;
        if ((int)g.gl < rem) v = buf[cnt + (int)g.gl];
        g.sync();
        if ((int)g.gl < rem) buf[(int)g.gl] = v;
        fill = rem;
        g.sync();
    }
};

// numpy's DOUBLE_pairwise_sum leaf (n <= 128) over a[0..n) in shared memory; 8 lanes hold the
// 8 accumulators; result is group-uniform.
template <int G>
__device__ __forceinline__ double coop_leaf(const Grp<G> &g, const double *a, int n)
{
    static_assert(G >= 8, "coop_leaf needs at least 8 lanes per group");
    if (n < 8) {
        double res = 0.;
        for (int k = 0; k < n; k++) res += a[k];
        return res;
    }
    const int j = (int)(g.gl & 7u);
    const int nb = n - (n % 8);
    double r = a[j];
    for (int k = 8; k < nb; k += 8) r += a[k + j];
    r += __shfl_xor_sync(g.gmask, r, 1);    // (r0+r1) (r2+r3) (r4+r5) (r6+r7)
    r += __shfl_xor_sync(g.gmask, r, 2);    // ((r0+r1)+(r2+r3)) ((r4+r5)+(r6+r7))
    r += __shfl_xor_sync(g.gmask, r, 4);
    double res = (G == 8) ? r : g.bcast(r, 0);
    for (int k = nb; k < n; k++) res += a[k];
    return res;
}

// pairwise sum of the next n samples of the stream (numpy's recursion, iterative post-order)
template <int G, class Ring>
__device__ __forceinline__ double coop_pw_sum(const Grp<G> &g, CoopSamples<G, Ring> &st, int n)
{
    int right_n[PCC_PW_STACK];
    double left_sum[PCC_PW_STACK];
    bool have_left[PCC_PW_STACK];
    int sp = 0;
    int cur = n;
    for (;;) {
        while (cur > PCC_LEAF) {
            int n2 = cur / 2;
            n2 -= n2 % 8;
            right_n[sp] = cur - n2; have_left[sp] = false; sp++;
            cur = n2;
        }
        st.fill_until(cur);
        double res = coop_leaf(g, st.buf, cur);
        st.consume(cur);
        for (;;) {
            if (sp == 0) return res;
            if (!have_left[sp - 1]) {
                left_sum[sp - 1] = res; have_left[sp - 1] = true;
                cur = right_n[sp - 1];
                break;
            }
            res = left_sum[sp - 1] + res;
            sp--;
        }
    }
}

template <int G, class Ring>
__device__ __forceinline__ void mi_means_coop(const Grp<G> &g, const MiOut &o, Ring &ring, double dl,
                                              double *buf, bool need_increase, double &avg_lat,
                                              double &lat_increase)
{
    const int n = o.acked;
    avg_lat = 0.0;
    lat_increase = 0.0;
    if (n <= 0) return;
    const int half = n / 2;
    if (n <= PCC_LEAF) {
        // everything fits the staging buffer: one pass over the ring, three leaves
        CoopSamples<G, Ring> st(g, ring, buf, o, dl);
        st.fill_until(n);
        double sum = 0.0;
        sum += coop_leaf(g, buf, n);
        avg_lat = sum / (double)n;                                          // sender_obs.py:119-122
        if (need_increase && half >= 1) {                                   // :138-142
            double s1 = 0.0, s2 = 0.0;
            s1 += coop_leaf(g, buf, half);
            s2 += coop_leaf(g, buf + half, n - half);
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
        g.sync();
        return;
    }
    {
        CoopSamples<G, Ring> st(g, ring, buf, o, dl);
        double sum = 0.0;
        sum += coop_pw_sum(g, st, n);
        avg_lat = sum / (double)n;
    }
    if (need_increase) {
        CoopSamples<G, Ring> st(g, ring, buf, o, dl);
        double s1 = 0.0, s2 = 0.0;
        s1 += coop_pw_sum(g, st, half);
        s2 += coop_pw_sum(g, st, n - half);
        lat_increase = s2 / (double)(n - half) - s1 / (double)half;
    }
}

}  // namespace pcc
