// pcc_coop.cuh -- group-cooperative monitor interval: G lanes of a warp work on ONE env, the
// 32/G groups of a warp move through the phases of the MI in lock step.
//
// Same algorithm and the same arithmetic as pcc_core.cuh::run_mi / mi_stats (which the host
// twin proves against the oracle); what changes is who does the work:
//   * queue recurrence (network_sim.py:72-84, inherently serial in binary64): lane 0 of each
//     group runs a branch-free loop over a chunk of 2G packets; the per-packet loss draws
//     are precomputed by all G lanes (one Philox4x32-10 block per lane) and handed over as
//     one bit mask; records are staged in shared memory and copied to the in-flight ring by
//     the whole group (coalesced 16-byte stores);
//   * ring scans (hop-1 / hop-2 cursors): W windows of G consecutive 16-byte records per
//     round, loads issued together (memory-level parallelism), predicate + ballot + ffs
//     instead of a dependent-load loop; the lines are prefetched before the send phase;
//   * MI-boundary cluster analysis (stragglers, tuple-order minimum): ballots and shuffle
//     min-reductions over the window;
//   * np.mean: acked latencies are compacted (ballot + popc prefix) into a shared-memory
//     staging buffer during the hop-2 scan; numpy's 8 accumulators live in 8 lanes and are
//     combined with a 3-level xor-shuffle tree -- bit-identical to DOUBLE_pairwise_sum
//     (SURVEY.md F2).
// Control flow is warp-uniform (loops run while ANY group needs them, finished groups are
// predicated off), so every *_sync primitive uses the full mask -- except the rare
// "more than 128 samples" path, which runs per group with group masks.
#pragma once
#include "pcc_core.cuh"

namespace pcc {

#define PCC_INF_BITS 0x7FF0000000000000ull
#define PCC_FULL 0xffffffffu
#define PCC_LEAF 128          // numpy's PW_BLOCKSIZE
#define PCC_SCAN_W 4          // windows per scan round

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
    // warp-wide vote, returns this group's G bits
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(PCC_FULL, p) >> gbase) & LOW; }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(PCC_FULL, v, (int)gbase + src); }
    __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(PCC_FULL, v, (int)gbase + src); }
    __device__ __forceinline__ unsigned bcast(unsigned v, int src) const { return __shfl_sync(PCC_FULL, v, (int)gbase + src); }
    __device__ __forceinline__ double min(double v) const
    {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(PCC_FULL, v, o));
        return v;
    }
    // group-mask variants for the divergent (per-group) slow path
    __device__ __forceinline__ unsigned gballot(bool p) const { return (__ballot_sync(gmask, p) >> gbase) & LOW; }
    __device__ __forceinline__ double gbcast(double v, int src) const { return __shfl_sync(gmask, v, (int)gbase + src); }
    __device__ __forceinline__ void gsync() const { __syncwarp(gmask); }
    static __device__ __forceinline__ int lead_ones(unsigned bm) { return (bm == LOW) ? G : (__ffs(~bm) - 1); }
    static __device__ __forceinline__ unsigned lowmask(int n) { return (n >= 32) ? 0xffffffffu : ((1u << n) - 1u); }
};

__device__ __forceinline__ void philox_block(uint64_t seed, uint64_t blk, uint32_t &c0, uint32_t &c1,
                                             uint32_t &c2, uint32_t &c3)
{
    c0 = (uint32_t)blk; c1 = (uint32_t)(blk >> 32); c2 = PCC_PHILOX_DOMAIN; c3 = 0u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// bit k of the result = bit k/2 of (k even ? even_bits : odd_bits)
__device__ __forceinline__ uint64_t interleave_bits(uint32_t even_bits, uint32_t odd_bits)
{
    uint64_t x = even_bits, y = odd_bits;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull; y = (y | (y << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;  y = (y | (y << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;  y = (y | (y << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;  y = (y | (y << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;  y = (y | (y << 1)) & 0x5555555555555555ull;
    return x | (y << 1);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// One instruction per scan round: lane gl asks L2 for the 128-byte line (8 records) that starts 8 * gl records past the
// round's windows, so the next 8 * G records -- two rounds -- are on their way while this round's loads are consumed.
template <int G, class Ring>
__device__ __forceinline__ void scan_prefetch(const Grp<G> &g, Ring &ring, uint32_t i, uint32_t lim, bool on)
{
    const uint32_t o = (uint32_t)(PCC_SCAN_W * G) + g.gl * 8u;
    if (on && o < (uint32_t)(lim - i)) prefetch_l2_line(ring.addr(i + o));
}

// window load: lane gl gets record i + gl if `on` and it lies before `lim`, else a neutral
// dummy (a = +inf: never consumable, not flagged; l = +1: not dropped)
template <int G, class Ring>
__device__ __forceinline__ Rec load_window(const Grp<G> &g, Ring &ring, uint32_t i, uint32_t lim, bool on, bool &valid)
{
    const uint32_t idx = i + g.gl;
    valid = on && ((int32_t)(idx - lim) < 0);
    Rec r;
    if (valid) r = ring.load(idx);
    else { r.a = u2d(PCC_INF_BITS); r.l = 1.0; }
    return r;
}

// argmin of (key1, key2, dropped False<True) over candidate lanes, merged into the running
// minimum (has, m_idx, m_k1, m_k2, m_d).  Results are group-uniform; warp-uniform call.
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

// Per-group shared-memory scratch
template <int G>
struct GroupSmem {
    static constexpr bool kSendV2 = false;
    double2 stage[2 * G + 1];    // records of one send chunk (slot k + 1 = packet k)
    double buf[PCC_LEAF + ((G == 32) ? 4 : 1) * G];    // acked-latency staging for np.mean
};

// Phase (1) of an MI for a group of G lanes: all sends with t < end.  Lane 0 of the group runs the
// branch-free queue recurrence over chunks of up to 2G packets; the loss draws of a chunk come from
// all G lanes (one Philox block each, two ballots); records are staged in shared memory and copied
// to the ring by the whole group.  Every lane holds the same (t, qd, t_upd, tail, draws, sent) on
// entry and on exit.  Warp-uniform control flow: all 32 lanes of the warp must call it together.
template <int G, class Ring>
__device__ __forceinline__ void coop_send_chunks(const Grp<G> &g, bool alive, const EnvState &s, Ring &ring,
                                                 uint64_t seed, uint64_t &draws, double end, double inv_rate,
                                                 double2 *stage, double &t, double &qd, double &t_upd,
                                                 uint32_t &tail, uint32_t h2, int32_t &sent, bool &ovf)
{
    const uint32_t cap = ring.capacity();
    bool more = alive && (t < end);
    while (__any_sync(PCC_FULL, more)) {
        const unsigned off = (unsigned)(draws & 1ull);
        uint32_t c0, c1, c2, c3;
        philox_block(seed, (draws >> 1) + g.gl, c0, c1, c2, c3);
        const unsigned be = g.ballot(res53(c0, c1) < s.lr);   // draw 2*blk   of lane's block  (:73)
        const unsigned bo = g.ballot(res53(c2, c3) < s.lr);   // draw 2*blk+1
        uint64_t dm = interleave_bits(be, bo) >> off;         // bit k = random drop of the chunk's k-th packet
        const int navail = 2 * G - (int)off;
        int k = 0;
        if (g.gl == 0 && more) {
            if ((uint32_t)(tail - h2) + (uint32_t)navail > cap) { ovf = true; }   // fatal, reported by the host
            else {
                // Measured on B200 (tools/chain_microbench.cu): the bare recurrence costs 42 cycles per packet;
                // a data-dependent exit test adds 27 (DADD -> DSETP -> unpredicted branch every iteration), send
                // times read back from shared memory add 35, the record store in the same iteration adds 25 --
                // a lone warp issues in order, so everything that waits stalls the recurrence behind it.  Hence:
                // (a) count-only pre-loop: how many send times t_k = fl(t_{k-1} + 1/rate) (:161) are < end
                //     (skipped when the whole chunk provably fits: the recurrence drifts from t + k/rate by at most
                //     k half-ulps, vastly less than the one extra 1/rate of margin)
                double tt = t;
                int cntk = 0;
                if (t + (double)(navail + 1) * inv_rate < end) cntk = navail;
                else {
#pragma unroll 16
                    for (int j = 0; j < 2 * G; ++j) {
                        cntk += (j < navail && tt < end) ? 1 : 0;  // monotone: counts a prefix
                        tt = tt + inv_rate;
                    }
                }
                // (b) counted main loop: send time by recurrence, the queue recurrence (:66-84) in speculative form
                //     (both candidates of q' from y = q - x before the selects), and the record of packet k-1
                //     finished while packet k's recurrence is in flight (stage[k] = record of packet k-1)
                const double k0 = (0.0 > s.w_full) ? 0.0 : s.d_bw;   // q' when the queue has drained (w = 0)
                const bool full0 = 0.0 > s.w_full;
                double q = qd, tu = t_upd;
                double pw = 0.0, pt = 0.0;
                bool pd = false;
                tt = t;
#pragma unroll 2
                for (; k < cntk; ++k) {
                    const bool rdrop = (dm & 1ull) != 0ull;                    // :73
                    dm >>= 1;
                    const double y = q - (tt - tu);                            // :66-67
                    // the previous packet is staged RAW -- send time and queue delay seen (>= +0.0, so its sign bit
                    // carries the drop flag); the two binary64 adds that turn it into a record are done by all lanes
                    // at copy-out, off the chain lane's in-order instruction stream
                    stage[k] = make_double2(pt, __longlong_as_double(__double_as_longlong(pw) | (pd ? (long long)PCC_SIGN : 0ll)));
                    const double cpos = s.d_bw + y;                            // :82 if 0 < y <= w_full
                    const bool pos = y > 0.0;
                    const bool fullp = y > s.w_full;                           // :77-79 (tail_drop_threshold)
                    const double w = pos ? y : 0.0;                            // max(0.0, y)
                    double qn = fullp ? y : cpos;
                    qn = pos ? qn : k0;
                    const bool full = pos ? fullp : full0;
                    q = rdrop ? q : qn;                                        // :74-82
                    tu = rdrop ? tu : tt;
                    pw = w; pt = tt; pd = rdrop || full;
                    tt = tt + inv_rate;                                        // :161
                }
                if (cntk > 0)
                    stage[cntk] = make_double2(pt, __longlong_as_double(__double_as_longlong(pw) | (pd ? (long long)PCC_SIGN : 0ll)));
                t = tt; qd = q; t_upd = tu;
            }
        }
        const int packed = g.bcast((int)(k | ((k == navail && t < end && !ovf) ? 0x100 : 0)), 0);
        k = packed & 0xff;
        __syncwarp();
        for (int j = (int)g.gl; j < k; j += G) {
            const double2 v = stage[j + 1];          // slot j + 1 holds packet j: (send time, +-queue delay seen)
            const double ll = s.dl + absd(v.y);                                // :69-70
            Rec r;
            r.a = v.x + ll;                                                    // :173-174
            r.l = __longlong_as_double(__double_as_longlong(ll) | (__double_as_longlong(v.y) & (long long)PCC_SIGN));   // :175
            ring.store(tail + (uint32_t)j, r);
        }
        __syncwarp();
        if (more) { tail += (uint32_t)k; draws += (uint64_t)k; sent += k; }
        more = more && ((packed & 0x100) != 0);
    }
    // make the chain lane's state the group's state
    t = g.bcast(t, 0); qd = g.bcast(qd, 0); t_upd = g.bcast(t_upd, 0);
    ovf = g.bcast((int)ovf, 0) != 0;
    __syncwarp();   // record stores above are read by other lanes below
}

// ---------------------------------------------------------------------------------------------------------------
// Send phase, second generation: G lanes per env, the 32/G groups of a warp in lock step (G = 32: a heavy env alone in
// its warp; G = 8: four envs per warp).  What is serial in binary64 is only the queue recurrence  q' = f(q - x)  over
// the packets that were NOT randomly dropped (network_sim.py:66-84); the first version's chain lane also ran the
// pacing-timer recurrence, tested the drop bit, built every record and stored it (115-200 cycles per packet in situ
// against 42 for the bare recurrence, profiles/r01_chain_microbench.txt).  Here, per chunk of 64 packets:
//   S1  loss draws: 32/G Philox blocks per lane, compared as 53-bit integers (loss_threshold), ballots -> drop mask;
//   S2  how many of the chunk's send times are < end (the times are already in shared memory, see S4);
//   S3  x_k = t_k - t_(last packet before k that reached the queue), COMPACTED over the packets that reach the queue;
//   S4  lane 0 of the group runs  y = q - x,  q = f(y)  in place over that compact array: one LDS, two DADD, two DSETP
//       and the selects per packet.  The SAME instruction stream, on lane 1, is the pacing-timer recurrence
//       t_(k+1) = fl(t_k + 1/rate) (:161) of the NEXT chunk: its array is pre-filled with x = -1/rate and its tail-drop
//       threshold is -inf, so it computes y = t + 1/rate and selects it -- exact, and free, because the dependency
//       chain of lane 0 leaves most issue slots idle.  With G = 8 the four chains and four timers of a warp share it;
//   S5  all lanes turn (t_k, y, drop bits) into records -- a randomly dropped packet sees y = f(y of the previous
//       packet that reached the queue) - x_k, recomputed here -- and store them to the ring, coalesced.
// Every lane of a group holds the same (t, qd, t_upd, tail, draws, sent) on entry and on exit.  Warp-uniform control
// flow: all 32 lanes call it together.
struct SoloSendSmem {
    double ts[2][72];      // send times of the current / next chunk: ts[b][k] = packet k, ts[b][navail] = the one after
    double xs[64];         // x of the packets that reach the queue, compact; overwritten in place by y
};

// CPython's MT19937 (state mt[0..623] + position mt[624], laid out like random.getstate()[1]) as the
// source of a chunk's loss draws, by a whole warp: draw k of the chunk is genrand_res53 of the stream words at positions
// pos + 2k and pos + 2k + 1 -- tempering is per word, so the draws of a chunk are independent of each other; the
// block regeneration (the "twist") runs in three phases of 32-wide steps whose reads all precede their writes.
// Exactly `cnt` draws are consumed, which is why the caller counts the chunk's packets first.
__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__device__ __forceinline__ void mt_twist_warp(uint32_t *mt)
{
    const int lane = (int)(threadIdx.x & 31u);
    for (int base = 0; base < 227; base += 32) {            // mt[kk] = mt[kk + 397] ^ f(mt[kk], mt[kk + 1]), old values
        const int kk = base + lane;
        const bool on = kk < 227;
        uint32_t a = 0, b = 0, c = 0;
        if (on) { a = mt[kk]; b = mt[kk + 1]; c = mt[kk + 397]; }
        __syncwarp();
        if (on) { const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu); mt[kk] = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
        __syncwarp();
    }
    for (int base = 227; base < 623; base += 32) {          // mt[kk] = NEW mt[kk - 227] ^ f(old mt[kk], old mt[kk + 1])
        const int kk = base + lane;
        const bool on = kk < 623;
        uint32_t a = 0, b = 0, c = 0;
        if (on) { a = mt[kk]; b = mt[kk + 1]; c = mt[kk - 227]; }
        __syncwarp();
        if (on) { const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu); mt[kk] = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
        __syncwarp();
    }
    if (lane == 0) {
        const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    __syncwarp();
}
// bit k of the result: draw k (k < cnt <= 64) of the chunk is < lr.  `tailbuf`: 128 words of shared scratch.
__device__ __forceinline__ uint64_t mt_chunk_mask(uint32_t *mt, int cnt, uint64_t thr, uint32_t *tailbuf)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t pos = mt[624];
    const bool cross = pos + 2u * (uint32_t)cnt > 624u;     // the chunk runs past the current block: regenerate
    if (cross) {
        const uint32_t ntail = 624u - pos;                   // < 128: the words of the old block still to be used
        for (uint32_t i = lane; i < ntail; i += 32u) tailbuf[i] = mt[pos + i];
        __syncwarp();
        mt_twist_warp(mt);
    }
    unsigned m[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int k = (int)lane + 32 * u;
        bool d = false;
        if (k < cnt) {
            const uint32_t i0 = pos + 2u * (uint32_t)k, i1 = i0 + 1u;
            const uint32_t w0 = i0 < 624u ? (cross ? tailbuf[i0 - pos] : mt[i0]) : mt[i0 - 624u];
            const uint32_t w1 = i1 < 624u ? (cross ? tailbuf[i1 - pos] : mt[i1]) : mt[i1 - 624u];
            d = u53(mt_temper(w0), mt_temper(w1)) < thr;    // genrand_res53: a = first word >> 5, b = second >> 6
        }
        m[u] = __ballot_sync(PCC_FULL, d);
    }
    __syncwarp();
    if (lane == 0 && cnt > 0) mt[624] = cross ? pos + 2u * (uint32_t)cnt - 624u : pos + 2u * (uint32_t)cnt;
    __syncwarp();
    return (uint64_t)m[0] | ((uint64_t)m[1] << 32);
}

template <int G>
struct GroupSmemV2 {
    static constexpr bool kSendV2 = true;
    SoloSendSmem send;
    double buf[PCC_LEAF + ((G == 32) ? 4 : 1) * G];    // acked-latency staging for np.mean
};

template <int G, class Ring>
__device__ __forceinline__ void group_send_chunks(const Grp<G> &g, bool alive, const EnvState &s, Ring &ring, uint64_t seed,
                                                  uint64_t &draws, double end, double inv_rate, SoloSendSmem &sm, double &t,
                                                  double &qd, double &t_upd, uint32_t &tail, uint32_t h2, int32_t &sent,
                                                  bool &ovf, uint32_t *mt = nullptr)
{
    // mt != nullptr (G == 32 only): the loss draws come from CPython's MT19937 state at `mt` instead of the Philox stream
    constexpr int NB = 32 / G;          // Philox blocks per lane and chunk
    constexpr int PPL = 64 / G;         // packets per lane and chunk
    const uint32_t cap = ring.capacity();
    bool more = alive && (t < end);
    if (!__any_sync(PCC_FULL, more)) return;                  // warp-uniform
    const uint64_t thr = loss_threshold(s.lr);
    const double k0 = (0.0 > s.w_full) ? 0.0 : s.d_bw;        // q' when the queue has drained (w = 0)
    const bool full0 = 0.0 > s.w_full;
    const bool timer = g.gl == 1;
    const double r_dbw = timer ? 0.0 : s.d_bw;
    const double r_wfull = timer ? __longlong_as_double((long long)PCC_NEG_INF) : s.w_full;
    const double neg_rate = -inv_rate;
    // send times of the first chunk: the recurrence (:161), once per MI; the timer lane's array of the next chunk
    if (g.gl == 0 && more) {
        double tt = t;
#pragma unroll 13
        for (int k = 0; k < 65; ++k) { sm.ts[0][k] = tt; tt = tt + inv_rate; }
    }
#pragma unroll
    for (int u = 0; u < PPL; u++) sm.ts[1][1 + (int)g.gl + u * G] = neg_rate;
    __syncwarp();
    int b = 0;
    double q = qd, tu = t_upd;
    while (__any_sync(PCC_FULL, more)) {
        const unsigned off = (G == 32 && mt != nullptr) ? 0u : (unsigned)(draws & 1ull);
        const int navail = 64 - (int)off;
        // S2: packets of this chunk that are sent in this MI (send times increase: a prefix)
        double tk[PPL];
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < PPL; u++) {
            const int k = (int)g.gl + u * G;
            tk[u] = sm.ts[b][k];
            cnt += __popc(g.ballot(k < navail && tk[u] < end));
        }
        bool fits = true;
        if (more && (uint32_t)(tail - h2) + (uint32_t)cnt > cap) { ovf = true; fits = false; }   // fatal, reported by the host
        const bool work = more && fits;
        if (!work) cnt = 0;
        // S1: bit k of dm = packet k of the chunk is randomly dropped (:73)
        uint64_t dm;
        if (G == 32 && mt != nullptr) {
            dm = mt_chunk_mask(mt, cnt, thr, reinterpret_cast<uint32_t *>(&sm.xs[0]));   // xs is free until S3
        } else {
            unsigned be = 0u, bo = 0u;
#pragma unroll
            for (int j = 0; j < NB; j++) {
                uint32_t c0, c1, c2, c3;
                philox_block(seed, (draws >> 1) + (uint64_t)(j * G) + g.gl, c0, c1, c2, c3);
                be |= g.ballot(u53(c0, c1) < thr) << (j * G);
                bo |= g.ballot(u53(c2, c3) < thr) << (j * G);
            }
            dm = interleave_bits(be, bo) >> off;
        }
        const double t_after = sm.ts[b][navail];
        const bool next = work && (cnt == navail) && (t_after < end);
        const uint64_t sentm = (cnt >= 64) ? ~0ull : ((1ull << cnt) - 1ull);
        const uint64_t ndm = ~dm & sentm;                     // sent packets that reach the queue
        const int cnt_nd = __popcll(ndm);
        // S3: x_k (:66-67 with :75-76: the update time is the send time of the last packet that reached the queue)
#pragma unroll
        for (int u = 0; u < PPL; u++) {
            const int k = (int)g.gl + u * G;
            const uint64_t before = ndm & ((1ull << k) - 1ull);
            if ((ndm >> k) & 1ull) {
                const double tuk = before ? sm.ts[b][63 - __clzll((long long)before)] : tu;
                sm.xs[__popcll(before)] = tk[u] - tuk;
            }
        }
        __syncwarp();
        // S4: the two recurrences of every group, one instruction stream, in place
        double state = q;
        {
            const int niter = timer ? (next ? 64 : 0) : cnt_nd;
            int kmax = (g.gl < 2) ? niter : 0, kmin = (g.gl < 2 && niter > 0) ? niter : 64;   // idle lanes do not shorten the fast part
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int v = __shfl_xor_sync(PCC_FULL, kmax, o), u = __shfl_xor_sync(PCC_FULL, kmin, o);
                kmax = v > kmax ? v : kmax;
                kmin = u < kmin ? u : kmin;
            }
            if (g.gl < 2) {
                double *io = timer ? &sm.ts[b ^ 1][1] : &sm.xs[0];
                if (timer) { state = t_after; if (next) sm.ts[b ^ 1][0] = t_after; }
                const double state0 = state;
                // [0, kmin): every WORKING chain and timer lane of the warp is inside its range -- no predicates on the
                // chain (a lane with nothing to do computes on scratch values and restores its state afterwards)
#pragma unroll 4
                for (int k = 0; k < kmin; ++k) {
                    const double y = state - io[k];                           // :66-67 | t + 1/rate
                    io[k] = y;
                    const double cpos = r_dbw + y;                            // :82 if 0 < y <= w_full
                    const bool pos = y > 0.0;
                    const bool fullp = y > r_wfull;                           // :77-79 (tail_drop_threshold)
                    const double qn = fullp ? y : cpos;
                    state = pos ? qn : k0;
                }
#pragma unroll 2
                for (int k = kmin; k < kmax; ++k) {
                    const bool on = k < niter;
                    const double y = state - io[k];
                    if (on) io[k] = y;
                    const double cpos = r_dbw + y;
                    const bool pos = y > 0.0;
                    const bool fullp = y > r_wfull;
                    double qn = fullp ? y : cpos;
                    qn = pos ? qn : k0;
                    state = on ? qn : state;
                }
                if (niter == 0) state = state0;
            }
        }
        __syncwarp();
        // S5: records (:173-175) and the chunk's carry
#pragma unroll
        for (int u = 0; u < PPL; u++) {
            const int k = (int)g.gl + u * G;
            if (k < cnt) {
                const uint64_t before = ndm & ((1ull << k) - 1ull);
                const int rank = __popcll(before);
                const bool rdrop = ((dm >> k) & 1ull) != 0ull;
                double y;
                if (!rdrop) y = sm.xs[rank];
                else {
                    // the queue as the last packet that reached it left it (:74: a random drop does not touch it)
                    double qp = q;
                    if (rank > 0) {
                        const double yp = sm.xs[rank - 1];
                        qp = (yp > 0.0) ? ((yp > s.w_full) ? yp : s.d_bw + yp) : k0;
                    }
                    const double tuk = before ? sm.ts[b][63 - __clzll((long long)before)] : tu;
                    y = qp - (tk[u] - tuk);
                }
                const bool pos = y > 0.0;
                const double w = pos ? y : 0.0;                           // max(0.0, y)
                const bool full = pos ? (y > s.w_full) : full0;
                const bool dropped = rdrop || full;
                const double ll = s.dl + w;                               // :69-70
                Rec r;
                r.a = tk[u] + ll;
                r.l = __longlong_as_double(__double_as_longlong(ll) | (dropped ? (long long)PCC_SIGN : 0ll));
                ring.store(tail + (uint32_t)k, r);
            }
        }
        double tu_n = tu, t_n = t;
        if (work) {
            if (ndm) tu_n = sm.ts[b][63 - __clzll((long long)ndm)];
            t_n = sm.ts[b][cnt];
        }
        const double qs = g.bcast(state, 0);
        __syncwarp();                                          // all reads of ts[b] / xs are done
#pragma unroll
        for (int u = 0; u < PPL; u++) sm.ts[b][1 + (int)g.gl + u * G] = neg_rate;   // the timer's array of the chunk after next
        if (work) { q = qs; tu = tu_n; t = t_n; tail += (uint32_t)cnt; draws += (uint64_t)cnt; sent += cnt; }
        more = next;
        b ^= 1;
        __syncwarp();
    }
    qd = q; t_upd = tu;
    __syncwarp();   // record stores above are read by other lanes below
}

// One monitor interval.  Every lane of a group holds the same EnvState copy on entry and on
// exit.  `alive` = this group has an env (warp-uniform control flow needs all lanes present).
// On exit, if out.acked <= PCC_LEAF, sm.buf[0..out.acked) holds the MI's samples in order.
#ifdef PCC_PROFILE
#define PCC_TICK(k) do { if (prof) prof[k] = clock64(); } while (0)
#else
#define PCC_TICK(k)
#endif
// `gbuf` (optional): global staging of ALL acked latencies of the MI (capacity gcap samples per group), read back by
// means_groups_from_buf when there are more than one numpy leaf of them.
template <int G, class Ring, class SM>
__device__ __forceinline__ void run_mi_coop(const Grp<G> &g, bool alive, EnvState &s, Ring &ring, uint64_t seed,
                                            uint64_t &draws, double dur, SM &sm, MiOut &out,
                                            long long *prof = nullptr, double *gbuf = nullptr, int gcap = 0)
{
    PCC_TICK(0);
    const double end = s.cur_time + dur;            // network_sim.py:124
    const double inv_rate = 1.0 / s.rate;           // :161
    const uint32_t cap = ring.capacity();
    int32_t sent = 0, acked = 0, lost = 0;
    out.start = s.cur_time;
    out.has_extra = false;
    out.extra = 0.0;
    out.s_begin = s.h2;
    double t = s.next_send, qd = s.qd, t_upd = s.t_upd;
    uint32_t tail = s.tail, h1 = s.h1, h2 = s.h2;
    bool ovf = false;

    // pull the lines the two cursors will walk first into L2 while the send phase runs
    if (alive) {
        const uint32_t pend1 = tail - h1, pend2 = tail - h2;   // records pending per stream
#pragma unroll
        for (int w = 0; w < 2; w++) {
            const uint32_t o = (uint32_t)((w * G + (int)g.gl) * 8);
            if (o < pend1) prefetch_l2_line(ring.addr(h1 + o));
            if (o < pend2) prefetch_l2_line(ring.addr(h2 + o));
        }
    }

    // ---- (1) sends with t < end: chain on lane 0, loss draws from all lanes ---------------
    if constexpr (SM::kSendV2)
        group_send_chunks(g, alive, s, ring, seed, draws, end, inv_rate, sm.send, t, qd, t_upd, tail, h2, sent, ovf);
    else
        coop_send_chunks(g, alive, s, ring, seed, draws, end, inv_rate, sm.stage, t, qd, t_upd, tail, h2, sent, ovf);
    PCC_TICK(1);

    // ---- (2) hop-1 events with a < end ------------------------------------------------------
    {
        bool scanning = alive;
        while (__any_sync(PCC_FULL, scanning)) {
            unsigned bm[PCC_SCAN_W];
            scan_prefetch(g, ring, h1, tail, scanning);
#pragma unroll
            for (int w = 0; w < PCC_SCAN_W; w++) {
                bool valid;
                const Rec r = load_window(g, ring, h1 + (uint32_t)(w * G), tail, scanning, valid);
                bm[w] = g.ballot(valid && (sgn(r.a) || r.a < end));
            }
            int adv = 0;
            bool stop = false;
#pragma unroll
            for (int w = 0; w < PCC_SCAN_W; w++) {
                const int nl = Grp<G>::lead_ones(bm[w]);
                if (!stop) adv += nl;
                stop = stop || (nl < G);
            }
            if (scanning) h1 += (uint32_t)adv;
            scanning = scanning && !stop;
        }
    }
    PCC_TICK(2);
    bool has1 = false;
    uint32_t m1 = 0; double m1a = 0.0, m1l = 0.0; bool m1d = false;
    {
        uint32_t kk = h1;
        bool open = alive && (kk != tail);
        while (__any_sync(PCC_FULL, open)) {
            bool valid;
            const Rec r = load_window(g, ring, kk, tail, open, valid);
            const bool dr = sgn(r.l);
            const unsigned validm = g.ballot(valid);
            const unsigned ndm = g.ballot(valid && !dr);
            const int pend = ndm ? (__ffs(ndm) - 1) : G;   // accepted record closes the cluster
            const bool pending = valid && (int)g.gl <= pend && !sgn(r.a);
            const bool strag = pending && (r.a < end);
            if (strag) ring.store_a(kk + g.gl, negd(r.a));
            window_argmin(g, pending && !strag, r.a, absd(r.l), dr, kk, has1, m1, m1a, m1l, m1d);
            open = open && (pend == G) && (validm == Grp<G>::LOW) && ((uint32_t)(kk + G) != tail);
            kk += G;
        }
        __syncwarp();   // straggler flags are read by the hop-2 scan
    }

    PCC_TICK(3);
    // ---- (3) hop-2 events with b < end; acked latencies staged for np.mean --------------------
    bool at_live = false;
    {
        bool scanning = alive;
        while (__any_sync(PCC_FULL, scanning)) {
            unsigned bm[PCC_SCAN_W], am[PCC_SCAN_W], lm[PCC_SCAN_W], lv[PCC_SCAN_W];
            double l2[PCC_SCAN_W];
            scan_prefetch(g, ring, h2, tail, scanning);
#pragma unroll
            for (int w = 0; w < PCC_SCAN_W; w++) {
                bool valid;
                const uint32_t i = h2 + (uint32_t)(w * G);
                const Rec r = load_window(g, ring, i, tail, scanning, valid);
                const bool dead = is_dead(r.a);
                const bool c1 = ((int32_t)(i + g.gl - h1) < 0) || sgn(r.a);
                const bool early = (absd(r.a) + s.dl) < end;         // :149-154, link 1 latency == dl
                const bool cons = valid && (dead || (c1 && early));
                bm[w] = g.ballot(cons);
                am[w] = g.ballot(cons && !dead && !sgn(r.l));        // :144-145
                lm[w] = g.ballot(cons && !dead && sgn(r.l));         // :141-142
                lv[w] = g.ballot(valid && !dead && c1 && !early);
                l2[w] = r.l + s.dl;                                  // rtt = fl(ll + dl)
            }
            int adv = 0;
            bool stop = false;
#pragma unroll
            for (int w = 0; w < PCC_SCAN_W; w++) {
                const int nl = Grp<G>::lead_ones(bm[w]);
                const unsigned lead = Grp<G>::lowmask(nl);
                const bool on = scanning && !stop;
                const unsigned a_w = on ? (am[w] & lead) : 0u;
                if ((a_w >> g.gl) & 1u) {
                    const int pos = acked + __popc(a_w & Grp<G>::lowmask((int)g.gl));
                    if (pos < PCC_LEAF) sm.buf[pos] = l2[w];
                    if (gbuf != nullptr && pos < gcap) gbuf[pos] = l2[w];
                }
                if (on) {
                    acked += __popc(a_w);
                    lost += __popc(lm[w] & lead);
                    adv += nl;
                    if (nl < G) at_live = ((lv[w] >> nl) & 1u) != 0u;
                }
                stop = stop || (nl < G);
            }
            if (scanning) h2 += (uint32_t)adv;
            scanning = scanning && !stop;
        }
    }
    out.s_end = h2;
    PCC_TICK(4);
    bool has2 = false;
    uint32_t m2 = 0; double m2b = 0.0, m2l = 0.0; bool m2d = false;
    {
        uint32_t kk = h2;
        bool open = alive && at_live;
        while (__any_sync(PCC_FULL, open)) {
            bool valid;
            const Rec r = load_window(g, ring, kk, tail, open, valid);
            const bool dr = sgn(r.l);
            const bool dead = is_dead(r.a);
            const bool c1 = ((int32_t)(kk + g.gl - h1) < 0) || sgn(r.a);
            const unsigned x1 = g.ballot(!valid || (!dead && !c1));   // stop BEFORE this record
            const unsigned x2 = g.ballot(valid && !dr);               // stop AFTER this record
            const int p1 = x1 ? (__ffs(x1) - 1) : G;
            const int p2 = x2 ? (__ffs(x2) - 1) : G;
            const bool act = open && (int)g.gl < p1 && (int)g.gl <= p2 && !dead;
            const double b = absd(r.a) + s.dl;
            const double l2 = absd(r.l) + s.dl;
            const bool strag = act && (b < end);
            const unsigned sa = g.ballot(strag && !dr), sl = g.ballot(strag && dr);
            const double ex = g.bcast(l2, sa ? (__ffs(sa) - 1) : 0);
            if (sa) { out.extra = ex; out.has_extra = true; }   // at most one acked per cluster
            acked += __popc(sa);
            lost += __popc(sl);
            if (strag) ring.store_a(kk + g.gl, u2d(PCC_NEG_INF));
            window_argmin(g, act && !strag, b, l2, dr, kk, has2, m2, m2b, m2l, m2d);
            open = open && (p1 == G) && (p2 == G);
            kk += G;
        }
    }

    PCC_TICK(5);
    // ---- (4) the event that crosses `end` (group-uniform values, no votes) --------------------
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
            if (w > s.w_full) dropped = true;
            else { qd += s.d_bw; dropped = false; }
        }
        Rec r; r.a = t + ll; r.l = dropped ? negd(ll) : ll;
        if ((uint32_t)(tail - out.s_begin) >= cap) ovf = true;   // s_begin, not the advanced h2: consumed records may be re-read for the means
        else { if (alive && g.gl == 0) ring.store(tail, r); tail++; }
        t = t + inv_rate;
    } else if (which == 1) {
        s.cur_time = m1a;
        if (m1 == h1) h1++;
        else if (alive && g.gl == 0) ring.store_a(m1, negd(m1a));
    } else {
        s.cur_time = m2b;
        if (m2d) lost++; else { acked++; out.extra = m2l; out.has_extra = true; }
        if (m2 == h2) h2++;
        else if (alive && g.gl == 0) ring.store_a(m2, u2d(PCC_NEG_INF));
    }
    // the one possible out-of-order sample is the MI's last sample
    if (out.has_extra && acked <= PCC_LEAF && g.gl == 0) sm.buf[acked - 1] = out.extra;
    if (out.has_extra && gbuf != nullptr && acked <= gcap && g.gl == 0) gbuf[acked - 1] = out.extra;
    __syncwarp();   // orders this MI's flag stores / staging writes before later reads
    s.next_send = t;
    s.qd = qd; s.t_upd = t_upd;
    s.tail = tail; s.h1 = h1; s.h2 = h2;
    out.sent = sent; out.acked = acked; out.lost = lost;
    out.end = s.cur_time;
    out.overflow = ovf;
}

// ---------------------------------------------------------------------------------------------
// np.mean over the MI's samples.
// ---------------------------------------------------------------------------------------------
// numpy's DOUBLE_pairwise_sum leaf (n <= 128) over a[0..n) in shared memory; 8 lanes hold the
// 8 accumulators; result is group-uniform.  Uses the group mask: callable from divergent code.
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
    double res = (G == 8) ? r : g.gbcast(r, 0);
    for (int k = nb; k < n; k++) res += a[k];
    return res;
}

// Streaming re-read of the MI's samples for n > PCC_LEAF: acked, not-dead records of
// [s_begin, s_end) in ring order, then the one possible out-of-order sample `extra`.
template <int G, class Ring, int W = 1>
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

    // buf must hold PCC_LEAF + W * G samples
    __device__ __forceinline__ void fill_until(int need)   // need <= PCC_LEAF
    {
        while (fill < need && (i != end || extra_pending)) {
            if (i != end) {
                Rec r[W];
                bool valid[W];
#pragma unroll
                for (int w = 0; w < W; w++) {       // W independent loads in flight
                    const uint32_t idx = i + (uint32_t)(w * G) + g.gl;
                    valid[w] = (int32_t)(idx - end) < 0;
                    r[w].a = 0.0; r[w].l = -1.0;
                    if (valid[w]) r[w] = ring.load(idx);
                }
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const bool keep = valid[w] && !is_dead(r[w].a) && !sgn(r[w].l);
                    const unsigned km = g.gballot(keep);
                    if (keep) buf[fill + __popc(km & Grp<G>::lowmask((int)g.gl))] = r[w].l + dl;   // rtt = fl(ll + dl)
                    fill += __popc(km);
                }
                i = ((uint32_t)(end - i) < (uint32_t)(W * G)) ? end : i + (uint32_t)(W * G);
            } else {
                if (g.gl == 0) buf[fill] = extra;
                fill++;
                extra_pending = false;
            }
        }
        g.gsync();
    }
    __device__ __forceinline__ void consume(int cnt)
    {
        const int rem = fill - cnt;   // < W * G by construction of fill_until
        double v[W];
#pragma unroll
        for (int w = 0; w < W; w++) {
            const int k = w * G + (int)g.gl;
            v[w] = (k < rem) ? buf[cnt + k] : 0.0;
        }
        g.gsync();
#pragma unroll
        for (int w = 0; w < W; w++) {
            const int k = w * G + (int)g.gl;
            if (k < rem) buf[k] = v[w];
        }
        fill = rem;
        g.gsync();
    }
};

// pairwise sum of the next n samples of the stream (numpy's recursion, iterative post-order)
template <int G, class Ring, int W>
__device__ __noinline__ double coop_pw_sum(const Grp<G> &g, CoopSamples<G, Ring, W> &st, int n)
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

// np.mean of n > 128 staged samples a[0..n) per group, the groups of the warp in lock step: every group walks the
// leaves of numpy's recursion in order (PwStream's descend / leaf_done with the leaf result supplied from outside), a leaf
// is summed by 8 lanes = numpy's 8 accumulators + the xor-shuffle tree; three walks per group (all, first half, second
// half).  Warp-uniform call; `on` selects the groups that take part.
struct NoAcc { __device__ __forceinline__ double get(int) const { return 0.0; } __device__ __forceinline__ void set(int, double) {} };
struct LocalPwStack {
    int *right_n; double *left_sum;
    __device__ __forceinline__ int &rn(int i) { return right_n[i]; }
    __device__ __forceinline__ double &ls(int i) { return left_sum[i]; }
};
// numpy's pairwise sum of a[0..n) per 8-lane subgroup, the four subgroups of the warp in lock step on four DIFFERENT
// arrays (four envs of a quad warp; or the three sums -- all, first half, second half -- of one env of a solo warp).
// Warp-uniform call; subgroups with !on idle along.
__device__ __noinline__ double pw_sum_subgroups(bool on, const double *a, int n)
{
    const unsigned lane = threadIdx.x & 31u;
    const int j = (int)(lane & 7u);
    int rn[PCC_PW_STACK];
    double ls[PCC_PW_STACK];
    PwStream<NoAcc, LocalPwStack> m;
    m.stk.right_n = rn; m.stk.left_sum = ls;
    bool active = on && n > 0;
    int consumed = 0;
    double total = 0.0;
    m.begin(active ? n : 0);
    while (__any_sync(PCC_FULL, active)) {
        const int c = active ? m.cur : 0;
        const double *p = a + consumed;
        const int nb = c - (c % 8);
        int nbmax = nb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_xor_sync(PCC_FULL, nbmax, o); nbmax = v > nbmax ? v : nbmax; }
        // the (up to 7) tail samples: loaded together with the blocks, added in order at the end
        const int t0 = (c >= 8) ? nb : 0;
        double tv[7];
#pragma unroll
        for (int k = 0; k < 7; k++) tv[k] = (t0 + k < c) ? p[t0 + k] : 0.0;
        double r = 0.0;
        if (c >= 8) r = p[j];
#pragma unroll 8
        for (int k = 8; k < nbmax; k += 8)
            if (k < nb) r += p[k + j];
        r += __shfl_xor_sync(PCC_FULL, r, 1);    // (r0+r1) (r2+r3) (r4+r5) (r6+r7)
        r += __shfl_xor_sync(PCC_FULL, r, 2);    // ((r0+r1)+(r2+r3)) ((r4+r5)+(r6+r7))
        r += __shfl_xor_sync(PCC_FULL, r, 4);
        double res = (c >= 8) ? r : 0.;
#pragma unroll
        for (int k = 0; k < 7; k++) if (t0 + k < c) res += tv[k];
        if (active) {
            consumed += c;
            m.res = res;
            m.leaf_done();
            if (m.done) { total = m.total; active = false; }
        }
    }
    __syncwarp();
    return total;
}

// np.mean of n > 128 staged samples a[0..n) per group of a quad warp: three lock-step sweeps (all, first half, second
// half).  Warp-uniform call; `on` selects the groups that take part.
template <int G>
__device__ __forceinline__ void means_groups_from_buf(const Grp<G> &g, bool on, const double *a, int n, bool need_increase,
                                                      double &avg_lat, double &lat_increase)
{
    static_assert(G == 8, "one env per 8-lane subgroup");
    const int half = n / 2;
    const bool inc = on && need_increase && half >= 1;
    const double sum0 = pw_sum_subgroups(on, a, n);
    double sum1 = 0.0, sum2 = 0.0;
    if (__any_sync(PCC_FULL, inc)) {
        sum1 = pw_sum_subgroups(inc, a, half);
        sum2 = pw_sum_subgroups(inc, a + half, n - half);
    }
    if (on) {
        double sum = 0.0;
        sum += sum0;
        avg_lat = sum / (double)n;                                              // sender_obs.py:119-122
        lat_increase = 0.0;
        if (inc) {                                                              // :138-142
            double s1 = 0.0, s2 = 0.0;
            s1 += sum1;
            s2 += sum2;
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
    }
}

// The same for ONE env owned by the whole warp (solo): its three sums run side by side on three subgroups.
__device__ __forceinline__ void means_solo_from_buf(const double *a, int n, bool need_increase, double &avg_lat,
                                                    double &lat_increase)
{
    const unsigned lane = threadIdx.x & 31u;
    const int sub = (int)(lane >> 3), half = n / 2;
    const bool inc = need_increase && half >= 1;
    const double *pa = (sub == 2) ? a + half : a;
    const int cnt = (sub == 0) ? n : (sub == 1) ? half : (sub == 2) ? n - half : 0;
    const bool on = sub == 0 || (inc && sub < 3);
    const double sres = pw_sum_subgroups(on, pa, cnt);
    const double tot = __shfl_sync(PCC_FULL, sres, 0), f1 = __shfl_sync(PCC_FULL, sres, 8), f2 = __shfl_sync(PCC_FULL, sres, 16);
    double sum = 0.0;
    sum += tot;
    avg_lat = sum / (double)n;                                                  // sender_obs.py:119-122
    lat_increase = 0.0;
    if (inc) {                                                                  // :138-142
        double s1 = 0.0, s2 = 0.0;
        s1 += f1;
        s2 += f2;
        lat_increase = s2 / (double)(n - half) - s1 / (double)half;
    }
}

// avg latency (sender_obs.py:119-122) and latency increase (:138-142) of the MI
template <int G, class Ring, class SM>
__device__ __forceinline__ void mi_means_coop(const Grp<G> &g, bool alive, const MiOut &o, Ring &ring, double dl,
                                              SM &sm, bool need_increase, double &avg_lat,
                                              double &lat_increase, const double *gbuf = nullptr, int gcap = 0)
{
    int n = alive ? o.acked : 0;
    avg_lat = 0.0;
    lat_increase = 0.0;
    if constexpr (G == 8) if (gbuf != nullptr) {   // warp-uniform: staged samples of MIs with more than one leaf
        const bool use_g = n > PCC_LEAF && n <= gcap;
        if (__any_sync(PCC_FULL, use_g)) means_groups_from_buf(g, use_g, gbuf, n, need_increase, avg_lat, lat_increase);
        if (use_g) n = 0;    // done
    }
    const int half = n / 2;
    if (n > 0 && n <= PCC_LEAF) {
        // common case: run_mi_coop left all n samples in sm.buf
        double sum = 0.0;
        sum += coop_leaf(g, sm.buf, n);
        avg_lat = sum / (double)n;
        if (need_increase && half >= 1) {
            double s1 = 0.0, s2 = 0.0;
            s1 += coop_leaf(g, sm.buf, half);
            s2 += coop_leaf(g, sm.buf + half, n - half);
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
    } else if (n > PCC_LEAF) {
        constexpr int SW = (G == 32) ? 4 : 1;   // windows per streaming round
        {
            CoopSamples<G, Ring, SW> st(g, ring, sm.buf, o, dl);
            double sum = 0.0;
            sum += coop_pw_sum(g, st, n);
            avg_lat = sum / (double)n;
        }
        if (need_increase) {
            CoopSamples<G, Ring, SW> st(g, ring, sm.buf, o, dl);
            double s1 = 0.0, s2 = 0.0;
            s1 += coop_pw_sum(g, st, half);
            s2 += coop_pw_sum(g, st, n - half);
            lat_increase = s2 / (double)(n - half) - s1 / (double)half;
        }
    }
    __syncwarp();
}

}  // namespace pcc
