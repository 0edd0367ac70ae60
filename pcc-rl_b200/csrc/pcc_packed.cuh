// pcc_packed.cuh -- "a lane owns an env": the packed execution of the monitor interval.
//
// A warp owns 32 envs of SIMILAR predicted work (the host sorts the batch by predicted packets every step, see
// pcc_b200.cu: pcc_cost_packed_kernel + a 16-bit radix sort), one per lane.  Every phase of pcc_core.cuh::run_mi runs per lane -- the scalar code
// the host twin proves against the heap oracle -- so all 32 lanes are busy with useful work:
//
//   phase A   sends with t < end: pacing timer, Philox loss draw (compared as a 53-bit integer against a per-env
//             threshold: no int->double conversions), the binary64 queue recurrence (network_sim.py:72-84,
//             156-178), one 16-byte record per packet;
//   phase B1  hop-1 cursor scan + MI-boundary cluster (network_sim.py:147-154);
//   phase B2  hop-2 cursor scan + boundary, counting acked / lost (network_sim.py:140-145);
//   phase B3  the event that crosses the MI end, then numpy's pairwise means of the acked latencies in PUSH form
//             (PwStream), accumulators in shared memory;
//   phase C   MI metrics, reward, state write-back (per lane), history / observation rows (cooperative, coalesced).
//
// What is NOT per lane is the memory traffic.  Rings are only touched through a shared-memory TILE of 16 records per
// env: the warp fills / flushes it cooperatively -- half a warp moves one env's 256 contiguous bytes, 16 `cp.async`
// (LDGSTS) instructions bring in 512 records with no register staging and all loads in flight at once -- and each
// lane then walks its own padded tile row (conflict-free 128-bit LDS).  A per-lane global load would cost one L1
// wavefront per record; the tile costs 1/8 per record for the fill and 1/8 for the LDS.
#pragma once
#include "pcc_core.cuh"
#include "pcc_coop.cuh"

namespace pcc {

#ifndef PCC_TILE_R
#define PCC_TILE_R 16                       // records per env and tile round (even, <= 16)
#endif

struct PackedSmem {                         // per warp: 12 800 bytes
    double2 tile[32][PCC_TILE_R + 1];       // row = env (lane); rows padded to 272 B: conflict-free both ways
    double acc[2][8][32];                   // PwStream accumulators [machine][k % 8][lane]
};

#if defined(__CUDACC__)

struct RingSet { Rec *rings; uint32_t cap, mask; };   // rings[env][cap]

// Loss draws of one env as drop decisions, one Philox block (two draws) at a time.
struct LaneDraws {
    uint64_t seed, draws, thr;
    bool d0, d1;                 // decisions of draws 2*(draws >> 1) and 2*(draws >> 1) + 1
    __device__ __forceinline__ void block(uint64_t blk, bool &e0, bool &e1) const
    {
        uint32_t c0, c1, c2, c3;
        philox_block(seed, blk, c0, c1, c2, c3);
        e0 = u53(c0, c1) < thr;                      // network_sim.py:73, random.random() < self.lr
        e1 = u53(c2, c3) < thr;
    }
    __device__ __forceinline__ void init(uint64_t seed_, uint64_t draws_, double lr)
    {
        seed = seed_; draws = draws_; thr = loss_threshold(lr);
        block(draws >> 1, d0, d1);
    }
    __device__ __forceinline__ bool next()           // one draw outside the pipelined send loop
    {
        const bool d = (draws & 1ull) ? d1 : d0;
        draws++;
        if (!(draws & 1ull)) block(draws >> 1, d0, d1);
        return d;
    }
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch of the 128-byte lines holding records [from, from + 16) of env e's ring (clipped at lim): issued two tile
// rounds ahead of their use, so that a round's fill finds its records in L2 instead of paying a DRAM round trip
__device__ __forceinline__ void ring_prefetch(const RingSet &rs, int e, uint32_t from, uint32_t lim)
{
    const Rec *base = rs.rings + (size_t)e * rs.cap;
    if ((int32_t)(from - lim) < 0) prefetch_l2(base + (from & rs.mask));
    if ((int32_t)(from + 8u - lim) < 0) prefetch_l2(base + ((from + 8u) & rs.mask));
    if ((int32_t)(from + 15u - lim) < 0) prefetch_l2(base + ((from + 15u) & rs.mask));
}

// Cooperative tile fill: for every lane j with `on`, records [cur_j, min(cur_j + 16, lim_j)) of env e_j's ring go to
// tile[j][0 ..).  Half a warp per env, two envs per instruction.  Warp-uniform call.
__device__ __forceinline__ void tile_fill(const RingSet &rs, PackedSmem &sm, int e, uint32_t cur, uint32_t lim, bool on)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned onm = __ballot_sync(PCC_FULL, on);
    const unsigned k = lane & 15u, hi = lane >> 4;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (((onm >> (2 * i)) & 3u) == 0u) continue;                 // warp-uniform
        const int j = 2 * i + (int)hi;
        const int ej = __shfl_sync(PCC_FULL, e, j);
        const uint32_t cj = __shfl_sync(PCC_FULL, cur, j);
        const uint32_t lj = __shfl_sync(PCC_FULL, lim, j);
        if (((onm >> j) & 1u) && k < (unsigned)PCC_TILE_R && (int32_t)(cj + k - lj) < 0)
            cp_async16(&sm.tile[j][k], rs.rings + (size_t)ej * rs.cap + ((cj + k) & rs.mask));
    }
    if (on) ring_prefetch(rs, e, cur + 2u * PCC_TILE_R, lim);
    cp_async_wait_all();
    __syncwarp();
}

// Cooperative tile flush: tile[j][0 .. n_j) -> records tail_j ... of env e_j's ring.
__device__ __forceinline__ void tile_flush(const RingSet &rs, PackedSmem &sm, int e, uint32_t tail, int nst)
{
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp();
    const unsigned have = __ballot_sync(PCC_FULL, nst > 0);
    const unsigned k = lane & 15u, hi = lane >> 4;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (((have >> (2 * i)) & 3u) == 0u) continue;                // warp-uniform
        const int j = 2 * i + (int)hi;
        const int ej = __shfl_sync(PCC_FULL, e, j);
        const uint32_t tj = __shfl_sync(PCC_FULL, tail, j);
        const int nj = __shfl_sync(PCC_FULL, nst, j);
        if ((int)k < nj)
            *reinterpret_cast<double2 *>(rs.rings + (size_t)ej * rs.cap + ((tj + k) & rs.mask)) = sm.tile[j][k];
    }
    __syncwarp();
}

// PwStream accumulators of one lane: a column of shared memory (conflict-free for any per-lane index), addressed in
// the shared state space explicitly (a generic pointer kept in a struct costs generic LD / ST)
struct SmemAcc {
    unsigned addr;                                                   // shared-space address of acc[m][0][lane]
    __device__ __forceinline__ double get(int j) const
    {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr + (unsigned)j * 256u) : "memory");
        return v;
    }
    __device__ __forceinline__ void set(int j, double v)
    {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr + (unsigned)j * 256u), "d"(v) : "memory");
    }
};
// PwStream recursion stack in caller-owned (local-memory) arrays: touched once per leaf of >= 64 samples
struct LocalStack {
    int *right_n; double *left_sum;
    __device__ __forceinline__ int &rn(int i) { return right_n[i]; }
    __device__ __forceinline__ double &ls(int i) { return left_sum[i]; }
};

// Ring view of one lane for the MI-boundary analysis: records inside the lane's current tile window come from shared
// memory (the record the scan stopped on is almost always the only one needed), the rest from global memory; flag
// stores go to both so that the window stays consistent.
struct TileRing {
    DevRing g;
    double2 *row;          // the lane's tile row
    uint32_t w0;           // ring position of row[0]
    uint32_t n;            // valid records in the row
    __device__ __forceinline__ uint32_t capacity() const { return g.capacity(); }
    __device__ __forceinline__ Rec load(uint32_t i) const
    {
        const uint32_t d = i - w0;
        if (d < n) { const double2 v = row[d]; Rec r; r.a = v.x; r.l = v.y; return r; }
        return g.load(i);
    }
    __device__ __forceinline__ void store_a(uint32_t i, double a)
    {
        const uint32_t d = i - w0;
        if (d < n) row[d].x = a;
        g.store_a(i, a);
    }
    __device__ __forceinline__ void prefetch(uint32_t) const {}
};

struct PackedOut {
    int32_t sent, acked, lost;
    double start, end;
    double avg_lat, lat_inc;
    bool overflow;
};

// One monitor interval for the (up to) 32 envs of this warp.  `owner` lanes hold their env's state in `s` and its draw
// stream in `rng`; non-owner lanes take part in the cooperative tile traffic only.  Warp-uniform call.
template <bool WANT_MEANS>
__device__ __forceinline__ void packed_mi(const RingSet &rs, PackedSmem &sm, bool owner, int e, EnvState &s, LaneDraws &rng,
                                          double dur, bool need_increase, PackedOut &out, long long *prof = nullptr)
{
#ifdef PCC_PROFILE
#define PCC_PTICK(k) do { if (prof) prof[k] = clock64(); } while (0)
#else
#define PCC_PTICK(k)
#endif
    const unsigned lane = threadIdx.x & 31u;
    DevRing ring{rs.rings + (size_t)e * rs.cap, rs.mask};
    const double end = s.cur_time + dur;             // network_sim.py:124
    const double inv_rate = 1.0 / s.rate;            // :161
    out.start = s.cur_time;                          // reset_obs :319-324
    double t = s.next_send, q = s.qd, tu = s.t_upd;
    uint32_t tail = s.tail, h1 = s.h1, h2 = s.h2;
    int32_t sent = 0, acked = 0, lost = 0;
    bool ovf = false;
    const double k0 = (0.0 > s.w_full) ? 0.0 : s.d_bw;    // q' when the queue has drained (w = 0)
    const bool full0 = 0.0 > s.w_full;
    PCC_PTICK(0);
    if (owner) {
        // the first two tiles of both cursor scans travel to L2 while the send phase runs
        ring_prefetch(rs, e, h1, tail); ring_prefetch(rs, e, h1 + PCC_TILE_R, tail);
        ring_prefetch(rs, e, h2, tail); ring_prefetch(rs, e, h2 + PCC_TILE_R, tail);
    }

    // One packet (network_sim.py:156-178 -> :66-84), branch-free, both candidates of q' from y = q - x before the
    // selects (see tail_drop_threshold for the exact form of the tail-drop test).
#define PCC_PACKED_SEND(RDROP, STORE)                                                                          \
    {                                                                                                          \
        const bool rdrop_ = (RDROP);                                                 /* :73 */                 \
        const double y_ = q - (t - tu);                                              /* :66-67 */              \
        const double cpos_ = s.d_bw + y_;                                            /* :82 */                 \
        const bool pos_ = y_ > 0.0;                                                                            \
        const bool fullp_ = y_ > s.w_full;                                           /* :77-79 */              \
        const double w_ = pos_ ? y_ : 0.0;                                           /* max(0.0, y) */         \
        const double ll_ = s.dl + w_;                                                /* :69-70 */              \
        double qn_ = fullp_ ? y_ : cpos_;                                                                      \
        qn_ = pos_ ? qn_ : k0;                                                                                 \
        const bool full_ = pos_ ? fullp_ : full0;                                                              \
        q = rdrop_ ? q : qn_;                                                        /* :74-82 */              \
        tu = rdrop_ ? tu : t;                                                                                  \
        const bool dropped_ = rdrop_ || full_;                                                                 \
        const double2 rec_ = make_double2(t + ll_,                                   /* :173-175 */            \
            __longlong_as_double(__double_as_longlong(ll_) | (dropped_ ? (long long)PCC_SIGN : 0ll)));         \
        STORE;                                                                                                 \
        t = t + inv_rate;                                                            /* :161 */                \
        sent++;                                                                                                \
    }
#define PCC_PACKED_STAGE                                                                                       \
    if ((uint32_t)(tail + (uint32_t)nst - h2) >= rs.cap) ovf = true;   /* fatal, reported by the host */       \
    else { sm.tile[lane][nst] = rec_; nst++; }

    // ---- phase A: sends with t < end, 16 records per lane and round ------------------------------------------
    {
        bool more = owner && (t < end);
        while (__any_sync(PCC_FULL, more)) {
            int nst = 0;
            if (more) {
                if (rng.draws & 1ull) {                               // finish the half-used block
                    PCC_PACKED_SEND(rng.d1, PCC_PACKED_STAGE);
                    rng.draws++;
                    rng.block(rng.draws >> 1, rng.d0, rng.d1);
                }
                // here `draws` is even and (d0, d1) = block draws >> 1.  The block of the NEXT two packets is
                // computed next to the queue recurrence of the current two: independent dependency chains.
                while (nst + 2 <= PCC_TILE_R && t < end) {
                    bool n0, n1;
                    rng.block((rng.draws >> 1) + 1, n0, n1);
                    PCC_PACKED_SEND(rng.d0, PCC_PACKED_STAGE);
                    rng.draws++;
                    if (!(t < end)) break;                            // (d0, d1) stay: draws is odd now
                    PCC_PACKED_SEND(rng.d1, PCC_PACKED_STAGE);
                    rng.draws++;
                    rng.d0 = n0; rng.d1 = n1;
                }
                more = t < end;
            }
            tile_flush(rs, sm, e, tail, nst);
            tail += (uint32_t)nst;
        }
    }

    PCC_PTICK(1);
    // ---- phase B1: hop-1 events with a < end (scan_hop1 of pcc_core.cuh, from the tile) ----------------------
    TileRing tring{ring, &sm.tile[lane][0], 0u, 0u};
    {
        bool scanning = owner && (h1 != tail);
        while (__any_sync(PCC_FULL, scanning)) {
            tile_fill(rs, sm, e, h1, tail, scanning);
            if (scanning) {
                const uint32_t left = tail - h1;
                const int avail = left < (uint32_t)PCC_TILE_R ? (int)left : PCC_TILE_R;
                tring.w0 = h1; tring.n = (uint32_t)avail;
                int k = 0;
                for (; k < avail; k++) {
                    const double a = sm.tile[lane][k].x;
                    if (!sgn(a) && !(a < end)) break;
                }
                h1 += (uint32_t)k;
                scanning = (k == avail) && (h1 != tail);
            }
            __syncwarp();
        }
    }
    Pending m1;
    m1.has = false; m1.idx = 0; m1.t = 0.0; m1.l = 0.0; m1.dropped = false;
    if (owner) boundary_hop1(tring, h1, tail, end, m1);
    __syncwarp();                                    // straggler flags are read by the hop-2 fill

    PCC_PTICK(2);
    // ---- phase B2: hop-2 events with b < end (scan_hop2) ---------------------------------------------------
    const uint32_t s_begin = h2;
    bool at_live = false;
    tring.n = 0u;
    int b2_rounds = 0;
    {
        bool scanning = owner && (h2 != tail);
        while (__any_sync(PCC_FULL, scanning)) {
            tile_fill(rs, sm, e, h2, tail, scanning);
            b2_rounds++;
            if (scanning) {
                const uint32_t left = tail - h2;
                const int avail = left < (uint32_t)PCC_TILE_R ? (int)left : PCC_TILE_R;
                tring.w0 = h2; tring.n = (uint32_t)avail;
                int k = 0;
                for (; k < avail; k++) {
                    const double2 r = sm.tile[lane][k];
                    if (!is_dead(r.x)) {
                        const bool c1 = ((int32_t)(h2 + (uint32_t)k - h1) < 0) || sgn(r.x);
                        if (!c1) break;
                        const double b = absd(r.x) + s.dl;           // link 1: latency == dl exactly (N1)
                        if (!(b < end)) { at_live = true; break; }
                        if (sgn(r.y)) lost++; else acked++;          // :141-145
                    }
                }
                h2 += (uint32_t)k;
                scanning = (k == avail) && (h2 != tail);
            }
            __syncwarp();
        }
    }
    const uint32_t s_end = h2;
    double extra = 0.0;
    bool has_extra = false;
    Pending m2;
    m2.has = false; m2.idx = 0; m2.t = 0.0; m2.l = 0.0; m2.dropped = false;
    if (owner && at_live) boundary_hop2(tring, h1, h2, tail, s.dl, end, acked, lost, extra, has_extra, m2);

    PCC_PTICK(3);
    // ---- the event that crosses `end` (run_mi phase 4) ---------------------------------------------------------
    if (owner) {
        int which;  // 0 = send, 1 = hop-1, 2 = hop-2
        if (m1.has && (!m2.has || m1.t <= m2.t)) which = (m1.t <= t) ? 1 : 0;
        else if (m2.has) which = (m2.t <= t) ? 2 : 0;
        else which = 0;
        if (which == 0) {
            s.cur_time = t;
            const bool d = rng.next();
            PCC_PACKED_SEND(d, if ((uint32_t)(tail - s_begin) >= rs.cap) ovf = true; else { ring.store(tail, Rec{rec_.x, rec_.y}); tail++; });   /* s_begin, not h2: the consumed records are re-read for the means */
        } else if (which == 1) {
            s.cur_time = m1.t;
            if (m1.idx == h1) h1++; else tring.store_a(m1.idx, negd(m1.t));
        } else {
            s.cur_time = m2.t;
            if (m2.dropped) lost++; else { acked++; extra = m2.l; has_extra = true; }
            if (m2.idx == h2) h2++; else tring.store_a(m2.idx, u2d(PCC_NEG_INF));
        }
    }
#undef PCC_PACKED_SEND
#undef PCC_PACKED_STAGE
    __syncwarp();

    PCC_PTICK(4);
    // ---- phase B3: np.mean of the acked latencies (sender_obs.py:119-122, 138-142), push form ------------------
    out.avg_lat = 0.0; out.lat_inc = 0.0;
    if (WANT_MEANS) {
        const int n = owner ? acked : 0;
        const int half = n / 2;
        const bool inc = need_increase && half >= 1;
        int rn_t[PCC_PW_STACK], rn_h[PCC_PW_STACK];
        double ls_t[PCC_PW_STACK], ls_h[PCC_PW_STACK];
        PwStream<SmemAcc, LocalStack> pt, ph;
        pt.acc.addr = (unsigned)__cvta_generic_to_shared(&sm.acc[0][0][lane]);
        ph.acc.addr = (unsigned)__cvta_generic_to_shared(&sm.acc[1][0][lane]);
        pt.stk.right_n = rn_t; pt.stk.left_sum = ls_t;
        ph.stk.right_n = rn_h; ph.stk.left_sum = ls_h;
        pt.begin(n);
        ph.begin(inc ? half : 0);
        int fed = 0;
        double first = 0.0;
#define PCC_PACKED_PUSH(X)                                                                                     \
        {                                                                                                      \
            const double x_ = (X);                                                                             \
            pt.push(x_);                                                                                       \
            if (inc) {                                                                                         \
                if (fed == half) { first = ph.mean(half); ph.begin(n - half); }                               \
                ph.push(x_);                                                                                   \
            }                                                                                                  \
            fed++;                                                                                             \
        }
        uint32_t i = s_begin;
        bool reading = (n > 0) && (i != s_end);
        // the hop-2 scan took a single tile round: every lane's row still holds its records from s_begin on
        bool reuse = b2_rounds == 1;
        while (__any_sync(PCC_FULL, reading)) {
            if (!reuse) tile_fill(rs, sm, e, i, s_end, reading);
            reuse = false;
            if (reading) {
                const uint32_t left = s_end - i;
                const int avail = left < (uint32_t)PCC_TILE_R ? (int)left : PCC_TILE_R;
                for (int k = 0; k < avail; k++) {
                    const double2 r = sm.tile[lane][k];
                    if (!is_dead(r.x) && !sgn(r.y)) PCC_PACKED_PUSH(r.y + s.dl);   // rtt = fl(ll + dl)
                }
                i += (uint32_t)avail;
                reading = i != s_end;
            }
            __syncwarp();
        }
        if (n > 0) {
            if (has_extra) PCC_PACKED_PUSH(extra);   // the one possible out-of-order sample is the MI's last
            out.avg_lat = pt.mean(n);
            if (inc) out.lat_inc = ph.mean(n - half) - first;
        }
#undef PCC_PACKED_PUSH
    }
    PCC_PTICK(5);
    s.next_send = t; s.qd = q; s.t_upd = tu;
    s.tail = tail; s.h1 = h1; s.h2 = h2;
    out.sent = sent; out.acked = acked; out.lost = lost;
    out.end = s.cur_time;
    out.overflow = ovf;
}

#endif  // __CUDACC__

}  // namespace pcc
