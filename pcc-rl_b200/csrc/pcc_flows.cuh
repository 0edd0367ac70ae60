// pcc_flows.cuh -- CUDA kernels + C ABI of the MI-sample ingestion path ("flow monitor"): batches of
// monitor-interval records from many live flows -> the 12 MI metrics -> per-flow history -> observation.
// Included at the end of pcc_b200.cu (one translation unit, one libpcc_b200.so).  Scalar semantics and the
// reference file:line map are in pcc_flows_core.cuh; the oracle is oracle/pcc_oracle_flows.c.
//
// This path is HBM-bound byte work: a record is ~76 B of fields plus 8 B per RTT sample (CSR), and the
// only arithmetic is numpy's pairwise mean over the samples (three ranges: all, first half, second half),
// a handful of binary64 divisions, and one history row.  Mapping:
//   * an 8-lane subgroup owns a record (4 records per warp): the 8 lanes ARE numpy's 8 accumulators
//     (r[j] += a[8k+j] in numpy's order), three xor-shuffles are its ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
//     tree, so np.mean is bit-exact; every load of a subgroup is a contiguous 64 B run of the sample array;
//   * the divisions of a record are spread over the subgroup's lanes (two rounds of one DDIV each instead
//     of twelve in sequence), the result row and the observation are written by the subgroup cooperatively;
//   * numpy's recursion (n > 128: split at n/2 rounded down to a multiple of 8) is walked per subgroup: records up
//     to 248 samples (ranges at most two leaves deep) take four warp-uniform leaf steps with full-mask shuffles,
//     longer ones a register-resident stack walk, and only lists beyond 1 800 samples a separate whole-warp kernel
//     (four leaves at a time, one per subgroup, folded in numpy's order);
//   * the four records of a pass are consecutive, hence ONE contiguous span of the sample array: a single
//     cp.async.bulk (TMA) per pass stages it in the warp's shared-memory stage, refilled for the next pass while
//     the sample-free tail of the current pass runs (pcc_flows_ingest_tma_kernel; the read-only-path variant
//     pcc_flows_ingest_kernel serves sample arrays that are not 16-byte aligned);
//   * persistent grid (SM count x resident blocks), every warp streams a contiguous run of records.
// State per flow (caller-owned workspace): hist[flow][H][F] (a ring over H with a per-flow head) and one 32-byte
// FlowState (conn-min dict entry, sending rate, head / flags, record counter, batch stamp).  The batch stamp is the
// unique-batch check (a flow seen twice in a batch declared unique is reported by pcc_flows_check); it is a 32-bit
// batch counter, so a flow idle for exactly 2^32 batches could be flagged spuriously.  The CSR offsets are trusted
// to be non-decreasing and within the sample array (negative counts and out-of-range flow ids are reported).
#pragma once
#include "pcc_flows_core.cuh"

namespace pccf {
using namespace pcc;

#define PCCF_THREADS 256
#ifndef PCCF_MINBLOCKS
#define PCCF_MINBLOCKS 4      // 64 registers, 32 warps per SM (measured: 3 -> 1.12 ms, 4 -> 1.04, 5 -> 1.01, 6 -> 1.01 per 1 Mi records)
#endif
#define PCCF_FULL 0xffffffffu
#define PCCF_HEAD_MASK 0xffu
#define PCCF_HAS_MIN 0x100u

// Per-flow scalars, one 32-byte sector per flow (flows arrive in arbitrary order: one random sector per record).
struct __align__(32) FlowState {
    double conn_min;            // _conn_min_latencies[flow] (valid when PCCF_HAS_MIN is set)
    double rate;                // the sending rate (PccGymDriver.rate / ShimNetworkEnv.rate)
    uint32_t flags;             // bits 0-7 head (slot of the OLDEST history row), bit 8 = dict entry exists
    uint32_t n_rec;             // records since the last reset (saturating); got_data = n_rec != 0
    uint32_t stamp;             // batch number of the flow's last record (unique-batch check)
    uint32_t pad;
};

struct FlowsDev {
    double *hist;               // [n_flows][H*F]; row `slot` at slot*F; head = slot of the OLDEST row
    FlowState *st;              // [n_flows]
    unsigned long long *meta;   // [0] duplicate flows in a batch declared unique, [1] flow index out of range
    int64_t n_flows;
    int32_t H, F;
    int32_t ids[PCC_MAX_FEATURES];
    int32_t touch_conn;         // the feature set reads/updates the conn-min entry
    double delta_scale, min_rate, max_rate;
    int32_t rate_style;
};

struct BatchDev {
    int64_t R;
    const int32_t *flow;
    const long long *bytes_sent, *bytes_acked, *bytes_lost, *packet_size;
    const double *send_start, *send_end, *recv_start, *recv_end;
    const long long *off;       // [R + 1]
    const double *rtt;
};

// One leaf of numpy's DOUBLE_pairwise_sum (n <= 128) on an 8-lane subgroup; the result is subgroup-uniform.
// LDG: the samples are read from global memory through the read-only path; !LDG: `a` is a generic pointer (the
// TMA-staged copy in shared memory, or global memory when a span did not fit the stage).
template <bool LDG>
__device__ __forceinline__ double ld_sample(const double *a) { return LDG ? __ldg(a) : *a; }

template <bool LDG>
__device__ __forceinline__ double sg_leaf(const double *__restrict__ a, int n, int j, unsigned mask)
{
    if (n < 8) {
        double res = 0.;
        for (int k = 0; k < n; k++) res += ld_sample<LDG>(a + k);
        return res;
    }
    const int nb = n - (n % 8);
    double r = ld_sample<LDG>(a + j);
#pragma unroll 8
    for (int k = 8; k < nb; k += 8) r += ld_sample<LDG>(a + k + j);
    r += __shfl_xor_sync(mask, r, 1);
    r += __shfl_xor_sync(mask, r, 2);
    r += __shfl_xor_sync(mask, r, 4);
    for (int k = nb; k < n; k++) r += ld_sample<LDG>(a + k);
    return r;
}

// The same leaf for code that all 32 lanes execute together (the four subgroups sum four different leaves of
// possibly different lengths, n == 0 included): no branch around the shuffles, which can then use the full-warp
// mask (a subgroup mask held in a register costs a WARPSYNC.COLLECTIVE bracket per shuffle).  n < 8 is numpy's
// sequential case: the accumulator part is skipped by nb = 0 and the "tail" adds all n elements to 0.0.
template <bool LDG>
__device__ __forceinline__ double sg_leaf_uniform(const double *__restrict__ a, int n, int j)
{
    const int nb = (n >= 8) ? n - (n & 7) : 0;
    double r = 0.0;
    if (nb) r = ld_sample<LDG>(a + j);
#pragma unroll 8
    for (int k = 8; k < nb; k += 8) r += ld_sample<LDG>(a + k + j);
    r += __shfl_xor_sync(PCCF_FULL, r, 1);
    r += __shfl_xor_sync(PCCF_FULL, r, 2);
    r += __shfl_xor_sync(PCCF_FULL, r, 4);
    const int nt = n - nb;                       // 0..7 elements, added in order
    const double *t = a + nb;
#pragma unroll
    for (int k = 0; k < 7; k++) if (k < nt) r += ld_sample<LDG>(t + k);
    return r;
}

// numpy's pairwise sum of a[0..n), n <= PCCF_SG_MAX_N, by ONE subgroup (subgroup-uniform control flow and result).
// Iterative post-order walk of the recursion (n > 128: n2 = n/2 rounded down to a multiple of 8; sum(left n2) +
// sum(right n - n2)); leaves tile [0, n) left to right.  A node at depth d holds at most n/2^d + 15 elements, so
// for n <= 1800 the walk is at most 4 frames deep: the frames live in registers (selected by compare chains),
// not in local memory.  Longer lists go to warp_pw_sum.
#define PCCF_SG_MAX_N 1800
template <bool LDG>
__device__ __forceinline__ double sg_pw_sum(const double *__restrict__ a, int n, int j, unsigned mask)
{
    if (n <= 128) return sg_leaf<LDG>(a, n, j, mask);
    int rn0 = 0, rn1 = 0, rn2 = 0, rn3 = 0;          // right-child sizes of the open frames
    double ls0 = 0., ls1 = 0., ls2 = 0., ls3 = 0.;   // left-child sums of the open frames
    unsigned have_left = 0u;
    int sp = 0, cur = n;
    const double *p = a;
    for (;;) {
        while (cur > 128) {
            int n2 = cur >> 1;
            n2 -= n2 & 7;
            const int rn = cur - n2;
            if (sp == 0) rn0 = rn; else if (sp == 1) rn1 = rn; else if (sp == 2) rn2 = rn; else rn3 = rn;
            have_left &= ~(1u << sp);
            sp++;
            cur = n2;
        }
        double res = sg_leaf<LDG>(p, cur, j, mask);
        p += cur;
        for (;;) {
            if (sp == 0) return res;
            const int t = sp - 1;
            if (!((have_left >> t) & 1u)) {
                if (t == 0) ls0 = res; else if (t == 1) ls1 = res; else if (t == 2) ls2 = res; else ls3 = res;
                have_left |= 1u << t;
                cur = (t == 0) ? rn0 : (t == 1) ? rn1 : (t == 2) ? rn2 : rn3;
                break;
            }
            res = ((t == 0) ? ls0 : (t == 1) ? ls1 : (t == 2) ? ls2 : ls3) + res;
            sp--;
        }
    }
}

// numpy's pairwise sum of a[0..n) for any n, by the whole warp (warp-uniform control flow and result).
// The recursion (n > 128: n2 = n/2 rounded down to a multiple of 8; sum(left n2) + sum(right n - n2)) is
// walked twice in lock step: a structure-only walk runs ahead and names the next four leaves -- leaves
// tile [0, n) left to right -- the subgroups sum them, and the value walk folds the sums in post-order.
struct PwWalk {
    int right_n[32];
    unsigned have_left;
    int sp, cur;
    __device__ __forceinline__ void init(int n) { sp = 0; cur = n; have_left = 0u; }
    __device__ __forceinline__ int next_leaf()     // descend to the next leaf, return its size
    {
        while (cur > 128) {
            int n2 = cur / 2;
            n2 -= n2 % 8;
            right_n[sp] = cur - n2; have_left &= ~(1u << sp); sp++;
            cur = n2;
        }
        return cur;
    }
    // after a leaf (or a completed subtree): climb; returns true when the whole sum is complete
    template <class OnLeft, class OnJoin>
    __device__ __forceinline__ bool climb(OnLeft on_left, OnJoin on_join)
    {
        for (;;) {
            if (sp == 0) return true;
            if (!(have_left & (1u << (sp - 1)))) {
                on_left(sp - 1);
                have_left |= 1u << (sp - 1);
                cur = right_n[sp - 1];
                return false;
            }
            on_join(sp - 1);
            sp--;
        }
    }
};

__device__ __noinline__ double warp_pw_sum(const double *__restrict__ a, long long n)
{
    const unsigned lane = threadIdx.x & 31u;
    const int sg = (int)(lane >> 3), j = (int)(lane & 7u);
    PwWalk ahead, fold;
    ahead.init((int)n);
    fold.init((int)n);
    double left_sum[32];
    long long off = 0;
    bool ahead_done = (n <= 0), done = (n <= 0);
    double result = 0.0;
    while (!done) {
        long long lo[4]; int lc[4]; int cnt = 0;
        while (cnt < 4 && !ahead_done) {
            const int c = ahead.next_leaf();
            lo[cnt] = off; lc[cnt] = c; cnt++;
            off += c;
            ahead_done = ahead.climb([](int) {}, [](int) {});
        }
        double mine = 0.0;
        {
            long long o = 0; int c = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) if (k == sg && k < cnt) { o = lo[k]; c = lc[k]; }
            mine = sg_leaf<true>(a + o, c, j, 0xffu << (sg * 8));   // subgroups diverge (idle ones have c == 0)
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double leaf = __shfl_sync(PCCF_FULL, mine, k * 8);
            if (k < cnt && !done) {
                fold.next_leaf();
                double res = leaf;
                done = fold.climb([&](int s) { left_sum[s] = res; },
                                  [&](int s) { res = left_sum[s] + res; });
                if (done) result = res;
            }
        }
    }
    return result;
}

// The three pairwise sums of a record (all, first half, second half) for the four subgroups of a warp together.
// Called by all 32 lanes.  A record whose ranges are at most two leaves deep (n <= 248: all = leaf + leaf,
// halves = one leaf each) takes four warp-uniform leaf steps; anything longer walks numpy's recursion per
// subgroup (sg_pw_sum).  `on` = the subgroup has a record to sum.
#define PCCF_FLAT_MAX_N 248
template <bool LDG>
__device__ __forceinline__ void pass_sums(const double *a, long long n, bool on, int j, unsigned sgmask,
                                          double &sum, double &s1, double &s2)
{
    sum = 0.0; s1 = 0.0; s2 = 0.0;
    const int nn = (int)n;
    const int half = nn / 2;
    const bool flat = on && nn <= PCCF_FLAT_MAX_N;
    int n2 = nn >> 1;
    n2 -= n2 & 7;                                    // numpy's split of the whole range when n > 128
    const int c0 = !flat ? 0 : (nn > 128 ? n2 : nn);
    const int c1 = !flat ? 0 : (nn > 128 ? nn - n2 : 0);
    const double l0 = sg_leaf_uniform<LDG>(a, c0, j);
    const double l1 = sg_leaf_uniform<LDG>(a + c0, c1, j);
    const double l2 = sg_leaf_uniform<LDG>(a, flat ? half : 0, j);
    const double l3 = sg_leaf_uniform<LDG>(a + half, flat ? nn - half : 0, j);
    if (flat) {
        sum = (nn > 128) ? l0 + l1 : l0;
        s1 = l2; s2 = l3;
    } else if (on) {                                 // subgroup-divergent: deeper recursions
        sum = sg_pw_sum<LDG>(a, nn, j, sgmask);
        if (half >= 1) { s1 = sg_pw_sum<LDG>(a, half, j, sgmask); s2 = sg_pw_sum<LDG>(a + half, nn - half, j, sgmask); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Ingest kernel.  UNIQUE: every flow appears at most once in the batch (checked): metrics, conn-min entry,
// history row and observation in this one kernel.  !UNIQUE: the per-record part only; rows[R][F] (scaled
// feature values) and avg[R] are handed to pcc_flows_apply_kernel, which walks each flow's records in batch
// order.
// ---------------------------------------------------------------------------------------------------------
// Everything of a record after its three pairwise sums: fields, the two rounds of lane-parallel divisions, the
// conn-min dict entry, the history row, the observation (UNIQUE) or the hand-over to the apply kernel (!UNIQUE).
// Called with the four subgroups of a warp converged; `good` is subgroup-uniform.
template <bool UNIQUE>
__device__ __forceinline__ void flows_finish_record(const FlowsDev &p, const BatchDev &b, long long r, int flow, long long n,
                                                    double sum, double s1, double s2, bool good, int j, int sg0,
                                                    unsigned sgmask, uint32_t batch_no, double *__restrict__ obs,
                                                    double *__restrict__ metrics, double *__restrict__ rows,
                                                    double *__restrict__ avg_out, double *xs)
{
    // xs: 16 doubles of shared memory private to the subgroup (the new history row, handed to the lanes that
    // write the observation)
    const int H = p.H, F = p.F, HF = H * F;
    const long long half = n / 2;
    // ---- per-flow state and the old history row: requested first, consumed after the divisions -----------
    uint32_t fl = 0; double cmin = 0.0, cm = 0.0;
    bool has_min = false;
    uint32_t n_rec = 0;
    double hold[4] = {0.0, 0.0, 0.0, 0.0};
    int rot = 0;
    uint32_t head = 0, nhead = 0;
    if (UNIQUE && good) {
        const FlowState fs = p.st[flow];              // one 32-byte sector
        fl = fs.flags; cmin = fs.conn_min; n_rec = fs.n_rec;
        has_min = (fl & PCCF_HAS_MIN) != 0;
        head = fl & PCCF_HEAD_MASK;
        nhead = (head + 1 == (uint32_t)H) ? 0u : head + 1;
        rot = (int)nhead * F;
        if (obs && HF <= 32) {                        // as_array (:68-73): oldest row first = the ring rotated by nhead
            const double *hrow = p.hist + (size_t)flow * HF;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int k = j + 8 * t;
                int src = k + rot;
                if (src >= HF) src -= HF;
                if (k < HF) hold[t] = hrow[src];
            }
        }
    }
    // ---- the record's scalar fields (loaded after the sums: they are not live across the sample passes) --
    long long bs = 0, ba = 0, bl = 0, ps = 0;
    double ss = 0., se = 0., rs = 0., re = 0.;
    if (good) {
        bs = __ldg(b.bytes_sent + r); ba = __ldg(b.bytes_acked + r); bl = __ldg(b.bytes_lost + r);
        ps = __ldg(b.packet_size + r);
        ss = __ldg(b.send_start + r); se = __ldg(b.send_end + r);
        rs = __ldg(b.recv_start + r); re = __ldg(b.recv_end + r);
    }

    // ---- round 1 of divisions: one per lane (sender_obs.py:110-142) ---------------------------------
    const double sdur = se - ss, rdur = re - rs;
    double num = 0.0, den = 1.0;
    bool ok = false;
    if (j == 0) { num = 8.0 * (double)bs; den = sdur; ok = sdur > 0.0; }                    // send rate
    else if (j == 1) { num = 8.0 * (double)(ba - ps); den = rdur; ok = rdur > 0.0; }        // recv rate
    else if (j == 2) { num = 0.0 + sum; den = (double)n; ok = n > 0; }                      // np.mean(all)
    else if (j == 3) { num = 0.0 + s1; den = (double)half; ok = half >= 1; }                // np.mean(first half)
    else if (j == 4) { num = 0.0 + s2; den = (double)(n - half); ok = half >= 1; }          // np.mean(second half)
    else if (j == 5) { num = (double)bl; den = (double)(bl + ba); ok = (bl + ba) > 0; }     // loss ratio
    double q = num / (ok ? den : 1.0);
    if (!ok) q = 0.0;
    const double send_rate = __shfl_sync(PCCF_FULL, q, sg0 + 0);
    const double recv_rate = __shfl_sync(PCCF_FULL, q, sg0 + 1);
    const double avg = __shfl_sync(PCCF_FULL, q, sg0 + 2);
    const double m1 = __shfl_sync(PCCF_FULL, q, sg0 + 3);
    const double m2 = __shfl_sync(PCCF_FULL, q, sg0 + 4);
    const double loss = __shfl_sync(PCCF_FULL, q, sg0 + 5);
    const double inc = (half >= 1) ? m2 - m1 : 0.0;

    // ---- conn-min dict entry (:158-176) -----------------------------------------------------------------
    if (UNIQUE && good) cm = flow_conn_min(avg, has_min, cmin, p.touch_conn != 0);

    // ---- round 2 --------------------------------------------------------------------------------------
    double dflt = 0.0;
    num = 0.0; den = 1.0; ok = false;
    if (j == 0) { num = inc; den = rdur; ok = rdur > 0.0; }                                  // ack latency inflation
    else if (j == 1) { num = inc; den = sdur; ok = sdur > 0.0; }                             // sent latency inflation
    else if (j == 2) { num = avg; den = cm; ok = cm > 0.0; dflt = 1.0; }                     // latency ratio
    else if (j == 3) { num = send_rate; den = recv_rate; dflt = 1.0;
                       ok = recv_rate > 0.0 && send_rate < 1000.0 * recv_rate; }             // send ratio
    else if (j == 4) { num = send_rate; den = 1e7; ok = true; }                              // scaled rates
    else if (j == 5) { num = recv_rate; den = 1e7; ok = true; }
    q = num / (ok ? den : 1.0);
    if (!ok) q = dflt;
    const double ack_infl = __shfl_sync(PCCF_FULL, q, sg0 + 0);
    const double sent_infl = __shfl_sync(PCCF_FULL, q, sg0 + 1);
    const double lat_ratio = __shfl_sync(PCCF_FULL, q, sg0 + 2);
    const double send_ratio = __shfl_sync(PCCF_FULL, q, sg0 + 3);
    const double send_rate_s = __shfl_sync(PCCF_FULL, q, sg0 + 4);
    const double recv_rate_s = __shfl_sync(PCCF_FULL, q, sg0 + 5);

    auto raw = [&](int id) -> double {
        switch (id) {
        case M_SEND_RATE: return send_rate;
        case M_RECV_RATE: return recv_rate;
        case M_RECV_DUR: return rdur;
        case M_SEND_DUR: return sdur;
        case M_AVG_LATENCY: return avg;
        case M_LOSS_RATIO: return loss;
        case M_ACK_LAT_INFL: return ack_infl;
        case M_SENT_LAT_INFL: return sent_infl;
        case M_CONN_MIN_LAT: return cm;
        case M_LAT_INCREASE: return inc;
        case M_LAT_RATIO: return lat_ratio;
        default: return send_ratio;
        }
    };
    if (!good) return;                                                  // (no warp-wide sync below this line)

    if (metrics) {
        metrics[r * N_METRICS + j] = raw(j);
        if (j + 8 < N_METRICS) metrics[r * N_METRICS + j + 8] = raw(j + 8);
    }
    if (UNIQUE) {
        double *hrow = p.hist + (size_t)flow * HF;
        for (int f = j; f < F; f += 8) {                                // SenderHistory.step (:64-66)
            const int id = p.ids[f];
            const double v = (id == M_SEND_RATE) ? send_rate_s : (id == M_RECV_RATE) ? recv_rate_s : raw(id);
            hrow[head * F + f] = v;
            xs[f] = v;
        }
        if (j == 0) {
            FlowState *fs = p.st + flow;
            if (atomicExch(&fs->stamp, batch_no) == batch_no) atomicAdd(&p.meta[0], 1ull);
            fs->flags = nhead | (has_min ? PCCF_HAS_MIN : 0u);
            if (p.touch_conn) fs->conn_min = cmin;
            if (n_rec != 0xffffffffu) fs->n_rec = n_rec + 1;
        }
        if (obs) {                                                      // as_array (:68-73): oldest row first
            __syncwarp(sgmask);
            double *ob = obs + (size_t)r * HF;
            const int lo = (int)head * F;                               // the slot that now holds the new row
            if (HF <= 32) {                                             // old rows were requested at the top
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int k = j + 8 * t;
                    int src = k + rot;
                    if (src >= HF) src -= HF;
                    if (k < HF) ob[k] = (src >= lo && src < lo + F) ? xs[src - lo] : hold[t];
                }
            } else {
                for (int k = j; k < HF; k += 8) {
                    int src = k + rot;
                    if (src >= HF) src -= HF;
                    ob[k] = hrow[src];
                }
            }
        }
    } else {
        for (int f = j; f < F; f += 8) {
            const int id = p.ids[f];
            rows[(size_t)r * F + f] = (id == M_SEND_RATE) ? send_rate_s : (id == M_RECV_RATE) ? recv_rate_s : raw(id);
        }
        if (j == 0) avg_out[r] = avg;
    }
}

template <bool UNIQUE>
__global__ void __launch_bounds__(PCCF_THREADS, PCCF_MINBLOCKS)
pcc_flows_ingest_kernel(FlowsDev p, BatchDev b, uint32_t batch_no, double *__restrict__ obs, double *__restrict__ metrics,
                        double *__restrict__ rows, double *__restrict__ avg_out)
{
    const unsigned lane = threadIdx.x & 31u;
    const int sg = (int)(lane >> 3), j = (int)(lane & 7u);
    const unsigned sgmask = 0xffu << (sg * 8);
    const int sg0 = sg * 8;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    __shared__ double xchg[(PCCF_THREADS / 32) * 4 * 16];

    // A warp owns a contiguous run of records (4 per pass, one per subgroup): its offsets, fields and samples are
    // one sequential stream; the next pass's offsets and flow id are loaded a pass ahead.  (Measured: an additional
    // prefetch.global.L2 of the next pass's sample span bought nothing and cost 35 % more DRAM reads.)
    long long chunk = (b.R + nwarps - 1) / nwarps;
    chunk = (chunk + 3) & ~3ll;
    const long long wbeg = warp * chunk;
    const long long wend = (wbeg + chunk < b.R) ? wbeg + chunk : b.R;
    long long o0n = 0, o1n = 0;
    int flown = 0;
    if (wbeg + sg < wend) {
        o0n = __ldg(b.off + wbeg + sg); o1n = __ldg(b.off + wbeg + sg + 1);
        flown = __ldg(b.flow + wbeg + sg);
    }
    for (long long base = wbeg; base < wend; base += 4) {                  // warp-uniform
        const long long r = base + sg;
        const bool valid = r < wend;
        const long long n0 = o1n - o0n;
        long long n = valid ? n0 : 0;
        const double *a = b.rtt + (valid ? o0n : 0);
        const int flow = valid ? flown : 0;
        {   // the next pass: load its offsets / flow id now, prefetch what it will touch
            const long long rn = r + 4;
            if (rn < wend) {
                o0n = __ldg(b.off + rn); o1n = __ldg(b.off + rn + 1);
                flown = __ldg(b.flow + rn);
            }
        }
        bool good = valid && flow >= 0 && (long long)flow < p.n_flows && n >= 0;
        if (valid && !good && j == 0) atomicAdd(&p.meta[1], 1ull);
        if (!good) n = 0;

        // ---- the three pairwise sums (all, first half, second half) ------------------------------------------
        // The halves re-read the samples the first range pulled from HBM (L1/L2 hits; numpy's summation trees of
        // the three ranges share nothing, so the sums cannot be shared either).
        double sum, s1, s2;
        if (n > PCCF_SG_MAX_N) good = false;      // a very long sample list: left to pcc_flows_long_kernel
        pass_sums<true>(a, n, good && n > 0, j, sgmask, sum, s1, s2);
        flows_finish_record<UNIQUE>(p, b, r, flow, n, sum, s1, s2, good, j, sg0, sgmask, batch_no, obs, metrics, rows, avg_out,
                                    xchg + ((threadIdx.x >> 5) * 4 + sg) * 16);
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same ingest with the sample stream staged by TMA.  The four records of a pass are consecutive, so their
// samples are ONE contiguous span of the CSR array: lane 0 issues a single cp.async.bulk (global -> shared,
// completion counted in bytes on an mbarrier) per pass.  One stage per warp: as soon as the three summation
// passes have read the stage, the copy of the NEXT pass is started, and it lands while the warp does the part of
// the current pass that needs no samples (fields, divisions, history row, observation -- 60 % of the
// instructions).  No registers are tied up by loads in flight and the summation passes read shared memory.
// cp.async.bulk moves 16-byte granules from a 16-byte aligned address: an odd first element index starts the
// copy one element early, an odd element count copies one element more (the last element of the whole array is
// fetched by hand instead of reading past it).  A span that does not fit the stage is read from global memory
// by the same code (generic pointers).
// ---------------------------------------------------------------------------------------------------------
#ifndef PCCF_TMA_WARPS
#define PCCF_TMA_WARPS 4         // warps per block
#endif
#ifndef PCCF_TMA_SLOT
#define PCCF_TMA_SLOT 768        // doubles per warp stage (6 KB): span of a pass + 1 must fit
#endif
#ifndef PCCF_TMA_MINBLOCKS
#define PCCF_TMA_MINBLOCKS 7     // 7 blocks x 4 warps per SM: 72 registers, 7 x 26.6 KB of shared memory
                                 // (measured per 1 Mi records: 6 -> 0.655 ms, 7 -> 0.641, 8 -> 0.665)
#endif
#define PCCF_TMA_SMEM ((size_t)PCCF_TMA_WARPS * PCCF_TMA_SLOT * 8 + (size_t)PCCF_TMA_WARPS * 8 + (size_t)PCCF_TMA_WARPS * 4 * 16 * 8)

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool UNIQUE>
__global__ void __launch_bounds__(PCCF_TMA_WARPS * 32, PCCF_TMA_MINBLOCKS)
pcc_flows_ingest_tma_kernel(FlowsDev p, BatchDev b, uint32_t batch_no, double *__restrict__ obs, double *__restrict__ metrics,
                            double *__restrict__ rows, double *__restrict__ avg_out)
{
    extern __shared__ __align__(128) unsigned char pccf_smem[];
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const int sg = (int)(lane >> 3), j = (int)(lane & 7u);
    const unsigned sgmask = 0xffu << (sg * 8);
    const int sg0 = sg * 8;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double *stage = reinterpret_cast<double *>(pccf_smem) + (size_t)wib * PCCF_TMA_SLOT;
    uint64_t *bar = reinterpret_cast<uint64_t *>(pccf_smem + (size_t)PCCF_TMA_WARPS * PCCF_TMA_SLOT * 8) + wib;
    double *xs = reinterpret_cast<double *>(pccf_smem + (size_t)PCCF_TMA_WARPS * PCCF_TMA_SLOT * 8 + (size_t)PCCF_TMA_WARPS * 8) +
                 (wib * 4 + sg) * 16;
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
    }
    __syncwarp();

    long long chunk = (b.R + nwarps - 1) / nwarps;
    chunk = (chunk + 3) & ~3ll;
    const long long wbeg = warp * chunk;
    const long long wend = (wbeg + chunk < b.R) ? wbeg + chunk : b.R;
    if (wbeg >= wend) return;                                              // warp-uniform
    const long long total = __ldg(b.off + b.R);                           // elements in the sample array

    // Starts the copy of the pass beginning at record `base` (offsets of its records in o0x / o1x of the subgroups).
    // Returns the element index the stage then starts at, or -1 when the pass has to read global memory.
    auto issue = [&](long long base, long long o0x, long long o1x) -> long long {
        const int nvalid = (wend - base >= 4) ? 4 : (int)(wend - base);
        const long long e0 = __shfl_sync(PCCF_FULL, o0x, 0);
        const long long e1 = __shfl_sync(PCCF_FULL, o1x, (nvalid - 1) * 8);
        const long long e0a = e0 & ~1ll;
        const long long nel = e1 - e0a;
        const bool ok = e0 >= 0 && nel >= 0 && nel + 1 <= PCCF_TMA_SLOT && e1 <= total;
        if (lane == 0) {
            long long ncopy = ok ? ((nel + 1) & ~1ll) : 0;
            if (ncopy > 0 && e0a + ncopy > total) {                        // never read past the end of the array
                ncopy -= 2;
                stage[nel - 1] = __ldg(b.rtt + e0a + nel - 1);
            }
            if (ncopy > 0) {
                // the stage was last READ through the generic proxy (all lanes are past __syncwarp); the bulk copy
                // writes it through the async proxy: order the two proxies before re-using the buffer
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, (uint32_t)(ncopy * 8));
                tma_load_1d(stage, b.rtt + e0a, (uint32_t)(ncopy * 8), bar);
            } else {
                mbar_arrive(bar);
            }
        }
        __syncwarp();
        return ok ? e0a : -1;
    };

    // offsets / flow id of this pass (c) and of the next one (n), per subgroup
    long long o0c = 0, o1c = 0, o0n = 0, o1n = 0;
    int flowc = 0, flown = 0;
    if (wbeg + sg < wend) { o0c = __ldg(b.off + wbeg + sg); o1c = __ldg(b.off + wbeg + sg + 1); flowc = __ldg(b.flow + wbeg + sg); }
    if (wbeg + 4 + sg < wend) { o0n = __ldg(b.off + wbeg + 4 + sg); o1n = __ldg(b.off + wbeg + 4 + sg + 1); flown = __ldg(b.flow + wbeg + 4 + sg); }
    long long e0a_c = issue(wbeg, o0c, o1c);

    int pass = 0;
    for (long long base = wbeg; base < wend; base += 4, pass++) {          // warp-uniform
        const long long r = base + sg;
        const bool valid = r < wend;
        // the pass after the next: offsets / flow id into registers (consumed two passes from now)
        long long o0f = 0, o1f = 0;
        int flowf = 0;
        if (r + 8 < wend) { o0f = __ldg(b.off + r + 8); o1f = __ldg(b.off + r + 9); flowf = __ldg(b.flow + r + 8); }
        long long n = valid ? o1c - o0c : 0;
        const int flow = valid ? flowc : 0;
        bool good = valid && flow >= 0 && (long long)flow < p.n_flows && n >= 0;
        if (valid && !good && j == 0) atomicAdd(&p.meta[1], 1ull);
        if (!good) n = 0;
        if (n > PCCF_SG_MAX_N) good = false;                               // left to pcc_flows_long_kernel
        mbar_wait(bar, (uint32_t)(pass & 1));
        const double *a = (e0a_c >= 0) ? stage + (o0c - e0a_c) : b.rtt + o0c;
        double sum, s1, s2;
        pass_sums<false>(a, n, good && n > 0, j, sgmask, sum, s1, s2);
        __syncwarp();                                                      // the stage has been read: refill it ...
        long long e0a_n = -1;
        if (base + 4 < wend) e0a_n = issue(base + 4, o0n, o1n);
        // ... while this pass's records are finished
        flows_finish_record<UNIQUE>(p, b, r, flow, n, sum, s1, s2, good, j, sg0, sgmask, batch_no, obs, metrics, rows, avg_out, xs);
        o0c = o0n; o1c = o1n; flowc = flown; e0a_c = e0a_n;
        o0n = o0f; o1n = o1f; flown = flowf;
    }
}

// Records with more than PCCF_SG_MAX_N samples (rare): one warp per record, numpy's recursion walked by the whole
// warp (warp_pw_sum), the scalar part by the core's flow_stats_finish on every lane redundantly.  Launched after
// the ingest kernel on the same stream; a warp looks at 32 consecutive records and usually finds nothing.
template <bool UNIQUE>
__global__ void __launch_bounds__(128)
pcc_flows_long_kernel(FlowsDev p, BatchDev b, uint32_t batch_no, double *__restrict__ obs, double *__restrict__ metrics,
                      double *__restrict__ rows, double *__restrict__ avg_out)
{
    const unsigned lane = threadIdx.x & 31u;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int H = p.H, F = p.F, HF = H * F;
    // the batch counter travels with the workspace: a handle attached to it later (checkpoint / resume) continues the
    // numbering, so the per-flow stamps of earlier batches can never equal a new batch's number
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta[2] = (unsigned long long)batch_no;
    for (long long base = warp * 32; base < b.R; base += nwarps * 32) {
        const long long mine = base + lane;
        bool is_long = false;
        if (mine < b.R) {
            const long long nn = __ldg(b.off + mine + 1) - __ldg(b.off + mine);
            const int fl = __ldg(b.flow + mine);
            is_long = nn > PCCF_SG_MAX_N && fl >= 0 && (long long)fl < p.n_flows;
        }
        unsigned m = __ballot_sync(PCCF_FULL, is_long);
        while (m) {                                                          // warp-uniform
            const long long r = base + (__ffs(m) - 1);
            m &= m - 1;
            const long long o0 = __ldg(b.off + r), n = __ldg(b.off + r + 1) - o0;
            const double *a = b.rtt + o0;
            const long long half = n / 2;
            const double sum = warp_pw_sum(a, n), s1 = warp_pw_sum(a, half), s2 = warp_pw_sum(a + half, n - half);
            const double avg = (0.0 + sum) / (double)n;
            const double m1 = (0.0 + s1) / (double)half, m2 = (0.0 + s2) / (double)(n - half);
            FlowRecord rec;
            rec.bytes_sent = __ldg(b.bytes_sent + r); rec.bytes_acked = __ldg(b.bytes_acked + r);
            rec.bytes_lost = __ldg(b.bytes_lost + r); rec.packet_size = __ldg(b.packet_size + r);
            rec.send_start = __ldg(b.send_start + r); rec.send_end = __ldg(b.send_end + r);
            rec.recv_start = __ldg(b.recv_start + r); rec.recv_end = __ldg(b.recv_end + r);
            const int flow = __ldg(b.flow + r);
            uint32_t fl = 0; double cmin = 0.0;
            bool has_min = false;
            uint32_t n_rec = 0;
            if (UNIQUE) { const FlowState fs = p.st[flow]; fl = fs.flags; cmin = fs.conn_min; n_rec = fs.n_rec; has_min = (fl & PCCF_HAS_MIN) != 0; }
            FlowStats st;
            flow_stats_finish(rec, n, avg, m1, m2, has_min, cmin, UNIQUE && p.touch_conn != 0, st);
            __syncwarp();
            if (metrics && lane < N_METRICS) metrics[r * N_METRICS + lane] = st.v[lane];
            if (UNIQUE) {
                const uint32_t head = fl & PCCF_HEAD_MASK;
                double *hrow = p.hist + (size_t)flow * HF;
                if ((int)lane < F) hrow[head * F + lane] = st.v[p.ids[lane]] / flow_metric_scale(p.ids[lane]);
                const uint32_t nhead = (head + 1 == (uint32_t)H) ? 0u : head + 1;
                if (lane == 0) {
                    FlowState *fs = p.st + flow;
                    if (atomicExch(&fs->stamp, batch_no) == batch_no) atomicAdd(&p.meta[0], 1ull);
                    fs->flags = nhead | (has_min ? PCCF_HAS_MIN : 0u);
                    if (p.touch_conn) fs->conn_min = cmin;
                    if (n_rec != 0xffffffffu) fs->n_rec = n_rec + 1;
                }
                if (obs) {
                    __syncwarp();
                    double *ob = obs + (size_t)r * HF;
                    const int rot = (int)nhead * F;
                    for (int k = (int)lane; k < HF; k += 32) {
                        int src = k + rot;
                        if (src >= HF) src -= HF;
                        ob[k] = hrow[src];
                    }
                }
            } else {
                if ((int)lane < F) rows[(size_t)r * F + lane] = st.v[p.ids[lane]] / flow_metric_scale(p.ids[lane]);
                if (lane == 0) avg_out[r] = avg;
            }
            __syncwarp();
        }
    }
}

// General batches: position i of the flow-sorted (stable) record list; the first record of each flow's run
// applies the whole run in batch order.
__global__ void pcc_flows_apply_kernel(FlowsDev p, int64_t R, const int32_t *__restrict__ sorted_flow,
                                       const int32_t *__restrict__ sorted_rec, const double *__restrict__ rows,
                                       const double *__restrict__ avg, double *__restrict__ obs,
                                       double *__restrict__ metrics)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const int flow = sorted_flow[i];
    if (flow < 0 || (int64_t)flow >= p.n_flows) return;
    if (i > 0 && sorted_flow[i - 1] == flow) return;
    const int H = p.H, F = p.F, HF = H * F;
    FlowState *fs = p.st + flow;
    uint32_t fl = fs->flags;
    double cmin = fs->conn_min;
    bool has_min = (fl & PCCF_HAS_MIN) != 0;
    uint32_t head = fl & PCCF_HEAD_MASK;
    double *hrow = p.hist + (size_t)flow * HF;
    uint32_t cnt = fs->n_rec;
    for (int64_t k = i; k < R && sorted_flow[k] == flow; k++) {
        const int64_t r = sorted_rec[k];
        const double a = avg[r];
        const double cm = flow_conn_min(a, has_min, cmin, p.touch_conn != 0);
        const double lat_ratio = (cm > 0.0) ? a / cm : 1.0;
        for (int f = 0; f < F; f++) {
            const int id = p.ids[f];
            double v = rows[(size_t)r * F + f];
            if (id == M_CONN_MIN_LAT) v = cm;
            else if (id == M_LAT_RATIO) v = lat_ratio;
            hrow[head * F + f] = v;
        }
        if (metrics) { metrics[r * N_METRICS + M_CONN_MIN_LAT] = cm; metrics[r * N_METRICS + M_LAT_RATIO] = lat_ratio; }
        head = (head + 1 == (uint32_t)H) ? 0u : head + 1;
        if (cnt != 0xffffffffu) cnt++;
        if (obs) {
            double *ob = obs + (size_t)r * HF;
            const int rot = (int)head * F;
            for (int q = 0; q < HF; q++) { int src = q + rot; if (src >= HF) src -= HF; ob[q] = hrow[src]; }
        }
    }
    fs->flags = head | (has_min ? PCCF_HAS_MIN : 0u);
    if (p.touch_conn) fs->conn_min = cmin;
    fs->n_rec = cnt;
}

__global__ void pcc_flows_iota_kernel(int32_t *__restrict__ v, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// History reset (mode: FLOW_RESET_*), one thread per flow.
__global__ void pcc_flows_reset_kernel(FlowsDev p, const uint8_t *__restrict__ mask, int mode)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_flows || (mask && !mask[e])) return;
    uint32_t fl = p.st[e].flags;
    double cmin = p.st[e].conn_min;
    if (mode == FLOW_RESET_NEW) { fl &= ~PCCF_HAS_MIN; cmin = 0.0; p.st[e].conn_min = 0.0; }
    const bool seen = (mode == FLOW_RESET_CLIENT) && (fl & PCCF_HAS_MIN);
    double *hrow = p.hist + (size_t)e * p.H * p.F;
    for (int h = 0; h < p.H; h++)
        for (int f = 0; f < p.F; f++) {
            const int id = p.ids[f];
            hrow[h * p.F + f] = flow_metric_empty(id, seen, cmin) / flow_metric_scale(id);
        }
    p.st[e].flags = fl & PCCF_HAS_MIN;     // head = 0
    p.st[e].n_rec = 0u;
}

// history.as_array() of every flow: obs[flow][H*F], oldest row first
__global__ void pcc_flows_obs_kernel(FlowsDev p, double *__restrict__ obs)
{
    const int HF = p.H * p.F;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_flows * HF) return;
    const int64_t e = i / HF;
    const int k = (int)(i - e * HF);
    int src = k + (int)(p.st[e].flags & PCCF_HEAD_MASK) * p.F;
    if (src >= HF) src -= HF;
    obs[i] = p.hist[(size_t)e * HF + src];
}

// PccGymDriver.get_rate (loaded_client.py:72-76): flows that have data apply the agent's action; shim style
// (ShimNetworkEnv.step, shim_env.py:107): every selected flow applies it.
__global__ void pcc_flows_rate_kernel(FlowsDev p, const double *__restrict__ actions, const uint8_t *__restrict__ mask,
                                      double *__restrict__ rates_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_flows) return;
    double rate = p.st[e].rate;
    const bool sel = !mask || mask[e];
    if (actions && sel && (p.rate_style == RATE_STYLE_SHIM || p.st[e].n_rec != 0u)) {
        rate = flow_apply_rate_delta(rate, actions[e], p.delta_scale, p.min_rate, p.max_rate, p.rate_style);
        p.st[e].rate = rate;
    }
    if (rates_out) rates_out[e] = rate;
}

__global__ void pcc_flows_set_rate_kernel(FlowsDev p, const uint8_t *__restrict__ mask, const double *__restrict__ rates,
                                          double scalar)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_flows || (mask && !mask[e])) return;
    p.st[e].rate = rates ? rates[e] : scalar;
}

// one column of the per-flow state: 0 = conn_min (0.0 when there is no dict entry), 1 = rate
__global__ void pcc_flows_column_kernel(FlowsDev p, int which, double *__restrict__ dst)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_flows) return;
    dst[e] = (which == 0) ? ((p.st[e].flags & PCCF_HAS_MIN) ? p.st[e].conn_min : 0.0) : p.st[e].rate;
}

}  // namespace pccf

// =========================================================================================================
// C ABI (include/pcc_b200.h, "MI-sample ingestion")
// =========================================================================================================
using namespace pccf;

// the agent on the device: action[flow] = MLP(history.as_array()) (loaded_agent.LoadedModelAgent.act with
// stochastic=False, loaded_client.py:72-75), the network of stable_solve.py:30-45
__global__ void pcc_flows_act_kernel(FlowsDev p, PolicyDev pol, double *__restrict__ actions)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_flows) return;
    const int head = (int)(p.st[e].flags & PCCF_HEAD_MASK);
    actions[e] = policy_action(pol, p.hist + (size_t)e * p.H * p.F, head, p.H, p.F, e, 0ull);
}

struct pcc_flows_handle_s {
    pcc_flows_config cfg;
    FlowsDev d;
    uint32_t batch_no;
    int64_t launches;
    int sm_count;
    int use_tma;          // 1: TMA-staged ingest kernel (default); 0: read-only-path loads (PCC_FLOWS_KERNEL=ldg)
    int tma_blocks_per_sm;
};

static void flows_layout(const pcc_flows_config *c, size_t off[8], size_t &total)
{
    const size_t n = (size_t)c->n_flows, HF = (size_t)c->history_len * c->n_features;
    size_t o = 0;
    off[0] = o; o = align_up(o + n * HF * 8);              // hist
    off[1] = o; o = align_up(o + n * sizeof(FlowState));   // per-flow state
    off[6] = o; o = align_up(o + 64);                      // meta
    total = o;
}

static int flows_validate(const pcc_flows_config *c)
{
    if (!c) return fail(PCC_EINVAL, "null config");
    if (c->abi_version != PCC_ABI_VERSION) return fail(PCC_EINVAL, "ABI version mismatch");
    if (c->n_flows < 1 || c->n_flows > 0x7fffffffLL) return fail(PCC_EINVAL, "n_flows out of range");
    if (c->history_len < 1 || c->history_len > PCC_MAX_HISTORY) return fail(PCC_EINVAL, "history_len out of range");
    if (c->n_features < 1 || c->n_features > PCC_MAX_FEATURES) return fail(PCC_EINVAL, "n_features out of range");
    for (int i = 0; i < c->n_features; i++)
        if (c->feature_ids[i] < 0 || c->feature_ids[i] >= PCC_N_METRICS) return fail(PCC_EINVAL, "unknown feature id");
    if (c->rate_style != PCC_RATE_CLIENT && c->rate_style != PCC_RATE_SHIM) return fail(PCC_EINVAL, "bad rate_style");
    return PCC_OK;
}

extern "C" {

void pcc_flows_default_config(pcc_flows_config *c)
{
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->abi_version = PCC_ABI_VERSION;
    c->history_len = 10;
    c->n_features = 3;
    c->feature_ids[0] = PCC_M_SENT_LATENCY_INFLATION;
    c->feature_ids[1] = PCC_M_LATENCY_RATIO;
    c->feature_ids[2] = PCC_M_SEND_RATIO;
    c->delta_scale = 0.05;      /* loaded_client.py:33-35 */
    c->min_rate = 0.5;
    c->max_rate = 300.0;
    c->rate_style = PCC_RATE_CLIENT;
}

int pcc_flows_workspace_bytes(const pcc_flows_config *cfg, uint64_t *bytes)
{
    int rc = flows_validate(cfg);
    if (rc) return rc;
    size_t off[8], total;
    flows_layout(cfg, off, total);
    if (bytes) *bytes = total;
    return PCC_OK;
}

static int flows_build(pcc_flows_handle *out, const pcc_flows_config *cfg, void *workspace_dev, bool init)
{
    int rc = flows_validate(cfg);
    if (rc) return rc;
    if (!out || !workspace_dev || ((uintptr_t)workspace_dev & 255)) return fail(PCC_EINVAL, "bad workspace pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PCC_ENODEV, "no CUDA device: libpcc_b200 has no CPU fallback");
    CUDA_TRY(cudaSetDevice(cfg->device));
    pcc_flows_handle h = new (std::nothrow) pcc_flows_handle_s();
    if (!h) return fail(PCC_EINVAL, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    size_t off[8], total;
    flows_layout(cfg, off, total);
    char *b = (char *)workspace_dev;
    FlowsDev &d = h->d;
    d.hist = (double *)(b + off[0]); d.st = (FlowState *)(b + off[1]);
    d.meta = (unsigned long long *)(b + off[6]);
    d.n_flows = cfg->n_flows; d.H = cfg->history_len; d.F = cfg->n_features;
    for (int i = 0; i < PCC_MAX_FEATURES; i++) d.ids[i] = i < cfg->n_features ? cfg->feature_ids[i] : 0;
    d.touch_conn = features_touch_conn_min(d.ids, d.F) ? 1 : 0;
    d.delta_scale = cfg->delta_scale; d.min_rate = cfg->min_rate; d.max_rate = cfg->max_rate;
    d.rate_style = cfg->rate_style;
    h->batch_no = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { delete h; return fail(PCC_ECUDA, "cudaGetDeviceProperties failed"); }
    h->sm_count = prop.multiProcessorCount;
    {
        const char *fk = getenv("PCC_FLOWS_KERNEL");
        h->use_tma = !(fk && !strcmp(fk, "ldg"));
        cudaError_t ce = cudaFuncSetAttribute(pcc_flows_ingest_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCCF_TMA_SMEM);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(pcc_flows_ingest_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCCF_TMA_SMEM);
        int nb = 0;
        if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pcc_flows_ingest_tma_kernel<true>, PCCF_TMA_WARPS * 32, PCCF_TMA_SMEM);
        if (ce != cudaSuccess || nb < 1) { delete h; return fail(PCC_ECUDA, "flows: TMA kernel configuration: %s", cudaGetErrorString(ce)); }
        h->tma_blocks_per_sm = nb;
    }
    if (init) {
        cudaError_t e = cudaMemset(b, 0, total);
        if (e == cudaSuccess) {
            pcc_flows_reset_kernel<<<(unsigned)((d.n_flows + 127) / 128), 128>>>(d, nullptr, FLOW_RESET_NEW);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { delete h; return fail(PCC_ECUDA, "flows init: %s", cudaGetErrorString(e)); }
        h->launches++;
    } else {
        unsigned long long bn = 0;
        cudaError_t e = cudaMemcpy(&bn, d.meta + 2, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { delete h; return fail(PCC_ECUDA, "flows attach: %s", cudaGetErrorString(e)); }
        h->batch_no = (uint32_t)bn;
    }
    *out = h;
    return PCC_OK;
}

int pcc_flows_create(pcc_flows_handle *out, const pcc_flows_config *cfg, void *workspace_dev)
{
    return flows_build(out, cfg, workspace_dev, true);
}
int pcc_flows_attach(pcc_flows_handle *out, const pcc_flows_config *cfg, void *workspace_dev)
{
    return flows_build(out, cfg, workspace_dev, false);
}
void pcc_flows_destroy(pcc_flows_handle h) { delete h; }

int pcc_flows_give_samples(pcc_flows_handle h, const pcc_mi_batch *batch, int32_t unique_flows, double *obs_dev,
                           double *metrics_dev, void *stream)
{
    if (!h || !batch) return fail(PCC_EINVAL, "null pointer");
    if (batch->n_records < 0 || batch->n_records > 0x7fffffffLL) return fail(PCC_EINVAL, "n_records out of range");
    if (batch->n_records == 0) return PCC_OK;
    if (!batch->flow || !batch->bytes_sent || !batch->bytes_acked || !batch->bytes_lost || !batch->packet_size ||
        !batch->send_start || !batch->send_end || !batch->recv_start || !batch->recv_end || !batch->rtt_offsets)
        return fail(PCC_EINVAL, "null field array in the batch");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    BatchDev b;
    b.R = batch->n_records; b.flow = batch->flow;
    b.bytes_sent = (const long long *)batch->bytes_sent; b.bytes_acked = (const long long *)batch->bytes_acked;
    b.bytes_lost = (const long long *)batch->bytes_lost; b.packet_size = (const long long *)batch->packet_size;
    b.send_start = batch->send_start; b.send_end = batch->send_end; b.recv_start = batch->recv_start;
    b.recv_end = batch->recv_end; b.off = (const long long *)batch->rtt_offsets; b.rtt = batch->rtt_samples;
    const int64_t R = b.R;
    // persistent grid: one resident wave (SM count x blocks per SM); every warp streams a contiguous run of records
    int64_t blocks = (R + 31) / 32;
    const int64_t cap = (int64_t)h->sm_count * PCCF_MINBLOCKS;
    if (blocks > cap) blocks = cap;
    int64_t lblocks = (R + 127) / 128;          // the long-list pass: 4 warps x 32 records per block pass
    if (lblocks > cap) lblocks = cap;
    h->batch_no++;
    if (h->batch_no == 0) h->batch_no = 1;
    // cp.async.bulk needs a 16-byte aligned source: any other sample array goes through the read-only-path kernel
    const bool tma = h->use_tma && (((uintptr_t)b.rtt) & 15) == 0;
    int64_t tblocks = (R + 4 * PCCF_TMA_WARPS - 1) / (4 * PCCF_TMA_WARPS);
    if (tblocks > (int64_t)h->sm_count * h->tma_blocks_per_sm) tblocks = (int64_t)h->sm_count * h->tma_blocks_per_sm;
    if (unique_flows) {
        if (tma)
            pcc_flows_ingest_tma_kernel<true><<<(unsigned)tblocks, PCCF_TMA_WARPS * 32, PCCF_TMA_SMEM, st>>>(
                h->d, b, h->batch_no, obs_dev, metrics_dev, nullptr, nullptr);
        else
            pcc_flows_ingest_kernel<true><<<(unsigned)blocks, PCCF_THREADS, 0, st>>>(h->d, b, h->batch_no, obs_dev, metrics_dev,
                                                                                       nullptr, nullptr);
        pcc_flows_long_kernel<true><<<(unsigned)lblocks, 128, 0, st>>>(h->d, b, h->batch_no, obs_dev, metrics_dev, nullptr, nullptr);
        CUDA_TRY(cudaGetLastError());
        h->launches += 2;
        return PCC_OK;
    }
    // general batch: per-record part, stable sort of record indices by flow, per-flow sequential apply
    const size_t F = (size_t)h->d.F;
    const size_t rows_b = align_up((size_t)R * F * 8), avg_b = align_up((size_t)R * 8), idx_b = align_up((size_t)R * 4);
    size_t tmp_b = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_b, (const int32_t *)nullptr, (int32_t *)nullptr,
                                             (const int32_t *)nullptr, (int32_t *)nullptr, (int)R, 0, 32, st));
    tmp_b = align_up(tmp_b);
    char *scratch = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&scratch, rows_b + avg_b + 3 * idx_b + tmp_b, st));
    double *rows = (double *)scratch;
    double *avg = (double *)(scratch + rows_b);
    int32_t *iota = (int32_t *)(scratch + rows_b + avg_b);
    int32_t *sflow = (int32_t *)(scratch + rows_b + avg_b + idx_b);
    int32_t *srec = (int32_t *)(scratch + rows_b + avg_b + 2 * idx_b);
    void *tmp = scratch + rows_b + avg_b + 3 * idx_b;
    if (tma)
        pcc_flows_ingest_tma_kernel<false><<<(unsigned)tblocks, PCCF_TMA_WARPS * 32, PCCF_TMA_SMEM, st>>>(
            h->d, b, h->batch_no, nullptr, metrics_dev, rows, avg);
    else
        pcc_flows_ingest_kernel<false><<<(unsigned)blocks, PCCF_THREADS, 0, st>>>(h->d, b, h->batch_no, nullptr, metrics_dev,
                                                                                    rows, avg);
    pcc_flows_long_kernel<false><<<(unsigned)lblocks, 128, 0, st>>>(h->d, b, h->batch_no, nullptr, metrics_dev, rows, avg);
    pcc_flows_iota_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(iota, R);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_b, b.flow, sflow, (const int32_t *)iota, srec, (int)R, 0, 32, st);
    if (e == cudaSuccess) {
        pcc_flows_apply_kernel<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(h->d, R, sflow, srec, rows, avg, obs_dev, metrics_dev);
        e = cudaGetLastError();
    }
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return fail(PCC_ECUDA, "give_samples: %s", cudaGetErrorString(e));
    h->launches += 4;   // ingest, long-list pass, iota, apply (+ cub's sort passes, not ours)
    return PCC_OK;
}

int pcc_flows_reset(pcc_flows_handle h, const uint8_t *mask_dev, int32_t mode, void *stream)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    if (mode < PCC_FLOW_RESET_NEW || mode > PCC_FLOW_RESET_SHIM) return fail(PCC_EINVAL, "bad reset mode");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    pcc_flows_reset_kernel<<<(unsigned)((h->d.n_flows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->d, mask_dev, mode);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return PCC_OK;
}

int pcc_flows_get_obs(pcc_flows_handle h, double *obs_dev, void *stream)
{
    if (!h || !obs_dev) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int64_t tot = h->d.n_flows * h->d.H * h->d.F;
    pcc_flows_obs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d, obs_dev);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return PCC_OK;
}

int pcc_flows_set_rates(pcc_flows_handle h, const uint8_t *mask_dev, const double *rates_dev, double rate, void *stream)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    pcc_flows_set_rate_kernel<<<(unsigned)((h->d.n_flows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->d, mask_dev, rates_dev, rate);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return PCC_OK;
}

int pcc_flows_get_rates(pcc_flows_handle h, const double *actions_dev, const uint8_t *mask_dev, double *rates_dev,
                        void *stream)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    pcc_flows_rate_kernel<<<(unsigned)((h->d.n_flows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->d, actions_dev, mask_dev, rates_dev);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return PCC_OK;
}

int pcc_flows_act(pcc_flows_handle h, const pcc_policy *policy, double *actions_dev, void *stream)
{
    if (!h || !policy || !policy->w1 || !actions_dev) return fail(PCC_EINVAL, "null pointer");
    if (policy->n_in != h->d.H * h->d.F || policy->n_in > 128 || policy->h1 < 1 || policy->h1 > PCC_POLICY_MAXH ||
        policy->h2 < 1 || policy->h2 > PCC_POLICY_MAXH)
        return fail(PCC_EINVAL, "policy shape does not fit (n_in = history_len * n_features <= 128, hidden <= 64)");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    PolicyDev pol;
    pol.w1 = policy->w1; pol.b1 = policy->b1; pol.w2 = policy->w2; pol.b2 = policy->b2; pol.w3 = policy->w3; pol.b3 = policy->b3;
    pol.n_in = policy->n_in; pol.h1 = policy->h1; pol.h2 = policy->h2;
    pol.log_std = 0.0; pol.noise_seed = 0ull; pol.stochastic = 0;      /* act(..., stochastic=False) */
    pcc_flows_act_kernel<<<(unsigned)((h->d.n_flows + 63) / 64), 64, 0, (cudaStream_t)stream>>>(h->d, pol, actions_dev);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return PCC_OK;
}

int pcc_flows_get_column(pcc_flows_handle h, const char *name, double *dst_dev, void *stream)
{
    if (!h || !name || !dst_dev) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    int which;
    if (!strcmp(name, "conn_min")) which = 0;
    else if (!strcmp(name, "rate")) which = 1;
    else return fail(PCC_EINVAL, "unknown column %s", name);
    pcc_flows_column_kernel<<<(unsigned)((h->d.n_flows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->d, which, dst_dev);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return PCC_OK;
}

int pcc_flows_check(pcc_flows_handle h, void *stream)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned long long meta[2];
    CUDA_TRY(cudaMemcpy(meta, h->d.meta, sizeof(meta), cudaMemcpyDeviceToHost));
    if (meta[1]) return fail(PCC_EINVAL, "a record named a flow index outside [0, n_flows) or a negative sample count (record ignored)");
    if (meta[0]) return fail(PCC_EINVAL, "a batch declared unique_flows held two records of one flow (that flow's history is undefined)");
    return PCC_OK;
}

int64_t pcc_flows_launch_count(pcc_flows_handle h) { return h ? h->launches : 0; }

}  // extern "C"
