// pcc_b200.cu -- libpcc_b200.so: CUDA kernels (sm_100a) + the C ABI of include/pcc_b200.h.
//
// Build (pcc-rl_b200/build.py):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
//        -Xcompiler -fPIC -shared -Iinclude -o pcc-rl_b200/libpcc_b200.so pcc-rl_b200/csrc/pcc_b200.cu
// -fmad=false is REQUIRED: the reference rounds every binary64 operation separately and the
// integer packet counts depend on float comparisons.
//
// Data layout in HBM (all owned by the caller, see pcc_workspace_bytes):
//   state block : structure-of-arrays, one column of n_envs elements per scalar of
//                 pcc::EnvState (coalesced: lane == env), then the MI-feature history
//                 hist[env][slot][feature] (a ring over `slot` with a handle-global head, so
//                 a step writes one new row instead of shifting H rows), then (MT19937 mode
//                 only) uint32[n_envs][625] generator states, then a small meta block.
//   ring block  : Rec[n_envs][ring_capacity], 16 B per in-flight packet (arrival time at
//                 hop 1, +-link-0 latency), addressed by monotonically increasing u32 cursors.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "pcc_b200.h"
#include "pcc_core.cuh"
#include "pcc_coop.cuh"
#include "pcc_warp.cuh"
#include "pcc_multi_core.cuh"
#include "pcc_multi_fast.cuh"
#include "pcc_multi_warp.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

using namespace pcc;

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) return fail(PCC_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// ---------------------------------------------------------------------------------------
// device-side view of the workspaces
// ---------------------------------------------------------------------------------------
struct DevState {
    double *d_bw, *bw, *dl, *lr, *max_qd, *qd, *t_upd, *rate, *next_send, *cur_time, *run_dur, *conn_min;
    double *w_full;              // tail-drop threshold in the queue-delay domain (pcc_core.cuh: tail_drop_threshold)
    double *ret_acc, *ret_last;  // reward_sum of the running / of the last finished episode (network_sim.py:390,443,483)
    unsigned long long *seed, *draws;
    uint32_t *tail, *h1, *h2;
    int32_t *steps;
    double *hist;              // [n][H][F]: per env a ring over H slots (handle-global head)
    uint32_t *mt;              // [n][625] or null
    unsigned long long *meta;  // [0] steps issued (history head), [1] overflow count, [2] first overflowed env + 1
    Rec *rings;                // [n][cap]
    double *mean_scratch;      // [warps][PCC_GSCRATCH] acked-latency staging of heavy MIs (global), or null
    uint32_t cap;
    int64_t n;
    int32_t H, F;
    int32_t ids[PCC_MAX_FEATURES];
    int32_t need_inc;
    Consts c;
};

enum { META_HEAD = 0, META_OVF_COUNT = 1, META_OVF_ENV = 2, META_PART_ERR = 3, META_WORDS = 8 };

struct DevRing {
    Rec *base;
    uint32_t mask;
    __device__ __forceinline__ uint32_t capacity() const { return mask + 1u; }
    __device__ __forceinline__ Rec load(uint32_t i) const
    {
        const double2 v = *reinterpret_cast<const double2 *>(base + (i & mask));
        Rec r; r.a = v.x; r.l = v.y;
        return r;
    }
    __device__ __forceinline__ void store(uint32_t i, Rec r)
    {
        *reinterpret_cast<double2 *>(base + (i & mask)) = make_double2(r.a, r.l);
    }
    __device__ __forceinline__ void store_a(uint32_t i, double a) { base[i & mask].a = a; }
    __device__ __forceinline__ const Rec *addr(uint32_t i) const { return base + (i & mask); }
    __device__ __forceinline__ void prefetch(uint32_t i) const { asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (i & mask))); }
};

__device__ __forceinline__ void load_env(const DevState &p, int64_t e, EnvState &s)
{
    s.d_bw = p.d_bw[e]; s.dl = p.dl[e]; s.lr = p.lr[e]; s.max_qd = p.max_qd[e]; s.w_full = p.w_full[e];
    s.qd = p.qd[e]; s.t_upd = p.t_upd[e]; s.rate = p.rate[e]; s.next_send = p.next_send[e];
    s.cur_time = p.cur_time[e]; s.run_dur = p.run_dur[e]; s.conn_min = p.conn_min[e];
    s.tail = p.tail[e]; s.h1 = p.h1[e]; s.h2 = p.h2[e]; s.steps = p.steps[e];
}
__device__ __forceinline__ void store_env_dynamic(const DevState &p, int64_t e, const EnvState &s)
{
    p.qd[e] = s.qd; p.t_upd[e] = s.t_upd; p.rate[e] = s.rate; p.next_send[e] = s.next_send;
    p.cur_time[e] = s.cur_time; p.run_dur[e] = s.run_dur; p.conn_min[e] = s.conn_min;
    p.tail[e] = s.tail; p.h1[e] = s.h1; p.h2[e] = s.h2; p.steps[e] = s.steps;
}
__device__ __forceinline__ void flag_overflow(const DevState &p, int64_t e)
{
    atomicAdd(&p.meta[META_OVF_COUNT], 1ull);
    atomicCAS(&p.meta[META_OVF_ENV], 0ull, (unsigned long long)(e + 1));
}

#include "pcc_packed.cuh"

// ---------------------------------------------------------------------------------------
// kernels (v1: one thread per env, every phase scalar; see DESIGN.md for the roofline)
// ---------------------------------------------------------------------------------------
template <int RNG>
__global__ void pcc_step_kernel(DevState p, const int32_t *__restrict__ perm, unsigned long long head_step,
                                const double *__restrict__ actions, double *__restrict__ obs,
                                double *__restrict__ reward, uint8_t *__restrict__ done,
                                int32_t *__restrict__ counts, double *__restrict__ info)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) p.meta[META_HEAD] = head_step + 1ull;
    if (tid >= p.n) return;
    const int64_t e = perm ? (int64_t)perm[tid] : tid;   // cost-sorted order: similar work per warp
    EnvState s;
    load_env(p, e, s);
    DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
    StepOut o;
    const double action = actions[e];
    if (RNG == PCC_RNG_PHILOX) {
        PhiloxRng rng;
        rng.init(p.seed[e], p.draws[e]);
        step_env(s, ring, rng, action, p.c, p.need_inc != 0, o);
        p.draws[e] = rng.draws;
    } else {
        Mt19937Rng rng;
        rng.init(p.mt + (size_t)e * 625);
        step_env(s, ring, rng, action, p.c, p.need_inc != 0, o);
    }
    store_env_dynamic(p, e, s);
    if (o.mi.overflow) flag_overflow(p, e);

    // history ring: the new row replaces the oldest slot; obs = oldest -> newest
    const int H = p.H, F = p.F;
    const int slot_new = (int)(head_step % (unsigned long long)H);
    double row[PCC_MAX_FEATURES];
#pragma unroll
    for (int f = 0; f < PCC_MAX_FEATURES; f++)
        if (f < F) row[f] = metric_value(o.st, p.ids[f]);
    double *ob = obs + (size_t)e * (size_t)(H * F);
    for (int h = 0; h < H - 1; h++) {
        int slot = slot_new + 1 + h;
        if (slot >= H) slot -= H;
        for (int f = 0; f < F; f++) ob[h * F + f] = p.hist[(size_t)e * (H * F) + slot * F + f];
    }
#pragma unroll
    for (int f = 0; f < PCC_MAX_FEATURES; f++)
        if (f < F) {
            ob[(H - 1) * F + f] = row[f];
            p.hist[(size_t)e * (H * F) + slot_new * F + f] = row[f];
        }
    reward[e] = o.st.reward;
    done[e] = o.done ? 1 : 0;
    {
        const double acc = p.ret_acc[e] + o.st.reward;   // self.reward_sum += reward  (:443)
        p.ret_acc[e] = acc;
        if (o.done) p.ret_last[e] = acc;
    }
    if (counts) {
        counts[3 * e + 0] = o.mi.sent; counts[3 * e + 1] = o.mi.acked; counts[3 * e + 2] = o.mi.lost;
    }
    if (info) {
        double *q = info + (size_t)e * PCC_INFO_WIDTH;
        q[0] = o.st.send_rate; q[1] = o.st.recv_rate; q[2] = o.st.avg_lat; q[3] = o.st.loss_ratio;
        q[4] = o.st.lat_infl; q[5] = o.st.lat_ratio; q[6] = o.st.send_ratio; q[7] = o.st.dur;
        q[8] = s.cur_time; q[9] = s.rate; q[10] = s.run_dur; q[11] = s.conn_min;
    }
}

// ---------------------------------------------------------------------------------------
// kernels (v2: G lanes per env, see pcc_coop.cuh) -- Philox streams only
// ---------------------------------------------------------------------------------------
#define PCC_COOP_THREADS 128

template <int G>
__device__ __forceinline__ void coop_emit(const Grp<G> &g, const DevState &p, int64_t e, unsigned long long head_step,
                                          const MiStats &st, double *__restrict__ obs)
{
    // history ring: the new row replaces the oldest slot; obs = oldest -> newest, written by
    // the group as one contiguous H*F*8-byte row
    const int H = p.H, F = p.F, HF = H * F;
    const int slot_new = (int)(head_step % (unsigned long long)H);
    double *hrow = p.hist + (size_t)e * HF;
    double *ob = obs + (size_t)e * HF;
    for (int base = (int)g.gl; base < HF; base += 4 * G) {     // four history reads in flight before their stores
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = base + u * G;
            v[u] = 0.0;
            if (k < HF) {
                const int h = k / F, f = k - h * F;
                if (h == H - 1) v[u] = metric_value(st, p.ids[f]);
                else {
                    int slot = slot_new + 1 + h;
                    if (slot >= H) slot -= H;
                    v[u] = hrow[slot * F + f];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = base + u * G;
            if (k < HF) ob[k] = v[u];
        }
    }
    g.gsync();
    if ((int)g.gl < F) hrow[slot_new * F + (int)g.gl] = metric_value(st, p.ids[g.gl]);
    if (G < PCC_MAX_FEATURES)
        for (int f = G + (int)g.gl; f < F; f += G) hrow[slot_new * F + f] = metric_value(st, p.ids[f]);
}

template <int G>
__global__ void __launch_bounds__(PCC_COOP_THREADS)
pcc_step_coop_kernel(DevState p, const int32_t *__restrict__ perm, unsigned long long head_step, const double *__restrict__ actions,
                     double *__restrict__ obs, double *__restrict__ reward, uint8_t *__restrict__ done,
                     int32_t *__restrict__ counts, double *__restrict__ info)
{
    __shared__ GroupSmem<G> smem[PCC_COOP_THREADS / G];
    const Grp<G> g;
    int64_t e = ((int64_t)blockIdx.x * PCC_COOP_THREADS + threadIdx.x) / G;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta[META_HEAD] = head_step + 1ull;
    const bool alive = e < p.n;      // whole groups; dead groups follow the warp's control flow
    if (!alive) e = p.n - 1;
    if (perm) e = perm[e];           // cost-sorted order: the heaviest envs start first
    GroupSmem<G> &sm = smem[threadIdx.x / G];
    EnvState s;
    load_env(p, e, s);
    DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
    const uint64_t seed = p.seed[e];
    uint64_t draws = p.draws[e];
    StepOut o;
    s.rate = apply_rate_delta(s.rate, actions[e], p.c);                          // :412
#ifdef PCC_PROFILE
    long long prof[8];
    run_mi_coop(g, alive, s, ring, seed, draws, s.run_dur, sm, o.mi, prof);
    prof[6] = clock64();
#else
    run_mi_coop(g, alive, s, ring, seed, draws, s.run_dur, sm, o.mi);           // :416
#endif
    double avg_lat, lat_inc;
    mi_means_coop(g, alive, o.mi, ring, s.dl, sm, p.need_inc != 0, avg_lat, lat_inc);
#ifdef PCC_PROFILE
    prof[7] = clock64();
#endif
    mi_stats_finish(o.mi, p.c, avg_lat, lat_inc, s.conn_min, true, o.st);
    s.steps += 1;                                                                // :419
    if (o.st.avg_lat > 0.0) s.run_dur = 0.5 * o.st.avg_lat;                      // :437-438
    o.done = s.steps >= p.c.max_steps;                                           // :444
    if (!alive) return;
    coop_emit(g, p, e, head_step, o.st, obs);
    if (g.gl == 0) {
        store_env_dynamic(p, e, s);
        p.draws[e] = draws;
        if (o.mi.overflow) flag_overflow(p, e);
        reward[e] = o.st.reward;
        done[e] = o.done ? 1 : 0;
        const double acc = p.ret_acc[e] + o.st.reward;                           // :443
        p.ret_acc[e] = acc;
        if (o.done) p.ret_last[e] = acc;
        if (counts) { counts[3 * e + 0] = o.mi.sent; counts[3 * e + 1] = o.mi.acked; counts[3 * e + 2] = o.mi.lost; }
        if (info) {
            double *q = info + (size_t)e * PCC_INFO_WIDTH;
            q[0] = o.st.send_rate; q[1] = o.st.recv_rate; q[2] = o.st.avg_lat; q[3] = o.st.loss_ratio;
            q[4] = o.st.lat_infl; q[5] = o.st.lat_ratio; q[6] = o.st.send_ratio; q[7] = o.st.dur;
            q[8] = s.cur_time; q[9] = s.rate; q[10] = s.run_dur; q[11] = s.conn_min;
#ifdef PCC_PROFILE   // profiling build: info = cycles per phase (send, hop1, bnd1, hop2, bnd2, cross, means), counts
            for (int k = 0; k < 7; k++) q[k] = (double)(prof[k + 1] - prof[k]);
            q[7] = (double)o.mi.sent; q[8] = (double)o.mi.acked; q[9] = (double)(prof[7] - prof[0]);
#endif
        }
    }
}

template <int G>
__global__ void __launch_bounds__(PCC_COOP_THREADS)
pcc_reset_coop_kernel(DevState p, const uint8_t *__restrict__ mask, const double *__restrict__ bw,
                      const double *__restrict__ delay, const long long *__restrict__ queue,
                      const double *__restrict__ loss, const double *__restrict__ start_rate,
                      double *__restrict__ obs)
{
    __shared__ GroupSmem<G> smem[PCC_COOP_THREADS / G];
    const Grp<G> g;
    int64_t e = ((int64_t)blockIdx.x * PCC_COOP_THREADS + threadIdx.x) / G;
    bool alive = e < p.n;
    if (!alive) e = p.n - 1;
    if (mask && !mask[e]) alive = false;
    GroupSmem<G> &sm = smem[threadIdx.x / G];
    EnvState s;
    const double bwv = bw[e], dlv = delay[e], sr = start_rate[e];
    // reset_env of pcc_core.cuh (network_sim.py:454-484), cooperatively
    s.d_bw = 1.0 / bwv; s.dl = dlv; s.lr = loss[e]; s.max_qd = (double)queue[e] / bwv;
    s.w_full = tail_drop_threshold(s.d_bw, s.max_qd);
    s.qd = 0.0; s.t_upd = 0.0; s.rate = sr; s.cur_time = 0.0; s.next_send = 1.0 / sr;
    s.run_dur = 3 * dlv; s.conn_min = 0.0;
    s.tail = p.tail[e]; s.h1 = s.tail; s.h2 = s.tail; s.steps = 0;
    DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
    const uint64_t seed = p.seed[e];
    uint64_t draws = p.draws[e];
    MiOut o;
    run_mi_coop(g, alive, s, ring, seed, draws, s.run_dur, sm, o);               // :478
    bool ovf = o.overflow;
    run_mi_coop(g, alive, s, ring, seed, draws, s.run_dur, sm, o);               // :479
    ovf = ovf || o.overflow;
    if (!alive) return;
    const int HF = p.H * p.F;
    for (int k = (int)g.gl; k < HF; k += G) {
        const double v = metric_empty(p.ids[k % p.F]);
        p.hist[(size_t)e * HF + k] = v;
        if (obs) obs[(size_t)e * HF + k] = v;
    }
    if (g.gl == 0) {
        p.d_bw[e] = s.d_bw; p.bw[e] = bwv; p.dl[e] = s.dl; p.lr[e] = s.lr; p.max_qd[e] = s.max_qd; p.w_full[e] = s.w_full;
        store_env_dynamic(p, e, s);
        p.draws[e] = draws;
        p.ret_acc[e] = 0.0;
        if (ovf) flag_overflow(p, e);
    }
}

// ---------------------------------------------------------------------------------------
// kernels (v4: a warp owns E envs, see pcc_warp.cuh) -- Philox streams only
// ---------------------------------------------------------------------------------------
#define PCC_WARP_THREADS 128
#define PCC_GSCRATCH 4096   // samples of global staging per warp (MIs with more acks than the shared buffer)
#ifndef PCC_STAGED_STORES
#define PCC_STAGED_STORES 0   // 1: stage ring records in shared memory, coalesced copy-out (measured: no gain)
#endif
#ifndef PCC_WARP_MINBLOCKS
#define PCC_WARP_MINBLOCKS 4
#endif

// Small batches (PAIR): every block is a worker warp plus a helper warp.  When the worker owns a single heavy env it
// offers the in-order consumption of the records that already exist (consume_scan_warp) to the helper, which runs it
// while the worker is in its send phase; two named barriers (offer visible / result visible) order the hand-over.
struct PairOffer {
    int32_t valid, use_g;
    uint32_t h1, h2, tail;
    double end, dl;
    long long e;
    ScanOut res;
};
// Producer / consumer named barriers between the two warps of a pair (64 threads): the producer signals with
// bar.arrive (does not wait), the consumer waits with bar.sync; memory written by the producer before its arrive is
// visible to the consumer after its sync.  Both are aligned instructions: the warp arrives converged (__syncwarp;
// the lane-0-only writes just before would otherwise leave it diverged).
__device__ __forceinline__ void pair_signal(int id)
{
    __syncwarp();
    asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void pair_wait(int id)
{
    __syncwarp();
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// One MI for the E envs of this warp.  `owner` lanes hold their env's state in `s`.
template <bool WANT_MEANS, bool DO_SEND, bool PAIR = false>
__device__ __forceinline__ void warp_mi(const Grp<32> &g, const DevState &p, bool owner, int cnt, int64_t e, EnvState &s,
                                        PhiloxRng &rng, double dur, double *buf, int wbuf, WarpStage &stage, MiOut &mo,
                                        double &avg_lat, double &lat_inc, int32_t sent_before = 0,
                                        long long *prof = nullptr, double *gscratch = nullptr, PairOffer *offer = nullptr,
                                        bool pair = PAIR, int bar0 = 1, uint32_t *mt = nullptr)
{
    const unsigned lane = g.gl;
    const double end = s.cur_time + dur;            // network_sim.py:124
    const double inv_rate = 1.0 / s.rate;           // :161
    DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
    LaneChain c;
    c.t = s.next_send; c.q = s.qd; c.tu = s.t_upd; c.tail = s.tail;
    c.sent = sent_before & 0x7fffffff; c.ovf = sent_before < 0;
    mo.start = s.cur_time;                          // reset_obs :319-324
    PCC_TICK(0);
    bool offered = false;
    int offered_use_g = 0;
    if (PAIR && pair) {
        // lane 0 owns the env when cnt == 1.  Worth offering when there is something to consume already.
        const uint32_t pend_old = __shfl_sync(PCC_FULL, (uint32_t)(s.tail - s.h2), 0);
        offered = (cnt == 1) && (buf != nullptr) && pend_old >= 64u;
        if (offered) {
            // the staging buffer both warps will use: pending hop-2 events bound this MI's acked samples, and this
            // MI adds at most (end - t) * rate + 3 records
            double est = (end - s.next_send) * s.rate + 3.0;
            est = __shfl_sync(PCC_FULL, est, 0);
            const bool big = !(est < 1e9) || ((double)pend_old + est > (double)wbuf);
            offered_use_g = (gscratch != nullptr && big) ? 1 : 0;
        }
        if (lane == 0) {
            offer->valid = offered ? 1 : 0;
            offer->use_g = offered_use_g;
            offer->h1 = s.h1; offer->h2 = s.h2; offer->tail = s.tail;
            offer->end = end; offer->dl = s.dl; offer->e = (long long)e;
        }
        pair_signal(bar0);                           // the offer is published to the helper warp
    }
    // phase A: E serial chains side by side (already done by pcc_send_kernel when !DO_SEND)
    if (DO_SEND) {
        if (cnt == 1 && buf != nullptr) {
            // a heavy env alone in its warp: the 32-lane chunked send phase of the group kernel (chain on
            // lane 0, Philox on all lanes, records staged in shared memory) -- 114 cycles per packet
            EnvState sb;
            sb.lr = __shfl_sync(PCC_FULL, s.lr, 0); sb.dl = __shfl_sync(PCC_FULL, s.dl, 0);
            sb.d_bw = __shfl_sync(PCC_FULL, s.d_bw, 0); sb.w_full = __shfl_sync(PCC_FULL, s.w_full, 0);
            double bt = __shfl_sync(PCC_FULL, c.t, 0), bq = __shfl_sync(PCC_FULL, c.q, 0), btu = __shfl_sync(PCC_FULL, c.tu, 0);
            uint32_t btail = __shfl_sync(PCC_FULL, c.tail, 0);
            const uint32_t bh2 = __shfl_sync(PCC_FULL, s.h2, 0);
            const unsigned long long bseed = __shfl_sync(PCC_FULL, (unsigned long long)rng.seed, 0);
            uint64_t bdraws = __shfl_sync(PCC_FULL, (unsigned long long)rng.draws, 0);
            const double bend = __shfl_sync(PCC_FULL, end, 0), binv = __shfl_sync(PCC_FULL, inv_rate, 0);
            const long long be0 = __shfl_sync(PCC_FULL, (long long)e, 0);
            DevRing r0{p.rings + (size_t)be0 * p.cap, p.cap - 1u};
            int32_t bsent = 0;
            bool bovf = false;
            double2 *stage2 = reinterpret_cast<double2 *>(reinterpret_cast<char *>(buf) + (size_t)(wbuf + 32) * 8);
#ifdef PCC_SOLO_V1
            coop_send_chunks(g, true, sb, r0, bseed, bdraws, bend, binv, stage2, bt, bq, btu, btail, bh2, bsent, bovf);
#else
            group_send_chunks(g, true, sb, r0, bseed, bdraws, bend, binv, *reinterpret_cast<SoloSendSmem *>(stage2), bt, bq,
                              btu, btail, bh2, bsent, bovf, mt);
#endif
            if (owner) {
                c.t = bt; c.q = bq; c.tu = btu; c.tail = btail; c.sent += bsent; c.ovf = c.ovf || bovf;
                rng.init(rng.seed, bdraws);
            }
        }
        else if (cnt == 1) coop_send_phase(c, s, ring, rng, owner, cnt, s.h2, p.cap, end, inv_rate);   // reset path (no smem)
#if PCC_STAGED_STORES
        else lane_send_phase_staged(c, s, ring.base, ring.mask, rng, owner, s.h2, p.cap, end, inv_rate, stage);
#else
        else if (owner) lane_send_phase(c, s, ring, rng, s.h2, p.cap, end, inv_rate);
#endif
        __syncwarp();
    }
    PCC_TICK(1);
    // phase B: the warp consumes each env's hop-1 / hop-2 events
    int which = 0, acked = 0, lost = 0;
    double ct = 0.0;
    avg_lat = 0.0; lat_inc = 0.0;
#pragma unroll 1
    for (int j = 0; j < cnt; j++) {
        if (!__shfl_sync(PCC_FULL, (int)owner, j)) continue;
#ifndef PCC_PREFETCH_AHEAD
#define PCC_PREFETCH_AHEAD 1
#endif
        const int jn = j + PCC_PREFETCH_AHEAD;
        if (j == 0 && PCC_PREFETCH_AHEAD > 1) {   // envs 1 .. AHEAD-1 are not covered by the steady-state prefetch
#pragma unroll
            for (int jj = 1; jj < PCC_PREFETCH_AHEAD; jj++)
                if (jj < cnt && __shfl_sync(PCC_FULL, (int)owner, jj)) {
                    const long long en = __shfl_sync(PCC_FULL, (long long)e, jj);
                    const uint32_t h1n = __shfl_sync(PCC_FULL, s.h1, jj), h2n = __shfl_sync(PCC_FULL, s.h2, jj);
                    const uint32_t tn = __shfl_sync(PCC_FULL, c.tail, jj);
                    const Rec *bn = p.rings + (size_t)en * p.cap;
                    const uint32_t o = lane * 8u;
                    if (o < (uint32_t)(tn - h1n)) prefetch_l1(bn + ((h1n + o) & (p.cap - 1u)));
                    if (o < (uint32_t)(tn - h2n)) prefetch_l1(bn + ((h2n + o) & (p.cap - 1u)));
                }
        }
        if (jn < cnt && __shfl_sync(PCC_FULL, (int)owner, jn)) {
            // pull a later env's cursor neighbourhoods into L1 while this one is processed
            const long long en = __shfl_sync(PCC_FULL, (long long)e, jn);
            const uint32_t h1n = __shfl_sync(PCC_FULL, s.h1, jn), h2n = __shfl_sync(PCC_FULL, s.h2, jn);
            const uint32_t tn = __shfl_sync(PCC_FULL, c.tail, jn);
            const Rec *bn = p.rings + (size_t)en * p.cap;
            const uint32_t o = lane * 8u;
            if (o < (uint32_t)(tn - h1n)) prefetch_l1(bn + ((h1n + o) & (p.cap - 1u)));
            if (o < (uint32_t)(tn - h2n)) prefetch_l1(bn + ((h2n + o) & (p.cap - 1u)));
        }
        ConsumeIn in;
        in.end = __shfl_sync(PCC_FULL, end, j);
        in.dl = __shfl_sync(PCC_FULL, s.dl, j);
        in.tnext = __shfl_sync(PCC_FULL, c.t, j);
        in.tail = __shfl_sync(PCC_FULL, c.tail, j);
        in.h1 = __shfl_sync(PCC_FULL, s.h1, j);
        in.h2 = __shfl_sync(PCC_FULL, s.h2, j);
        in.acked0 = 0; in.lost0 = 0; in.s_begin = in.h2;
        // staging of this MI's acked latencies: shared memory, or -- when more than wbuf packets could be acked
        // (pending hop-2 events bound it) -- the warp's global scratch
        double *sbuf_j = buf;
        in.wbuf = wbuf;
        if (PAIR && offered) {
            // the helper warp has scanned the records that existed before the send phase: resume from its cursors
            pair_wait(bar0 + 1);                     // its result (and the staged samples) are visible
            if (offered_use_g) { sbuf_j = gscratch; in.wbuf = PCC_GSCRATCH; }
            in.h1 = offer->res.h1; in.h2 = offer->res.h2;
            in.acked0 = offer->res.acked; in.lost0 = offer->res.lost;
        } else
        if (gscratch != nullptr && (uint32_t)(in.tail - in.h2) > (uint32_t)wbuf) { sbuf_j = gscratch; in.wbuf = PCC_GSCRATCH; }
#ifdef PCC_PROFILE
        const long long tc0 = clock64();
#endif
        const long long ej = __shfl_sync(PCC_FULL, (long long)e, j);
        DevRing rj{p.rings + (size_t)ej * p.cap, p.cap - 1u};
        ConsumeOut co;
        consume_mi_warp(g, in, rj, sbuf_j, co);
        double a = 0.0, li = 0.0;
#ifdef PCC_PROFILE
        const long long tc1 = clock64();
#endif
        if (WANT_MEANS) mi_means_warp(g, co, rj, in.dl, sbuf_j, in.wbuf, buf,
                                      p.need_inc != 0, a, li);
#ifdef PCC_PROFILE
        if (prof && (int)lane == j) { prof[2] = tc1 - tc0; prof[3] = clock64() - tc1; }
#endif
        if ((int)lane == j) {
            s.h1 = co.h1; s.h2 = co.h2;
            acked = co.acked; lost = co.lost; which = co.which; ct = co.cur_time;
            avg_lat = a; lat_inc = li;
        }
    }
    PCC_TICK(4);
    // phase C (per lane): the crossing event if it is the pacing timer (:156-178)
    if (owner) {
        if (which == 0) {
            s.cur_time = c.t;
            double u;
            if (mt != nullptr) { Mt19937Rng mr; mr.init(mt); u = mr.next(); }
            else u = rng.next();
            lane_send_one(c, s, ring, s.h2, p.cap, inv_rate, u);
        } else {
            s.cur_time = ct;
        }
    }
    s.next_send = c.t; s.qd = c.q; s.t_upd = c.tu; s.tail = c.tail;
    mo.sent = c.sent; mo.acked = acked; mo.lost = lost;
    mo.end = s.cur_time;
    mo.overflow = c.ovf;
}

// Work partition of a step: warp w owns the envs perm[starts[w] .. starts[w+1]) of the cost-sorted
// list (at most 32), or -- without a partition -- the static_e consecutive envs w*static_e ...
struct WarpPartition {
    const int32_t *perm;      // cost-sorted env ids, or null
    const int32_t *starts;    // [n_warps + 1] offsets into perm, or null
    const int32_t *n_warps;   // device scalar
    int32_t static_e;
    int32_t wbuf;             // staging capacity (samples) per warp
    int32_t n_pair_blocks;    // PAIRMODE 2: the first n_pair_blocks blocks (the heaviest slots) are worker + helper blocks
};

// SPLIT: the sends of this MI were already done by pcc_send_kernel.  PAIRMODE 1 (small batches): every block is
// workers + helpers (64 threads: worker of partition slot blockIdx.x and its helper, see PairOffer).  PAIRMODE 2 (big
// batches): only the first part.n_pair_blocks blocks -- the heaviest slots of the cost-sorted list, whose residency is
// the kernel's critical path -- are worker + helper blocks; the others are all workers as in PAIRMODE 0 (a helper for
// every slot would halve the occupancy that big batches live on).
template <bool SPLIT, int PAIRMODE = 0>
__global__ void __launch_bounds__(PCC_WARP_THREADS, PCC_WARP_MINBLOCKS)
pcc_step_warp_kernel(DevState p, WarpPartition part, const int32_t *__restrict__ sent_tmp, unsigned long long head_step,
                     const double *__restrict__ actions, double *__restrict__ obs, double *__restrict__ reward,
                     uint8_t *__restrict__ done, int32_t *__restrict__ counts, double *__restrict__ info)
{
    extern __shared__ double dyn_smem[];      // per worker warp: warp_smem_bytes(part.wbuf)
    __shared__ WarpStage sstage[(SPLIT || !PCC_STAGED_STORES) ? 1 : PCC_WARP_THREADS / 32];
    __shared__ PairOffer pair_offers[PCC_WARP_THREADS / 64];
    constexpr bool PAIR = PAIRMODE != 0;
    const bool pair_blk = PAIRMODE == 1 || (PAIRMODE == 2 && (int)blockIdx.x < part.n_pair_blocks);
    const int wib = (int)(threadIdx.x >> 5), wpb = (int)(blockDim.x >> 5), half = wpb >> 1;
    // pair block: warps 0 .. half-1 are workers, warp half + i is the helper of worker i
    const int pi = pair_blk ? wib % half : 0;
    const bool helper = pair_blk && wib >= half;
    PairOffer &pair_offer = pair_offers[pi];
    const int bar0 = 1 + 2 * pi;
    double *wsm = dyn_smem + (size_t)(pair_blk ? pi : wib) * (warp_smem_bytes(part.wbuf) / 8);
    const Grp<32> g;
    const unsigned lane = threadIdx.x & 31u;
    const int64_t w = pair_blk ? (int64_t)blockIdx.x * half + pi
                    : PAIRMODE == 2 ? (int64_t)part.n_pair_blocks * half + ((int64_t)blockIdx.x - part.n_pair_blocks) * wpb + wib
                                    : (int64_t)blockIdx.x * wpb + wib;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta[META_HEAD] = head_step + 1ull;
    int64_t first;
    int cnt;
    if (part.starts) {
        const int nw = *part.n_warps;
        if (w >= nw) return;                       // whole warp (PAIR: the whole block)
        first = part.starts[w];
        cnt = (int)(part.starts[w + 1] - first);
    } else {
        first = w * part.static_e;
        if (first >= p.n) return;
        cnt = (int)((p.n - first < part.static_e) ? (p.n - first) : part.static_e);
    }
    if (PAIR && helper) {                            // the helper warp
        pair_wait(bar0);
        if (pair_offer.valid) {
            DevRing rj{p.rings + (size_t)pair_offer.e * p.cap, p.cap - 1u};
            double *sb = pair_offer.use_g ? p.mean_scratch + (size_t)w * PCC_GSCRATCH : wsm;
            ScanOut so;
            consume_scan_warp(g, pair_offer.end, pair_offer.dl, pair_offer.h1, pair_offer.h2, pair_offer.tail, rj, sb,
                              pair_offer.use_g ? PCC_GSCRATCH : part.wbuf, so);
            if (lane == 0) pair_offer.res = so;
            pair_signal(bar0 + 1);
        }
        return;
    }
    const bool owner = (int)lane < cnt;
    const int64_t e = owner ? (part.perm ? (int64_t)part.perm[first + lane] : first + lane) : 0;
    EnvState s;
    load_env(p, e, s);
    PhiloxRng rng;
    rng.init(p.seed[e], p.draws[e]);
    if (!SPLIT) s.rate = apply_rate_delta(s.rate, actions[e], p.c);              // :412
    StepOut o;
    double avg_lat, lat_inc;
#ifdef PCC_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tk0 = clock64();
#else
    long long *prof = nullptr;
#endif
    warp_mi<true, !SPLIT, PAIR>(g, p, owner, cnt, e, s, rng, s.run_dur, wsm, part.wbuf,
                                sstage[(SPLIT || !PCC_STAGED_STORES || PAIR) ? 0 : (threadIdx.x >> 5)], o.mi, avg_lat, lat_inc,
                                SPLIT ? sent_tmp[e] : 0, prof,
                                p.mean_scratch ? p.mean_scratch + (size_t)w * PCC_GSCRATCH : nullptr, &pair_offer, pair_blk, bar0);   // :416
    if (!owner) return;
    mi_stats_finish(o.mi, p.c, avg_lat, lat_inc, s.conn_min, true, o.st);
    s.steps += 1;                                                                // :419
    if (o.st.avg_lat > 0.0) s.run_dur = 0.5 * o.st.avg_lat;                      // :437-438
    o.done = s.steps >= p.c.max_steps;                                           // :444
    store_env_dynamic(p, e, s);
    p.draws[e] = rng.draws;
    if (o.mi.overflow) flag_overflow(p, e);
    // history ring + obs row (oldest -> newest), each lane writes its env's contiguous row
    {
        const int H = p.H, F = p.F, HF = H * F;
        const int slot_new = (int)(head_step % (unsigned long long)H);
        double *hrow = p.hist + (size_t)e * HF;
        double *ob = obs + (size_t)e * HF;
        for (int h = 0; h < H - 1; h++) {
            int sl = slot_new + 1 + h;
            if (sl >= H) sl -= H;
            for (int f = 0; f < F; f++) ob[h * F + f] = hrow[sl * F + f];
        }
        for (int f = 0; f < F; f++) {
            const double v = metric_value(o.st, p.ids[f]);
            ob[(H - 1) * F + f] = v;
            hrow[slot_new * F + f] = v;
        }
    }
    reward[e] = o.st.reward;
    done[e] = o.done ? 1 : 0;
    const double acc = p.ret_acc[e] + o.st.reward;                               // :443
    p.ret_acc[e] = acc;
    if (o.done) p.ret_last[e] = acc;
    if (counts) { counts[3 * e + 0] = o.mi.sent; counts[3 * e + 1] = o.mi.acked; counts[3 * e + 2] = o.mi.lost; }
    if (info) {
        double *q = info + (size_t)e * PCC_INFO_WIDTH;
        q[0] = o.st.send_rate; q[1] = o.st.recv_rate; q[2] = o.st.avg_lat; q[3] = o.st.loss_ratio;
        q[4] = o.st.lat_infl; q[5] = o.st.lat_ratio; q[6] = o.st.send_ratio; q[7] = o.st.dur;
        q[8] = s.cur_time; q[9] = s.rate; q[10] = s.run_dur; q[11] = s.conn_min;
#ifdef PCC_PROFILE   // profiling build: info = cycles (send phase of the warp, consume, means of this env, phase B of the warp), counts
        q[0] = (double)(prof[1] - prof[0]); q[1] = (double)prof[2]; q[2] = (double)prof[3];
        q[3] = (double)(prof[4] - prof[1]); q[4] = (double)cnt; q[5] = 0; q[6] = 0;
        q[7] = (double)o.mi.sent; q[8] = (double)o.mi.acked; q[9] = (double)(clock64() - tk0);
#endif
    }
}


// ---------------------------------------------------------------------------------------
// kernels (v5: a lane owns an env, see pcc_packed.cuh) -- Philox streams only
// ---------------------------------------------------------------------------------------
// Work partition of a packed step: `perm` is the batch sorted by predicted packets of THIS step (descending).  The
// first n_solo entries (predicted packets above the solo threshold; their serial chain would make a lane-per-env warp
// the step's critical path) get a warp each -- the cooperative single-env path of pcc_warp.cuh --, the rest is cut
// into groups of 32 consecutive entries, one warp per group, one env per lane: similar work on every lane.
struct PackedPartition {
    const int32_t *perm;      // [n] env ids, heaviest first
    const uint32_t *sched;    // [units] launch order: (role << 28) | unit index, longest estimated run time first
    const int32_t *counts;    // device: [0] envs above the solo threshold, [1] envs above the quad threshold (and not solo),
                              //         [2] number of work units (warps with work)
    int32_t n_solo_cap;       // launch geometry covers at most this many solo warps ...
    int32_t n_quad_cap;       // ... and this many quad ENVS (four per warp)
    int32_t wbuf;             // staging capacity (samples) of a solo warp
};
#define PCC_PACKED_THREADS 128
#define PCC_MT_WBUF 4096           // staging capacity (samples) of the MT19937 solo warp
#ifndef PCC_PACKED_SOLO_WBUF
#define PCC_PACKED_SOLO_WBUF 1280
#endif
__host__ __device__ inline size_t packed_warp_smem_bytes()
{
    const size_t a = sizeof(PackedSmem), b = warp_smem_bytes(PCC_PACKED_SOLO_WBUF), c = 4 * sizeof(GroupSmemV2<8>);
    const size_t m = a > b ? (a > c ? a : c) : (b > c ? b : c);
    return (m + 127) & ~(size_t)127;
}

// Phase C of a lane-per-env warp: MI metrics, reward, state write-back per lane; the history ring and the observation
// row (oldest -> newest) of every owned env are written by the whole warp, one contiguous H*F row per instruction.
__device__ __forceinline__ void packed_emit(const DevState &p, double *rowbuf, bool owner, int64_t e, EnvState &s,
                                            const MiOut &mo, double avg_lat, double lat_inc, unsigned long long draws,
                                            unsigned long long head_step, double *__restrict__ obs,
                                            double *__restrict__ reward, uint8_t *__restrict__ done,
                                            int32_t *__restrict__ counts, double *__restrict__ info)
{
    const unsigned lane = threadIdx.x & 31u;
    const int H = p.H, F = p.F, HF = H * F;
    if (owner) {
        MiStats st;
        mi_stats_finish(mo, p.c, avg_lat, lat_inc, s.conn_min, true, st);
        s.steps += 1;                                                                // :419
        if (st.avg_lat > 0.0) s.run_dur = 0.5 * st.avg_lat;                          // :437-438
        const bool dn = s.steps >= p.c.max_steps;                                    // :444
        store_env_dynamic(p, e, s);
        p.draws[e] = draws;
        if (mo.overflow) flag_overflow(p, e);
        for (int f = 0; f < F; f++) rowbuf[lane * PCC_MAX_FEATURES + f] = metric_value(st, p.ids[f]);
        reward[e] = st.reward;
        done[e] = dn ? 1 : 0;
        const double acc = p.ret_acc[e] + st.reward;                                 // :443
        p.ret_acc[e] = acc;
        if (dn) p.ret_last[e] = acc;
        if (counts) { counts[3 * e + 0] = mo.sent; counts[3 * e + 1] = mo.acked; counts[3 * e + 2] = mo.lost; }
        if (info) {
            double *q = info + (size_t)e * PCC_INFO_WIDTH;
            q[0] = st.send_rate; q[1] = st.recv_rate; q[2] = st.avg_lat; q[3] = st.loss_ratio;
            q[4] = st.lat_infl; q[5] = st.lat_ratio; q[6] = st.send_ratio; q[7] = st.dur;
            q[8] = s.cur_time; q[9] = s.rate; q[10] = s.run_dur; q[11] = s.conn_min;
        }
    }
    __syncwarp();
    const unsigned om = __ballot_sync(PCC_FULL, owner);
    const int slot_new = (int)(head_step % (unsigned long long)H);
    double *__restrict__ hist = p.hist;
    if (HF <= 32) {
        // one lane per element of the row; the history reads of eight envs are in flight before their stores
        const int k = (int)lane;
        const int h = k / F, f = k - h * F;
        int sl = slot_new + 1 + h;
        if (sl >= H) sl -= H;
        const bool act = k < HF, isnew = act && (h == H - 1);
        const int src = isnew ? 0 : sl * F + f;
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += 8) {
            if (!((om >> j0) & 0xFFu)) continue;                                     // warp-uniform
            double v[8];
            long long ej[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                ej[u] = __shfl_sync(PCC_FULL, (long long)e, j0 + u);
                v[u] = 0.0;
                if (((om >> (j0 + u)) & 1u) && act && !isnew) v[u] = hist[(size_t)ej[u] * HF + src];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (!((om >> (j0 + u)) & 1u) || !act) continue;
                if (isnew) {
                    v[u] = rowbuf[(j0 + u) * PCC_MAX_FEATURES + f];
                    hist[(size_t)ej[u] * HF + slot_new * F + f] = v[u];            // the new row replaces the oldest slot
                }
                obs[(size_t)ej[u] * HF + k] = v[u];
            }
        }
    } else {
        for (int j = 0; j < 32; j++) {
            if (!((om >> j) & 1u)) continue;                                         // warp-uniform
            const long long ej = __shfl_sync(PCC_FULL, (long long)e, j);
            double *hrow = hist + (size_t)ej * HF;
            double *ob = obs + (size_t)ej * HF;
            for (int k = (int)lane; k < HF; k += 32) {
                const int h = k / F, f = k - h * F;
                double v;
                if (h == H - 1) {
                    v = rowbuf[j * PCC_MAX_FEATURES + f];
                    hrow[slot_new * F + f] = v;                                      // the new row replaces the oldest slot
                } else {
                    int sl = slot_new + 1 + h;
                    if (sl >= H) sl -= H;
                    v = hrow[sl * F + f];
                }
                ob[k] = v;
            }
        }
    }
    __syncwarp();
}

#ifndef PCC_PACKED_MINBLOCKS
#define PCC_PACKED_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(PCC_PACKED_THREADS, PCC_PACKED_MINBLOCKS)
pcc_step_packed_kernel(DevState p, PackedPartition part, unsigned long long head_step,
                       const double *__restrict__ actions, double *__restrict__ obs, double *__restrict__ reward,
                       uint8_t *__restrict__ done, int32_t *__restrict__ counts, double *__restrict__ info)
{
    extern __shared__ double dyn_smem[];      // per warp: packed_warp_smem_bytes()
    const int wib = (int)(threadIdx.x >> 5);
    double *wsm = dyn_smem + (size_t)wib * (packed_warp_smem_bytes() / 8);
    const unsigned lane = threadIdx.x & 31u;
    const int64_t w = (int64_t)blockIdx.x * (PCC_PACKED_THREADS / 32) + wib;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta[META_HEAD] = head_step + 1ull;
    int n_solo = part.counts[0], n_quad = part.counts[1];
    if (n_solo > part.n_solo_cap) n_solo = part.n_solo_cap;
    if ((int64_t)n_solo > p.n) n_solo = (int)p.n;
    if (n_quad > part.n_quad_cap) n_quad = part.n_quad_cap;
    const int nqw = (n_quad + 3) / 4;                                                 // quad warps
    int64_t pos_packed = (int64_t)n_solo + 4 * (int64_t)nqw;                          // first sorted position of the packed part
    if (pos_packed > p.n) pos_packed = p.n;
    // launch order: the unit this warp runs.  Default: role order -- solo, quad, lanes, each heaviest first --, computed
    // here; with a schedule (pcc_schedule_kernel: longest estimated run time first across the roles) looked up
    int role;
    int64_t uj;
    if (part.sched != nullptr) {
        if (w >= part.counts[2]) return;                                              // whole warp
        const uint32_t unit = part.sched[w];
        role = (int)(unit >> 28);
        uj = (int64_t)(unit & 0x0fffffffu);
    } else {
        const int nqw_eff = (int)((pos_packed - n_solo + 3) / 4);
        if (w < n_solo) { role = 0; uj = w; }
        else if (w < (int64_t)n_solo + nqw_eff) { role = 1; uj = w - n_solo; }
        else { role = 2; uj = w - n_solo - nqw_eff; }
    }
    if (role == 1) {
        // ---- four envs of similar (medium / heavy) work, 8 lanes each: the group-cooperative MI of pcc_coop.cuh with the
        // second-generation send phase; the four chain lanes share one instruction stream
        const Grp<8> g;
        const int grp = (int)(lane >> 3);
        int64_t pos = (int64_t)n_solo + 4 * uj + grp;
        const bool alive = pos < p.n;
        if (!alive) pos = p.n - 1;
        const int64_t e = (int64_t)part.perm[pos];
        GroupSmemV2<8> &sm = reinterpret_cast<GroupSmemV2<8> *>(wsm)[grp];
        EnvState s;
        load_env(p, e, s);
        {   // what the epilogue reads (history row, episode return) travels to L2 while the MI runs
            const int HFq = p.H * p.F;
            if (g.gl * 16 < (unsigned)HFq) prefetch_l2_line(p.hist + (size_t)e * HFq + g.gl * 16);
            if (g.gl == 7) prefetch_l2_line(p.ret_acc + e);
        }
        DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
        const uint64_t seed = p.seed[e];
        uint64_t draws = p.draws[e];
        StepOut o;
        s.rate = apply_rate_delta(s.rate, actions[e], p.c);                          // :412
        double *gbuf = p.mean_scratch ? p.mean_scratch + (size_t)e * PCC_GSCRATCH : nullptr;
#ifdef PCC_PROFILE
        const long long tq0 = clock64();
        unsigned long long gtq0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gtq0));
        long long qprof[8];
        run_mi_coop(g, alive, s, ring, seed, draws, s.run_dur, sm, o.mi, qprof, gbuf, PCC_GSCRATCH);
        qprof[6] = clock64();
#else
        run_mi_coop(g, alive, s, ring, seed, draws, s.run_dur, sm, o.mi, nullptr, gbuf, PCC_GSCRATCH);   // :416
#endif
        double avg_lat, lat_inc;
        mi_means_coop(g, alive, o.mi, ring, s.dl, sm, p.need_inc != 0, avg_lat, lat_inc, gbuf, PCC_GSCRATCH);
#ifdef PCC_PROFILE
        qprof[7] = clock64();
#endif
        mi_stats_finish(o.mi, p.c, avg_lat, lat_inc, s.conn_min, true, o.st);
        s.steps += 1;                                                                // :419
        if (o.st.avg_lat > 0.0) s.run_dur = 0.5 * o.st.avg_lat;                      // :437-438
        o.done = s.steps >= p.c.max_steps;                                           // :444
        if (!alive) return;
        coop_emit(g, p, e, head_step, o.st, obs);
        if (g.gl == 0) {
            store_env_dynamic(p, e, s);
            p.draws[e] = draws;
            if (o.mi.overflow) flag_overflow(p, e);
            reward[e] = o.st.reward;
            done[e] = o.done ? 1 : 0;
            const double acc = p.ret_acc[e] + o.st.reward;                           // :443
            p.ret_acc[e] = acc;
            if (o.done) p.ret_last[e] = acc;
            if (counts) { counts[3 * e + 0] = o.mi.sent; counts[3 * e + 1] = o.mi.acked; counts[3 * e + 2] = o.mi.lost; }
            if (info) {
                double *q = info + (size_t)e * PCC_INFO_WIDTH;
                q[0] = o.st.send_rate; q[1] = o.st.recv_rate; q[2] = o.st.avg_lat; q[3] = o.st.loss_ratio;
                q[4] = o.st.lat_infl; q[5] = o.st.lat_ratio; q[6] = o.st.send_ratio; q[7] = o.st.dur;
                q[8] = s.cur_time; q[9] = s.rate; q[10] = s.run_dur; q[11] = s.conn_min;
#ifdef PCC_PROFILE   // profiling build: cycles of send, hop1+bnd1, hop2+bnd2+cross, means; role -4, counts, times
                unsigned long long gt1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
                q[0] = (double)(qprof[1] - qprof[0]); q[1] = (double)(qprof[3] - qprof[1]); q[2] = (double)(qprof[6] - qprof[3]);
                q[3] = 0.0; q[4] = (double)(qprof[7] - qprof[6]); q[5] = (double)(clock64() - qprof[7]);
                q[6] = -4.0; q[7] = (double)o.mi.sent; q[8] = (double)o.mi.acked; q[9] = (double)(clock64() - tq0);
                q[10] = (double)gtq0; q[11] = (double)gt1;
#endif
            }
        }
        return;
    }
    if (role == 0) {
        // ---- a heavy env alone in its warp: cooperative path (chain on lane 0, Philox / scans / means on all lanes)
        const Grp<32> g;
        const bool owner = lane == 0;
        const int64_t e = owner ? (int64_t)part.perm[uj] : 0;
        EnvState s;
        load_env(p, e, s);
        PhiloxRng rng;
        rng.init(p.seed[e], p.draws[e]);
        s.rate = apply_rate_delta(s.rate, actions[e], p.c);                          // :412
        MiOut mo;
        double avg_lat, lat_inc;
#ifdef PCC_PROFILE
        const long long tsolo0 = clock64();
        unsigned long long gtsolo0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gtsolo0));
#endif
#ifdef PCC_PROFILE
        long long sprof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#else
        long long *sprof = nullptr;
#endif
        warp_mi<true, true, false>(g, p, owner, 1, e, s, rng, s.run_dur, wsm, part.wbuf,
                                   *reinterpret_cast<WarpStage *>(wsm), mo, avg_lat, lat_inc, 0, sprof,
                                   p.mean_scratch ? p.mean_scratch + (size_t)__shfl_sync(PCC_FULL, (long long)e, 0) * PCC_GSCRATCH
                                                  : nullptr);   // :416 (scratch rows are per env)
        __syncwarp();
#ifdef PCC_PROFILE
        const long long tsolo1 = clock64();
#endif
        packed_emit(p, wsm, owner, e, s, mo, avg_lat, lat_inc, rng.draws, head_step, obs, reward, done, counts, info);
#ifdef PCC_PROFILE
        if (owner && info) {
            unsigned long long gt1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
            double *q = info + (size_t)e * PCC_INFO_WIDTH;
            for (int k = 0; k < 6; k++) q[k] = 0.0;
            q[0] = (double)(sprof[1] - sprof[0]); q[1] = (double)sprof[2]; q[2] = (double)sprof[3];
            q[6] = -1.0; q[7] = (double)mo.sent; q[8] = (double)mo.acked; q[9] = (double)(tsolo1 - tsolo0);
            q[10] = (double)gtsolo0; q[11] = (double)gt1;
        }
#endif
        return;
    }
    // ---- 32 envs of similar predicted work, one per lane ------------------------------------------------------------
    const int64_t first = pos_packed + uj * 32;
    if (first >= p.n) return;                                                        // whole warp
    const int cnt = (int)((p.n - first < 32) ? (p.n - first) : 32);
    const bool owner = (int)lane < cnt;
    const int64_t e = owner ? (int64_t)part.perm[first + lane] : 0;
    EnvState s;
    load_env(p, e, s);
    if (owner) {   // what the epilogue reads travels to L2 while the MI runs
        const int HFp = p.H * p.F;
        prefetch_l2_line(p.hist + (size_t)e * HFp);
        prefetch_l2_line(p.hist + (size_t)e * HFp + 16);
        if (HFp > 32) prefetch_l2_line(p.hist + (size_t)e * HFp + 32);
        prefetch_l2_line(p.ret_acc + e);
    }
    s.rate = apply_rate_delta(s.rate, actions[e], p.c);                              // :412
    LaneDraws rng;
    rng.init(p.seed[e], p.draws[e], s.lr);
    PackedSmem &sm = *reinterpret_cast<PackedSmem *>(wsm);
    const RingSet rs{p.rings, p.cap, p.cap - 1u};
    PackedOut po;
#ifdef PCC_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long gt0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
    packed_mi<true>(rs, sm, owner, (int)e, s, rng, s.run_dur, p.need_inc != 0, po, prof);
#else
    packed_mi<true>(rs, sm, owner, (int)e, s, rng, s.run_dur, p.need_inc != 0, po);  // :416
#endif
    MiOut mo;
    mo.sent = po.sent; mo.acked = po.acked; mo.lost = po.lost; mo.start = po.start; mo.end = po.end;
    mo.overflow = po.overflow;
    packed_emit(p, wsm, owner, e, s, mo, po.avg_lat, po.lat_inc, rng.draws, head_step, obs, reward, done, counts, info);
#ifdef PCC_PROFILE   // profiling build: info = cycles of the warp's phases A, B1, B2, crossing, B3, emit; role, counts, times
    if (owner && info) {
        const long long tend = clock64();
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
        double *q = info + (size_t)e * PCC_INFO_WIDTH;
        for (int k = 0; k < 5; k++) q[k] = (double)(prof[k + 1] - prof[k]);
        q[5] = (double)(tend - prof[5]); q[6] = (double)cnt; q[7] = (double)mo.sent; q[8] = (double)mo.acked;
        q[9] = (double)(tend - prof[0]); q[10] = (double)gt0; q[11] = (double)gt1;
    }
#endif
}

// MT19937 fidelity mode (the drop-in SimulatedNetworkEnv: CPython's generator, one env per call): one warp per env on the
// cooperative single-env path, the loss draws of a chunk tempered in parallel from the generator state.  Round 1 ran this
// mode on the scalar one-thread kernel (~300 us per step of dependent ring loads).
__global__ void __launch_bounds__(32)
pcc_step_solo_mt_kernel(DevState p, int wbuf, unsigned long long head_step, const double *__restrict__ actions,
                        double *__restrict__ obs, double *__restrict__ reward, uint8_t *__restrict__ done,
                        int32_t *__restrict__ counts, double *__restrict__ info)
{
    extern __shared__ double dyn_smem[];      // warp_smem_bytes(wbuf)
    const Grp<32> g;
    const unsigned lane = threadIdx.x & 31u;
    const int64_t e = (int64_t)blockIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.meta[META_HEAD] = head_step + 1ull;
    const bool owner = lane == 0;
    EnvState s;
    load_env(p, e, s);
    PhiloxRng rng;
    rng.seed = 0ull; rng.draws = 0ull; rng.w2 = 0u; rng.w3 = 0u;
    s.rate = apply_rate_delta(s.rate, actions[e], p.c);                              // :412
    MiOut mo;
    double avg_lat, lat_inc;
    warp_mi<true, true, false>(g, p, owner, 1, e, s, rng, s.run_dur, dyn_smem, wbuf, *reinterpret_cast<WarpStage *>(dyn_smem),
                               mo, avg_lat, lat_inc, 0, nullptr, p.mean_scratch ? p.mean_scratch + (size_t)e * PCC_GSCRATCH : nullptr,
                               nullptr, false, 1, p.mt + (size_t)e * 625);           // :416
    __syncwarp();
    packed_emit(p, dyn_smem, owner, e, s, mo, avg_lat, lat_inc, p.draws[e], head_step, obs, reward, done, counts, info);
}

// predicted packets of the step about to run (its action applied): the sort key of the packed partition
__global__ void pcc_cost_packed_kernel(DevState p, const double *__restrict__ actions, uint32_t solo_packets,
                                       uint32_t quad_packets, uint32_t *__restrict__ keys, int32_t *__restrict__ vals,
                                       int32_t *__restrict__ cnt_w, int32_t *__restrict__ cnt_clear)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) { cnt_clear[0] = 0; cnt_clear[1] = 0; }   // the counters of the NEXT rebalance
    if (e >= p.n) return;
    const double rate = apply_rate_delta(p.rate[e], actions[e], p.c);
    double pk = (p.cur_time[e] + p.run_dur[e] - p.next_send[e]) * rate + 1.0;
    if (!(pk > 0.0)) pk = 0.0;
    const uint32_t key = (uint32_t)fmin(pk, 65535.0);
    keys[e] = key;
    vals[e] = (int32_t)e;
    if (key > solo_packets) atomicAdd(cnt_w, 1);
    else if (key > quad_packets) atomicAdd(cnt_w + 1, 1);
}

// Launch order of the packed step: every warp-sized work unit -- a solo env, four quad envs, 32 lane-per-env envs --
// gets an estimated run time (cycles per packet of its heaviest env, measured per role: tools/phase_profile_packed.py)
// and the units are ranked longest first across the roles (each role's list is already sorted: a rank is the unit's own
// index plus a binary search in each of the two other lists).  Blocks are dispatched in index order, so the long units
// start at once and the short ones fill the tail.
struct SchedCost { float solo, quad, lanes; };
__device__ __forceinline__ int sched_count_greater(const uint32_t *__restrict__ keys, int64_t base, int stride, int count,
                                                   float a, float mine, bool ties_first)
{
    int lo = 0, hi = count;                        // units [0, lo) of the other role run before mine
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float c = a * (float)keys[base + (int64_t)mid * stride];
        if (c > mine || (ties_first && c == mine)) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__global__ void pcc_schedule_kernel(const uint32_t *__restrict__ sorted_keys, int64_t n, int32_t *__restrict__ counts,
                                    int n_solo_cap, int n_quad_cap, SchedCost sc, uint32_t *__restrict__ sched)
{
    int n_solo = counts[0], n_quad = counts[1];
    if (n_solo > n_solo_cap) n_solo = n_solo_cap;
    if ((int64_t)n_solo > n) n_solo = (int)n;
    if (n_quad > n_quad_cap) n_quad = n_quad_cap;
    const int nqw = (n_quad + 3) / 4;
    int64_t pos_packed = (int64_t)n_solo + 4 * (int64_t)nqw;
    if (pos_packed > n) pos_packed = n;
    const int nqw_eff = (int)((pos_packed - n_solo + 3) / 4);
    const int npw = (int)((n - pos_packed + 31) / 32);
    const int units = n_solo + nqw_eff + npw;
    const int u = (int)((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (u == 0) counts[2] = units;
    if (u >= units) return;
    int role, j;
    if (u < n_solo) { role = 0; j = u; }
    else if (u < n_solo + nqw_eff) { role = 1; j = u - n_solo; }
    else { role = 2; j = u - n_solo - nqw_eff; }
    const int64_t base[3] = {0, (int64_t)n_solo, pos_packed};
    const int stride[3] = {1, 4, 32};
    const int count[3] = {n_solo, nqw_eff, npw};
    const float a[3] = {sc.solo, sc.quad, sc.lanes};
    const float mine = a[role] * (float)sorted_keys[base[role] + (int64_t)j * stride[role]];
    int rank = j;
#pragma unroll
    for (int r = 0; r < 3; r++)
        if (r != role) rank += sched_count_greater(sorted_keys, base[r], stride[r], count[r], a[r], mine, r < role);
    sched[rank] = ((uint32_t)role << 28) | (uint32_t)j;
}

// ---------------------------------------------------------------------------------------
// Rollout: K monitor intervals per call with the policy, the value head and the auto-reset on
// the device -- what PPO1's traj_segment_generator does around SimulatedNetworkEnv.step (stable_solve.py:30-58).
// Round 1 fused the K steps into ONE persistent kernel; its work partition was fixed for the whole call, went stale
// and the kernel ended up slower than the K launches it replaced.  The rollout is now a device-side SEQUENCE per step
// -- policy/value kernel, (re-sort +) the regular step kernel, bank gather + masked reset -- enqueued back to back on
// the caller's stream: nothing returns to the host between steps, and every step gets the step kernel's own freshly
// balanced partition.
// ---------------------------------------------------------------------------------------
struct PolicyDev {
    const double *w1, *b1, *w2, *b2, *w3, *b3;   // row-major [out][in]; null w1 = no policy
    const double *vw1, *vb1, *vw2, *vb2, *vw3, *vb3;   // value head of the same shape; null vw1 = none
    int32_t n_in, h1, h2;
    double log_std;
    unsigned long long noise_seed;
    int32_t stochastic;
};

#define PCC_POLICY_MAXH 64
// obs -> tanh(W1 obs + b1) -> tanh(W2 . + b2) -> w3 . + b3   (MlpPolicy of PPO1: two tanh layers, linear output)
__device__ __noinline__ double mlp_eval(const double *w1, const double *b1, const double *w2, const double *b2,
                                        const double *w3, const double *b3, int n_in, int h1, int h2, const double *obs_row)
{
    double a1[PCC_POLICY_MAXH], a2[PCC_POLICY_MAXH];
    // explicit fma(): the policy networks are not part of the bit-exact simulator path (-fmad=false is for that), and a
    // fused multiply-add halves their binary64 instruction count; both evaluation kernels fuse the same way
    for (int i = 0; i < h1; i++) {
        double acc = b1[i];
        for (int j = 0; j < n_in; j++) acc = fma(w1[i * n_in + j], obs_row[j], acc);
        a1[i] = tanh(acc);
    }
    for (int i = 0; i < h2; i++) {
        double acc = b2[i];
        for (int j = 0; j < h1; j++) acc = fma(w2[i * h1 + j], a1[j], acc);
        a2[i] = tanh(acc);
    }
    double out = b3[0];
    for (int j = 0; j < h2; j++) out = fma(w3[j], a2[j], out);
    return out;
}

// deterministic action from a history ring (rows oldest -> newest starting at slot_oldest): the flow monitor's agent
__device__ __noinline__ double policy_action(PolicyDev pol, const double *hrow, int slot_oldest, int H, int F, int64_t e,
                                             unsigned long long t)
{
    double obs_row[128];
    for (int h = 0; h < H; h++) {
        int sl = slot_oldest + h;
        if (sl >= H) sl -= H;
        for (int f = 0; f < F; f++) obs_row[h * F + f] = hrow[sl * F + f];
    }
    (void)e; (void)t;
    return mlp_eval(pol.w1, pol.b1, pol.w2, pol.b2, pol.w3, pol.b3, pol.n_in, pol.h1, pol.h2, obs_row);
}

// The same network by a whole warp for ONE env: lane i owns hidden unit i (and i + 32), the inputs of a layer travel by
// shuffle, every sum runs over j ascending from the bias (the order of mlp_eval).  ~80 dependent steps instead of ~1 500.
__device__ __forceinline__ double mlp_eval_warp(const double *__restrict__ w1, const double *__restrict__ b1,
                                                const double *__restrict__ w2, const double *__restrict__ b2,
                                                const double *__restrict__ w3, const double *__restrict__ b3, int n_in, int h1,
                                                int h2, const double (&x)[4])
{
    const int lane = (int)(threadIdx.x & 31u);
    double a1[2], a2[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {                       // layer 1: units lane, lane + 32
        const int i = lane + 32 * u;
        double acc = (i < h1) ? b1[i] : 0.0;
        for (int j = 0; j < n_in; j++) {
            const double xj = __shfl_sync(PCC_FULL, (j >> 5) == 0 ? x[0] : (j >> 5) == 1 ? x[1] : (j >> 5) == 2 ? x[2] : x[3], j & 31);
            if (i < h1) acc = fma(w1[i * n_in + j], xj, acc);
        }
        a1[u] = (i < h1) ? tanh(acc) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {                       // layer 2
        const int i = lane + 32 * u;
        double acc = (i < h2) ? b2[i] : 0.0;
        for (int j = 0; j < h1; j++) {
            const double aj = __shfl_sync(PCC_FULL, (j >> 5) ? a1[1] : a1[0], j & 31);
            if (i < h2) acc = fma(w2[i * h1 + j], aj, acc);
        }
        a2[u] = (i < h2) ? tanh(acc) : 0.0;
    }
    double out = b3[0];                                  // output: the same ascending sum on every lane
    for (int j = 0; j < h2; j++) out = fma(w3[j], __shfl_sync(PCC_FULL, (j >> 5) ? a2[1] : a2[0], j & 31), out);
    return out;
}

// One THREAD per env (big batches: every weight load is a warp-wide broadcast, all lanes busy)
__global__ void pcc_policy_thread_kernel(PolicyDev pol, int64_t n, const double *__restrict__ obs, unsigned long long t,
                                         double *__restrict__ act_out, double *__restrict__ vpred_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double row[128];
    for (int j = 0; j < pol.n_in; j++) row[j] = obs[(size_t)e * pol.n_in + j];
    if (act_out && pol.w1) {
        double out = mlp_eval(pol.w1, pol.b1, pol.w2, pol.b2, pol.w3, pol.b3, pol.n_in, pol.h1, pol.h2, row);
        if (pol.stochastic) {
            uint32_t c0 = (uint32_t)e, c1 = (uint32_t)(e >> 32), c2 = (uint32_t)t, c3 = 0x4e4f4953u;   // 'NOIS'
            philox4x32_10(c0, c1, c2, c3, (uint32_t)pol.noise_seed, (uint32_t)(pol.noise_seed >> 32));
            const double u1 = (res53(c0, c1) + 1.1102230246251565e-16), u2 = res53(c2, c3);
            out += exp(pol.log_std) * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        }
        act_out[e] = out;
    }
    if (vpred_out && pol.vw1)
        vpred_out[e] = mlp_eval(pol.vw1, pol.vb1, pol.vw2, pol.vb2, pol.vw3, pol.vb3, pol.n_in, pol.h1, pol.h2, row);
}

// One WARP per env (small batches: latency): action = pi(obs) [+ exp(log_std) * N(0,1), Box-Muller on a Philox block
// keyed by (noise_seed; env, step t)], vpred = V(obs).  obs rows are the current observations, oldest -> newest.  Both
// kernels sum in the same order, so a batch gives the same actions whichever runs.
__global__ void __launch_bounds__(128)
pcc_policy_kernel(PolicyDev pol, int64_t n, const double *__restrict__ obs, unsigned long long t,
                  double *__restrict__ act_out, double *__restrict__ vpred_out)
{
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= n) return;                                  // whole warp
    const int lane = (int)(threadIdx.x & 31u);
    double x[4];
#pragma unroll
    for (int u = 0; u < 4; u++) x[u] = (lane + 32 * u < pol.n_in) ? obs[(size_t)e * pol.n_in + lane + 32 * u] : 0.0;
    if (act_out && pol.w1) {
        double out = mlp_eval_warp(pol.w1, pol.b1, pol.w2, pol.b2, pol.w3, pol.b3, pol.n_in, pol.h1, pol.h2, x);
        if (pol.stochastic) {
            uint32_t c0 = (uint32_t)e, c1 = (uint32_t)(e >> 32), c2 = (uint32_t)t, c3 = 0x4e4f4953u;   // 'NOIS'
            philox4x32_10(c0, c1, c2, c3, (uint32_t)pol.noise_seed, (uint32_t)(pol.noise_seed >> 32));
            const double u1 = (res53(c0, c1) + 1.1102230246251565e-16), u2 = res53(c2, c3);
            out += exp(pol.log_std) * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        }
        if (lane == 0) act_out[e] = out;
    }
    if (vpred_out && pol.vw1) {
        const double v = mlp_eval_warp(pol.vw1, pol.vb1, pol.vw2, pol.vb2, pol.vw3, pol.vb3, pol.n_in, pol.h1, pol.h2, x);
        if (lane == 0) vpred_out[e] = v;
    }
}

// SenderHistory.as_array of every env (oldest -> newest) from the history ring: the rollout's first observation
__global__ void pcc_obs_from_hist_kernel(DevState p, unsigned long long head, double *__restrict__ obs)
{
    const int HF = p.H * p.F;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n * HF) return;
    const int64_t e = i / HF;
    const int k = (int)(i - e * HF), h = k / p.F, f = k - h * p.F;
    int sl = (int)(head % (unsigned long long)p.H) + h;     // the slot the next step will overwrite is the oldest
    if (sl >= p.H) sl -= p.H;
    obs[i] = p.hist[(size_t)e * HF + sl * p.F + f];
}

// Auto-reset of a rollout, part 1: for every env that finished its episode in this step (done != 0) fetch the link
// parameters of its next episode from the bank [n_episodes][5][n] (row = resets of this env so far in this rollout)
__global__ void pcc_bank_gather_kernel(DevState p, const uint8_t *__restrict__ done, const double *__restrict__ bank,
                                       int32_t n_episodes, int32_t *__restrict__ ep, uint8_t *__restrict__ mask,
                                       double *__restrict__ bw, double *__restrict__ dl, long long *__restrict__ queue,
                                       double *__restrict__ loss, double *__restrict__ rate)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    uint8_t m = 0;
    if (done[e]) {
        const int j = ep[e];
        if (j >= n_episodes) atomicAdd(&p.meta[META_PART_ERR], 1ull);   // the bank is too short: reported by pcc_check
        else {
            const size_t n = (size_t)p.n;
            const double *b = bank + (size_t)j * 5 * n;
            bw[e] = b[0 * n + e]; dl[e] = b[1 * n + e]; queue[e] = (long long)b[2 * n + e];
            loss[e] = b[3 * n + e]; rate[e] = b[4 * n + e];
            ep[e] = j + 1;
            m = 1;
        }
    }
    mask[e] = m;
}

template <int E>
__global__ void __launch_bounds__(PCC_WARP_THREADS)
pcc_reset_warp_kernel(DevState p, const uint8_t *__restrict__ mask, const double *__restrict__ bw,
                      const double *__restrict__ delay, const long long *__restrict__ queue,
                      const double *__restrict__ loss, const double *__restrict__ start_rate,
                      double *__restrict__ obs)
{
    __shared__ WarpStage sstage[PCC_STAGED_STORES ? PCC_WARP_THREADS / 32 : 1];
    const Grp<32> g;
    const unsigned lane = threadIdx.x & 31u;
    const int64_t warp_global = (int64_t)blockIdx.x * (PCC_WARP_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t slot = warp_global * E + lane;
    if (warp_global * E >= p.n) return;   // whole warp
    bool owner = lane < (unsigned)E && slot < p.n;
    const int64_t e = owner ? slot : 0;
    if (mask && !mask[e]) owner = false;
    if (!__any_sync(PCC_FULL, owner)) return;   // nothing to reset here (the per-step masked reset of a rollout)
    EnvState s;
    const double bwv = bw[e], dlv = delay[e], sr = start_rate[e];
    // reset_env of pcc_core.cuh (network_sim.py:454-484)
    s.d_bw = 1.0 / bwv; s.dl = dlv; s.lr = loss[e]; s.max_qd = (double)queue[e] / bwv;
    s.w_full = tail_drop_threshold(s.d_bw, s.max_qd);
    s.qd = 0.0; s.t_upd = 0.0; s.rate = sr; s.cur_time = 0.0; s.next_send = 1.0 / sr;
    s.run_dur = 3 * dlv; s.conn_min = 0.0;
    s.tail = p.tail[e]; s.h1 = s.tail; s.h2 = s.tail; s.steps = 0;
    PhiloxRng rng;
    rng.init(p.seed[e], p.draws[e]);
    MiOut mo;
    double a, li;
    const int cnt = (int)((p.n - warp_global * E < E) ? (p.n - warp_global * E) : E);
    warp_mi<false, true>(g, p, owner, cnt, e, s, rng, s.run_dur, nullptr, 0, sstage[PCC_STAGED_STORES ? (threadIdx.x >> 5) : 0], mo, a, li);   // :478
    bool ovf = mo.overflow;
    warp_mi<false, true>(g, p, owner, cnt, e, s, rng, s.run_dur, nullptr, 0, sstage[PCC_STAGED_STORES ? (threadIdx.x >> 5) : 0], mo, a, li);   // :479
    ovf = ovf || mo.overflow;
    if (!owner) return;
    const int HF = p.H * p.F;
    for (int k = 0; k < HF; k++) {
        const double v = metric_empty(p.ids[k % p.F]);
        p.hist[(size_t)e * HF + k] = v;
        if (obs) obs[(size_t)e * HF + k] = v;
    }
    p.d_bw[e] = s.d_bw; p.bw[e] = bwv; p.dl[e] = s.dl; p.lr[e] = s.lr; p.max_qd[e] = s.max_qd; p.w_full[e] = s.w_full;
    store_env_dynamic(p, e, s);
    p.draws[e] = rng.draws;
    p.ret_acc[e] = 0.0;
    if (ovf) flag_overflow(p, e);
}

// Phase A as a kernel of its own (split mode): all sends of the MI for every env, one env per lane
// in cost-sorted order -- 32 chains of similar length per warp.  The first `heavy_warps` warps take
// only 4 envs each (the heaviest ones) and draw their Philox blocks with 8 lanes per env.
__global__ void __launch_bounds__(PCC_WARP_THREADS)
pcc_send_kernel(DevState p, const int32_t *__restrict__ perm, int heavy_warps, const double *__restrict__ actions,
                int32_t *__restrict__ sent_tmp)
{
    const unsigned lane = threadIdx.x & 31u;
    const int64_t w = (int64_t)blockIdx.x * (PCC_WARP_THREADS / 32) + (threadIdx.x >> 5);
    int64_t first;
    int cnt;
    if (w < heavy_warps) { first = 4 * w; cnt = 4; }
    else { first = 4 * (int64_t)heavy_warps + 32 * (w - heavy_warps); cnt = 32; }
    if (first >= p.n) return;   // whole warp
    if (first + cnt > p.n) cnt = (int)(p.n - first);
    const bool owner = (int)lane < cnt;
    const int64_t e = owner ? (perm ? (int64_t)perm[first + lane] : first + lane) : 0;
    EnvState s;
    load_env(p, e, s);
    PhiloxRng rng;
    rng.init(p.seed[e], p.draws[e]);
    s.rate = apply_rate_delta(s.rate, actions[e], p.c);                          // :412
    const double end = s.cur_time + s.run_dur;                                   // :124
    const double inv_rate = 1.0 / s.rate;
    DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
    LaneChain c;
    c.t = s.next_send; c.q = s.qd; c.tu = s.t_upd; c.tail = s.tail; c.sent = 0; c.ovf = false;
    if (cnt <= 8) coop_send_phase(c, s, ring, rng, owner, cnt, s.h2, p.cap, end, inv_rate);
    else if (owner) lane_send_phase(c, s, ring, rng, s.h2, p.cap, end, inv_rate);
    if (!owner) return;
    p.rate[e] = s.rate; p.next_send[e] = c.t; p.qd[e] = c.q; p.t_upd[e] = c.tu; p.tail[e] = c.tail;
    p.draws[e] = rng.draws;
    sent_tmp[e] = c.sent | (c.ovf ? (int32_t)0x80000000 : 0);
}

// ---- work-balanced partition (rebalance) -----------------------------------------------------
// cost model of one env-MI in SM cycles: fixed cooperative overhead + per-packet work
struct CostModel { float c0, c1; int32_t target_warps; float heavy_packets; int32_t max_envs; };

__global__ void pcc_cost_kernel(DevState p, CostModel cm, uint32_t *__restrict__ keys, int32_t *__restrict__ vals)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    const double pk = p.rate[e] * p.run_dur[e];            // packets the next MI will send, roughly
    const double c = (double)cm.c0 + (double)cm.c1 * (pk > 0.0 ? pk : 0.0);
    keys[e] = (uint32_t)fmin(c, 4.0e9);
    vals[e] = (int32_t)e;
}

// one block: T = max(heaviest env, total / target_warps); per-env cost floor T/32 caps a warp at 32 envs
__global__ void pcc_target_kernel(const uint32_t *__restrict__ sorted_cost, int64_t n, CostModel cm,
                                  unsigned long long *__restrict__ target)
{
    __shared__ unsigned long long part[32];
    unsigned long long acc = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += sorted_cost[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) tot += part[k];
        unsigned long long t = tot / (unsigned long long)(cm.target_warps > 0 ? cm.target_warps : 1);
        const unsigned long long mx = sorted_cost[0];
        if (t < mx) t = mx;
        if (t < 32) t = 32;
        target[0] = t;
    }
}

__global__ void pcc_costfloor_kernel(const uint32_t *__restrict__ sorted_cost, int64_t n, CostModel cm,
                                     const unsigned long long *__restrict__ target,
                                     unsigned long long *__restrict__ cost64)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long t = target[0];
    const unsigned long long fl = t / (unsigned long long)cm.max_envs + 1ull;   // caps a warp at max_envs envs
    const unsigned long long c = sorted_cost[i];
    // an env expected to send more than heavy_packets gets a warp to itself (cost = T): its chain then
    // runs with 32-lane Philox support (114 cycles/packet measured) instead of stalling 31 co-tenants
    const float pk = ((float)c - cm.c0) / cm.c1;
    cost64[i] = (pk > cm.heavy_packets) ? t : (c > fl ? c : fl);
}

// warp id of sorted position i = floor(exclusive prefix cost / T); ids are gap-free (cost' <= T)
__global__ void pcc_heads_kernel(const unsigned long long *__restrict__ cum_excl, int64_t n,
                                 const unsigned long long *__restrict__ target, int32_t *__restrict__ starts,
                                 int32_t *__restrict__ n_warps, long long max_warps, unsigned long long *meta)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long t = target[0];
    const long long wid = (long long)(cum_excl[i] / t);
    const long long prev = (i == 0) ? -1 : (long long)(cum_excl[i - 1] / t);
    if (wid != prev) starts[wid] = (int32_t)i;
    if (i == n - 1) {
        starts[wid + 1] = (int32_t)n;
        n_warps[0] = (int32_t)(wid + 1);
        if (wid + 1 > max_warps) meta[META_PART_ERR] = (unsigned long long)(wid + 1);   // launch grid too small
    }
}

template <int RNG>
__global__ void pcc_reset_kernel(DevState p, const uint8_t *__restrict__ mask,
                                 const double *__restrict__ bw, const double *__restrict__ delay,
                                 const long long *__restrict__ queue, const double *__restrict__ loss,
                                 const double *__restrict__ start_rate, double *__restrict__ obs)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    if (mask && !mask[e]) return;
    EnvState s;
    s.tail = p.tail[e];
    DevRing ring{p.rings + (size_t)e * p.cap, p.cap - 1u};
    bool ovf;
    if (RNG == PCC_RNG_PHILOX) {
        PhiloxRng rng;
        rng.init(p.seed[e], p.draws[e]);
        ovf = reset_env(s, ring, rng, bw[e], delay[e], loss[e], (int64_t)queue[e], start_rate[e]);
        p.draws[e] = rng.draws;
    } else {
        Mt19937Rng rng;
        rng.init(p.mt + (size_t)e * 625);
        ovf = reset_env(s, ring, rng, bw[e], delay[e], loss[e], (int64_t)queue[e], start_rate[e]);
    }
    p.d_bw[e] = s.d_bw; p.bw[e] = bw[e]; p.dl[e] = s.dl; p.lr[e] = s.lr; p.max_qd[e] = s.max_qd; p.w_full[e] = s.w_full;
    store_env_dynamic(p, e, s);
    p.ret_acc[e] = 0.0;                                 // self.reward_sum = 0.0  (:483)
    if (ovf) flag_overflow(p, e);
    const int H = p.H, F = p.F;
    for (int h = 0; h < H; h++)
        for (int f = 0; f < F; f++) {
            double v = metric_empty(p.ids[f]);
            p.hist[(size_t)e * (H * F) + h * F + f] = v;
            if (obs) obs[(size_t)e * (size_t)(H * F) + h * F + f] = v;
        }
}

__global__ void pcc_seed_philox_kernel(DevState p, const unsigned long long *__restrict__ seeds,
                                       const uint8_t *__restrict__ mask)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n || (mask && !mask[e])) return;
    p.seed[e] = seeds[e];
    p.draws[e] = 0ull;
}

// CPython random.seed(int): init_by_array over the 32-bit limbs of the (non-negative) seed.
__global__ void pcc_seed_mt_kernel(DevState p, const unsigned long long *__restrict__ seeds,
                                   const uint8_t *__restrict__ mask)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n || (mask && !mask[e])) return;
    uint32_t *mt = p.mt + (size_t)e * 625;
    const unsigned long long sd = seeds[e];
    const uint32_t key[2] = {(uint32_t)sd, (uint32_t)(sd >> 32)};
    const int len = key[1] ? 2 : 1;
    mt[0] = 19650218u;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    for (int k = 624; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (int k = 623; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    mt[624] = 624u;
    p.seed[e] = sd;
    p.draws[e] = 0ull;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct pcc_handle_s {
    pcc_config cfg;
    DevState d;
    unsigned long long head;  // steps issued so far (history ring head)
    int64_t launches;
    int block;
    int group;   // lanes per env of the group-cooperative kernels (0 = off)
    int epw;     // envs per warp of the warp-owned kernels (0 = off); the default path
    int rebalance_every;       // re-sort envs by work every this many steps (0 = never)
    bool rebalance_now;
    int64_t steps_since_rebalance;
    uint32_t *sort_keys_in, *sort_keys_out;
    int32_t *sort_vals_in, *perm;
    void *sort_tmp;
    size_t sort_tmp_bytes;
    unsigned long long *cost64, *cum_excl, *target;
    int32_t *starts, *n_warps, *sent_tmp;
    bool split;
    bool mt_scalar;           // MT19937 mode on the scalar one-thread kernel (PCC_B200_MODE=scalar) instead of the solo warp
    bool scalar_sorted;
    int wbuf, warp_threads;   // warp kernel: staging capacity per warp, threads per block
    bool pair;                // small batches: worker + helper warp per partition slot (pcc_step_warp_kernel<false, 1>)
    int n_pair;               // big batches: helpers for the n_pair heaviest slots only (pcc_step_warp_kernel<false, 2>)
    int64_t max_warps;
    CostModel cm;
    // packed mode (pcc_step_packed_kernel): a lane owns an env, the batch re-sorted by predicted packets
    bool packed;
    int packed_every;         // re-sort every this many steps (1 = every step)
    int solo_packets;         // predicted packets above which an env gets a warp to itself
    int quad_packets;         // predicted packets above which an env is run by 8 lanes (four envs per warp)
    int n_solo_cap, n_quad_cap;
    int32_t *n_solo_dev;      // [2][4] device counters (solo, quad, units), alternating between rebalances
    uint32_t *sched;          // launch order of the work units (experiment: PCC_B200_SCHED_COST=solo;quad;lanes cycles per packet)
    SchedCost sched_cost;
    bool use_sched;
    int64_t uniform_steps;    // steps every env has taken since a reset of ALL envs, or -1 when the envs may differ
    int reb_parity;
    // scratch of pcc_rollout: current observations, parameters / mask / episode index of the per-step masked reset
    double *ro_obs, *ro_par;      // [n][H*F]; [4][n] bw, delay, loss, start_rate
    long long *ro_queue;
    int32_t *ro_ep;
    uint8_t *ro_mask;
    // staging for pcc_step_host / pcc_step_host_submit: two slots, device -> host copies on a stream of their own
    double *st_actions[2], *st_obs[2], *st_reward[2], *st_info[2];
    uint8_t *st_done[2];
    int32_t *st_counts[2];
    cudaStream_t copy_stream;
    cudaEvent_t ev_step[2], ev_copy[2];
    int64_t tickets;          // steps submitted through the host path so far
    bool host_small;          // the last submission took the single-stream path (see pcc_step_host_submit)
    cudaStream_t host_stream;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout {
    size_t off_d[15], off_u64[2], off_u32[4], off_hist, off_mt, off_meta, total;
};

static Layout make_layout(const pcc_config *c)
{
    Layout L;
    size_t o = 0, n = (size_t)c->n_envs;
    for (int i = 0; i < 15; i++) { L.off_d[i] = o; o = align_up(o + 8 * n); }
    for (int i = 0; i < 2; i++) { L.off_u64[i] = o; o = align_up(o + 8 * n); }
    for (int i = 0; i < 4; i++) { L.off_u32[i] = o; o = align_up(o + 4 * n); }
    L.off_hist = o; o = align_up(o + 8 * n * (size_t)c->history_len * (size_t)c->n_features);
    L.off_mt = o;
    if (c->rng_kind == PCC_RNG_MT19937) o = align_up(o + 4 * n * 625);
    L.off_meta = o; o = align_up(o + 8 * META_WORDS);
    L.total = o;
    return L;
}

static int validate(const pcc_config *c)
{
    if (!c) return fail(PCC_EINVAL, "null config");
    if (c->abi_version != PCC_ABI_VERSION) return fail(PCC_EINVAL, "abi_version mismatch");
    if (c->n_envs < 1 || c->n_envs > (int64_t)1 << 31) return fail(PCC_EINVAL, "n_envs out of range");
    if (c->history_len < 1 || c->history_len > PCC_MAX_HISTORY) return fail(PCC_EINVAL, "history_len out of range");
    if (c->n_features < 1 || c->n_features > PCC_MAX_FEATURES) return fail(PCC_EINVAL, "n_features out of range");
    for (int i = 0; i < c->n_features; i++)
        if (c->feature_ids[i] < 0 || c->feature_ids[i] >= PCC_N_METRICS) return fail(PCC_EINVAL, "bad feature id");
    if (c->rng_kind != PCC_RNG_MT19937 && c->rng_kind != PCC_RNG_PHILOX) return fail(PCC_EINVAL, "bad rng_kind");
    if (c->ring_capacity < 16 || c->ring_capacity > ((int64_t)1 << 30) || (c->ring_capacity & (c->ring_capacity - 1)))
        return fail(PCC_EINVAL, "ring_capacity must be a power of two in [16, 2^30]");
    if (!(c->consts.max_rate > 0) || !(c->consts.min_rate > 0) || c->consts.max_steps < 1 || c->consts.bytes_per_packet < 1)
        return fail(PCC_EINVAL, "bad consts");
    return PCC_OK;
}

extern "C" {

const char *pcc_last_error(void) { return g_err; }
int pcc_abi_version(void) { return PCC_ABI_VERSION; }

void pcc_default_consts(pcc_consts *c)
{
    c->max_rate = 1000.0; c->min_rate = 40.0; c->delta_scale = 0.025; c->reward_scale = 0.001;
    c->max_steps = 400; c->bytes_per_packet = 1500;
}

void pcc_default_config(pcc_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = PCC_ABI_VERSION;
    cfg->history_len = 10;
    cfg->n_features = 3;
    cfg->feature_ids[0] = PCC_M_SENT_LATENCY_INFLATION;
    cfg->feature_ids[1] = PCC_M_LATENCY_RATIO;
    cfg->feature_ids[2] = PCC_M_SEND_RATIO;
    cfg->rng_kind = PCC_RNG_PHILOX;
    pcc_default_consts(&cfg->consts);
}

int64_t pcc_ring_capacity_for(double max_rate, double min_bw, double max_delay, double max_queue)
{
    // packets in flight <= max_rate * max RTT; an MI lasts at most max(0.5 * max RTT, 3 * delay)
    double rtt = 2.0 * max_delay + max_queue / min_bw;
    double mi = 0.5 * rtt > 3.0 * max_delay ? 0.5 * rtt : 3.0 * max_delay;
    double need = max_rate * (rtt + mi) + 64.0;
    int64_t cap = 16;
    while ((double)cap < need && cap < ((int64_t)1 << 30)) cap <<= 1;
    return cap;
}

int pcc_workspace_bytes(const pcc_config *cfg, uint64_t *state_bytes, uint64_t *ring_bytes)
{
    int rc = validate(cfg);
    if (rc) return rc;
    Layout L = make_layout(cfg);
    if (state_bytes) *state_bytes = L.total;
    if (ring_bytes) *ring_bytes = (uint64_t)cfg->n_envs * (uint64_t)cfg->ring_capacity * sizeof(Rec);
    return PCC_OK;
}

static int build_handle(pcc_handle *out, const pcc_config *cfg, void *state_dev, void *ring_dev, bool init)
{
    int rc = validate(cfg);
    if (rc) return rc;
    if (!out || !state_dev || !ring_dev) return fail(PCC_EINVAL, "null pointer");
    if (((uintptr_t)state_dev & 255) || ((uintptr_t)ring_dev & 15)) return fail(PCC_EINVAL, "workspace misaligned");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PCC_ENODEV, "no CUDA device: libpcc_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PCC_EINVAL, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(cfg->device));
    pcc_handle h = new (std::nothrow) pcc_handle_s();
    if (!h) return fail(PCC_EINVAL, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    Layout L = make_layout(cfg);
    char *b = (char *)state_dev;
    DevState &d = h->d;
    double **dcols[15] = {&d.d_bw, &d.bw, &d.dl, &d.lr, &d.max_qd, &d.qd, &d.t_upd, &d.rate,
                          &d.next_send, &d.cur_time, &d.run_dur, &d.conn_min, &d.ret_acc, &d.ret_last, &d.w_full};
    for (int i = 0; i < 15; i++) *dcols[i] = (double *)(b + L.off_d[i]);
    d.seed = (unsigned long long *)(b + L.off_u64[0]);
    d.draws = (unsigned long long *)(b + L.off_u64[1]);
    d.tail = (uint32_t *)(b + L.off_u32[0]);
    d.h1 = (uint32_t *)(b + L.off_u32[1]);
    d.h2 = (uint32_t *)(b + L.off_u32[2]);
    d.steps = (int32_t *)(b + L.off_u32[3]);
    d.hist = (double *)(b + L.off_hist);
    d.mt = cfg->rng_kind == PCC_RNG_MT19937 ? (uint32_t *)(b + L.off_mt) : nullptr;
    d.meta = (unsigned long long *)(b + L.off_meta);
    d.rings = (Rec *)ring_dev;
    d.cap = (uint32_t)cfg->ring_capacity;
    d.n = cfg->n_envs;
    d.H = cfg->history_len;
    d.F = cfg->n_features;
    for (int i = 0; i < PCC_MAX_FEATURES; i++) d.ids[i] = i < cfg->n_features ? cfg->feature_ids[i] : 0;
    d.need_inc = features_need_increase(d.ids, d.F) ? 1 : 0;
    d.c.max_rate = cfg->consts.max_rate; d.c.min_rate = cfg->consts.min_rate;
    d.c.delta_scale = cfg->consts.delta_scale; d.c.reward_scale = cfg->consts.reward_scale;
    d.c.max_steps = cfg->consts.max_steps; d.c.bytes_per_packet = cfg->consts.bytes_per_packet;
    const char *blk = getenv("PCC_B200_BLOCK");
    h->block = blk ? atoi(blk) : 32;
    if (h->block < 32 || h->block > 1024 || (h->block & 31)) h->block = 32;
    const char *grp = getenv("PCC_B200_GROUP");
    h->group = grp ? atoi(grp) : 8;
    if (h->group != 0 && h->group != 8 && h->group != 16 && h->group != 32) h->group = 8;
    // execution mode: "warp" (default: a warp owns one heavy env or several light ones), "group" (G lanes
    // per env, all envs alike), "scalar" (1 thread per env).
    const char *mode = getenv("PCC_B200_MODE");
    const char *epw = getenv("PCC_B200_EPW");
    const bool small_batch = cfg->n_envs <= 16384;
    h->epw = 0;
    if (!mode || !strcmp(mode, "warp") || !strcmp(mode, "packed")) {
        h->epw = epw ? atoi(epw) : 8;   // static envs per warp, used only when rebalancing is off
        if (h->epw != 4 && h->epw != 8 && h->epw != 16 && h->epw != 32) h->epw = 8;
        h->group = 0;
        // big batches: lane-per-env packed execution (throughput); small batches: warp-per-heavy-env (latency)
        h->packed = mode ? !strcmp(mode, "packed") : !small_batch;
    } else if (!strcmp(mode, "scalar")) {
        h->group = 0;
    } else if (!grp) {
        h->group = small_batch ? 32 : 8;
    }
    if (cfg->rng_kind != PCC_RNG_PHILOX) {   // MT19937 (fidelity mode): one warp per env (pcc_step_solo_mt_kernel)
        h->group = 0; h->epw = 0; h->packed = false;
        h->mt_scalar = mode && !strcmp(mode, "scalar");
        cudaError_t ce = cudaFuncSetAttribute(pcc_step_solo_mt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)warp_smem_bytes(PCC_MT_WBUF));
        if (ce != cudaSuccess) h->mt_scalar = true;
    }
    {
        const char *wb = getenv("PCC_B200_WBUF");
        h->wbuf = wb ? atoi(wb) : (small_batch ? 4096 : 512);   // big batches: small shared buffer -> 16 warps/SM; heavy MIs stage to global scratch
        if (h->wbuf < 128) h->wbuf = 128;
        const char *wt = getenv("PCC_B200_WARP_THREADS");
        h->warp_threads = wt ? atoi(wt) : ((h->wbuf > 1024) ? 64 : PCC_WARP_THREADS);
        if (h->warp_threads != 32 && h->warp_threads != 64 && h->warp_threads != 128) h->warp_threads = 64;
        if (h->epw) {
            const size_t dyn = (size_t)(h->warp_threads / 32) * warp_smem_bytes(h->wbuf);
            {
                const char *pe = getenv("PCC_B200_PAIR");
                h->pair = pe ? atoi(pe) != 0 : small_batch;
                cudaError_t cp = cudaFuncSetAttribute(pcc_step_warp_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)warp_smem_bytes(h->wbuf));
                if (cp != cudaSuccess) h->pair = false;
                const char *np_ = getenv("PCC_B200_NPAIR");
                h->n_pair = np_ ? atoi(np_) : 0;
                if (h->n_pair < 0) h->n_pair = 0;
                cp = cudaFuncSetAttribute(pcc_step_warp_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
                if (cp != cudaSuccess) h->n_pair = 0;
            }
            cudaError_t ce = cudaFuncSetAttribute(pcc_step_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(pcc_step_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            if (ce != cudaSuccess) { delete h; return fail(PCC_ECUDA, "dynamic shared memory: %s", cudaGetErrorString(ce)); }
        }
    }
    const char *reb = getenv("PCC_B200_REBALANCE");
    h->rebalance_every = reb ? atoi(reb) : 16;
    h->rebalance_now = true;
    h->steps_since_rebalance = 0;
    // scalar and group modes can also visit the envs in cost-sorted order (sort only, no partition)
    h->scalar_sorted = (h->epw == 0) && cfg->rng_kind == PCC_RNG_PHILOX && h->rebalance_every > 0 &&
                       getenv("PCC_B200_SORT") != nullptr;   // measured: no gain for these modes
    if ((h->epw || h->scalar_sorted) && h->rebalance_every > 0) {
        const size_t n = (size_t)cfg->n_envs;
        cudaError_t ce = cudaMalloc(&h->sort_keys_in, 4 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sort_keys_out, 4 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sort_vals_in, 4 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->perm, 4 * n);
        h->sort_tmp_bytes = 0;
        if (ce == cudaSuccess)
            ce = cub::DeviceRadixSort::SortPairsDescending(nullptr, h->sort_tmp_bytes, h->sort_keys_in, h->sort_keys_out,
                                                           h->sort_vals_in, h->perm, (int)n);
        size_t scan_bytes = 0;
        if (ce == cudaSuccess)
            ce = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, h->cost64, h->cum_excl, (int)n);
        if (scan_bytes > h->sort_tmp_bytes) h->sort_tmp_bytes = scan_bytes;
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sort_tmp, h->sort_tmp_bytes);
        const char *c0 = getenv("PCC_B200_COST0"), *c1 = getenv("PCC_B200_COST1"), *tw = getenv("PCC_B200_TARGET_WARPS");
        h->cm.c0 = c0 ? (float)atof(c0) : 7000.0f;
        h->cm.c1 = c1 ? (float)atof(c1) : 60.0f;
        const char *me = getenv("PCC_B200_MAX_ENVS");
        h->cm.max_envs = me ? atoi(me) : (cfg->n_envs <= 16384 ? 8 : 32);   // small batches: short warps, the chip is not full anyway
        if (h->cm.max_envs < 1 || h->cm.max_envs > 32) h->cm.max_envs = 32;
        const char *hp = getenv("PCC_B200_HEAVY");
        h->cm.heavy_packets = hp ? (float)atof(hp) : (cfg->n_envs <= 16384 ? 128.0f : 1024.0f);   // small batches: latency first
        h->cm.target_warps = tw ? atoi(tw) : 148 * 32;
        h->max_warps = (int64_t)n;   // worst case: every env heavy; idle warps exit at once
        if (ce == cudaSuccess) ce = cudaMalloc(&h->cost64, 8 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->cum_excl, 8 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->target, 8);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->starts, 4 * (n + 2));
        if (ce == cudaSuccess) ce = cudaMalloc(&h->n_warps, 4);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sent_tmp, 4 * n);
        if (ce == cudaSuccess && (h->wbuf < PCC_GSCRATCH || h->packed) && !getenv("PCC_B200_NO_GSCRATCH"))
            ce = cudaMalloc(&h->d.mean_scratch, (size_t)n * PCC_GSCRATCH * 8);   // one row per (possible) warp
        const char *sp = getenv("PCC_B200_SPLIT");
        h->split = sp ? atoi(sp) != 0 : false;   // two-kernel variant: measured slower, kept for experiments
        if (h->packed) {
            const char *pe = getenv("PCC_B200_PACKED_EVERY"), *so = getenv("PCC_B200_SOLO");
            h->packed_every = pe ? atoi(pe) : 4;   // measured: 1: 0.62, 3-4: 0.59, 8: 0.62, 16: 0.71 ms per step (predictions age slowly, the sort is not free)
            if (h->packed_every < 1) h->packed_every = 1;
            const char *qu = getenv("PCC_B200_QUAD");
            h->solo_packets = so ? atoi(so) : 1600;
            if (h->solo_packets < 32) h->solo_packets = 32;
            if (h->solo_packets > 65534) h->solo_packets = 65534;
            h->quad_packets = qu ? atoi(qu) : 128;
            if (h->quad_packets < 8) h->quad_packets = 8;
            if (h->quad_packets > h->solo_packets) h->quad_packets = h->solo_packets;
            h->n_solo_cap = (int)(n / 32 > 64 ? n / 32 : 64);
            if ((int64_t)h->n_solo_cap > (int64_t)n) h->n_solo_cap = (int)n;
            h->n_quad_cap = (int)(n / 2 > 256 ? n / 2 : 256);
            if ((int64_t)h->n_quad_cap > (int64_t)n) h->n_quad_cap = (int)n;
            if (ce == cudaSuccess) ce = cudaMalloc(&h->n_solo_dev, 8 * sizeof(int32_t));
            if (ce == cudaSuccess) ce = cudaMemset(h->n_solo_dev, 0, 8 * sizeof(int32_t));
            if (ce == cudaSuccess) ce = cudaMalloc(&h->sched, sizeof(uint32_t) * (size_t)(h->n_solo_cap + (h->n_quad_cap + 3) / 4 + (n + 31) / 32 + 8));
            // estimated cycles per packet of a unit's heaviest env, per role (B200, 65 536 envs, profiles/r02_phase_profile_packed.txt)
            const char *cs = getenv("PCC_B200_SCHED_COST");
            h->sched_cost = SchedCost{225.0f, 755.0f, 2900.0f};
            h->use_sched = cs != nullptr;    // default: role order (a merged longest-first order measured no better)
            if (cs && sscanf(cs, "%f;%f;%f", &h->sched_cost.solo, &h->sched_cost.quad, &h->sched_cost.lanes) != 3)
                sscanf(cs, "%f,%f,%f", &h->sched_cost.solo, &h->sched_cost.quad, &h->sched_cost.lanes);
            if (ce == cudaSuccess)
                ce = cudaFuncSetAttribute(pcc_step_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)((PCC_PACKED_THREADS / 32) * packed_warp_smem_bytes()));
        }
        if (ce != cudaSuccess) { delete h; return fail(PCC_ECUDA, "rebalance scratch: %s", cudaGetErrorString(ce)); }
    }
    if (init) {
        cudaError_t e = cudaMemset(state_dev, 0, L.total);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { delete h; return fail(PCC_ECUDA, "state init: %s", cudaGetErrorString(e)); }
        h->head = 0;
    } else {
        unsigned long long head = 0;
        cudaError_t e = cudaMemcpy(&head, d.meta + META_HEAD, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { delete h; return fail(PCC_ECUDA, "attach: %s", cudaGetErrorString(e)); }
        h->head = head;
        h->uniform_steps = -1;    // an adopted workspace: the envs may be anywhere in their episodes
    }
    *out = h;
    return PCC_OK;
}

int pcc_create(pcc_handle *out, const pcc_config *cfg, void *state_dev, void *ring_dev)
{
    return build_handle(out, cfg, state_dev, ring_dev, true);
}
int pcc_attach(pcc_handle *out, const pcc_config *cfg, void *state_dev, void *ring_dev)
{
    return build_handle(out, cfg, state_dev, ring_dev, false);
}

void pcc_destroy(pcc_handle h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    for (int i = 0; i < 2; i++) {
        cudaFree(h->st_actions[i]); cudaFree(h->st_obs[i]);      // st_obs heads the slot's single staging block
        if (h->ev_step[i]) cudaEventDestroy(h->ev_step[i]);
        if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaFree(h->sort_keys_in); cudaFree(h->sort_keys_out); cudaFree(h->sort_vals_in); cudaFree(h->perm); cudaFree(h->sort_tmp);
    cudaFree(h->cost64); cudaFree(h->cum_excl); cudaFree(h->target); cudaFree(h->starts); cudaFree(h->n_warps); cudaFree(h->sent_tmp); cudaFree(h->d.mean_scratch); cudaFree(h->n_solo_dev); cudaFree(h->sched);
    cudaFree(h->ro_obs); cudaFree(h->ro_par); cudaFree(h->ro_queue); cudaFree(h->ro_ep); cudaFree(h->ro_mask);
    delete h;
}

static inline unsigned grid_for(pcc_handle h) { return (unsigned)((h->cfg.n_envs + h->block - 1) / h->block); }

int pcc_seed(pcc_handle h, const uint64_t *seeds_dev, const uint8_t *mask_dev, void *stream)
{
    if (!h || !seeds_dev) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (h->cfg.rng_kind == PCC_RNG_PHILOX)
        pcc_seed_philox_kernel<<<grid_for(h), h->block, 0, st>>>(h->d, (const unsigned long long *)seeds_dev, mask_dev);
    else
        pcc_seed_mt_kernel<<<grid_for(h), h->block, 0, st>>>(h->d, (const unsigned long long *)seeds_dev, mask_dev);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return PCC_OK;
}

int pcc_get_mt_state(pcc_handle h, int64_t env, uint32_t *state_host)
{
    if (!h || !state_host) return fail(PCC_EINVAL, "null pointer");
    if (h->cfg.rng_kind != PCC_RNG_MT19937) return fail(PCC_EINVAL, "handle is not in MT19937 mode");
    if (env < 0 || env >= h->cfg.n_envs) return fail(PCC_EINVAL, "env out of range");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(state_host, h->d.mt + (size_t)env * 625, 625 * 4, cudaMemcpyDeviceToHost));
    return PCC_OK;
}

int pcc_set_mt_state(pcc_handle h, int64_t env, const uint32_t *state_host)
{
    if (!h || !state_host) return fail(PCC_EINVAL, "null pointer");
    if (h->cfg.rng_kind != PCC_RNG_MT19937) return fail(PCC_EINVAL, "handle is not in MT19937 mode");
    if (env < 0 || env >= h->cfg.n_envs) return fail(PCC_EINVAL, "env out of range");
    if (state_host[624] > 624u) return fail(PCC_EINVAL, "bad MT position");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(h->d.mt + (size_t)env * 625, state_host, 625 * 4, cudaMemcpyHostToDevice));
    return PCC_OK;
}

int pcc_reset(pcc_handle h, const uint8_t *mask_dev, const double *bw_dev, const double *delay_dev,
              const int64_t *queue_dev, const double *loss_dev, const double *start_rate_dev,
              double *obs_dev, void *stream)
{
    if (!h || !bw_dev || !delay_dev || !queue_dev || !loss_dev || !start_rate_dev)
        return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned cgrid = h->group ? (unsigned)((h->cfg.n_envs * h->group + PCC_COOP_THREADS - 1) / PCC_COOP_THREADS) : 0u;
#define PCC_RESET_COOP(G_) pcc_reset_coop_kernel<G_><<<cgrid, PCC_COOP_THREADS, 0, st>>>( \
        h->d, mask_dev, bw_dev, delay_dev, (const long long *)queue_dev, loss_dev, start_rate_dev, obs_dev)
    const unsigned wgrid = h->epw ? (unsigned)((h->cfg.n_envs + (int64_t)h->epw * 4 - 1) / ((int64_t)h->epw * 4)) : 0u;
#define PCC_RESET_WARP(E_) pcc_reset_warp_kernel<E_><<<wgrid, PCC_WARP_THREADS, 0, st>>>( \
        h->d, mask_dev, bw_dev, delay_dev, (const long long *)queue_dev, loss_dev, start_rate_dev, obs_dev)
    h->rebalance_now = true;
    h->uniform_steps = mask_dev ? -1 : 0;
    if (h->epw == 4) PCC_RESET_WARP(4);
    else if (h->epw == 8) PCC_RESET_WARP(8);
    else if (h->epw == 16) PCC_RESET_WARP(16);
    else if (h->epw == 32) PCC_RESET_WARP(32);
    else if (h->group == 8) PCC_RESET_COOP(8);
    else if (h->group == 16) PCC_RESET_COOP(16);
    else if (h->group == 32) PCC_RESET_COOP(32);
    else if (h->cfg.rng_kind == PCC_RNG_PHILOX)
        pcc_reset_kernel<PCC_RNG_PHILOX><<<grid_for(h), h->block, 0, st>>>(
            h->d, mask_dev, bw_dev, delay_dev, (const long long *)queue_dev, loss_dev, start_rate_dev, obs_dev);
    else
        pcc_reset_kernel<PCC_RNG_MT19937><<<grid_for(h), h->block, 0, st>>>(
            h->d, mask_dev, bw_dev, delay_dev, (const long long *)queue_dev, loss_dev, start_rate_dev, obs_dev);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return PCC_OK;
}

// cost-sort the envs (descending) and cut the list into warps of ~equal cost (see WarpPartition)
static int rebalance(pcc_handle h, cudaStream_t st)
{
    const int64_t n = h->cfg.n_envs;
    const unsigned kg = (unsigned)((n + 255) / 256);
    pcc_cost_kernel<<<kg, 256, 0, st>>>(h->d, h->cm, h->sort_keys_in, h->sort_vals_in);
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(h->sort_tmp, h->sort_tmp_bytes, h->sort_keys_in,
                                                       h->sort_keys_out, h->sort_vals_in, h->perm, (int)n, 0, 32, st));
    pcc_target_kernel<<<1, 1024, 0, st>>>(h->sort_keys_out, n, h->cm, h->target);
    pcc_costfloor_kernel<<<kg, 256, 0, st>>>(h->sort_keys_out, n, h->cm, h->target, h->cost64);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(h->sort_tmp, h->sort_tmp_bytes, h->cost64, h->cum_excl, (int)n, st));
    pcc_heads_kernel<<<kg, 256, 0, st>>>(h->cum_excl, n, h->target, h->starts, h->n_warps, (long long)h->max_warps, h->d.meta);
    h->rebalance_now = false;
    h->steps_since_rebalance = 0;
    h->launches += 6;
    return PCC_OK;
}

int pcc_step(pcc_handle h, const double *actions_dev, double *obs_dev, double *reward_dev,
             uint8_t *done_dev, int32_t *counts_dev, double *info_dev, void *stream)
{
    if (!h || !actions_dev || !obs_dev || !reward_dev || !done_dev) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned cgrid = h->group ? (unsigned)((h->cfg.n_envs * h->group + PCC_COOP_THREADS - 1) / PCC_COOP_THREADS) : 0u;
#define PCC_STEP_COOP(G_) pcc_step_coop_kernel<G_><<<cgrid, PCC_COOP_THREADS, 0, st>>>( \
        h->d, scalar_perm, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev)
    const int32_t *scalar_perm = nullptr;
    if (h->scalar_sorted) {
        if (h->rebalance_now || h->steps_since_rebalance >= h->rebalance_every) {
            const int64_t n = h->cfg.n_envs;
            pcc_cost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d, h->cm, h->sort_keys_in, h->sort_vals_in);
            CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(h->sort_tmp, h->sort_tmp_bytes, h->sort_keys_in,
                                                               h->sort_keys_out, h->sort_vals_in, h->perm, (int)n, 0, 32, st));
            h->rebalance_now = false;
            h->steps_since_rebalance = 0;
            h->launches += 2;
        }
        h->steps_since_rebalance++;
        scalar_perm = h->perm;
    }
    if (h->packed && h->perm) {
        const int64_t n = h->cfg.n_envs;
        if (h->rebalance_now || h->steps_since_rebalance >= h->packed_every) {
            // sort the batch by the packets this step will send (16-bit keys: two radix passes)
            h->reb_parity ^= 1;
            pcc_cost_packed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
                h->d, actions_dev, (uint32_t)h->solo_packets, (uint32_t)h->quad_packets, h->sort_keys_in, h->sort_vals_in,
                h->n_solo_dev + 4 * h->reb_parity, h->n_solo_dev + 4 * (h->reb_parity ^ 1));
            CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(h->sort_tmp, h->sort_tmp_bytes, h->sort_keys_in,
                                                               h->sort_keys_out, h->sort_vals_in, h->perm, (int)n, 0, 16, st));
            if (h->use_sched) {
                const int64_t max_units = (int64_t)h->n_solo_cap + (h->n_quad_cap + 3) / 4 + (n + 31) / 32;
                pcc_schedule_kernel<<<(unsigned)((max_units + 255) / 256), 256, 0, st>>>(
                    h->sort_keys_out, n, h->n_solo_dev + 4 * h->reb_parity, h->n_solo_cap, h->n_quad_cap, h->sched_cost, h->sched);
                h->launches++;
            }
            h->rebalance_now = false;
            h->steps_since_rebalance = 0;
            h->launches += 2;
        }
        h->steps_since_rebalance++;
        PackedPartition pp{h->perm, h->use_sched ? h->sched : nullptr, h->n_solo_dev + 4 * h->reb_parity, h->n_solo_cap,
                           h->n_quad_cap, PCC_PACKED_SOLO_WBUF};
        const int wpb = PCC_PACKED_THREADS / 32;
        const int64_t nwarps = (int64_t)h->n_solo_cap + (h->n_quad_cap + 3) / 4 + (n + 31) / 32;
        pcc_step_packed_kernel<<<(unsigned)((nwarps + wpb - 1) / wpb), PCC_PACKED_THREADS, wpb * packed_warp_smem_bytes(), st>>>(
            h->d, pp, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev);
    }
    else if (h->epw) {
        WarpPartition part{nullptr, nullptr, nullptr, h->epw, h->wbuf, 0};
        const int wpb = h->warp_threads / 32;
        const size_t dyn = (size_t)wpb * warp_smem_bytes(h->wbuf);
        int64_t nwarps = (h->cfg.n_envs + h->epw - 1) / h->epw;
        if (h->rebalance_every > 0) {
            if (h->rebalance_now || h->steps_since_rebalance >= h->rebalance_every) {
                int rc = rebalance(h, st);
                if (rc) return rc;
            }
            h->steps_since_rebalance++;
            part.perm = h->perm; part.starts = h->starts; part.n_warps = h->n_warps;
            nwarps = h->max_warps;
        }
        const unsigned wgrid = (unsigned)((nwarps + wpb - 1) / wpb);
        if (h->split && part.perm) {
            const int64_t n = h->cfg.n_envs;
            int64_t heavy_envs = n / 64;
            if (heavy_envs > 4096) heavy_envs = 4096;
            const int heavy_warps = (int)(heavy_envs / 4);
            const int64_t sw = heavy_warps + (n - 4 * (int64_t)heavy_warps + 31) / 32;
            pcc_send_kernel<<<(unsigned)((sw + 3) / 4), PCC_WARP_THREADS, 0, st>>>(h->d, h->perm, heavy_warps, actions_dev,
                                                                                 h->sent_tmp);
            pcc_step_warp_kernel<true><<<wgrid, h->warp_threads, dyn, st>>>(h->d, part, h->sent_tmp, h->head, actions_dev,
                                                                           obs_dev, reward_dev, done_dev, counts_dev,
                                                                           info_dev);
            h->launches++;
        } else if (h->pair && part.perm) {
            pcc_step_warp_kernel<false, 1><<<(unsigned)nwarps, 64, warp_smem_bytes(h->wbuf), st>>>(
                h->d, part, nullptr, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev);
        } else if (h->n_pair > 0 && part.perm && wpb >= 2 && nwarps > h->n_pair) {
            const int half = wpb / 2;
            part.n_pair_blocks = (h->n_pair + half - 1) / half;
            const int64_t paired = (int64_t)part.n_pair_blocks * half;
            const unsigned mgrid = (unsigned)(part.n_pair_blocks + (nwarps - paired + wpb - 1) / wpb);
            pcc_step_warp_kernel<false, 2><<<mgrid, h->warp_threads, dyn, st>>>(
                h->d, part, nullptr, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev);
        } else {
            pcc_step_warp_kernel<false><<<wgrid, h->warp_threads, dyn, st>>>(h->d, part, nullptr, h->head, actions_dev,
                                                                            obs_dev, reward_dev, done_dev, counts_dev,
                                                                            info_dev);
        }
    }
    else if (h->group == 8) PCC_STEP_COOP(8);
    else if (h->group == 16) PCC_STEP_COOP(16);
    else if (h->group == 32) PCC_STEP_COOP(32);
    else if (h->cfg.rng_kind == PCC_RNG_PHILOX)
        pcc_step_kernel<PCC_RNG_PHILOX><<<grid_for(h), h->block, 0, st>>>(
            h->d, scalar_perm, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev);
    else if (h->mt_scalar)
        pcc_step_kernel<PCC_RNG_MT19937><<<grid_for(h), h->block, 0, st>>>(
            h->d, nullptr, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev);
    else
        pcc_step_solo_mt_kernel<<<(unsigned)h->cfg.n_envs, 32, warp_smem_bytes(PCC_MT_WBUF), st>>>(
            h->d, PCC_MT_WBUF, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev, info_dev);
    h->head++;
    h->launches++;
    if (h->uniform_steps >= 0) h->uniform_steps++;
    CUDA_TRY(cudaGetLastError());
    return PCC_OK;
}

int pcc_rollout(pcc_handle h, int32_t n_steps, const double *actions_dev, const pcc_policy *policy,
                const double *reset_params_dev, int32_t n_episodes, double *obs_dev, double *actions_out_dev,
                double *reward_dev, uint8_t *done_dev, int32_t *counts_dev, double *vpred_dev, void *stream)
{
    if (!h || !reward_dev || !done_dev) return fail(PCC_EINVAL, "null pointer");
    if (n_steps < 1) return fail(PCC_EINVAL, "n_steps must be positive");
    const bool use_policy = !actions_dev;
    if (use_policy && !(policy && policy->w1)) return fail(PCC_EINVAL, "pcc_rollout needs actions or a policy");
    if (use_policy && !actions_out_dev) return fail(PCC_EINVAL, "pcc_rollout with a policy needs actions_out_dev");
    if (vpred_dev && !(policy && policy->vw1)) return fail(PCC_EINVAL, "vpred_dev needs a policy with a value head");
    if (h->cfg.rng_kind != PCC_RNG_PHILOX) return fail(PCC_EINVAL, "pcc_rollout needs Philox streams");
    if (n_episodes < 0 || (n_episodes > 0 && !reset_params_dev)) return fail(PCC_EINVAL, "bad reset parameter bank");
    const int HF = h->cfg.history_len * h->cfg.n_features;
    if (policy && (policy->w1 || policy->vw1) &&
        (policy->n_in != HF || policy->h1 < 1 || policy->h1 > PCC_POLICY_MAXH || policy->h2 < 1 ||
         policy->h2 > PCC_POLICY_MAXH || HF > 128))
        return fail(PCC_EINVAL, "policy shape does not fit (n_in = history_len * n_features <= 128, hidden <= 64)");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.n_envs;
    if (!h->ro_obs) {
        CUDA_TRY(cudaMalloc(&h->ro_obs, 8 * n * (size_t)HF));
        CUDA_TRY(cudaMalloc(&h->ro_par, 8 * n * 4));
        CUDA_TRY(cudaMalloc(&h->ro_queue, 8 * n));
        CUDA_TRY(cudaMalloc(&h->ro_ep, 4 * n));
        CUDA_TRY(cudaMalloc(&h->ro_mask, n));
    }
    PolicyDev pol;
    memset(&pol, 0, sizeof(pol));
    if (policy) {
        pol.w1 = use_policy ? policy->w1 : nullptr; pol.b1 = policy->b1; pol.w2 = policy->w2; pol.b2 = policy->b2;
        pol.w3 = policy->w3; pol.b3 = policy->b3;
        pol.vw1 = vpred_dev ? policy->vw1 : nullptr; pol.vb1 = policy->vb1; pol.vw2 = policy->vw2; pol.vb2 = policy->vb2;
        pol.vw3 = policy->vw3; pol.vb3 = policy->vb3;
        pol.n_in = policy->n_in; pol.h1 = policy->h1; pol.h2 = policy->h2;
        pol.log_std = policy->log_std; pol.noise_seed = policy->noise_seed; pol.stochastic = policy->stochastic;
    }
    const bool eval = pol.w1 || pol.vw1;
    const unsigned eg = (unsigned)((n * 32 + 127) / 128);   // one warp per env ...
    const bool per_thread = n >= 16384;                     // ... or one thread per env when there are enough envs to fill the GPU that way
    if (eval) {   // the observation the first action is computed from
        pcc_obs_from_hist_kernel<<<(unsigned)((n * HF + 255) / 256), 256, 0, st>>>(h->d, h->head, h->ro_obs);
        h->launches++;
    }
    if (n_episodes > 0) CUDA_TRY(cudaMemsetAsync(h->ro_ep, 0, 4 * n, st));
    double *bw = h->ro_par, *dl = h->ro_par + n, *loss = h->ro_par + 2 * n, *rate = h->ro_par + 3 * n;
    for (int32_t k = 0; k < n_steps; k++) {
        const size_t kn = (size_t)k * n;
        const double *act = actions_dev ? actions_dev + kn : actions_out_dev + kn;
        if (eval) {
            if (per_thread)
                pcc_policy_thread_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(pol, (int64_t)n, h->ro_obs, h->head,
                    pol.w1 ? actions_out_dev + kn : nullptr, pol.vw1 ? vpred_dev + kn : nullptr);
            else
                pcc_policy_kernel<<<eg, 128, 0, st>>>(pol, (int64_t)n, h->ro_obs, h->head, pol.w1 ? actions_out_dev + kn : nullptr,
                                                     pol.vw1 ? vpred_dev + kn : nullptr);
            h->launches++;
        }
        if (actions_dev && actions_out_dev)
            CUDA_TRY(cudaMemcpyAsync(actions_out_dev + kn, actions_dev + kn, 8 * n, cudaMemcpyDeviceToDevice, st));
        int rc = pcc_step(h, act, h->ro_obs, reward_dev + kn, done_dev + kn, counts_dev ? counts_dev + 3 * kn : nullptr,
                          nullptr, stream);
        if (rc) return rc;
        // when every env has taken the same number of steps (the usual case: all reset together) the host knows at
        // which step the episodes end and skips the two launches everywhere else
        const bool may_finish = h->uniform_steps < 0 || h->uniform_steps >= (int64_t)h->cfg.consts.max_steps;
        if (n_episodes > 0 && may_finish) {
            const bool all_finish = h->uniform_steps >= 0;
            // finished envs start their next episode with the next row of the bank (network_sim.py:469-484)
            pcc_bank_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d, done_dev + kn, reset_params_dev, n_episodes,
                                                                             h->ro_ep, h->ro_mask, bw, dl, h->ro_queue, loss, rate);
            h->launches++;
            const bool reb = h->rebalance_now;
            rc = pcc_reset(h, h->ro_mask, bw, dl, (const int64_t *)h->ro_queue, loss, rate, h->ro_obs, stream);
            if (rc) return rc;
            if (!h->packed && !all_finish) h->rebalance_now = reb;   // a stale partition is only slower; the periodic rebalance picks the resets up
            if (all_finish) h->uniform_steps = 0;            // the masked reset just reset every env
        }
        if (obs_dev) CUDA_TRY(cudaMemcpyAsync(obs_dev + kn * HF, h->ro_obs, 8 * n * (size_t)HF, cudaMemcpyDeviceToDevice, st));
    }
    if (pol.vw1) {   // V(observation after the last step): PPO1's nextvpred (before the (1 - new) factor)
        if (per_thread)
            pcc_policy_thread_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(pol, (int64_t)n, h->ro_obs, h->head, nullptr,
                                                                           vpred_dev + (size_t)n_steps * n);
        else
            pcc_policy_kernel<<<eg, 128, 0, st>>>(pol, (int64_t)n, h->ro_obs, h->head, nullptr, vpred_dev + (size_t)n_steps * n);
        h->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return PCC_OK;
}

// Device staging of one slot: ONE allocation laid out [obs | reward | info | counts | done] so that a caller whose host
// buffers have the same layout (the drop-in single env) gets its results with a single copy.
struct StageLayout { size_t obs, reward, info, counts, done, total; };
static StageLayout stage_layout(size_t n, size_t hf)
{
    StageLayout L;
    L.obs = 0; L.reward = 8 * n * hf; L.info = L.reward + 8 * n; L.counts = L.info + 8 * n * PCC_INFO_WIDTH;
    L.done = L.counts + 12 * n; L.total = L.done + n;
    return L;
}
#define PCC_HOST_SMALL_BYTES (64u << 10)   // below this a step's results are latency-, not bandwidth-bound

int pcc_step_host_submit(pcc_handle h, const double *actions_host, double *obs_host, double *reward_host,
                         uint8_t *done_host, int32_t *counts_host, double *info_host, void *stream, int64_t *ticket)
{
    if (!h || !actions_host || !obs_host || !reward_host || !done_host) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.n_envs, hf = (size_t)h->cfg.history_len * h->cfg.n_features;
    const StageLayout SL = stage_layout(n, hf);
    if (!h->copy_stream) {
        for (int i = 0; i < 2; i++) {
            char *blk = nullptr;
            CUDA_TRY(cudaMalloc(&h->st_actions[i], 8 * n));
            CUDA_TRY(cudaMalloc(&blk, (SL.total + 255) & ~(size_t)255));
            h->st_obs[i] = (double *)(blk + SL.obs); h->st_reward[i] = (double *)(blk + SL.reward);
            h->st_info[i] = (double *)(blk + SL.info); h->st_counts[i] = (int32_t *)(blk + SL.counts);
            h->st_done[i] = (uint8_t *)(blk + SL.done);
            CUDA_TRY(cudaEventCreateWithFlags(&h->ev_step[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    }
    const int slot = (int)(h->tickets & 1);
    const bool small = SL.total <= PCC_HOST_SMALL_BYTES;
    // small batches: everything on the caller's stream (no cross-stream events: each costs microseconds that a 30 us
    // step notices); big ones: results travel on the copy stream so that the next submission does not queue behind them
    cudaStream_t cs = small ? st : h->copy_stream;
    // the slot's previous occupant (two submissions ago) must have left for the host before the kernel overwrites it
    if (!small && h->tickets >= 2) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_copy[slot], 0));
    CUDA_TRY(cudaMemcpyAsync(h->st_actions[slot], actions_host, 8 * n, cudaMemcpyHostToDevice, st));
    int rc = pcc_step(h, h->st_actions[slot], h->st_obs[slot], h->st_reward[slot], h->st_done[slot],
                      counts_host ? h->st_counts[slot] : nullptr, info_host ? h->st_info[slot] : nullptr, stream);
    if (rc) return rc;
    if (!small) {
        CUDA_TRY(cudaEventRecord(h->ev_step[slot], st));
        CUDA_TRY(cudaStreamWaitEvent(cs, h->ev_step[slot], 0));
    }
    char *host0 = (char *)obs_host;
    const bool one_block = info_host && counts_host && (char *)reward_host == host0 + SL.reward &&
                           (char *)info_host == host0 + SL.info && (char *)counts_host == host0 + SL.counts &&
                           (char *)done_host == host0 + SL.done;
    if (one_block) {
        CUDA_TRY(cudaMemcpyAsync(host0, h->st_obs[slot], SL.total, cudaMemcpyDeviceToHost, cs));
    } else {
        CUDA_TRY(cudaMemcpyAsync(reward_host, h->st_reward[slot], 8 * n, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(done_host, h->st_done[slot], n, cudaMemcpyDeviceToHost, cs));
        if (counts_host) CUDA_TRY(cudaMemcpyAsync(counts_host, h->st_counts[slot], 12 * n, cudaMemcpyDeviceToHost, cs));
        if (info_host) CUDA_TRY(cudaMemcpyAsync(info_host, h->st_info[slot], 8 * n * PCC_INFO_WIDTH, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(obs_host, h->st_obs[slot], 8 * n * hf, cudaMemcpyDeviceToHost, cs));
    }
    if (!small) CUDA_TRY(cudaEventRecord(h->ev_copy[slot], cs));
    h->host_small = small;
    h->host_stream = st;
    if (ticket) *ticket = h->tickets;
    h->tickets++;
    return PCC_OK;
}

int pcc_step_host_wait(pcc_handle h, int64_t ticket)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    if (ticket < 0 || ticket >= h->tickets) return fail(PCC_EINVAL, "unknown ticket");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->host_small) { CUDA_TRY(cudaStreamSynchronize(h->host_stream)); return PCC_OK; }
    if (ticket + 2 < h->tickets) return PCC_OK;          // its slot has been reused: those copies finished long ago
    CUDA_TRY(cudaEventSynchronize(h->ev_copy[ticket & 1]));
    return PCC_OK;
}

int pcc_step_host(pcc_handle h, const double *actions_host, double *obs_host, double *reward_host,
                  uint8_t *done_host, int32_t *counts_host, void *stream)
{
    int64_t t = 0;
    int rc = pcc_step_host_submit(h, actions_host, obs_host, reward_host, done_host, counts_host, nullptr, stream, &t);
    if (rc) return rc;
    return pcc_step_host_wait(h, t);
}

int pcc_check(pcc_handle h, void *stream)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned long long meta[META_WORDS];
    CUDA_TRY(cudaMemcpy(meta, h->d.meta, sizeof(meta), cudaMemcpyDeviceToHost));
    if (meta[META_PART_ERR])
        return fail(PCC_EINVAL, "internal: work partition produced more warps than were launched");
    if (meta[META_OVF_COUNT]) {
        snprintf(g_err, sizeof(g_err),
                 "in-flight ring overflow in %llu env-MI(s), first env %llu: ring_capacity %lld is too small "
                 "for these link parameters (see pcc_ring_capacity_for)",
                 meta[META_OVF_COUNT], meta[META_OVF_ENV] - 1, (long long)h->cfg.ring_capacity);
        return PCC_EOVERFLOW;
    }
    return PCC_OK;
}

int pcc_get_column(pcc_handle h, const char *name, double *dst_dev, void *stream)
{
    if (!h || !name || !dst_dev) return fail(PCC_EINVAL, "null pointer");
    const DevState &d = h->d;
    const double *src = nullptr;
    if (!strcmp(name, "cur_time")) src = d.cur_time;
    else if (!strcmp(name, "run_dur")) src = d.run_dur;
    else if (!strcmp(name, "rate")) src = d.rate;
    else if (!strcmp(name, "next_send")) src = d.next_send;
    else if (!strcmp(name, "queue_delay")) src = d.qd;
    else if (!strcmp(name, "conn_min")) src = d.conn_min;
    else if (!strcmp(name, "bw")) src = d.bw;
    else if (!strcmp(name, "delay")) src = d.dl;
    else if (!strcmp(name, "loss")) src = d.lr;
    else if (!strcmp(name, "max_queue_delay")) src = d.max_qd;
    else if (!strcmp(name, "episode_return")) src = d.ret_acc;
    else if (!strcmp(name, "last_episode_return")) src = d.ret_last;
    else return fail(PCC_EINVAL, "unknown column %s", name);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(dst_dev, src, 8 * (size_t)h->cfg.n_envs, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PCC_OK;
}

int64_t pcc_launch_count(pcc_handle h) { return h ? h->launches : 0; }

}  // extern "C"

// =======================================================================================
// Several senders per link (BASELINE config 5): generic heap simulator, one env per thread.
// Workspace per env: MEnv header, hist[S][H][F], heap[heap_cap] events, samples[S][cap_s].
// =======================================================================================
struct MEnv {
    MNet net;
    MSender snd[PCC_MAX_SENDERS];
    unsigned long long seed, draws;
};
struct MultiDev {
    MEnv *envs;
    double *hist;        // [n][S][H][F], per env a ring over H slots (handle-global head)
    MEvent *heaps;       // [n][heap_cap]
    double *samples;     // [n][S][cap_s]
    unsigned long long *meta;   // [1] = overflow count
    int64_t n;
    int32_t S, H, F, heap_cap, cap_s;
    int32_t ids[PCC_MAX_FEATURES];
    int32_t need_inc;
    Consts c;
    Variant v;
    MFast *fast;         // [n] heap-free mode: timers + ring cursors; the heap region then holds Rec[cap] + sender ids[cap]
    int32_t ring_cap;    // records per env in heap-free mode (power of two)
};
// heap-free mode: the env's slice of the heap region reinterpreted as the shared in-flight ring (DevSidRing, pcc_multi_warp.cuh)
__device__ __forceinline__ DevSidRing multi_ring(const MultiDev &p, int64_t e)
{
    char *b = reinterpret_cast<char *>(p.heaps + (size_t)e * p.heap_cap);
    return DevSidRing{reinterpret_cast<Rec *>(b), reinterpret_cast<uint8_t *>(b + (size_t)p.ring_cap * sizeof(Rec)),
                      (uint32_t)p.ring_cap - 1u};
}
struct DevHeap {
    MEvent *base; int cap;
    __device__ __forceinline__ int capacity() const { return cap; }
    __device__ __forceinline__ MEvent get(int i) const { return base[i]; }
    __device__ __forceinline__ void set(int i, const MEvent &e) { base[i] = e; }
};

__global__ void pcc_multi_reset_kernel(MultiDev p, const uint8_t *__restrict__ mask, const double *__restrict__ bw,
                                       const double *__restrict__ delay, const long long *__restrict__ queue,
                                       const double *__restrict__ loss, const double *__restrict__ rates,
                                       double *__restrict__ obs)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n || (mask && !mask[e])) return;
    MEnv &me = p.envs[e];
    MNet net = me.net;
    MSender snd[PCC_MAX_SENDERS];
    PhiloxRng rng;
    rng.init(me.seed, me.draws);
    DevHeap heap{p.heaps + (size_t)e * p.heap_cap, p.heap_cap};
    double r[PCC_MAX_SENDERS];
    for (int i = 0; i < p.S; i++) r[i] = rates[(size_t)e * p.S + i];
    const bool ok = multi_reset(net, snd, p.S, heap, p.samples + (size_t)e * p.S * p.cap_s, p.cap_s, rng, bw[e], delay[e],
                                (int64_t)queue[e], loss[e], r, p.v);
    me.net = net;
    for (int i = 0; i < p.S; i++) me.snd[i] = snd[i];
    me.draws = rng.draws;
    if (!ok) atomicAdd(&p.meta[1], 1ull);
    const int HF = p.H * p.F;
    for (int i = 0; i < p.S; i++)
        for (int k = 0; k < HF; k++) {
            const double v = metric_empty(p.ids[k % p.F]);
            p.hist[((size_t)e * p.S + i) * HF + k] = v;
            if (obs) obs[((size_t)e * p.S + i) * HF + k] = v;
        }
}

__global__ void pcc_multi_step_kernel(MultiDev p, unsigned long long head_step, const double *__restrict__ actions,
                                      const double *__restrict__ cwnd_actions,
                                      double *__restrict__ obs, double *__restrict__ reward, uint8_t *__restrict__ done,
                                      int32_t *__restrict__ counts, int32_t *__restrict__ cwnd_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    MEnv &me = p.envs[e];
    MNet net = me.net;
    MSender snd[PCC_MAX_SENDERS];
    for (int i = 0; i < p.S; i++) snd[i] = me.snd[i];
    PhiloxRng rng;
    rng.init(me.seed, me.draws);
    DevHeap heap{p.heaps + (size_t)e * p.heap_cap, p.heap_cap};
    double acts[PCC_MAX_SENDERS], cacts[PCC_MAX_SENDERS], rows[PCC_MAX_SENDERS * PCC_MAX_FEATURES], rew[PCC_MAX_SENDERS];
    int32_t cnt[PCC_MAX_SENDERS * 3];
    for (int i = 0; i < p.S; i++) {
        acts[i] = actions[(size_t)e * p.S + i];
        cacts[i] = cwnd_actions ? cwnd_actions[(size_t)e * p.S + i] : 0.0;
    }
    bool dn;
    const bool ok = multi_step(net, snd, p.S, heap, p.samples + (size_t)e * p.S * p.cap_s, p.cap_s, rng, acts,
                               cwnd_actions ? cacts : nullptr, p.c, p.v, p.ids, p.F, p.need_inc != 0, rows, rew, cnt, dn);
    if (cwnd_out) for (int i = 0; i < p.S; i++) cwnd_out[(size_t)e * p.S + i] = snd[i].cwnd;
    me.net = net;
    for (int i = 0; i < p.S; i++) me.snd[i] = snd[i];
    me.draws = rng.draws;
    if (!ok) atomicAdd(&p.meta[1], 1ull);
    const int H = p.H, F = p.F, HF = H * F;
    const int slot_new = (int)(head_step % (unsigned long long)H);
    for (int i = 0; i < p.S; i++) {
        double *hrow = p.hist + ((size_t)e * p.S + i) * HF;
        double *ob = obs + ((size_t)e * p.S + i) * HF;
        for (int f = 0; f < F; f++) hrow[slot_new * F + f] = rows[i * F + f];
        for (int h = 0; h < H; h++) {
            int sl = slot_new + 1 + h;
            if (sl >= H) sl -= H;
            for (int f = 0; f < F; f++) ob[h * F + f] = hrow[sl * F + f];
        }
        reward[(size_t)e * p.S + i] = rew[i];
        if (counts) for (int k = 0; k < 3; k++) counts[((size_t)e * p.S + i) * 3 + k] = cnt[3 * i + k];
    }
    done[e] = dn ? 1 : 0;
}

__global__ void pcc_multi_seed_kernel(MultiDev p, const unsigned long long *__restrict__ seeds)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    p.envs[e].seed = seeds[e];
    p.envs[e].draws = 0ull;
}

// heap-free mode (pcc_multi_fast.cuh): the same kernels' work on the streaming MI, one env per thread
__global__ void pcc_mfast_reset_kernel(MultiDev p, const uint8_t *__restrict__ mask, const double *__restrict__ bw,
                                       const double *__restrict__ delay, const long long *__restrict__ queue,
                                       const double *__restrict__ loss, const double *__restrict__ rates,
                                       double *__restrict__ obs)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n || (mask && !mask[e])) return;
    MEnv &me = p.envs[e];
    MNet net = me.net;
    MSender snd[PCC_MAX_SENDERS];
    MFast f = p.fast[e];
    PhiloxRng rng;
    rng.init(me.seed, me.draws);
    DevSidRing ring = multi_ring(p, e);
    double r[PCC_MAX_SENDERS];
    for (int i = 0; i < p.S; i++) r[i] = rates[(size_t)e * p.S + i];
    const bool ok = mfast_reset(net, snd, p.S, f, ring, p.samples + (size_t)e * p.S * p.cap_s, p.cap_s, rng, bw[e], delay[e],
                                (int64_t)queue[e], loss[e], r);
    me.net = net;
    for (int i = 0; i < p.S; i++) me.snd[i] = snd[i];
    me.draws = rng.draws;
    p.fast[e] = f;
    if (!ok) atomicAdd(&p.meta[1], 1ull);
    const int HF = p.H * p.F;
    for (int i = 0; i < p.S; i++)
        for (int k = 0; k < HF; k++) {
            const double v = metric_empty(p.ids[k % p.F]);
            p.hist[((size_t)e * p.S + i) * HF + k] = v;
            if (obs) obs[((size_t)e * p.S + i) * HF + k] = v;
        }
}

__global__ void pcc_mfast_step_kernel(MultiDev p, unsigned long long head_step, const double *__restrict__ actions,
                                      double *__restrict__ obs, double *__restrict__ reward, uint8_t *__restrict__ done,
                                      int32_t *__restrict__ counts)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    MEnv &me = p.envs[e];
    MNet net = me.net;
    MSender snd[PCC_MAX_SENDERS];
    for (int i = 0; i < p.S; i++) snd[i] = me.snd[i];
    MFast f = p.fast[e];
    PhiloxRng rng;
    rng.init(me.seed, me.draws);
    DevSidRing ring = multi_ring(p, e);
    double acts[PCC_MAX_SENDERS], rows[PCC_MAX_SENDERS * PCC_MAX_FEATURES], rew[PCC_MAX_SENDERS];
    int32_t cnt[PCC_MAX_SENDERS * 3];
    for (int i = 0; i < p.S; i++) acts[i] = actions[(size_t)e * p.S + i];
    bool dn;
    const bool ok = mfast_step(net, snd, p.S, f, ring, p.samples + (size_t)e * p.S * p.cap_s, p.cap_s, rng, acts, p.c,
                               p.ids, p.F, p.need_inc != 0, rows, rew, cnt, dn);
    me.net = net;
    for (int i = 0; i < p.S; i++) me.snd[i] = snd[i];
    me.draws = rng.draws;
    p.fast[e] = f;
    if (!ok) atomicAdd(&p.meta[1], 1ull);
    const int H = p.H, F = p.F, HF = H * F;
    const int slot_new = (int)(head_step % (unsigned long long)H);
    for (int i = 0; i < p.S; i++) {
        double *hrow = p.hist + ((size_t)e * p.S + i) * HF;
        double *ob = obs + ((size_t)e * p.S + i) * HF;
        for (int k = 0; k < F; k++) hrow[slot_new * F + k] = rows[i * F + k];
        for (int h = 0; h < H; h++) {
            int sl = slot_new + 1 + h;
            if (sl >= H) sl -= H;
            for (int k = 0; k < F; k++) ob[h * F + k] = hrow[sl * F + k];
        }
        reward[(size_t)e * p.S + i] = rew[i];
        if (counts) for (int k = 0; k < 3; k++) counts[((size_t)e * p.S + i) * 3 + k] = cnt[3 * i + k];
    }
    done[e] = dn ? 1 : 0;
}

// heap-free mode, one link per warp (pcc_multi_warp.cuh): the default engine of the streaming MI.  `perm` = links in
// descending predicted cost (the heaviest links start first), or null.
template <int S>
__global__ void __launch_bounds__(128, 4) pcc_mwarp_step_kernel(MultiDev p, const int32_t *__restrict__ perm,
                                                             unsigned long long head_step, const double *__restrict__ actions,
                                                             double *__restrict__ obs, double *__restrict__ reward,
                                                             uint8_t *__restrict__ done, int32_t *__restrict__ counts)
{
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= p.n) return;
    const unsigned lane = threadIdx.x & 31u;
    const int64_t e = perm ? (int64_t)perm[w] : w;
    MEnv &me = p.envs[e];
    MNet net = me.net;
    MSender snd[S];
#pragma unroll
    for (int i = 0; i < S; i++) {
        snd[i] = me.snd[i];
        snd[i].rate = apply_rate_delta(snd[i].rate, actions[(size_t)e * S + i], p.c);   // network_sim.py:409-412
    }
    MFast f = p.fast[e];
    const uint64_t seed = me.seed;
    uint64_t draws = me.draws;
    DevSidRing ring = multi_ring(p, e);
    double *smp = p.samples + (size_t)e * S * p.cap_s;
    __shared__ MwSendSmem send_sm[4];
    const bool ok = mwarp_run_for_dur<S>(net, snd, f, ring, smp, p.cap_s, seed, draws, net.run_dur,
                                         send_sm[threadIdx.x >> 5]);   // :416
    double avg[S], inc[S];
    mwarp_means<S>(snd, smp, p.cap_s, p.need_inc != 0, avg, inc);
    const int H = p.H, F = p.F, HF = H * F;
    const int slot_new = (int)(head_step % (unsigned long long)H);
    double avg0 = 0.0;
#pragma unroll
    for (int i = 0; i < S; i++) {
        MiOut o;
        o.sent = snd[i].sent; o.acked = snd[i].acked; o.lost = snd[i].lost; o.start = snd[i].obs_start; o.end = net.cur_time;
        MiStats st;
        mi_stats_finish(o, p.c, avg[i], inc[i], snd[i].conn_min, true, st);
        double *hrow = p.hist + ((size_t)e * S + i) * HF;
        for (int k = 0; k < F; k++) {
            const double v = metric_value(st, p.ids[k]);
            if (lane == (unsigned)k) hrow[slot_new * F + k] = v;
        }
        if (lane == 0) {
            reward[(size_t)e * S + i] = st.reward;
            if (counts) {
                int32_t *c = counts + ((size_t)e * S + i) * 3;
                c[0] = snd[i].sent; c[1] = snd[i].acked; c[2] = snd[i].lost;
            }
        }
        if (i == 0) avg0 = st.avg_lat;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < S; i++) {
        const double *hrow = p.hist + ((size_t)e * S + i) * HF;
        double *ob = obs + ((size_t)e * S + i) * HF;
        for (int idx = (int)lane; idx < HF; idx += 32) {
            const int hh = idx / F, k = idx - hh * F;
            int sl = slot_new + 1 + hh;
            if (sl >= H) sl -= H;
            ob[idx] = hrow[sl * F + k];
        }
    }
    net.steps += 1;                                                  // :419
    if (avg0 > 0.0) net.run_dur = 0.5 * avg0;                        // :437-438 (sender 0, as written)
    if (lane == 0) {
        me.net = net;
#pragma unroll
        for (int i = 0; i < S; i++) me.snd[i] = snd[i];
        me.draws = draws;
        p.fast[e] = f;
        if (!ok) atomicAdd(&p.meta[1], 1ull);
        done[e] = net.steps >= p.c.max_steps ? 1 : 0;                // :444
    }
}

// reset of the links in `mask` (all if null), one link per warp: fresh link + senders and the two discarded warm-up MIs
// (network_sim.py:454-484; mfast_reset of pcc_multi_fast.cuh on the warp machinery)
template <int S>
__global__ void __launch_bounds__(128, 4) pcc_mwarp_reset_kernel(MultiDev p, const uint8_t *__restrict__ mask,
                                                              const double *__restrict__ bw, const double *__restrict__ delay,
                                                              const long long *__restrict__ queue,
                                                              const double *__restrict__ loss, const double *__restrict__ rates,
                                                              double *__restrict__ obs)
{
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= p.n || (mask && !mask[e])) return;                      // whole warp
    const unsigned lane = threadIdx.x & 31u;
    MEnv &me = p.envs[e];
    MNet net = me.net;
    MFast f = p.fast[e];
    MSender snd[S];
    const double dl = delay[e];
    net.d_bw = 1.0 / bw[e]; net.dl = dl; net.lr = loss[e]; net.max_qd = (double)queue[e] / bw[e];
    net.w_full = tail_drop_threshold(net.d_bw, net.max_qd);
    net.qd = 0.0; net.t_upd = 0.0; net.cur_time = 0.0; net.run_dur = 3 * dl; net.steps = 0; net.heap_n = 0;
    f.h1 = f.tail; f.h2 = f.tail;                                    // drop everything in flight; positions keep counting
#pragma unroll
    for (int i = 0; i < S; i++) {
        const double r = rates[(size_t)e * S + i];
        snd[i].rate = r; snd[i].conn_min = 0.0;
        snd[i].sent = snd[i].acked = snd[i].lost = snd[i].n_rtt = 0; snd[i].obs_start = 0.0;
        snd[i].cwnd = 0; snd[i].inflight = 0;
        f.next_send[i] = 1.0 / r;                                    // queue_initial_packets :107-111
    }
    const uint64_t seed = me.seed;
    uint64_t draws = me.draws;
    DevSidRing ring = multi_ring(p, e);
    double *smp = p.samples + (size_t)e * S * p.cap_s;
    __shared__ MwSendSmem send_sm[4];
    MwSendSmem &sm = send_sm[threadIdx.x >> 5];
    bool ok = mwarp_run_for_dur<S>(net, snd, f, ring, smp, p.cap_s, seed, draws, net.run_dur, sm);   // :478
    ok = mwarp_run_for_dur<S>(net, snd, f, ring, smp, p.cap_s, seed, draws, net.run_dur, sm) && ok;  // :479
    if (lane == 0) {
        me.net = net;
#pragma unroll
        for (int i = 0; i < S; i++) me.snd[i] = snd[i];
        me.draws = draws;
        p.fast[e] = f;
        if (!ok) atomicAdd(&p.meta[1], 1ull);
    }
    const int HF = p.H * p.F;
    for (int k = (int)lane; k < S * HF; k += 32) {
        const double v = metric_empty(p.ids[(k % HF) % p.F]);
        p.hist[(size_t)e * S * HF + k] = v;
        if (obs) obs[(size_t)e * S * HF + k] = v;
    }
}

// predicted packets of the next MI per link (sort key, 16 bits) + identity values for the sort
__global__ void pcc_mcost_kernel(MultiDev p, uint32_t *__restrict__ keys, int32_t *__restrict__ vals)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    const MEnv &me = p.envs[e];
    double r = 0.0;
    for (int i = 0; i < p.S; i++) r += me.snd[i].rate;
    const double pk = me.net.run_dur * r;
    keys[e] = pk >= 65535.0 ? 65535u : (uint32_t)pk;
    vals[e] = (int32_t)e;
}

struct pcc_multi_handle_s {
    pcc_config cfg;
    MultiDev d;
    unsigned long long head;
    int mode;            // 0 undecided (before the first reset), 1 heap-free streaming MI, 2 event heap
    bool heap_forced;    // PCC_MULTI_MODE=heap
    bool thread_forced;  // PCC_MULTI_MODE=thread: the streaming MI with one link per thread (pcc_mfast_step_kernel)
    // link-per-warp engine: links visited in descending predicted cost, re-sorted every `sort_every` steps
    uint32_t *sort_keys_in, *sort_keys_out;
    int32_t *sort_vals_in, *perm;
    void *sort_tmp;
    size_t sort_tmp_bytes;
    int sort_every, since_sort;
    bool sort_now;
    int64_t launches;
};

static void multi_free_sort(pcc_multi_handle_s *h)
{
    cudaFree(h->sort_keys_in); cudaFree(h->sort_keys_out); cudaFree(h->sort_vals_in); cudaFree(h->perm); cudaFree(h->sort_tmp);
    h->sort_keys_in = h->sort_keys_out = nullptr; h->sort_vals_in = h->perm = nullptr; h->sort_tmp = nullptr;
}

static void multi_layout(const pcc_config *cfg, int S, size_t off[6], size_t &total)
{
    const size_t n = (size_t)cfg->n_envs, HF = (size_t)cfg->history_len * cfg->n_features;
    size_t o = 0;
    off[0] = o; o = align_up(o + n * sizeof(MEnv));
    off[1] = o; o = align_up(o + n * S * HF * 8);
    off[2] = o; o = align_up(o + n * (size_t)S * (size_t)cfg->ring_capacity * sizeof(MEvent));
    off[3] = o; o = align_up(o + n * (size_t)S * (size_t)cfg->ring_capacity * 8);
    off[4] = o; o = align_up(o + 64);
    off[5] = o; o = align_up(o + n * sizeof(MFast));
    total = o;
}

extern "C" {

int pcc_multi_workspace_bytes(const pcc_config *cfg, int32_t n_senders, uint64_t *bytes)
{
    int rc = validate(cfg);
    if (rc) return rc;
    if (n_senders < 1 || n_senders > PCC_MAX_SENDERS) return fail(PCC_EINVAL, "n_senders out of range (1..4)");
    size_t off[6], total;
    multi_layout(cfg, n_senders, off, total);
    if (bytes) *bytes = total;
    return PCC_OK;
}

int pcc_multi_create(pcc_multi_handle *out, const pcc_config *cfg, int32_t n_senders, void *workspace_dev)
{
    uint64_t bytes = 0;
    int rc = pcc_multi_workspace_bytes(cfg, n_senders, &bytes);
    if (rc) return rc;
    if (!out || !workspace_dev || ((uintptr_t)workspace_dev & 255)) return fail(PCC_EINVAL, "bad workspace pointer");
    if (cfg->rng_kind != PCC_RNG_PHILOX) return fail(PCC_EINVAL, "the multi-sender path supports Philox streams only");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(PCC_ENODEV, "no CUDA device: libpcc_b200 has no CPU fallback");
    CUDA_TRY(cudaSetDevice(cfg->device));
    pcc_multi_handle h = new (std::nothrow) pcc_multi_handle_s();
    if (!h) return fail(PCC_EINVAL, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    size_t off[6], total;
    multi_layout(cfg, n_senders, off, total);
    char *b = (char *)workspace_dev;
    MultiDev &d = h->d;
    d.envs = (MEnv *)(b + off[0]); d.hist = (double *)(b + off[1]); d.heaps = (MEvent *)(b + off[2]);
    d.samples = (double *)(b + off[3]); d.meta = (unsigned long long *)(b + off[4]);
    d.fast = (MFast *)(b + off[5]); d.ring_cap = (int32_t)cfg->ring_capacity;
    {
        const char *mm = getenv("PCC_MULTI_MODE");
        h->heap_forced = mm && !strcmp(mm, "heap");
        h->thread_forced = mm && !strcmp(mm, "thread");
        h->mode = 0;
        const char *se = getenv("PCC_MULTI_SORT_EVERY");
        h->sort_every = se ? atoi(se) : 8;       // 0: no sort (links in index order)
        h->sort_now = true;
    }
    if (!h->heap_forced && !h->thread_forced && h->sort_every > 0) {
        const size_t n = (size_t)cfg->n_envs;
        cudaError_t ce = cudaMalloc(&h->sort_keys_in, 4 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sort_keys_out, 4 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sort_vals_in, 4 * n);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->perm, 4 * n);
        if (ce == cudaSuccess)
            ce = cub::DeviceRadixSort::SortPairsDescending(nullptr, h->sort_tmp_bytes, h->sort_keys_in, h->sort_keys_out,
                                                           h->sort_vals_in, h->perm, (int)n, 0, 16);
        if (ce == cudaSuccess) ce = cudaMalloc(&h->sort_tmp, h->sort_tmp_bytes);
        if (ce != cudaSuccess) {
            multi_free_sort(h);
            delete h;
            return fail(PCC_ECUDA, "multi create: %s", cudaGetErrorString(ce));
        }
    }
    d.n = cfg->n_envs; d.S = n_senders; d.H = cfg->history_len; d.F = cfg->n_features;
    d.heap_cap = (int32_t)(cfg->ring_capacity * n_senders); d.cap_s = (int32_t)cfg->ring_capacity;
    for (int i = 0; i < PCC_MAX_FEATURES; i++) d.ids[i] = i < cfg->n_features ? cfg->feature_ids[i] : 0;
    d.need_inc = features_need_increase(d.ids, d.F) ? 1 : 0;
    d.c.max_rate = cfg->consts.max_rate; d.c.min_rate = cfg->consts.min_rate; d.c.delta_scale = cfg->consts.delta_scale;
    d.c.reward_scale = cfg->consts.reward_scale; d.c.max_steps = cfg->consts.max_steps;
    d.c.bytes_per_packet = cfg->consts.bytes_per_packet;
    d.v = default_variant();
    cudaError_t e = cudaMemset(b + off[0], 0, off[1] - off[0]);
    if (e == cudaSuccess) e = cudaMemset(b + off[4], 0, 64);
    if (e == cudaSuccess) e = cudaMemset(b + off[5], 0, total - off[5]);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { multi_free_sort(h); delete h; return fail(PCC_ECUDA, "multi init: %s", cudaGetErrorString(e)); }
    *out = h;
    return PCC_OK;
}

void pcc_multi_destroy(pcc_multi_handle h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    multi_free_sort(h);
    delete h;
}

int pcc_multi_seed(pcc_multi_handle h, const uint64_t *seeds_dev, void *stream)
{
    if (!h || !seeds_dev) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    pcc_multi_seed_kernel<<<(unsigned)((h->d.n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->d, (const unsigned long long *)seeds_dev);
    CUDA_TRY(cudaGetLastError());
    return PCC_OK;
}

int pcc_multi_reset(pcc_multi_handle h, const uint8_t *mask_dev, const double *bw_dev, const double *delay_dev,
                    const int64_t *queue_dev, const double *loss_dev, const double *start_rates_dev, double *obs_dev,
                    void *stream)
{
    if (!h || !bw_dev || !delay_dev || !queue_dev || !loss_dev || !start_rates_dev) return fail(PCC_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    // The streaming MI (no heap) serves the shipped configuration; the cwnd / latency-noise variants need the heap.
    // The two keep different per-env state, so the mode is fixed by the first reset.
    const int want = (h->heap_forced || h->d.v.use_cwnd || h->d.v.use_noise) ? 2 : 1;
    if (h->mode == 0) h->mode = want;
    else if (h->mode != want) return fail(PCC_EINVAL, "the variant switches changed after the first reset (create a new handle)");
    if (h->mode == 1 && !h->thread_forced) {
        const unsigned grid = (unsigned)((h->d.n + 3) / 4);
#define PCC_MWARP_RESET(S_)                                                                                      \
        pcc_mwarp_reset_kernel<S_><<<grid, 128, 0, (cudaStream_t)stream>>>(h->d, mask_dev, bw_dev, delay_dev,           \
            (const long long *)queue_dev, loss_dev, start_rates_dev, obs_dev)
        switch (h->d.S) {
            case 1: PCC_MWARP_RESET(1); break;
            case 2: PCC_MWARP_RESET(2); break;
            case 3: PCC_MWARP_RESET(3); break;
            default: PCC_MWARP_RESET(4); break;
        }
#undef PCC_MWARP_RESET
    } else if (h->mode == 1)
        pcc_mfast_reset_kernel<<<(unsigned)((h->d.n + 31) / 32), 32, 0, (cudaStream_t)stream>>>(
            h->d, mask_dev, bw_dev, delay_dev, (const long long *)queue_dev, loss_dev, start_rates_dev, obs_dev);
    else
    pcc_multi_reset_kernel<<<(unsigned)((h->d.n + 31) / 32), 32, 0, (cudaStream_t)stream>>>(
        h->d, mask_dev, bw_dev, delay_dev, (const long long *)queue_dev, loss_dev, start_rates_dev, obs_dev);
    CUDA_TRY(cudaGetLastError());
    h->sort_now = true;
    h->launches++;
    return PCC_OK;
}

int pcc_multi_step_cwnd(pcc_multi_handle h, const double *actions_dev, const double *cwnd_actions_dev, double *obs_dev,
                        double *reward_dev, uint8_t *done_dev, int32_t *counts_dev, int32_t *cwnd_dev, void *stream)
{
    if (!h || !actions_dev || !obs_dev || !reward_dev || !done_dev) return fail(PCC_EINVAL, "null pointer");
    if (cwnd_actions_dev && !h->d.v.use_cwnd) return fail(PCC_EINVAL, "cwnd actions given but use_cwnd is off (pcc_multi_set_variant)");
    if (h->mode == 0) return fail(PCC_EINVAL, "pcc_multi_step before pcc_multi_reset");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->mode == 1) {
        cudaStream_t st = (cudaStream_t)stream;
        if (h->thread_forced)
            pcc_mfast_step_kernel<<<(unsigned)((h->d.n + 31) / 32), 32, 0, st>>>(
                h->d, h->head, actions_dev, obs_dev, reward_dev, done_dev, counts_dev);
        else {
            const int64_t n = h->d.n;
            if (h->perm && (h->sort_now || h->since_sort >= h->sort_every)) {
                pcc_mcost_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->d, h->sort_keys_in, h->sort_vals_in);
                h->launches++;
                CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(h->sort_tmp, h->sort_tmp_bytes, h->sort_keys_in,
                                                                   h->sort_keys_out, h->sort_vals_in, h->perm, (int)n, 0, 16, st));
                h->sort_now = false;
                h->since_sort = 0;
            }
            h->since_sort++;
            const unsigned grid = (unsigned)((n + 3) / 4);
#define PCC_MWARP_LAUNCH(S_)                                                                                     \
            pcc_mwarp_step_kernel<S_><<<grid, 128, 0, st>>>(h->d, h->perm, h->head, actions_dev, obs_dev, reward_dev, \
                                                            done_dev, counts_dev)
            switch (h->d.S) {
                case 1: PCC_MWARP_LAUNCH(1); break;
                case 2: PCC_MWARP_LAUNCH(2); break;
                case 3: PCC_MWARP_LAUNCH(3); break;
                default: PCC_MWARP_LAUNCH(4); break;
            }
#undef PCC_MWARP_LAUNCH
        }
        if (cwnd_dev) CUDA_TRY(cudaMemsetAsync(cwnd_dev, 0, (size_t)h->d.n * h->d.S * sizeof(int32_t), (cudaStream_t)stream));
    } else
    pcc_multi_step_kernel<<<(unsigned)((h->d.n + 31) / 32), 32, 0, (cudaStream_t)stream>>>(
        h->d, h->head, actions_dev, cwnd_actions_dev, obs_dev, reward_dev, done_dev, counts_dev, cwnd_dev);
    h->head++;
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return PCC_OK;
}

int pcc_multi_step(pcc_multi_handle h, const double *actions_dev, double *obs_dev, double *reward_dev, uint8_t *done_dev,
                   int32_t *counts_dev, void *stream)
{
    return pcc_multi_step_cwnd(h, actions_dev, nullptr, obs_dev, reward_dev, done_dev, counts_dev, nullptr, stream);
}

void pcc_default_variant(pcc_variant *v)
{
    if (!v) return;
    memset(v, 0, sizeof(*v));
    v->max_latency_noise = 1.1; v->initial_cwnd = 25; v->min_cwnd = 4; v->max_cwnd = 5000;
}

int pcc_multi_set_variant(pcc_multi_handle h, const pcc_variant *v)
{
    if (!h || !v) return fail(PCC_EINVAL, "null pointer");
    if (v->min_cwnd < 1 || v->max_cwnd < v->min_cwnd || v->max_cwnd > (1 << 30) || v->initial_cwnd < 1 ||
        !(v->max_latency_noise >= 1.0))
        return fail(PCC_EINVAL, "bad variant constants");
    h->d.v.use_cwnd = v->use_cwnd ? 1 : 0; h->d.v.use_noise = v->use_latency_noise ? 1 : 0;
    h->d.v.max_noise = v->max_latency_noise; h->d.v.initial_cwnd = v->initial_cwnd;
    h->d.v.min_cwnd = v->min_cwnd; h->d.v.max_cwnd = v->max_cwnd;
    return PCC_OK;
}

int64_t pcc_multi_launch_count(pcc_multi_handle h) { return h ? h->launches : 0; }

int pcc_multi_check(pcc_multi_handle h, void *stream)
{
    if (!h) return fail(PCC_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned long long meta[2];
    CUDA_TRY(cudaMemcpy(meta, h->d.meta, sizeof(meta), cudaMemcpyDeviceToHost));
    if (meta[1]) return fail(PCC_EOVERFLOW, "event heap / sample buffer overflow: ring_capacity too small for these link parameters");
    return PCC_OK;
}

}  // extern "C"

#include "pcc_flows.cuh"
