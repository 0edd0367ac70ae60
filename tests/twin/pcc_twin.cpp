// pcc_twin.cpp -- TEST HARNESS: the product's per-env MI code (pcc-rl_b200/csrc/pcc_core.cuh)
// compiled for the HOST with g++, one env at a time, so the streaming three-cursor algorithm
// can be checked against the heap-based oracle in a container without a GPU.  It is not a
// CPU fallback: nothing in the product loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../pcc-rl_b200/csrc/pcc_core.cuh"

using namespace pcc;

struct HostRing {
    Rec *buf; uint32_t mask;
    uint32_t capacity() const { return mask + 1; }
    Rec load(uint32_t i) const { return buf[i & mask]; }
    void store(uint32_t i, Rec r) { buf[i & mask] = r; }
    void store_a(uint32_t i, double a) { buf[i & mask].a = a; }
    void prefetch(uint32_t) const {}
};

struct Twin {
    int H, F; int ids[PCC_MAX_FEATURES];
    bool need_inc;
    Consts c;
    EnvState s;
    std::vector<Rec> ringbuf;
    HostRing ring;
    int rng_kind;  // 0 mt, 1 philox
    uint32_t mt[625];
    PhiloxRng ph;
    std::vector<double> hist;  // [H][F], oldest first
    bool overflow;
    bool prescan;              // two-stage consumption (see run_mi)
};

extern "C" {

Twin *twin_create(int history_len, const int *feature_ids, int n_features, int ring_capacity)
{
    Twin *t = new Twin();
    t->H = history_len; t->F = n_features;
    for (int i = 0; i < n_features; i++) t->ids[i] = feature_ids[i];
    t->need_inc = features_need_increase(t->ids, t->F);
    t->c.max_rate = 1000.0; t->c.min_rate = 40.0; t->c.delta_scale = 0.025; t->c.reward_scale = 0.001;
    t->c.max_steps = 400; t->c.bytes_per_packet = 1500;
    t->ringbuf.resize(ring_capacity);
    t->ring.buf = t->ringbuf.data(); t->ring.mask = (uint32_t)ring_capacity - 1;
    t->rng_kind = 1; t->ph.init(0, 0);
    t->hist.assign((size_t)history_len * n_features, 0.0);
    t->overflow = false;
    t->prescan = false;
    memset(&t->s, 0, sizeof(t->s));
    return t;
}
void twin_destroy(Twin *t) { delete t; }
void twin_seed_philox(Twin *t, uint64_t seed) { t->rng_kind = 1; t->ph.init(seed, 0); }
void twin_mt_setstate(Twin *t, const uint32_t *st) { t->rng_kind = 0; memcpy(t->mt, st, sizeof(t->mt)); }
void twin_mt_getstate(Twin *t, uint32_t *st) { memcpy(st, t->mt, sizeof(t->mt)); }
void twin_set_max_steps(Twin *t, int n) { t->c.max_steps = n; }
void twin_set_prescan(Twin *t, int on) { t->prescan = on != 0; }
void twin_set_ring_cursor(Twin *t, uint32_t base) { t->s.tail = t->s.h1 = t->s.h2 = base; }

void twin_get_obs(Twin *t, double *obs) { memcpy(obs, t->hist.data(), sizeof(double) * t->hist.size()); }

void twin_reset(Twin *t, double bw, double dl, int64_t queue, double lr, double start_rate)
{
    bool ovf;
    if (t->rng_kind == 0) { Mt19937Rng r; r.init(t->mt); ovf = reset_env(t->s, t->ring, r, bw, dl, lr, queue, start_rate); }
    else ovf = reset_env(t->s, t->ring, t->ph, bw, dl, lr, queue, start_rate);
    t->overflow |= ovf;
    for (int h = 0; h < t->H; h++)
        for (int f = 0; f < t->F; f++) t->hist[(size_t)h * t->F + f] = metric_empty(t->ids[f]);
}

// counts[3], info[8] as in the oracle
void twin_step(Twin *t, double action, double *obs, double *reward, int *done, int64_t *counts, double *info)
{
    StepOut o;
    if (t->rng_kind == 0) { Mt19937Rng r; r.init(t->mt); step_env(t->s, t->ring, r, action, t->c, true, o, t->prescan); }
    else step_env(t->s, t->ring, t->ph, action, t->c, true, o, t->prescan);
    t->overflow |= o.mi.overflow;
    memmove(t->hist.data(), t->hist.data() + t->F, sizeof(double) * (size_t)(t->H - 1) * t->F);
    for (int f = 0; f < t->F; f++) t->hist[(size_t)(t->H - 1) * t->F + f] = metric_value(o.st, t->ids[f]);
    if (obs) twin_get_obs(t, obs);
    *reward = o.st.reward;
    *done = o.done;
    if (counts) { counts[0] = o.mi.sent; counts[1] = o.mi.acked; counts[2] = o.mi.lost; }
    if (info) {
        info[0] = o.st.send_rate; info[1] = o.st.recv_rate; info[2] = o.st.avg_lat; info[3] = o.st.loss_ratio;
        info[4] = o.st.lat_infl; info[5] = o.st.lat_ratio; info[6] = o.st.send_ratio; info[7] = o.st.dur;
    }
}
double twin_cur_time(Twin *t) { return t->s.cur_time; }
double twin_run_dur(Twin *t) { return t->s.run_dur; }
double twin_rate(Twin *t) { return t->s.rate; }
int twin_overflow(Twin *t) { return t->overflow; }
uint32_t twin_inflight(Twin *t) { return t->s.tail - t->s.h2; }
}

// PwStream (push-style pairwise sum) against pw_sum (pull-style, the one the oracle tests pin to numpy): returns the
// number of prefix lengths n in [1, n_max] whose total / first-half / second-half means differ in any bit.
struct ArrReader { const double *a; int i; double next() { return a[i++]; } };
struct ArrAcc { double r[8]; double get(int j) const { return r[j]; } void set(int j, double v) { r[j] = v; } };
extern "C" int twin_pw_stream_mismatches(const double *a, int n_max)
{
    int bad = 0;
    for (int n = 1; n <= n_max; n++) {
        const int half = n / 2;
        ArrReader r0{a, 0};
        const double m_all = np_mean_stream(r0, n);
        double m_lo = 0.0, m_hi = 0.0;
        if (half >= 1) { ArrReader r1{a, 0}; m_lo = np_mean_stream(r1, half); m_hi = np_mean_stream(r1, n - half); }
        PwStream<ArrAcc> t, h;
        t.begin(n); h.begin(half >= 1 ? half : 0);
        double s_lo = 0.0, s_hi = 0.0;
        for (int i = 0; i < n; i++) {
            t.push(a[i]);
            if (half >= 1) {
                if (i == half) { s_lo = h.mean(half); h.begin(n - half); }
                h.push(a[i]);
            }
        }
        if (half >= 1) s_hi = h.mean(n - half);
        if (!t.done || (half >= 1 && !h.done)) { bad++; continue; }
        const double tm = t.mean(n);
        if (memcmp(&m_all, &tm, 8) != 0) bad++;
        else if (half >= 1 && (memcmp(&m_lo, &s_lo, 8) != 0 || memcmp(&m_hi, &s_hi, 8) != 0)) bad++;
    }
    return bad;
}

// integer form of the loss draw against the double form, for loss rates `lr[i]` and word pairs (a[j], b[j])
extern "C" long twin_loss_threshold_mismatches(const double *lr, int n_lr, const uint32_t *a, const uint32_t *b, int n_w)
{
    long bad = 0;
    for (int i = 0; i < n_lr; i++) {
        const uint64_t thr = loss_threshold(lr[i]);
        for (int j = 0; j < n_w; j++)
            if ((res53(a[j], b[j]) < lr[i]) != (u53(a[j], b[j]) < thr)) bad++;
    }
    return bad;
}

extern "C" double twin_tail_drop_threshold(double d_bw, double max_qd) { return pcc::tail_drop_threshold(d_bw, max_qd); }

// ---- several senders on one bottleneck (pcc_multi_core.cuh), host build -------------------------------------
#include "../../pcc-rl_b200/csrc/pcc_multi_core.cuh"

struct HostHeap {
    std::vector<MEvent> *v;
    int capacity() const { return (int)v->size(); }
    MEvent get(int i) const { return (*v)[i]; }
    void set(int i, const MEvent &e) { (*v)[i] = e; }
};
struct TwinMulti {
    int S, H, F, cap_s; int ids[PCC_MAX_FEATURES]; bool need_inc;
    Consts c; Variant v; MNet net; MSender snd[PCC_MAX_SENDERS];
    std::vector<MEvent> heap; std::vector<double> samples, hist;   // hist [S][H][F], oldest first
    PhiloxRng ph; bool ok;
};
extern "C" {
TwinMulti *twin_multi_create(int n_senders, int history_len, const int *feature_ids, int n_features, int capacity)
{
    TwinMulti *t = new TwinMulti();
    t->S = n_senders; t->H = history_len; t->F = n_features; t->cap_s = capacity;
    for (int i = 0; i < n_features; i++) t->ids[i] = feature_ids[i];
    t->need_inc = features_need_increase(t->ids, t->F);
    t->c.max_rate = 1000.0; t->c.min_rate = 40.0; t->c.delta_scale = 0.025; t->c.reward_scale = 0.001;
    t->c.max_steps = 400; t->c.bytes_per_packet = 1500;
    t->heap.resize((size_t)capacity * n_senders + 8);
    t->samples.resize((size_t)capacity * n_senders);
    t->hist.assign((size_t)n_senders * history_len * n_features, 0.0);
    t->ph.init(0, 0); t->ok = true;
    t->v = default_variant();
    return t;
}
void twin_multi_destroy(TwinMulti *t) { delete t; }
void twin_multi_set_variant(TwinMulti *t, int use_cwnd, int use_noise) { t->v.use_cwnd = use_cwnd; t->v.use_noise = use_noise; }
int twin_multi_cwnd(TwinMulti *t, int i) { return t->snd[i].cwnd; }
void twin_multi_step_cwnd(TwinMulti *t, const double *actions, const double *cwnd_actions, double *obs, double *rewards,
                          int *done, int32_t *counts);
void twin_multi_seed(TwinMulti *t, uint64_t seed) { t->ph.init(seed, 0); }
void twin_multi_reset(TwinMulti *t, double bw, double dl, int64_t queue, double lr, const double *rates)
{
    HostHeap hp{&t->heap};
    t->ok = multi_reset(t->net, t->snd, t->S, hp, t->samples.data(), t->cap_s, t->ph, bw, dl, queue, lr, rates, t->v) && t->ok;
    for (int i = 0; i < t->S; i++)
        for (int h = 0; h < t->H; h++)
            for (int f = 0; f < t->F; f++) t->hist[((size_t)i * t->H + h) * t->F + f] = metric_empty(t->ids[f]);
}
void twin_multi_step(TwinMulti *t, const double *actions, double *obs, double *rewards, int *done, int32_t *counts)
{
    twin_multi_step_cwnd(t, actions, nullptr, obs, rewards, done, counts);
}
void twin_multi_step_cwnd(TwinMulti *t, const double *actions, const double *cwnd_actions, double *obs, double *rewards,
                          int *done, int32_t *counts)
{
    HostHeap hp{&t->heap};
    double rows[PCC_MAX_SENDERS * PCC_MAX_FEATURES];
    bool dn;
    t->ok = multi_step(t->net, t->snd, t->S, hp, t->samples.data(), t->cap_s, t->ph, actions, cwnd_actions, t->c, t->v,
                       t->ids, t->F, t->need_inc, rows, rewards, counts, dn) && t->ok;
    const size_t hf = (size_t)t->H * t->F;
    for (int i = 0; i < t->S; i++) {
        double *hs = t->hist.data() + i * hf;
        memmove(hs, hs + t->F, sizeof(double) * (hf - t->F));
        for (int f = 0; f < t->F; f++) hs[hf - t->F + f] = rows[i * t->F + f];
    }
    memcpy(obs, t->hist.data(), sizeof(double) * t->hist.size());
    *done = dn;
}
double twin_multi_cur_time(TwinMulti *t) { return t->net.cur_time; }
double twin_multi_run_dur(TwinMulti *t) { return t->net.run_dur; }
int twin_multi_ok(TwinMulti *t) { return t->ok; }
}

// ---- MI-sample ingestion (pcc_flows_core.cuh), host build: one flow, scalar ------------------------------------
#include "../../pcc-rl_b200/csrc/pcc_flows_core.cuh"

struct TwinFlow {
    int H, F; int ids[PCC_MAX_FEATURES]; bool touch;
    bool has_min; double conn_min;
    std::vector<double> hist;   // [H][F], oldest first
};
extern "C" {
TwinFlow *twin_flow_create(int history_len, const int *feature_ids, int n_features)
{
    TwinFlow *t = new TwinFlow();
    t->H = history_len; t->F = n_features;
    for (int i = 0; i < n_features; i++) t->ids[i] = feature_ids[i];
    t->touch = features_touch_conn_min(t->ids, t->F);
    t->has_min = false; t->conn_min = 0.0;
    t->hist.assign((size_t)history_len * n_features, 0.0);
    for (int h = 0; h < t->H; h++)
        for (int f = 0; f < t->F; f++) t->hist[h * t->F + f] = flow_metric_empty(t->ids[f], false, 0.0) / flow_metric_scale(t->ids[f]);
    return t;
}
void twin_flow_destroy(TwinFlow *t) { delete t; }
void twin_flow_reset(TwinFlow *t, int mode)
{
    if (mode == FLOW_RESET_NEW) { t->has_min = false; t->conn_min = 0.0; }
    const bool seen = (mode == FLOW_RESET_CLIENT) && t->has_min;
    for (int h = 0; h < t->H; h++)
        for (int f = 0; f < t->F; f++)
            t->hist[h * t->F + f] = flow_metric_empty(t->ids[f], seen, t->conn_min) / flow_metric_scale(t->ids[f]);
}
void twin_flow_give_sample(TwinFlow *t, int64_t bytes_sent, int64_t bytes_acked, int64_t bytes_lost, double send_start,
                           double send_end, double recv_start, double recv_end, const double *rtt, int64_t n,
                           int64_t packet_size, double *metrics12)
{
    FlowRecord r{bytes_sent, bytes_acked, bytes_lost, packet_size, send_start, send_end, recv_start, recv_end};
    FlowStats st;
    flow_record_stats(r, rtt, n, t->has_min, t->conn_min, t->touch, st);
    for (int h = 0; h + 1 < t->H; h++)
        for (int f = 0; f < t->F; f++) t->hist[h * t->F + f] = t->hist[(h + 1) * t->F + f];
    for (int f = 0; f < t->F; f++) t->hist[(t->H - 1) * t->F + f] = st.v[t->ids[f]] / flow_metric_scale(t->ids[f]);
    if (metrics12) memcpy(metrics12, st.v, sizeof(st.v));
}
void twin_flow_get_obs(TwinFlow *t, double *obs) { memcpy(obs, t->hist.data(), sizeof(double) * t->hist.size()); }
double twin_flow_apply_rate_delta(double rate, double action, double scale, double mn, double mx, int style)
{
    return flow_apply_rate_delta(rate, action, scale, mn, mx, style);
}
}

// ---- several senders without the heap (pcc_multi_fast.cuh), host build ------------------------------------------
#include "../../pcc-rl_b200/csrc/pcc_multi_fast.cuh"

struct HostSidRing {
    Rec *buf; uint8_t *sids; uint32_t mask;
    uint32_t capacity() const { return mask + 1; }
    Rec load(uint32_t i) const { return buf[i & mask]; }
    void store(uint32_t i, Rec r) { buf[i & mask] = r; }
    void store_a(uint32_t i, double a) { buf[i & mask].a = a; }
    int sid(uint32_t i) const { return sids[i & mask]; }
    void set_sid(uint32_t i, int s) { sids[i & mask] = (uint8_t)s; }
};
struct TwinMultiFast {
    int S, H, F, cap_s; int ids[PCC_MAX_FEATURES]; bool need_inc;
    Consts c; MNet net; MSender snd[PCC_MAX_SENDERS]; MFast f;
    std::vector<Rec> ring; std::vector<uint8_t> sids; std::vector<double> samples, hist;
    PhiloxRng ph; bool ok;
};
extern "C" {
TwinMultiFast *twin_mfast_create(int n_senders, int history_len, const int *feature_ids, int n_features, int capacity,
                                 uint32_t ring_base)
{
    TwinMultiFast *t = new TwinMultiFast();
    t->S = n_senders; t->H = history_len; t->F = n_features; t->cap_s = capacity;
    for (int i = 0; i < n_features; i++) t->ids[i] = feature_ids[i];
    t->need_inc = features_need_increase(t->ids, t->F);
    t->c.max_rate = 1000.0; t->c.min_rate = 40.0; t->c.delta_scale = 0.025; t->c.reward_scale = 0.001;
    t->c.max_steps = 400; t->c.bytes_per_packet = 1500;
    t->ring.resize((size_t)capacity); t->sids.resize((size_t)capacity);
    t->samples.resize((size_t)capacity * n_senders);
    t->hist.assign((size_t)n_senders * history_len * n_features, 0.0);
    t->f.tail = t->f.h1 = t->f.h2 = ring_base;                      // cursors may start anywhere (u32 wrap test)
    t->ph.init(0, 0); t->ok = true;
    return t;
}
void twin_mfast_destroy(TwinMultiFast *t) { delete t; }
void twin_mfast_seed(TwinMultiFast *t, uint64_t seed) { t->ph.init(seed, 0); }
void twin_mfast_reset(TwinMultiFast *t, double bw, double dl, int64_t queue, double lr, const double *rates)
{
    HostSidRing rg{t->ring.data(), t->sids.data(), (uint32_t)t->ring.size() - 1u};
    t->ok = mfast_reset(t->net, t->snd, t->S, t->f, rg, t->samples.data(), t->cap_s, t->ph, bw, dl, queue, lr, rates) && t->ok;
    for (int i = 0; i < t->S; i++)
        for (int h = 0; h < t->H; h++)
            for (int k = 0; k < t->F; k++) t->hist[((size_t)i * t->H + h) * t->F + k] = metric_empty(t->ids[k]);
}
void twin_mfast_step(TwinMultiFast *t, const double *actions, double *obs, double *rewards, int *done, int32_t *counts)
{
    HostSidRing rg{t->ring.data(), t->sids.data(), (uint32_t)t->ring.size() - 1u};
    double rows[PCC_MAX_SENDERS * PCC_MAX_FEATURES];
    bool dn;
    t->ok = mfast_step(t->net, t->snd, t->S, t->f, rg, t->samples.data(), t->cap_s, t->ph, actions, t->c, t->ids, t->F,
                       t->need_inc, rows, rewards, counts, dn) && t->ok;
    const size_t hf = (size_t)t->H * t->F;
    for (int i = 0; i < t->S; i++) {
        double *hs = t->hist.data() + i * hf;
        memmove(hs, hs + t->F, sizeof(double) * (hf - t->F));
        for (int k = 0; k < t->F; k++) hs[hf - t->F + k] = rows[i * t->F + k];
    }
    memcpy(obs, t->hist.data(), sizeof(double) * t->hist.size());
    *done = dn;
}
double twin_mfast_cur_time(TwinMultiFast *t) { return t->net.cur_time; }
double twin_mfast_run_dur(TwinMultiFast *t) { return t->net.run_dur; }
int twin_mfast_ok(TwinMultiFast *t) { return t->ok; }
}

// ---- executable specification of the flows kernels' lane-level summation (pcc_flows.cuh) --------------------------
// The CUDA code cannot run here (shuffles), so this is a lane-by-lane emulation of the SAME algorithm: 8 lanes =
// numpy's 8 accumulators, three xor-exchange rounds = its combination tree, sequential tail; records up to 248 samples
// by four flat leaf steps, longer ones by the 4-frame register stack walk (PCCF_SG_MAX_N = 1800).  tests/test_twin_flows.py
// checks it against numpy for every n; the GPU tests check the real kernels against the same oracle.
static double emu_sg_leaf(const double *a, int n)
{
    const int nb = (n >= 8) ? n - (n & 7) : 0;
    double r[8];
    for (int j = 0; j < 8; j++) {
        r[j] = 0.0;
        if (nb) r[j] = a[j];
        for (int k = 8; k < nb; k += 8) r[j] += a[k + j];
    }
    for (int x = 1; x <= 4; x <<= 1) {                 // r += shfl_xor(r, x), all lanes at once
        double t[8];
        for (int j = 0; j < 8; j++) t[j] = r[j] + r[j ^ x];
        for (int j = 0; j < 8; j++) r[j] = t[j];
    }
    double res = r[0];
    for (int k = nb; k < n; k++) res += a[k];          // the (up to 7) tail elements, in order
    return res;
}
static double emu_sg_pw_sum(const double *a, int n)
{
    if (n <= 128) return emu_sg_leaf(a, n);
    int rn[4] = {0, 0, 0, 0}; double ls[4] = {0, 0, 0, 0};
    unsigned have_left = 0u; int sp = 0, cur = n; const double *p = a;
    for (;;) {
        while (cur > 128) {
            int n2 = cur >> 1; n2 -= n2 & 7;
            if (sp >= 4) return 0.0 / 0.0;              // deeper than the register stack: not this path's job
            rn[sp] = cur - n2; have_left &= ~(1u << sp); sp++; cur = n2;
        }
        double res = emu_sg_leaf(p, cur); p += cur;
        for (;;) {
            if (sp == 0) return res;
            const int t = sp - 1;
            if (!((have_left >> t) & 1u)) { ls[t] = res; have_left |= 1u << t; cur = rn[t]; break; }
            res = ls[t] + res; sp--;
        }
    }
}
extern "C" void twin_flows_pass_sums(const double *a, int n, double *out3)
{
    const int half = n / 2;
    if (n <= 248) {                                     // PCCF_FLAT_MAX_N: four leaf steps
        int n2 = n >> 1; n2 -= n2 & 7;
        const int c0 = n > 128 ? n2 : n, c1 = n > 128 ? n - n2 : 0;
        const double l0 = emu_sg_leaf(a, c0), l1 = emu_sg_leaf(a + c0, c1);
        out3[0] = n > 128 ? l0 + l1 : l0;
        out3[1] = emu_sg_leaf(a, half);
        out3[2] = emu_sg_leaf(a + half, n - half);
    } else {
        out3[0] = emu_sg_pw_sum(a, n);
        out3[1] = emu_sg_pw_sum(a, half);
        out3[2] = emu_sg_pw_sum(a + half, n - half);
    }
}
