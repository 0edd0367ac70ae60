/*
 * pcc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU, binary64 restatement of the PCC-RL gym hot path, written so that it can
 * be read side by side with the reference.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may link or call this file.  The product
 * (pcc-rl_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs this file against the
 * unmodified reference imported from /root/reference (thousands of env-steps, bit-exact on
 * every output), and tests/test_oracle_golden.py checks it against the .npz files in tests/golden,
 * which oracle/gen_golden.py produced by executing the reference itself.
 *
 * What is restated (reference file:line, all under /root/reference/src):
 *   Link                      gym/network_sim.py:56-96
 *   Network.run_for_dur       gym/network_sim.py:123-205   (heapq event loop + reward)
 *   Sender                    gym/network_sim.py:207-342
 *   SimulatedNetworkEnv glue  gym/network_sim.py:406-484   (step, reset, warm-up MIs)
 *   MI metrics / history      common/sender_obs.py:20-206
 *   DELTA_SCALE               common/config.py:17
 * Third-party arithmetic restated from its published algorithm:
 *   CPython `random` (MT19937, init_by_array seeding, genrand_res53)  -- Python 3.12
 *   numpy `np.mean` over a float64 list (pairwise summation, PW_BLOCKSIZE 128) -- numpy 2.3.5
 *   Python `heapq` ordering = lexicographic tuple order
 *       (time, sender, type 'A'<'S', next_hop, cur_latency, dropped False<True)
 * The event queue here is a textbook binary heap on that tuple order; identical tuples are
 * interchangeable, so any exact priority queue gives the reference's outputs.
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).  -ffp-contract=off
 * matters: the reference rounds every operation separately.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PCCO_MAX_SENDERS 8
#define PCCO_N_METRICS 12
#define PCCO_MAX_HISTORY 64
#define PCCO_MAX_FEATURES 12

/* network_sim.py:33-54 */
#define MAX_RATE 1000.0
#define MIN_RATE 40.0
#define REWARD_SCALE 0.001
#define MAX_STEPS 400
#define BYTES_PER_PACKET 1500
/* common/config.py:17 */
#define DELTA_SCALE 0.025
/* the compiled-out variants of the same loop (network_sim.py:33-34, 51-54): congestion window and latency noise */
#define MAX_CWND 5000
#define MIN_CWND 4
#define MAX_LATENCY_NOISE 1.1
#define INITIAL_CWND 25      /* Sender.__init__ default, network_sim.py:209 */

/* ------------------------------------------------------------------------------------ */
/* RNG streams                                                                          */
/* ------------------------------------------------------------------------------------ */
enum { PCCO_RNG_MT19937 = 0, PCCO_RNG_PHILOX = 1 };

typedef struct {
    int kind;
    /* MT19937 (CPython Modules/_randommodule.c) */
    uint32_t mt[624];
    int mti;
    /* Philox4x32-10, counter = draw index, key = 64-bit seed */
    uint64_t seed;
    uint64_t draws;
} pcco_rng;

static void mt_init_genrand(pcco_rng *r, uint32_t s)
{
    r->mt[0] = s;
    for (int i = 1; i < 624; i++)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->mti = 624;
}

static void mt_init_by_array(pcco_rng *r, const uint32_t *key, int len)
{
    mt_init_genrand(r, 19650218u);
    int i = 1, j = 0;
    int k = 624 > len ? 624 : len;
    for (; k; k--) {
        r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; k--) {
        r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
    }
    r->mt[0] = 0x80000000u;
}

static uint32_t mt_genrand(pcco_rng *r)
{
    if (r->mti >= 624) {
        uint32_t *mt = r->mt;
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
        r->mti = 0;
    }
    uint32_t y = r->mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* Philox4x32-10 (Salmon et al., SC'11), restated from the paper's round function. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

#define PCCO_PHILOX_DOMAIN 0x50434352u /* 'PCCR' */

/* Uniform double in [0,1) with 53 bits, CPython's genrand_res53 construction for both kinds:
 * ((a>>5)*2^26 + (b>>6)) / 2^53 from two successive 32-bit words. */
static double rng_random(pcco_rng *r)
{
    uint32_t a, b;
    if (r->kind == PCCO_RNG_MT19937) {
        a = mt_genrand(r);
        b = mt_genrand(r);
    } else {
        uint64_t blk = r->draws >> 1;
        uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), PCCO_PHILOX_DOMAIN, 0u};
        philox4x32_10(c, (uint32_t)r->seed, (uint32_t)(r->seed >> 32));
        if (r->draws & 1u) { a = c[2]; b = c[3]; } else { a = c[0]; b = c[1]; }
        r->draws++;
    }
    a >>= 5; b >>= 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------------------ */
/* numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum)*/
/* ------------------------------------------------------------------------------------ */
static double np_pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* np.mean(list_of_floats): add.reduce with the identity 0.0 as the initial value, then a
 * true division by the count (numpy/_core/_methods.py:_mean). */
static double np_mean(const double *a, long n)
{
    double s = 0.0;
    s += np_pairwise_sum(a, n);
    return s / (double)n;
}

/* exported for direct unit tests against numpy */
double pcco_np_mean(const double *a, long n) { return np_mean(a, n); }

/* ------------------------------------------------------------------------------------ */
/* Link  (network_sim.py:56-96)                                                         */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double bw, dl, lr;
    double queue_delay, queue_delay_update_time, max_queue_delay;
} pcco_link;

static void link_init(pcco_link *l, double bandwidth, double delay, long queue_size, double loss)
{
    l->bw = bandwidth;                          /* :59 */
    l->dl = delay;
    l->lr = loss;
    l->queue_delay = 0.0;
    l->queue_delay_update_time = 0.0;
    l->max_queue_delay = (double)queue_size / l->bw; /* :64 */
}

static double py_max0(double x) { return (x > 0.0) ? x : 0.0; } /* max(0.0, x) */

static double link_cur_queue_delay(const pcco_link *l, double t)
{
    return py_max0(l->queue_delay - (t - l->queue_delay_update_time)); /* :66-67 */
}

static double link_cur_latency(const pcco_link *l, double t)
{
    return l->dl + link_cur_queue_delay(l, t); /* :69-70 */
}

static int link_packet_enters(pcco_link *l, double t, pcco_rng *rng)
{
    if (rng_random(rng) < l->lr) return 0;                 /* :73-74 */
    l->queue_delay = link_cur_queue_delay(l, t);           /* :75 */
    l->queue_delay_update_time = t;                        /* :76 */
    double extra_delay = 1.0 / l->bw;                      /* :77 */
    if (extra_delay + l->queue_delay > l->max_queue_delay) /* :79 */
        return 0;
    l->queue_delay += extra_delay;                         /* :82 */
    return 1;
}

/* ------------------------------------------------------------------------------------ */
/* Sender, MI, history  (network_sim.py:207-342, sender_obs.py)                          */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double v[PCCO_MAX_FEATURES]; /* feature values already divided by scale (memoised) */
} pcco_mi_row;

typedef struct {
    double rate, starting_rate;
    long sent, acked, lost;
    double obs_start_time;
    double *rtt; long n_rtt, cap_rtt;
    /* _conn_min_latencies[sender_id] (sender_obs.py:158): has_min == key present */
    int has_min; double conn_min;
    pcco_mi_row hist[PCCO_MAX_HISTORY];
    long cwnd;               /* :225, int after set_cwnd (:283-289) */
    long bytes_in_flight;    /* :215, +-BYTES_PER_PACKET on sent / acked / lost (:260-273) */
} pcco_sender;

/* metric ids = position in SENDER_MI_METRICS (sender_obs.py:193-206) */
enum {
    M_SEND_RATE = 0, M_RECV_RATE, M_RECV_DUR, M_SEND_DUR, M_AVG_LATENCY, M_LOSS_RATIO,
    M_ACK_LAT_INFL, M_SENT_LAT_INFL, M_CONN_MIN_LAT, M_LAT_INCREASE, M_LAT_RATIO, M_SEND_RATIO
};
static const double metric_scale[PCCO_N_METRICS] = {1e7, 1e7, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
/* value of each metric on an empty MI (SenderMonitorInterval(sender_id) with no samples),
 * i.e. the initial history rows: rates 0, durs 0, lat 0, loss 0, inflations 0, conn min 0,
 * increase 0, latency ratio 1, send ratio 1. */
static const double metric_empty[PCCO_N_METRICS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 1.0};

typedef struct {
    long bytes_sent, bytes_acked, bytes_lost;
    double start, end;
    const double *rtt; long n;
} pcco_mi;

static double mi_dur(const pcco_mi *m) { return m->end - m->start; }          /* :116-117,130-131 */
static double mi_send_rate(const pcco_mi *m)                                    /* :124-128 */
{
    double dur = mi_dur(m);
    if (dur > 0.0) return 8.0 * (double)m->bytes_sent / dur;
    return 0.0;
}
static double mi_recv_rate(const pcco_mi *m)                                    /* :110-114 */
{
    double dur = mi_dur(m);
    if (dur > 0.0) return 8.0 * (double)(m->bytes_acked - BYTES_PER_PACKET) / dur;
    return 0.0;
}
static double mi_avg_latency(const pcco_mi *m)                                  /* :119-122 */
{
    if (m->n > 0) return np_mean(m->rtt, m->n);
    return 0.0;
}
static double mi_loss_ratio(const pcco_mi *m)                                   /* :133-136 */
{
    if (m->bytes_lost + m->bytes_acked > 0)
        return (double)m->bytes_lost / (double)(m->bytes_lost + m->bytes_acked);
    return 0.0;
}
static double mi_latency_increase(const pcco_mi *m)                             /* :138-142 */
{
    long half = m->n / 2;
    if (half >= 1) return np_mean(m->rtt + half, m->n - half) - np_mean(m->rtt, half);
    return 0.0;
}
static double mi_latency_inflation(const pcco_mi *m)                            /* :144-156 */
{
    double dur = mi_dur(m);
    double inc = mi_latency_increase(m);
    if (dur > 0.0) return inc / dur;
    return 0.0;
}
static double mi_conn_min_latency(const pcco_mi *m, pcco_sender *s)             /* :158-176 */
{
    double latency = mi_avg_latency(m);
    if (s->has_min) {
        double prev = s->conn_min;
        if (latency == 0.0) return prev;
        if (latency < prev) { s->conn_min = latency; return latency; }
        return prev;
    }
    if (latency > 0.0) { s->has_min = 1; s->conn_min = latency; return latency; }
    return 0.0;
}
static double mi_send_ratio(const pcco_mi *m)                                   /* :179-184 */
{
    double thpt = mi_recv_rate(m), send_rate = mi_send_rate(m);
    if (thpt > 0.0 && send_rate < 1000.0 * thpt) return send_rate / thpt;
    return 1.0;
}
static double mi_latency_ratio(const pcco_mi *m, pcco_sender *s)                /* :186-191 */
{
    double min_lat = mi_conn_min_latency(m, s);
    double cur_lat = mi_avg_latency(m);
    if (min_lat > 0.0) return cur_lat / min_lat;
    return 1.0;
}
static double mi_metric(const pcco_mi *m, pcco_sender *s, int id)
{
    switch (id) {
    case M_SEND_RATE: return mi_send_rate(m);
    case M_RECV_RATE: return mi_recv_rate(m);
    case M_RECV_DUR: case M_SEND_DUR: return mi_dur(m);
    case M_AVG_LATENCY: return mi_avg_latency(m);
    case M_LOSS_RATIO: return mi_loss_ratio(m);
    case M_ACK_LAT_INFL: case M_SENT_LAT_INFL: return mi_latency_inflation(m);
    case M_CONN_MIN_LAT: return mi_conn_min_latency(m, s);
    case M_LAT_INCREASE: return mi_latency_increase(m);
    case M_LAT_RATIO: return mi_latency_ratio(m, s);
    case M_SEND_RATIO: return mi_send_ratio(m);
    }
    return 0.0;
}

/* ------------------------------------------------------------------------------------ */
/* Event heap  (heapq on tuples, network_sim.py:111,129,161,178)                        */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    double time;
    int sender;       /* index; reference compares Sender objects only on full ties of time */
    int type;         /* 0 = 'A' (ACK), 1 = 'S' (SEND): 'A' < 'S' */
    int next_hop;
    double cur_latency;
    int dropped;      /* False < True */
} pcco_event;

static int ev_less(const pcco_event *a, const pcco_event *b)
{
    if (a->time != b->time) return a->time < b->time;
    if (a->sender != b->sender) return a->sender < b->sender;
    if (a->type != b->type) return a->type < b->type;
    if (a->next_hop != b->next_hop) return a->next_hop < b->next_hop;
    if (a->cur_latency != b->cur_latency) return a->cur_latency < b->cur_latency;
    return a->dropped < b->dropped;
}

typedef struct { pcco_event *e; long n, cap; } pcco_heap;

static void heap_push(pcco_heap *h, pcco_event ev)
{
    if (h->n == h->cap) {
        h->cap = h->cap ? 2 * h->cap : 256;
        h->e = (pcco_event *)realloc(h->e, (size_t)h->cap * sizeof(pcco_event));
    }
    long i = h->n++;
    while (i > 0) {
        long p = (i - 1) / 2;
        if (!ev_less(&ev, &h->e[p])) break;
        h->e[i] = h->e[p];
        i = p;
    }
    h->e[i] = ev;
}

static pcco_event heap_pop(pcco_heap *h)
{
    pcco_event top = h->e[0];
    pcco_event last = h->e[--h->n];
    long i = 0;
    for (;;) {
        long c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && ev_less(&h->e[c + 1], &h->e[c])) c++;
        if (!ev_less(&h->e[c], &last)) break;
        h->e[i] = h->e[c];
        i = c;
    }
    if (h->n > 0) h->e[i] = last;
    return top;
}

/* ------------------------------------------------------------------------------------ */
/* Env                                                                                  */
/* ------------------------------------------------------------------------------------ */
typedef struct pcco_env {
    int history_len, n_features;
    int feature_ids[PCCO_MAX_FEATURES];
    int n_senders;
    pcco_rng rng;
    pcco_link links[2];
    pcco_sender senders[PCCO_MAX_SENDERS];
    pcco_heap q;
    double cur_time;
    double run_dur;
    long steps_taken;
    long max_steps;
    long long total_events; /* heap pops, for events/s reporting */
    int use_cwnd;           /* USE_CWND (:54); the reference ships False */
    int use_latency_noise;  /* USE_LATENCY_NOISE (:51); the reference ships False */
} pcco_env;

pcco_env *pcco_create(int history_len, const int *feature_ids, int n_features)
{
    if (history_len < 1 || history_len > PCCO_MAX_HISTORY) return NULL;
    if (n_features < 1 || n_features > PCCO_MAX_FEATURES) return NULL;
    pcco_env *e = (pcco_env *)calloc(1, sizeof(pcco_env));
    e->history_len = history_len;
    e->n_features = n_features;
    for (int i = 0; i < n_features; i++) {
        if (feature_ids[i] < 0 || feature_ids[i] >= PCCO_N_METRICS) { free(e); return NULL; }
        e->feature_ids[i] = feature_ids[i];
    }
    e->n_senders = 1;
    e->max_steps = MAX_STEPS;
    e->rng.kind = PCCO_RNG_MT19937;
    mt_init_genrand(&e->rng, 5489u);
    return e;
}

void pcco_destroy(pcco_env *e)
{
    if (!e) return;
    for (int i = 0; i < PCCO_MAX_SENDERS; i++) free(e->senders[i].rtt);
    free(e->q.e);
    free(e);
}

void pcco_set_max_steps(pcco_env *e, long n) { e->max_steps = n; }

/* The module-level switches USE_CWND / USE_LATENCY_NOISE of network_sim.py:51-54 (both False as shipped). */
void pcco_set_variant(pcco_env *e, int use_cwnd, int use_latency_noise)
{
    e->use_cwnd = use_cwnd;
    e->use_latency_noise = use_latency_noise;
}
long pcco_cwnd(const pcco_env *e, int sender) { return e->senders[sender].cwnd; }

/* random.seed(int) of CPython: init_by_array over the 32-bit little-endian limbs of |seed| */
void pcco_seed_mt(pcco_env *e, uint64_t seed)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    e->rng.kind = PCCO_RNG_MT19937;
    mt_init_by_array(&e->rng, key, key[1] ? 2 : 1);
}

void pcco_seed_philox(pcco_env *e, uint64_t seed)
{
    e->rng.kind = PCCO_RNG_PHILOX;
    e->rng.seed = seed;
    e->rng.draws = 0;
}

/* state[0..623] = mt words, state[624] = index (random.getstate()[1] layout) */
void pcco_mt_getstate(const pcco_env *e, uint32_t *state)
{
    memcpy(state, e->rng.mt, 624 * sizeof(uint32_t));
    state[624] = (uint32_t)e->rng.mti;
}
void pcco_mt_setstate(pcco_env *e, const uint32_t *state)
{
    e->rng.kind = PCCO_RNG_MT19937;
    memcpy(e->rng.mt, state, 624 * sizeof(uint32_t));
    e->rng.mti = (int)state[624];
}
uint64_t pcco_philox_draws(const pcco_env *e) { return e->rng.draws; }

/* One draw from the env's stream: lets a host-side harness sample link parameters from the
 * same stream the per-packet loss draws use, as the reference's global `random` does. */
double pcco_random(pcco_env *e) { return rng_random(&e->rng); }

static void sender_reset_obs(pcco_env *e, pcco_sender *s) /* network_sim.py:319-324 */
{
    s->sent = 0; s->acked = 0; s->lost = 0;
    s->n_rtt = 0;
    s->obs_start_time = e->cur_time;
}

static void sender_append_rtt(pcco_sender *s, double rtt)
{
    if (s->n_rtt == s->cap_rtt) {
        s->cap_rtt = s->cap_rtt ? 2 * s->cap_rtt : 256;
        s->rtt = (double *)realloc(s->rtt, (size_t)s->cap_rtt * sizeof(double));
    }
    s->rtt[s->n_rtt++] = rtt;
}

static void sender_set_rate(pcco_sender *s, double new_rate) /* :275-281 */
{
    s->rate = new_rate;
    if (s->rate > MAX_RATE) s->rate = MAX_RATE;
    if (s->rate < MIN_RATE) s->rate = MIN_RATE;
}

static void sender_apply_rate_delta(pcco_sender *s, double delta) /* :235-241 */
{
    delta *= DELTA_SCALE;
    if (delta >= 0.0) sender_set_rate(s, s->rate * (1.0 + delta));
    else sender_set_rate(s, s->rate / (1.0 - delta));
}

static void sender_set_cwnd(pcco_sender *s, double new_cwnd) /* :283-289 */
{
    s->cwnd = (long)new_cwnd;                /* int(): truncation toward zero */
    if (s->cwnd > MAX_CWND) s->cwnd = MAX_CWND;
    if (s->cwnd < MIN_CWND) s->cwnd = MIN_CWND;
}

static void sender_apply_cwnd_delta(pcco_sender *s, double delta) /* :243-249 */
{
    delta *= DELTA_SCALE;
    if (delta >= 0.0) sender_set_cwnd(s, (double)s->cwnd * (1.0 + delta));
    else sender_set_cwnd(s, (double)s->cwnd / (1.0 - delta));
}

static int sender_can_send_packet(const pcco_env *e, const pcco_sender *s) /* :251-255 */
{
    if (e->use_cwnd) return (double)s->bytes_in_flight / (double)BYTES_PER_PACKET < (double)s->cwnd;
    return 1;
}

/* random.uniform(1.0, MAX_LATENCY_NOISE) = a + (b - a) * random() (CPython Lib/random.py) */
static double latency_noise(pcco_env *e)
{
    return 1.0 + (MAX_LATENCY_NOISE - 1.0) * rng_random(&e->rng);
}

static pcco_mi sender_get_run_data(const pcco_env *e, const pcco_sender *s) /* :298-317 */
{
    pcco_mi m;
    m.bytes_sent = s->sent * BYTES_PER_PACKET;
    m.bytes_acked = s->acked * BYTES_PER_PACKET;
    m.bytes_lost = s->lost * BYTES_PER_PACKET;
    m.start = s->obs_start_time;
    m.end = e->cur_time;
    m.rtt = s->rtt;
    m.n = s->n_rtt;
    return m;
}

/* Network.run_for_dur, network_sim.py:123-205.  Returns the reward of senders[0]. */
static double run_for_dur(pcco_env *e, double dur)
{
    double end_time = e->cur_time + dur;                                   /* :124 */
    for (int i = 0; i < e->n_senders; i++) sender_reset_obs(e, &e->senders[i]); /* :125-126 */

    while (e->cur_time < end_time) {                                       /* :128 */
        pcco_event ev = heap_pop(&e->q);                                   /* :129 */
        e->total_events++;
        pcco_sender *sender = &e->senders[ev.sender];
        e->cur_time = ev.time;                                             /* :131 */
        pcco_event nw = ev;                                                /* :132-136 */
        int push_new_event = 0;

        if (ev.type == 0) { /* ACK :139 */
            if (ev.next_hop == 2) {                 /* len(sender.path) == 2 :140 */
                if (ev.dropped) sender->lost += 1;  /* :141-142, :271-273 */
                else {                              /* :144-145, :264-269 */
                    sender->acked += 1;
                    sender_append_rtt(sender, ev.cur_latency);
                }
                sender->bytes_in_flight -= BYTES_PER_PACKET;   /* :269, :273 */
            } else {                                /* :147-154 */
                nw.next_hop = ev.next_hop + 1;
                double link_latency = link_cur_latency(&e->links[ev.next_hop], e->cur_time);
                if (e->use_latency_noise) link_latency *= latency_noise(e);   /* :150-151 */
                nw.cur_latency += link_latency;
                nw.time += link_latency;
                push_new_event = 1;
            }
        }
        if (ev.type == 1) { /* SEND :155 */
            if (ev.next_hop == 0) {                 /* :156 */
                if (sender_can_send_packet(e, sender)) {   /* :158; always True as shipped */
                    sender->sent += 1;              /* :159-160, :260-262 */
                    sender->bytes_in_flight += BYTES_PER_PACKET;
                    push_new_event = 1;
                }
                pcco_event timer = {e->cur_time + (1.0 / sender->rate), ev.sender, 1, 0, 0.0, 0};
                heap_push(&e->q, timer);            /* :161 */
            } else {
                push_new_event = 1;                 /* :163-164 (unreachable with dest == 0) */
            }
            if (ev.next_hop == 0) nw.type = 0;      /* next_hop == sender.dest (== 0) :166-167 */
            nw.next_hop = ev.next_hop + 1;          /* :168 */
            double link_latency = link_cur_latency(&e->links[ev.next_hop], e->cur_time); /* :170 */
            if (e->use_latency_noise) link_latency *= latency_noise(e);                  /* :171-172 */
            nw.cur_latency += link_latency;         /* :173 */
            nw.time += link_latency;                /* :174 */
            nw.dropped = !link_packet_enters(&e->links[ev.next_hop], e->cur_time, &e->rng); /* :175 */
        }
        if (push_new_event) heap_push(&e->q, nw);   /* :177-178 */
    }

    pcco_mi m = sender_get_run_data(e, &e->senders[0]);                    /* :180 */
    double throughput = mi_recv_rate(&m);                                  /* :181 */
    double latency = mi_avg_latency(&m);                                   /* :182 */
    double loss = mi_loss_ratio(&m);                                       /* :183 */
    double reward = (10.0 * throughput / (8 * BYTES_PER_PACKET) - 1e3 * latency - 2e3 * loss); /* :194 */
    return reward * REWARD_SCALE;                                          /* :205 */
}

/* reset(): network_sim.py:469-484 with the five parameter draws (:455-466) done by the
 * caller (they are inputs).  queue_size = 1 + int(np.exp(u)) already applied. */
void pcco_reset(pcco_env *e, double bw, double lat, long queue_size, double loss, double start_rate)
{
    link_init(&e->links[0], bw, lat, queue_size, loss);   /* :463 two identical links */
    link_init(&e->links[1], bw, lat, queue_size, loss);
    e->n_senders = 1;
    pcco_sender *s = &e->senders[0];                      /* :466 */
    s->rate = start_rate; s->starting_rate = start_rate;
    s->cwnd = INITIAL_CWND; s->bytes_in_flight = 0;       /* a fresh Sender (:209-226) */
    s->has_min = 0; s->conn_min = 0.0;                    /* fresh sender id => no dict entry */
    for (int h = 0; h < e->history_len; h++)              /* SenderHistory.__init__ sender_obs.py:57-62 */
        for (int f = 0; f < e->n_features; f++) {
            int id = e->feature_ids[f];
            s->hist[h].v[f] = metric_empty[id] / metric_scale[id];
        }
    e->run_dur = 3 * lat;                                 /* :467 */
    e->q.n = 0;                                           /* Network.__init__ :100-105 */
    e->cur_time = 0.0;
    sender_reset_obs(e, s);                               /* queue_initial_packets :107-111 */
    pcco_event first = {1.0 / s->rate, 0, 1, 0, 0.0, 0};
    heap_push(&e->q, first);
    e->steps_taken = 0;                                   /* :470 */
    run_for_dur(e, e->run_dur);                           /* :478 */
    run_for_dur(e, e->run_dur);                           /* :479 */
}

void pcco_get_obs(const pcco_env *e, double *obs) /* _get_all_sender_obs :400-404 */
{
    const pcco_sender *s = &e->senders[0];
    for (int h = 0; h < e->history_len; h++)
        for (int f = 0; f < e->n_features; f++)
            obs[h * e->n_features + f] = s->hist[h].v[f];
}

/* step(): network_sim.py:406-444.
 * counts[3] = sent, acked, lost of this MI; info[8] = reward's three inputs and the event-log
 * fields: send rate, throughput (recv rate), avg latency, loss ratio, latency inflation,
 * latency ratio, send ratio, MI duration. */
static void step_impl(pcco_env *e, double action, double cwnd_action, double *obs, double *reward, int *done,
                      long *counts, double *info);

void pcco_step(pcco_env *e, double action, double *obs, double *reward, int *done,
               long *counts, double *info)
{
    step_impl(e, action, 0.0, obs, reward, done, counts, info);
}

/* step([rate_action, cwnd_action]) of the USE_CWND variant (:409-414) */
void pcco_step_cwnd(pcco_env *e, double action, double cwnd_action, double *obs, double *reward, int *done,
                    long *counts, double *info)
{
    step_impl(e, action, cwnd_action, obs, reward, done, counts, info);
}

static void step_impl(pcco_env *e, double action, double cwnd_action, double *obs, double *reward, int *done,
                      long *counts, double *info)
{
    pcco_sender *s = &e->senders[0];
    sender_apply_rate_delta(s, action);                   /* :412 */
    if (e->use_cwnd) sender_apply_cwnd_delta(s, cwnd_action);   /* :413-414 */
    double r = run_for_dur(e, e->run_dur);                /* :416 */
    /* record_run (:418 -> :291-293): features are memoised per MI object the first time the
     * history is turned into an array (sender_obs.py:44-54, 68-73). */
    pcco_mi m = sender_get_run_data(e, s);
    for (int h = 0; h + 1 < e->history_len; h++) s->hist[h] = s->hist[h + 1]; /* SenderHistory.step */
    pcco_mi_row *row = &s->hist[e->history_len - 1];
    for (int f = 0; f < e->n_features; f++) {
        int id = e->feature_ids[f];
        row->v[f] = mi_metric(&m, s, id) / metric_scale[id];
    }
    e->steps_taken += 1;                                  /* :419 */
    if (obs) pcco_get_obs(e, obs);                        /* :420 */
    /* event record :421-436 (a second MI object; evaluating "latency ratio" on it touches
     * _conn_min_latencies again with the same value, which is idempotent) */
    double avg_latency = mi_avg_latency(&m);
    double lat_ratio = mi_latency_ratio(&m, s);
    if (info) {
        info[0] = mi_send_rate(&m);
        info[1] = mi_recv_rate(&m);
        info[2] = avg_latency;
        info[3] = mi_loss_ratio(&m);
        info[4] = mi_latency_inflation(&m);
        info[5] = lat_ratio;
        info[6] = mi_send_ratio(&m);
        info[7] = mi_dur(&m);
    }
    if (avg_latency > 0.0) e->run_dur = 0.5 * avg_latency; /* :437-438 */
    if (counts) { counts[0] = s->sent; counts[1] = s->acked; counts[2] = s->lost; }
    *reward = r;
    *done = (e->steps_taken >= e->max_steps);             /* :444 */
}

double pcco_cur_time(const pcco_env *e) { return e->cur_time; }
double pcco_run_dur(const pcco_env *e) { return e->run_dur; }
double pcco_rate(const pcco_env *e) { return e->senders[0].rate; }
long pcco_queue_len(const pcco_env *e) { return e->q.n; }
long long pcco_total_events(const pcco_env *e) { return e->total_events; }

/* ------------------------------------------------------------------------------------ */
/* Several senders on the same two links (BASELINE config 5).              */
/* The reference's Network handles lists of senders (network_sim.py:100-126, 140-178);   */
/* its env only ever creates one.  Semantics fixed here (and pinned against the reference */
/* classes with the external patch Sender.__lt__ = id order, tests/test_oracle_multi.py): */
/*   - all senders share links [l0, l1] (one shared bottleneck queue), dest = 0;          */
/*   - heap ties are broken by sender index (position 2 of the event tuple);              */
/*   - step(actions[S]) applies actions[i] to sender i (network_sim.py:409-412 generalised),*/
/*     reward i uses sender i's own MI with the formula of :194,205, every sender records */
/*     its MI, run_dur follows sender 0's average latency (:437-438 as written);          */
/*   - one uniform draw per sent packet from the env's single stream, in event order.     */
/* ------------------------------------------------------------------------------------ */
static double reward_of(const pcco_mi *m)
{
    double throughput = mi_recv_rate(m), latency = mi_avg_latency(m), loss = mi_loss_ratio(m);
    double reward = (10.0 * throughput / (8 * BYTES_PER_PACKET) - 1e3 * latency - 2e3 * loss);
    return reward * REWARD_SCALE;
}

void pcco_reset_multi(pcco_env *e, int n_senders, double bw, double lat, long queue_size, double loss,
                      const double *start_rates)
{
    link_init(&e->links[0], bw, lat, queue_size, loss);
    link_init(&e->links[1], bw, lat, queue_size, loss);
    e->n_senders = n_senders;
    e->run_dur = 3 * lat;
    e->q.n = 0;
    e->cur_time = 0.0;
    for (int i = 0; i < n_senders; i++) {
        pcco_sender *s = &e->senders[i];
        s->rate = start_rates[i]; s->starting_rate = start_rates[i];
        s->cwnd = INITIAL_CWND; s->bytes_in_flight = 0;
        s->has_min = 0; s->conn_min = 0.0;
        for (int h = 0; h < e->history_len; h++)
            for (int f = 0; f < e->n_features; f++) {
                int id = e->feature_ids[f];
                s->hist[h].v[f] = metric_empty[id] / metric_scale[id];
            }
        sender_reset_obs(e, s);                               /* queue_initial_packets :107-111 */
        pcco_event first = {1.0 / s->rate, i, 1, 0, 0.0, 0};
        heap_push(&e->q, first);
    }
    e->steps_taken = 0;
    run_for_dur(e, e->run_dur);
    run_for_dur(e, e->run_dur);
}

/* obs [S][H*F], rewards [S], counts [S][3] */
static void step_multi_impl(pcco_env *e, const double *actions, const double *cwnd_actions, double *obs, double *rewards,
                            int *done, long *counts);
void pcco_step_multi(pcco_env *e, const double *actions, double *obs, double *rewards, int *done, long *counts)
{
    step_multi_impl(e, actions, NULL, obs, rewards, done, counts);
}
void pcco_step_multi_cwnd(pcco_env *e, const double *actions, const double *cwnd_actions, double *obs, double *rewards,
                          int *done, long *counts)
{
    step_multi_impl(e, actions, cwnd_actions, obs, rewards, done, counts);
}
static void step_multi_impl(pcco_env *e, const double *actions, const double *cwnd_actions, double *obs, double *rewards,
                            int *done, long *counts)
{
    const int hf = e->history_len * e->n_features;
    for (int i = 0; i < e->n_senders; i++) {
        sender_apply_rate_delta(&e->senders[i], actions[i]);
        if (e->use_cwnd && cwnd_actions) sender_apply_cwnd_delta(&e->senders[i], cwnd_actions[i]);
    }
    run_for_dur(e, e->run_dur);
    double avg0 = 0.0;
    for (int i = 0; i < e->n_senders; i++) {
        pcco_sender *s = &e->senders[i];
        pcco_mi m = sender_get_run_data(e, s);
        rewards[i] = reward_of(&m);
        for (int h = 0; h + 1 < e->history_len; h++) s->hist[h] = s->hist[h + 1];
        pcco_mi_row *row = &s->hist[e->history_len - 1];
        for (int f = 0; f < e->n_features; f++) {
            int id = e->feature_ids[f];
            row->v[f] = mi_metric(&m, s, id) / metric_scale[id];
        }
        (void)mi_latency_ratio(&m, s);   /* event record of step(): touches the conn-min entry (idempotent) */
        for (int h = 0; h < e->history_len; h++)
            for (int f = 0; f < e->n_features; f++) obs[i * hf + h * e->n_features + f] = s->hist[h].v[f];
        if (counts) { counts[3 * i] = s->sent; counts[3 * i + 1] = s->acked; counts[3 * i + 2] = s->lost; }
        if (i == 0) avg0 = mi_avg_latency(&m);
    }
    e->steps_taken += 1;
    if (avg0 > 0.0) e->run_dur = 0.5 * avg0;
    *done = (e->steps_taken >= e->max_steps);
}
