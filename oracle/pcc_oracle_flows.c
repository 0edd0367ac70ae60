/*
 * pcc_oracle_flows.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as pcc_oracle.c).
 *
 * CPU restatement of the reference's MI-sample ingestion path: monitor-interval records measured on
 * REAL flows (by the PCC sender in C++) are handed to Python, turned into SenderMonitorInterval
 * objects, pushed into a SenderHistory, and the history array is what the agent sees:
 *     give_sample(...)            udt-plugins/testing/loaded_client.py:111-138  (PccGymDriver)
 *     ShimNetworkEnv.step         gym/online/shim_env.py:102-139               (same record off a socket)
 *     SenderMonitorInterval       common/sender_obs.py:20-54
 *     SenderHistory               common/sender_obs.py:56-73
 *     the 12 metrics              common/sender_obs.py:110-191, table :193-206
 *     apply_rate_delta            loaded_client.py:147-168 / shim_env.py:80-95
 * Unlike the simulator's MIs, these records carry independent send and recv windows, arbitrary byte
 * counts and their own packet size, so the metrics are restated here in their general form.
 *
 * Memoisation.  The reference evaluates a metric the first time an MI's array is taken and caches it
 * (sender_obs.py:44-54); `conn min latency` reads and updates the module-global dict
 * _conn_min_latencies[sender_id] at that moment (:158-176).  This restatement evaluates every MI when
 * it is ingested, and the empty MIs of a fresh history when the history is created.  That is what the
 * reference computes as long as history.as_array() is taken at least once per history_len ingested
 * records (ShimNetworkEnv.step does it after every record; PccGymDriver.get_rate after every record
 * once the flow has data).  The dict is keyed by flow id and never cleared by the reference, so a
 * history reset keeps conn_min; see pcco_flows_reset for the three variants.
 *
 * Parity status: PINNED -- tests/test_oracle_flows.py drives the unmodified common/sender_obs.py and
 * loaded_client.py (with loaded_agent stubbed: TensorFlow is absent) next to this file, and
 * tests/golden/flows_*.npz hold outputs of those reference modules (oracle/gen_golden_flows.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

double pcco_np_mean(const double *a, long n);   /* pcc_oracle.c: numpy's pairwise np.mean */

#define PCCF_N_METRICS 12
#define PCCF_MAX_HISTORY 64
#define PCCF_MAX_FEATURES 12

enum {
    F_SEND_RATE = 0, F_RECV_RATE, F_RECV_DUR, F_SEND_DUR, F_AVG_LATENCY, F_LOSS_RATIO,
    F_ACK_LAT_INFL, F_SENT_LAT_INFL, F_CONN_MIN_LAT, F_LAT_INCREASE, F_LAT_RATIO, F_SEND_RATIO
};
static const double f_scale[PCCF_N_METRICS] = {1e7, 1e7, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};   /* :193-206 */

typedef struct {
    long long bytes_sent, bytes_acked, bytes_lost, packet_size;
    double send_start, send_end, recv_start, recv_end;
    const double *rtt; long n;
} pccf_mi;

typedef struct {
    int has_min; double conn_min;            /* _conn_min_latencies[flow id] */
    double rows[PCCF_MAX_HISTORY][PCCF_MAX_FEATURES];   /* oldest first, memoised values / scale */
    long long n_records;                     /* records since the last reset (got_data = n_records > 0) */
    double rate;
} pccf_flow;

typedef struct pccf_flows {
    long n_flows;
    int history_len, n_features;
    int feature_ids[PCCF_MAX_FEATURES];
    pccf_flow *f;
} pccf_flows;

/* ---- the metrics, general form ------------------------------------------------------------------- */
static double g_recv_dur(const pccf_mi *m) { return m->recv_end - m->recv_start; }     /* :116-117 */
static double g_send_dur(const pccf_mi *m) { return m->send_end - m->send_start; }     /* :130-131 */
static double g_recv_rate(const pccf_mi *m)                                             /* :110-114 */
{
    double dur = g_recv_dur(m);
    if (dur > 0.0) return 8.0 * (double)(m->bytes_acked - m->packet_size) / dur;
    return 0.0;
}
static double g_send_rate(const pccf_mi *m)                                             /* :124-128 */
{
    double dur = g_send_dur(m);
    if (dur > 0.0) return 8.0 * (double)m->bytes_sent / dur;
    return 0.0;
}
static double g_avg_latency(const pccf_mi *m)                                           /* :119-122 */
{
    if (m->n > 0) return pcco_np_mean(m->rtt, m->n);
    return 0.0;
}
static double g_loss_ratio(const pccf_mi *m)                                            /* :133-136 */
{
    if (m->bytes_lost + m->bytes_acked > 0)
        return (double)m->bytes_lost / (double)(m->bytes_lost + m->bytes_acked);
    return 0.0;
}
static double g_latency_increase(const pccf_mi *m)                                      /* :138-142 */
{
    long half = m->n / 2;
    if (half >= 1) return pcco_np_mean(m->rtt + half, m->n - half) - pcco_np_mean(m->rtt, half);
    return 0.0;
}
static double g_ack_latency_inflation(const pccf_mi *m)                                 /* :144-149 */
{
    double dur = g_recv_dur(m);
    double inc = g_latency_increase(m);
    if (dur > 0.0) return inc / dur;
    return 0.0;
}
static double g_sent_latency_inflation(const pccf_mi *m)                                /* :151-156 */
{
    double dur = g_send_dur(m);
    double inc = g_latency_increase(m);
    if (dur > 0.0) return inc / dur;
    return 0.0;
}
static double g_conn_min_latency(const pccf_mi *m, pccf_flow *fl)                       /* :158-176 */
{
    double latency = g_avg_latency(m);
    if (fl->has_min) {
        double prev = fl->conn_min;
        if (latency == 0.0) return prev;
        if (latency < prev) { fl->conn_min = latency; return latency; }
        return prev;
    }
    if (latency > 0.0) { fl->has_min = 1; fl->conn_min = latency; return latency; }
    return 0.0;
}
static double g_send_ratio(const pccf_mi *m)                                            /* :179-184 */
{
    double thpt = g_recv_rate(m), send_rate = g_send_rate(m);
    if (thpt > 0.0 && send_rate < 1000.0 * thpt) return send_rate / thpt;
    return 1.0;
}
/* All 12 metrics of one MI, raw (not divided by scale).  `conn min latency` is memoised per MI in
 * the reference, so it is evaluated once and `latency ratio` reuses the value (:186-191). */
static void g_all_metrics(const pccf_mi *m, pccf_flow *fl, int touch_conn_min, double out[PCCF_N_METRICS])
{
    out[F_SEND_RATE] = g_send_rate(m);
    out[F_RECV_RATE] = g_recv_rate(m);
    out[F_RECV_DUR] = g_recv_dur(m);
    out[F_SEND_DUR] = g_send_dur(m);
    out[F_AVG_LATENCY] = g_avg_latency(m);
    out[F_LOSS_RATIO] = g_loss_ratio(m);
    out[F_ACK_LAT_INFL] = g_ack_latency_inflation(m);
    out[F_SENT_LAT_INFL] = g_sent_latency_inflation(m);
    out[F_LAT_INCREASE] = g_latency_increase(m);
    out[F_SEND_RATIO] = g_send_ratio(m);
    double cm;
    if (touch_conn_min) cm = g_conn_min_latency(m, fl);
    else { pccf_flow tmp = *fl; cm = g_conn_min_latency(m, &tmp); }
    out[F_CONN_MIN_LAT] = cm;
    out[F_LAT_RATIO] = (cm > 0.0) ? out[F_AVG_LATENCY] / cm : 1.0;                       /* :186-191 */
}

/* ---- flows ---------------------------------------------------------------------------------------- */
/* Rows of an empty history (SenderHistory.__init__ :57-62 builds history_len default MIs: all byte
 * counts 0.0, all times 0.0, no samples).  Their conn-min-dependent metrics see the dict entry of
 * `seen` (NULL or an entry-less flow: conn min 0.0, latency ratio 1.0; an entry c: conn min c,
 * latency ratio 0.0 / c = 0.0). */
static void fill_empty_rows(pccf_flows *fs, pccf_flow *fl, const pccf_flow *seen)
{
    for (int h = 0; h < fs->history_len; h++)
        for (int k = 0; k < fs->n_features; k++) {
            int id = fs->feature_ids[k];
            double v = 0.0;
            if (id == F_SEND_RATIO) v = 1.0;
            else if (id == F_CONN_MIN_LAT) v = (seen && seen->has_min) ? seen->conn_min : 0.0;
            else if (id == F_LAT_RATIO) v = (seen && seen->has_min && seen->conn_min > 0.0) ? 0.0 / seen->conn_min : 1.0;
            fl->rows[h][k] = v / f_scale[id];
        }
}

pccf_flows *pcco_flows_create(long n_flows, int history_len, const int *feature_ids, int n_features)
{
    if (n_flows < 1 || history_len < 1 || history_len > PCCF_MAX_HISTORY) return NULL;
    if (n_features < 1 || n_features > PCCF_MAX_FEATURES) return NULL;
    pccf_flows *fs = (pccf_flows *)calloc(1, sizeof(*fs));
    fs->n_flows = n_flows; fs->history_len = history_len; fs->n_features = n_features;
    for (int i = 0; i < n_features; i++) fs->feature_ids[i] = feature_ids[i];
    fs->f = (pccf_flow *)calloc((size_t)n_flows, sizeof(pccf_flow));
    for (long i = 0; i < n_flows; i++) fill_empty_rows(fs, &fs->f[i], NULL);
    return fs;
}
void pcco_flows_destroy(pccf_flows *fs) { if (fs) { free(fs->f); free(fs); } }

/* mode 0: a new flow id          -- no dict entry, empty rows (conn min 0, latency ratio 1)
 * mode 1: PccGymDriver.reset_history (loaded_client.py:97-101): same id, the dict entry survives and the
 *         new empty MIs (which carry that id) see it when first evaluated
 * mode 2: ShimNetworkEnv.reset (shim_env.py:141-149): the empty MIs carry sender id 0, records carry the
 *         wire flow id: the entry survives, the empty rows do not see it (flow id != 0) */
void pcco_flows_reset(pccf_flows *fs, long flow, int mode)
{
    pccf_flow *fl = &fs->f[flow];
    if (mode == 0) { fl->has_min = 0; fl->conn_min = 0.0; }
    fill_empty_rows(fs, fl, mode == 1 ? fl : NULL);
    fl->n_records = 0;
}

void pcco_flows_set_rate(pccf_flows *fs, long flow, double rate) { fs->f[flow].rate = rate; }
double pcco_flows_rate(const pccf_flows *fs, long flow) { return fs->f[flow].rate; }
double pcco_flows_conn_min(const pccf_flows *fs, long flow) { return fs->f[flow].has_min ? fs->f[flow].conn_min : 0.0; }
long long pcco_flows_n_records(const pccf_flows *fs, long flow) { return fs->f[flow].n_records; }

/* give_sample (loaded_client.py:111-138): record_observation -> history.step (sender_obs.py:64-66).
 * metrics12 (optional) receives the 12 raw metric values of this MI. */
void pcco_flows_give_sample(pccf_flows *fs, long flow, long long bytes_sent, long long bytes_acked, long long bytes_lost,
                            double send_start, double send_end, double recv_start, double recv_end,
                            const double *rtt, long n_rtt, long long packet_size, double *metrics12)
{
    pccf_flow *fl = &fs->f[flow];
    pccf_mi m = {bytes_sent, bytes_acked, bytes_lost, packet_size, send_start, send_end, recv_start, recv_end, rtt, n_rtt};
    double all[PCCF_N_METRICS];
    int touches = 0;
    for (int k = 0; k < fs->n_features; k++)
        if (fs->feature_ids[k] == F_CONN_MIN_LAT || fs->feature_ids[k] == F_LAT_RATIO) touches = 1;
    g_all_metrics(&m, fl, touches, all);
    for (int h = 0; h + 1 < fs->history_len; h++) memcpy(fl->rows[h], fl->rows[h + 1], sizeof(fl->rows[h]));
    for (int k = 0; k < fs->n_features; k++) {
        int id = fs->feature_ids[k];
        fl->rows[fs->history_len - 1][k] = all[id] / f_scale[id];                        /* :53-54 */
    }
    fl->n_records++;
    if (metrics12) memcpy(metrics12, all, sizeof(all));
}

void pcco_flows_get_obs(const pccf_flows *fs, long flow, double *obs)                    /* :68-73 */
{
    const pccf_flow *fl = &fs->f[flow];
    for (int h = 0; h < fs->history_len; h++)
        for (int k = 0; k < fs->n_features; k++) obs[h * fs->n_features + k] = fl->rows[h][k];
}

/* style 0: loaded_client.apply_rate_delta (:147-168): delta *= scale; > 0 multiply, < 0 divide, == 0 keep;
 *          clamp min then max.
 * style 1: ShimNetworkEnv.apply_action / set_rate (shim_env.py:80-95): >= 0 multiply, else divide; clamp max
 *          then min. */
double pcco_flows_apply_rate_delta(double rate, double delta, double delta_scale, double min_rate, double max_rate,
                                   int style)
{
    delta *= delta_scale;
    if (style == 0) {
        if (delta > 0) rate *= (1.0 + delta);
        else if (delta < 0) rate /= (1.0 - delta);
        if (rate < min_rate) rate = min_rate;
        if (rate > max_rate) rate = max_rate;
    } else {
        if (delta >= 0.0) rate = rate * (1.0 + delta);
        else rate = rate / (1.0 - delta);
        if (rate > max_rate) rate = max_rate;
        if (rate < min_rate) rate = min_rate;
    }
    return rate;
}

/* A whole batch in C (the cpu_baseline leg of bench.py --workload flows): records in batch order. */
void pcco_flows_give_batch(pccf_flows *fs, long n_records, const int *flow, const long long *bytes_sent,
                           const long long *bytes_acked, const long long *bytes_lost, const double *send_start,
                           const double *send_end, const double *recv_start, const double *recv_end,
                           const long long *packet_size, const long long *rtt_off, const double *rtt)
{
    for (long r = 0; r < n_records; r++)
        pcco_flows_give_sample(fs, flow[r], bytes_sent[r], bytes_acked[r], bytes_lost[r], send_start[r], send_end[r],
                               recv_start[r], recv_end[r], rtt + rtt_off[r], (long)(rtt_off[r + 1] - rtt_off[r]),
                               packet_size[r], NULL);
}

/* The same batch on n_threads host threads (bench.py --impl reference --workload flows): a flow's records stay on
 * one thread (flow % n_threads), so per-flow order is preserved and no state is shared. */
#include <pthread.h>
typedef struct {
    pccf_flows *fs; long n_records; int tid, n_threads;
    const int *flow; const long long *bs, *ba, *bl; const double *ss, *se, *rs, *re; const long long *ps, *off;
    const double *rtt;
} pccf_job;
static void *pccf_worker(void *arg)
{
    pccf_job *j = (pccf_job *)arg;
    for (long r = 0; r < j->n_records; r++) {
        if (j->flow[r] % j->n_threads != j->tid) continue;
        pcco_flows_give_sample(j->fs, j->flow[r], j->bs[r], j->ba[r], j->bl[r], j->ss[r], j->se[r], j->rs[r], j->re[r],
                               j->rtt + j->off[r], (long)(j->off[r + 1] - j->off[r]), j->ps[r], NULL);
    }
    return NULL;
}
void pcco_flows_give_batch_mt(pccf_flows *fs, int n_threads, long n_records, const int *flow, const long long *bytes_sent,
                              const long long *bytes_acked, const long long *bytes_lost, const double *send_start,
                              const double *send_end, const double *recv_start, const double *recv_end,
                              const long long *packet_size, const long long *rtt_off, const double *rtt)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 512) n_threads = 512;
    pthread_t th[512];
    pccf_job jobs[512];
    for (int t = 0; t < n_threads; t++) {
        pccf_job jb = {fs, n_records, t, n_threads, flow, bytes_sent, bytes_acked, bytes_lost, send_start, send_end,
                       recv_start, recv_end, packet_size, rtt_off, rtt};
        jobs[t] = jb;
        pthread_create(&th[t], NULL, pccf_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}
