// pcc_flows_core.cuh -- MI-sample ingestion ("flow monitor"), scalar core shared by the CUDA kernels of
// pcc_flows.cuh and by the host-compiled twin harness (tests/twin).
//
// What this replaces (reference file:line, under /root/reference/src): monitor-interval records measured
// on real flows are pushed into a per-flow SenderHistory whose array is the agent's observation
//   give_sample -> record_observation          udt-plugins/testing/loaded_client.py:84-86, 111-138
//   ShimNetworkEnv.step (record off a socket)  gym/online/shim_env.py:102-139
//   SenderMonitorInterval / SenderHistory      common/sender_obs.py:20-73
//   the 12 metrics in their general form       common/sender_obs.py:110-191 (table :193-206)
//   apply_rate_delta                           loaded_client.py:147-168, shim_env.py:80-95
// Unlike the simulator's MIs (pcc_core.cuh), a record has independent send and recv windows, arbitrary
// byte counts (|bytes| < 2^53, exact in binary64 like Python's int -> float) and its own packet size.
//
// Memoisation (sender_obs.py:44-54) and the module-global _conn_min_latencies dict (:158-176): a record's
// metrics are evaluated when it is ingested, the empty MIs of a history when it is (re)created -- what the
// reference computes provided history.as_array() is taken at least once per history_len records.  The
// dict entry of a flow survives a history reset (the reference never clears it).
#pragma once
#include "pcc_core.cuh"

namespace pcc {

struct FlowRecord {
    int64_t bytes_sent, bytes_acked, bytes_lost, packet_size;
    double send_start, send_end, recv_start, recv_end;
};

// history reset variants (see pcc_b200.h)
enum { FLOW_RESET_NEW = 0, FLOW_RESET_CLIENT = 1, FLOW_RESET_SHIM = 2 };
enum { RATE_STYLE_CLIENT = 0, RATE_STYLE_SHIM = 1 };

// Raw values of the 12 metrics, indexed by metric id.
struct FlowStats { double v[N_METRICS]; };

// _mi_metric_conn_min_latency (sender_obs.py:158-176) on the flow's dict entry (has_min, conn_min).
PCC_HD double flow_conn_min(double avg_lat, bool &has_min, double &conn_min, bool update)
{
    if (has_min) {
        const double prev = conn_min;
        if (avg_lat == 0.0) return prev;
        if (avg_lat < prev) { if (update) conn_min = avg_lat; return avg_lat; }
        return prev;
    }
    if (avg_lat > 0.0) { if (update) { has_min = true; conn_min = avg_lat; } return avg_lat; }
    return 0.0;
}

// Everything scalar: takes np.mean(all), np.mean(first half), np.mean(second half) as inputs
// (m_first / m_second are ignored when half == 0).
PCC_HD void flow_stats_finish(const FlowRecord &r, int64_t n, double avg_lat, double m_first, double m_second,
                              bool &has_min, double &conn_min, bool update_conn_min, FlowStats &st)
{
    const double sdur = r.send_end - r.send_start;                                          // :130-131
    const double rdur = r.recv_end - r.recv_start;                                          // :116-117
    const double send_rate = (sdur > 0.0) ? 8.0 * (double)r.bytes_sent / sdur : 0.0;        // :124-128
    const double recv_rate = (rdur > 0.0) ? 8.0 * (double)(r.bytes_acked - r.packet_size) / rdur : 0.0;   // :110-114
    const double loss = (r.bytes_lost + r.bytes_acked > 0)
        ? (double)r.bytes_lost / (double)(r.bytes_lost + r.bytes_acked) : 0.0;              // :133-136
    const double inc = (n / 2 >= 1) ? m_second - m_first : 0.0;                             // :138-142
    st.v[M_SEND_RATE] = send_rate;
    st.v[M_RECV_RATE] = recv_rate;
    st.v[M_RECV_DUR] = rdur;
    st.v[M_SEND_DUR] = sdur;
    st.v[M_AVG_LATENCY] = avg_lat;                                                          // :119-122
    st.v[M_LOSS_RATIO] = loss;
    st.v[M_ACK_LAT_INFL] = (rdur > 0.0) ? inc / rdur : 0.0;                                 // :144-149
    st.v[M_SENT_LAT_INFL] = (sdur > 0.0) ? inc / sdur : 0.0;                                // :151-156
    st.v[M_LAT_INCREASE] = inc;
    const double cm = flow_conn_min(avg_lat, has_min, conn_min, update_conn_min);
    st.v[M_CONN_MIN_LAT] = cm;
    st.v[M_LAT_RATIO] = (cm > 0.0) ? avg_lat / cm : 1.0;                                    // :186-191
    st.v[M_SEND_RATIO] = (recv_rate > 0.0 && send_rate < 1000.0 * recv_rate) ? send_rate / recv_rate : 1.0;   // :179-184
}

PCC_HD double flow_metric_scale(int id) { return (id == M_SEND_RATE || id == M_RECV_RATE) ? 1e7 : 1.0; }

// Metric `id` of an EMPTY MI (SenderHistory.__init__, sender_obs.py:57-62: no bytes, all times 0.0, no
// samples) evaluated while the dict entry it sees is (seen_min, conn_min): conn min latency = the entry
// (:162-164 with latency == 0.0), latency ratio = 0.0 / entry; without an entry 0.0 and 1.0.
PCC_HD double flow_metric_empty(int id, bool seen_min, double conn_min)
{
    if (id == M_SEND_RATIO) return 1.0;
    if (id == M_CONN_MIN_LAT) return seen_min ? conn_min : 0.0;
    if (id == M_LAT_RATIO) return (seen_min && conn_min > 0.0) ? 0.0 / conn_min : 1.0;
    return 0.0;
}

PCC_HD bool features_touch_conn_min(const int *ids, int n)
{
    for (int k = 0; k < n; k++) if (ids[k] == M_CONN_MIN_LAT || ids[k] == M_LAT_RATIO) return true;
    return false;
}

// loaded_client.apply_rate_delta (:147-168; > 0 multiply, < 0 divide, clamp min then max) and
// ShimNetworkEnv.apply_action + set_rate (shim_env.py:80-95; >= 0 multiply, clamp max then min).
PCC_HD double flow_apply_rate_delta(double rate, double action, double delta_scale, double min_rate, double max_rate,
                                    int style)
{
    const double delta = action * delta_scale;
    if (style == RATE_STYLE_CLIENT) {
        if (delta > 0) rate *= (1.0 + delta);
        else if (delta < 0) rate /= (1.0 - delta);
        if (rate < min_rate) rate = min_rate;
        if (rate > max_rate) rate = max_rate;
    } else {
        rate = (delta >= 0.0) ? rate * (1.0 + delta) : rate / (1.0 - delta);
        if (rate > max_rate) rate = max_rate;
        if (rate < min_rate) rate = min_rate;
    }
    return rate;
}

// np.mean over a plain array, numpy's pairwise summation (scalar; the kernels use 8-lane leaves)
struct FlowArrayReader {
    const double *a; int64_t i;
    PCC_HD double next() { return a[i++]; }
};

// Scalar reference of one record: used by the host twin and by nothing on the GPU hot path.
PCC_HD void flow_record_stats(const FlowRecord &r, const double *rtt, int64_t n, bool &has_min, double &conn_min,
                              bool update_conn_min, FlowStats &st)
{
    double avg = 0.0, m1 = 0.0, m2 = 0.0;
    if (n > 0) { FlowArrayReader rd{rtt, 0}; avg = np_mean_stream(rd, (int)n); }
    const int64_t half = n / 2;
    if (half >= 1) {
        FlowArrayReader rd{rtt, 0};
        m1 = np_mean_stream(rd, (int)half);
        m2 = np_mean_stream(rd, (int)(n - half));
    }
    flow_stats_finish(r, n, avg, m1, m2, has_min, conn_min, update_conn_min, st);
}

}  // namespace pcc
