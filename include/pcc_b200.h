/*
 * pcc_b200.h -- C ABI of libpcc_b200.so: a batched, B200-native replacement for the hot path
 * of PCC-RL's gym environment (one monitor interval of packet/link simulation + MI feature,
 * history and reward computation per env per call).
 *
 * The reference is 100 % Python and has no FFI of its own; this header is the boundary a
 * maintainer would bind with ctypes.  Each entry point names the
 * reference interface it replaces (file:line under /root/reference/src):
 *
 *   pcc_create / pcc_destroy   SimulatedNetworkEnv.__init__ / close     gym/network_sim.py:346-394, 489-492
 *   pcc_seed                   random.seed() feeding network_sim.py:73 (the per-packet loss draw)
 *   pcc_reset                  SimulatedNetworkEnv.reset                gym/network_sim.py:469-484
 *                              (create_new_links_and_senders :454-467 stays on the host: link
 *                              parameters and the start rate are inputs)
 *   pcc_step / pcc_step_host   SimulatedNetworkEnv.step                 gym/network_sim.py:406-444
 *                                -> Sender.apply_rate_delta             :235-241, 275-281
 *                                -> Network.run_for_dur                 :123-205
 *                                -> Sender.record_run / SenderHistory   :291-293; common/sender_obs.py:56-73
 *                                -> MI metrics                          common/sender_obs.py:110-191
 *   pcc_get_mt_state / pcc_set_mt_state   random.getstate() / random.setstate()
 *   pcc_rollout                the env side of PPO1's rollout loop      gym/stable_solve.py:52-58 (+ policy :30-45)
 *
 * Conventions
 *   - Every function returns 0 on success, a negative PCC_E* code otherwise; the message is
 *     available from pcc_last_error() (thread-local).  Nothing throws across the ABI.
 *   - Pointers named *_dev are CUDA device pointers on the handle's device, borrowed for the
 *     duration of the stream work the call enqueues.  `stream` is a cudaStream_t passed as
 *     void* (NULL = legacy default stream).  Calls are asynchronous unless stated otherwise.
 *   - All floating point is IEEE binary64, evaluated in the reference's operation order.
 *   - One host thread per handle; no global state.
 *   - There is no CPU fallback: without a CUDA device pcc_create fails.
 */
#ifndef PCC_B200_H
#define PCC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCC_ABI_VERSION 2

/* error codes */
#define PCC_OK 0
#define PCC_EINVAL (-1)   /* bad argument */
#define PCC_ECUDA (-2)    /* CUDA runtime error */
#define PCC_EOVERFLOW (-3)/* an env's in-flight ring overflowed (results of that env are invalid) */
#define PCC_ENODEV (-4)   /* no usable CUDA device */

/* metric ids = position in the reference's SENDER_MI_METRICS (common/sender_obs.py:193-206) */
#define PCC_M_SEND_RATE 0
#define PCC_M_RECV_RATE 1
#define PCC_M_RECV_DUR 2
#define PCC_M_SEND_DUR 3
#define PCC_M_AVG_LATENCY 4
#define PCC_M_LOSS_RATIO 5
#define PCC_M_ACK_LATENCY_INFLATION 6
#define PCC_M_SENT_LATENCY_INFLATION 7
#define PCC_M_CONN_MIN_LATENCY 8
#define PCC_M_LATENCY_INCREASE 9
#define PCC_M_LATENCY_RATIO 10
#define PCC_M_SEND_RATIO 11
#define PCC_N_METRICS 12
#define PCC_MAX_FEATURES 12
#define PCC_MAX_HISTORY 64

/* RNG feeding the per-packet loss draw (network_sim.py:73) */
#define PCC_RNG_MT19937 0 /* CPython's `random`: bit-compatible state, for drop-in fidelity */
#define PCC_RNG_PHILOX 1  /* Philox4x32-10, counter = draw index, key = per-env seed (fast path) */

/* width of one row of the optional `info` output of pcc_step:
 *  0 send rate  1 throughput (recv rate)  2 avg latency  3 loss ratio  4 latency inflation
 *  5 latency ratio  6 send ratio          (the event-log fields of network_sim.py:422-436)
 *  7 MI duration  8 cur_time after the MI  9 sending rate  10 next run_dur  11 conn min latency */
#define PCC_INFO_WIDTH 12

/* Simulation constants (gym/network_sim.py:33-54, common/config.py:17). */
typedef struct pcc_consts {
    double max_rate;          /* MAX_RATE          1000  */
    double min_rate;          /* MIN_RATE          40    */
    double delta_scale;       /* DELTA_SCALE       0.025 */
    double reward_scale;      /* REWARD_SCALE      0.001 */
    int32_t max_steps;        /* MAX_STEPS         400   */
    int32_t bytes_per_packet; /* BYTES_PER_PACKET  1500  */
} pcc_consts;

typedef struct pcc_config {
    int32_t abi_version;      /* PCC_ABI_VERSION */
    int32_t device;           /* CUDA device ordinal */
    int64_t n_envs;           /* independent environments stepped in lock step */
    int32_t history_len;      /* --history-len, network_sim.py:347 (default 10) */
    int32_t n_features;       /* --input-features, network_sim.py:348-351 (default 3) */
    int32_t feature_ids[PCC_MAX_FEATURES];
    int32_t rng_kind;         /* PCC_RNG_* */
    int32_t reserved0;
    int64_t ring_capacity;    /* in-flight records per env (16 B each), power of two; must cover
                                 packets in flight + packets sent in one MI (see pcc_ring_capacity_for) */
    pcc_consts consts;
} pcc_config;

typedef struct pcc_handle_s *pcc_handle;

/* Fills `c` with the reference's constants. */
void pcc_default_consts(pcc_consts *c);

/* Fills `cfg` with the reference's defaults (history 10, the three default features, Philox,
 * reference constants); n_envs, device and ring_capacity are left for the caller. */
void pcc_default_config(pcc_config *cfg);

/* Smallest power-of-two ring capacity that cannot overflow for link parameters within the given
 * bounds: 1.5 * max_rate * (2*max_delay + max_queue/min_bw), plus the warm-up MIs of reset. */
int64_t pcc_ring_capacity_for(double max_rate, double min_bw, double max_delay, double max_queue);

/* Bytes of device memory the caller must provide to pcc_create: `state_bytes` for the
 * structure-of-arrays env state (+ history, RNG state), `ring_bytes` for the in-flight rings.
 * The two blocks together are a complete checkpoint of the simulator. */
int pcc_workspace_bytes(const pcc_config *cfg, uint64_t *state_bytes, uint64_t *ring_bytes);

/* Creates a handle over caller-owned device memory (256-byte aligned; e.g. torch.empty(uint8)).
 * The state block is zero-initialised by this call (asynchronously on the default stream,
 * followed by a device synchronise). */
int pcc_create(pcc_handle *out, const pcc_config *cfg, void *state_dev, void *ring_dev);
void pcc_destroy(pcc_handle h);

/* Attaches a handle to workspaces that already hold a checkpoint (no initialisation). */
int pcc_attach(pcc_handle *out, const pcc_config *cfg, void *state_dev, void *ring_dev);

/* Seeds the per-env loss-draw streams.  seeds_dev: uint64[n_envs].  mask_dev: uint8[n_envs] or
 * NULL (= all).  PHILOX: key = seed, draw counter = 0.  MT19937: CPython's random.seed(int). */
int pcc_seed(pcc_handle h, const uint64_t *seeds_dev, const uint8_t *mask_dev, void *stream);

/* Synchronous transfer of one env's MT19937 state in random.getstate()[1] layout: 624 words +
 * position (uint32[625], host memory). */
int pcc_get_mt_state(pcc_handle h, int64_t env, uint32_t *state_host);
int pcc_set_mt_state(pcc_handle h, int64_t env, const uint32_t *state_host);

/* reset(): builds a fresh link pair + sender for every env selected by mask_dev (NULL = all) and
 * runs the two discarded warm-up MIs.  Parameter arrays are double[n_envs] / int64[n_envs] on the
 * device, in the reference's units (bw: packets/s, delay: s, queue: packets, loss: probability,
 * start_rate: packets/s).  obs_dev (optional): double[n_envs][history_len*n_features]; rows of
 * unselected envs are left untouched. */
int pcc_reset(pcc_handle h, const uint8_t *mask_dev, const double *bw_dev, const double *delay_dev,
              const int64_t *queue_dev, const double *loss_dev, const double *start_rate_dev,
              double *obs_dev, void *stream);

/* step(): one monitor interval for every env.
 *   actions_dev  double[n_envs]                        (action[0] of the reference; finite)
 *   obs_dev      double[n_envs][history_len*n_features] oldest MI first, feature-minor
 *   reward_dev   double[n_envs]
 *   done_dev     uint8[n_envs]    steps_taken >= max_steps   (the env is NOT reset automatically)
 *   counts_dev   int32[n_envs][3] packets sent / acked / lost in this MI      (optional)
 *   info_dev     double[n_envs][PCC_INFO_WIDTH]                               (optional) */
int pcc_step(pcc_handle h, const double *actions_dev, double *obs_dev, double *reward_dev,
             uint8_t *done_dev, int32_t *counts_dev, double *info_dev, void *stream);

/* The same step through HOST buffers (the call a gym-style user makes): copies actions to the
 * device, steps, copies obs / reward / done (/ counts) back and synchronises the stream.  Host
 * buffers should be page-locked for full copy bandwidth. */
int pcc_step_host(pcc_handle h, const double *actions_host, double *obs_host, double *reward_host,
                  uint8_t *done_host, int32_t *counts_host, void *stream);

/* The same, split in two so that a caller can keep two steps in flight (e.g. two half-batches of a vector env, or
 * actions that do not depend on the previous observation): submit enqueues upload, step and download -- the download
 * runs on a stream of the handle's own, from one of two device staging slots, so it overlaps the NEXT submission's
 * upload and kernel -- and returns a ticket; wait blocks until that step's host buffers are complete.  At most two
 * tickets may be outstanding; each needs its own host buffers.  pcc_step_host == submit + wait. */
int pcc_step_host_submit(pcc_handle h, const double *actions_host, double *obs_host, double *reward_host,
                         uint8_t *done_host, int32_t *counts_host, double *info_host /* optional, [n][PCC_INFO_WIDTH] */,
                         void *stream, int64_t *ticket);
int pcc_step_host_wait(pcc_handle h, int64_t ticket);

/* On-device policy for pcc_rollout: the MLP of stable_solve.py:30-45 (obs -> h1 -> h2 -> 1, tanh hidden
 * layers, linear output = mean of the Gaussian action) and, optionally, the value network of the same shape that
 * PPO1 trains beside it (MlpPolicy: vf head; vw1 == NULL = none).  All pointers are device pointers to binary64,
 * row-major [out][in].  stochastic != 0 adds exp(log_std) * N(0,1) drawn from a Philox stream keyed by
 * (noise_seed; env, step). */
typedef struct pcc_policy {
    const double *w1, *b1, *w2, *b2, *w3, *b3;
    int32_t n_in, h1, h2, stochastic;
    double log_std;
    uint64_t noise_seed;
    const double *vw1, *vb1, *vw2, *vb2, *vw3, *vb3;
} pcc_policy;

/* Rollout: n_steps monitor intervals for every env without returning to the host -- what PPO1's
 * traj_segment_generator does around SimulatedNetworkEnv.step (stable_solve.py:52-58: ob, ac, vpred, rew, new per
 * step plus nextvpred).  Per step, enqueued back to back on `stream`: the policy / value kernel on the current
 * observation, exactly pcc_step, and for finished envs exactly pcc_reset with the next row of the parameter bank.
 *   actions_dev       double[n_steps][n_envs], or NULL to use `policy` (then actions_out_dev is required)
 *   reset_params_dev  double[n_episodes][5][n_envs]: bw, delay, queue, loss, start_rate of the
 *                     episodes each env starts DURING this rollout, in order (row 0 = its first reset);
 *                     n_episodes == 0: no auto-reset
 *   obs_dev           double[n_steps][n_envs][history_len*n_features]  (optional) observation AFTER the
 *                     step -- for a finished env the first observation of its next episode
 *   actions_out_dev   double[n_steps][n_envs] (optional with actions_dev) the actions taken
 *   reward_dev, done_dev, counts_dev   [n_steps][n_envs] (counts optional, [..][3])
 *   vpred_dev         double[n_steps + 1][n_envs] (optional, needs policy->vw1): row k = V(observation BEFORE step k),
 *                     row n_steps = V(observation after the last step) */
int pcc_rollout(pcc_handle h, int32_t n_steps, const double *actions_dev, const pcc_policy *policy,
                const double *reset_params_dev, int32_t n_episodes, double *obs_dev, double *actions_out_dev,
                double *reward_dev, uint8_t *done_dev, int32_t *counts_dev, double *vpred_dev, void *stream);

/* Synchronises `stream` and reports sticky device-side errors (PCC_EOVERFLOW). */
int pcc_check(pcc_handle h, void *stream);

/* Copies one per-env scalar column of the state to dst_dev (double[n_envs]); names: "cur_time",
 * "run_dur", "rate", "next_send", "queue_delay", "conn_min", "bw", "delay", "loss", "max_queue_delay",
 * "episode_return" (reward_sum of the running episode, network_sim.py:443) and "last_episode_return".
 * For tests and checkpoints of derived quantities. */
int pcc_get_column(pcc_handle h, const char *name, double *dst_dev, void *stream);

/* Number of kernel launches issued by this handle so far (pcc_step = 1 launch). */
int64_t pcc_launch_count(pcc_handle h);

/* ---- several senders per link (BASELINE config 5) ------------------------------------------------------
 * The reference's Network takes lists of senders (gym/network_sim.py:100-126, 140-178) although its env
 * creates one.  Semantics here: all senders share links [l0, l1]; heap ties are broken by sender index;
 * step applies actions[i] to sender i (:409-412 generalised); every sender has its own MI, history, obs and
 * reward (:194,205 on its own MI); the MI duration follows sender 0 (:437-438 as written).  Two engines with identical
 * results, fixed at the first pcc_multi_reset: the heap-free streaming MI (default: one shared in-flight ring with a
 * sender id per record, timers merged by (time, sender), three cursors; one link per warp, links visited in
 * descending predicted cost -- environment PCC_MULTI_MODE=thread runs it with one link per thread) and the per-env
 * event heap, one env per thread (PCC_MULTI_MODE=heap, and always for the cwnd / latency-noise variants below).
 * cfg->ring_capacity
 * (a power of two) sizes the in-flight ring / the heap (x n_senders events) and the per-sender RTT sample buffers;
 * Philox streams only.  Arrays: [n_envs][n_senders]... row-major. */
typedef struct pcc_multi_handle_s *pcc_multi_handle;
int pcc_multi_workspace_bytes(const pcc_config *cfg, int32_t n_senders, uint64_t *bytes);
int pcc_multi_create(pcc_multi_handle *out, const pcc_config *cfg, int32_t n_senders, void *workspace_dev);
void pcc_multi_destroy(pcc_multi_handle h);
int pcc_multi_seed(pcc_multi_handle h, const uint64_t *seeds_dev, void *stream);
int pcc_multi_reset(pcc_multi_handle h, const uint8_t *mask_dev, const double *bw_dev, const double *delay_dev,
                    const int64_t *queue_dev, const double *loss_dev, const double *start_rates_dev /*[n][S]*/,
                    double *obs_dev /*[n][S][H*F], optional*/, void *stream);
int pcc_multi_step(pcc_multi_handle h, const double *actions_dev /*[n][S]*/, double *obs_dev /*[n][S][H*F]*/,
                   double *reward_dev /*[n][S]*/, uint8_t *done_dev /*[n]*/, int32_t *counts_dev /*[n][S][3], optional*/,
                   void *stream);
int pcc_multi_check(pcc_multi_handle h, void *stream);
/* kernels of this library launched for the handle so far (steps, resets, cost keys of the link sort) */
int64_t pcc_multi_launch_count(pcc_multi_handle h);

/* The two variants of the event loop that the reference compiles out behind module switches:
 *   use_cwnd           USE_CWND = True  (network_sim.py:54): congestion window, Sender.can_send_packet :251-255,
 *                      apply_cwnd_delta / set_cwnd :243-249, 283-289, second action component :413-414
 *   use_latency_noise  USE_LATENCY_NOISE = True (:51-52): every hop's latency times random.uniform(1.0, 1.1), :150-151, 171-172
 * They run on the generic heap path (n_senders = 1 is the reference's own single-sender env with the switch on).
 * Set before pcc_multi_reset. */
typedef struct pcc_variant {
    int32_t use_cwnd;
    int32_t use_latency_noise;
    double max_latency_noise;   /* MAX_LATENCY_NOISE 1.1 (:52) */
    int32_t initial_cwnd;       /* 25   (Sender.__init__, :209) */
    int32_t min_cwnd;           /* MIN_CWND 4    (:34) */
    int32_t max_cwnd;           /* MAX_CWND 5000 (:33) */
    int32_t reserved0;
} pcc_variant;
void pcc_default_variant(pcc_variant *v);   /* both switches off, the reference's constants */
int pcc_multi_set_variant(pcc_multi_handle h, const pcc_variant *v);
/* pcc_multi_step with the window action: cwnd_actions_dev double[n][S] (NULL = leave the windows alone),
 * cwnd_dev int32[n][S] (optional) = the windows after the update. */
int pcc_multi_step_cwnd(pcc_multi_handle h, const double *actions_dev, const double *cwnd_actions_dev, double *obs_dev,
                        double *reward_dev, uint8_t *done_dev, int32_t *counts_dev, int32_t *cwnd_dev, void *stream);

/* ---- MI-sample ingestion ("flow monitor") -------------------------------------------
 * The other producer of SenderMonitorIntervals: records measured on REAL flows by the PCC sender, handed to
 * Python one at a time and turned into the agent's observation:
 *   pcc_flows_give_samples   give_sample -> SenderMonitorInterval -> SenderHistory.step
 *                              udt-plugins/testing/loaded_client.py:84-86, 111-138; gym/online/shim_env.py:102-139
 *                              (wire format udt-plugins/training/shim.py:31-42); common/sender_obs.py:20-73, 110-191
 *   pcc_flows_get_obs        SenderHistory.as_array                     common/sender_obs.py:68-73
 *   pcc_flows_reset          PccGymDriver.reset_history / ShimNetworkEnv.reset   loaded_client.py:97-101; shim_env.py:141-149
 *   pcc_flows_get_rates      PccGymDriver.get_rate -> apply_rate_delta  loaded_client.py:72-76, 147-168
 *                            ShimNetworkEnv.apply_action / set_rate     shim_env.py:80-95
 *   pcc_flows_act            LoadedModelAgent.act (the saved MLP policy) loaded_client.py:74; stable_solve.py:30-45
 * Here a batch carries records of many flows (structure of arrays + CSR sample lists, device pointers); each
 * record's metrics are evaluated when it is ingested and appended to its flow's history, exactly what the
 * reference computes when history.as_array() is taken at least once per history_len records of a flow (it
 * memoises metrics per MI and updates the global _conn_min_latencies dict on first evaluation).  Byte counts
 * and packet sizes are integers below 2^53.  Records of one flow inside a batch apply in batch order. */
#define PCC_RATE_CLIENT 0     /* loaded_client.apply_rate_delta: >0 multiply, <0 divide; only flows that have data */
#define PCC_RATE_SHIM 1       /* ShimNetworkEnv.apply_action:    >=0 multiply, else divide; every selected flow   */
#define PCC_FLOW_RESET_NEW 0     /* a new flow id: no conn-min dict entry, empty history                          */
#define PCC_FLOW_RESET_CLIENT 1  /* PccGymDriver.reset_history: entry survives, the new empty MIs see it          */
#define PCC_FLOW_RESET_SHIM 2    /* ShimNetworkEnv.reset: entry survives, the empty MIs (sender id 0) do not see it */

typedef struct pcc_flows_config {
    int32_t abi_version;      /* PCC_ABI_VERSION */
    int32_t device;
    int64_t n_flows;
    int32_t history_len;      /* --history-len (default 10) */
    int32_t n_features;       /* --input-features (default: the three of network_sim.py:348-351) */
    int32_t feature_ids[PCC_MAX_FEATURES];
    double delta_scale;       /* DELTA_SCALE  0.05 (loaded_client.py:35)  / 0.025 (shim_env.py:43) */
    double min_rate;          /* MIN_RATE     0.5  (loaded_client.py:33)  / 0.25  (shim_env.py:40) */
    double max_rate;          /* MAX_RATE     300  (loaded_client.py:34)  / 1000  (shim_env.py:39) */
    int32_t rate_style;       /* PCC_RATE_* */
    int32_t reserved0;
} pcc_flows_config;

/* One batch of MI records; every pointer is a device pointer, arrays have n_records elements. */
typedef struct pcc_mi_batch {
    int64_t n_records;
    const int32_t *flow;          /* index of the record's flow, 0 .. n_flows-1 */
    const int64_t *bytes_sent, *bytes_acked, *bytes_lost, *packet_size;
    const double *send_start, *send_end, *recv_start, *recv_end;   /* seconds */
    const int64_t *rtt_offsets;   /* [n_records + 1]: record r owns rtt_samples[rtt_offsets[r] .. rtt_offsets[r+1]) */
    const double *rtt_samples;    /* seconds, in arrival order */
} pcc_mi_batch;

typedef struct pcc_flows_handle_s *pcc_flows_handle;

/* loaded_client's defaults: history 10, the three default features, rate control 0.05 / 0.5 / 300. */
void pcc_flows_default_config(pcc_flows_config *cfg);
int pcc_flows_workspace_bytes(const pcc_flows_config *cfg, uint64_t *bytes);
/* Caller-owned, 256-byte aligned device workspace; create initialises every flow as PCC_FLOW_RESET_NEW with
 * rate 0 (synchronous); attach adopts a workspace that already holds state (checkpoint). */
int pcc_flows_create(pcc_flows_handle *out, const pcc_flows_config *cfg, void *workspace_dev);
int pcc_flows_attach(pcc_flows_handle *out, const pcc_flows_config *cfg, void *workspace_dev);
void pcc_flows_destroy(pcc_flows_handle h);

/* Ingests a batch.  unique_flows != 0: the caller guarantees at most one record per flow (one fused kernel;
 * violations are detected and reported by pcc_flows_check); 0: any batch (records of a flow apply in batch order).
 *   obs_dev      double[n_records][history_len*n_features]  (optional) the flow's observation right after the record
 *   metrics_dev  double[n_records][PCC_N_METRICS]           (optional) the record's 12 raw metric values */
int pcc_flows_give_samples(pcc_flows_handle h, const pcc_mi_batch *batch, int32_t unique_flows, double *obs_dev,
                           double *metrics_dev, void *stream);
/* mask_dev: uint8[n_flows] or NULL (= all); mode: PCC_FLOW_RESET_*.  Rates are not touched (the reference's
 * reset_rate sets an unused attribute, loaded_client.py:94-95); use pcc_flows_set_rates. */
int pcc_flows_reset(pcc_flows_handle h, const uint8_t *mask_dev, int32_t mode, void *stream);
/* obs_dev: double[n_flows][history_len*n_features], oldest MI first. */
int pcc_flows_get_obs(pcc_flows_handle h, double *obs_dev, void *stream);
/* rate[flow] = rates_dev[flow], or `rate` for every selected flow when rates_dev is NULL. */
int pcc_flows_set_rates(pcc_flows_handle h, const uint8_t *mask_dev, const double *rates_dev, double rate, void *stream);
/* Applies actions_dev[flow] (optional) to the selected flows' rates (PCC_RATE_CLIENT: only flows that have
 * received a record since their last reset) and writes the rates to rates_dev[n_flows] (optional; the
 * reference's get_rate returns rate * 1e6). */
int pcc_flows_get_rates(pcc_flows_handle h, const double *actions_dev, const uint8_t *mask_dev, double *rates_dev,
                        void *stream);
/* actions_dev[flow] = policy(observation of the flow), deterministic (stochastic = False). */
int pcc_flows_act(pcc_flows_handle h, const pcc_policy *policy, double *actions_dev, void *stream);
/* Copies a per-flow column ("conn_min", "rate") to dst_dev[n_flows]. */
int pcc_flows_get_column(pcc_flows_handle h, const char *name, double *dst_dev, void *stream);
/* Synchronises and reports sticky errors: out-of-range flow index, duplicate flow in a unique batch. */
int pcc_flows_check(pcc_flows_handle h, void *stream);
int64_t pcc_flows_launch_count(pcc_flows_handle h);

const char *pcc_last_error(void);
int pcc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PCC_B200_H */
