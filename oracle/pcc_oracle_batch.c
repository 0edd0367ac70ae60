/*
 * pcc_oracle_batch.c -- TEST INFRASTRUCTURE (CPU baseline driver), NOT PRODUCT CODE.
 *
 * Runs many independent oracle envs (pcc_oracle.c) on host threads so that bench.py can
 * time "the reference's algorithm on this box's cores" next to the GPU number.  Envs are
 * statically partitioned over pthreads; there is no shared state.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

typedef struct pcco_env pcco_env;
pcco_env *pcco_create(int history_len, const int *feature_ids, int n_features);
void pcco_destroy(pcco_env *e);
void pcco_seed_philox(pcco_env *e, uint64_t seed);
void pcco_reset(pcco_env *e, double bw, double lat, long queue_size, double loss, double start_rate);
void pcco_step(pcco_env *e, double action, double *obs, double *reward, int *done, long *counts, double *info);

typedef struct {
    int tid, n_threads;
    long n_envs, n_steps;
    int history_len, n_features;
    const int *feature_ids;
    const double *bw, *lat, *loss, *start_rate;
    const long *queue;
    const uint64_t *seeds;
    const double *actions; /* [n_steps][n_envs] or NULL (=0.0) */
    double *obs_out;       /* [n_envs][H*F] of the last step, or NULL */
    double *reward_sum;    /* [n_envs] */
    long *count_sum;       /* [n_envs][3] */
    double *reward_traj;   /* [n_steps][n_envs] or NULL */
    int *count_traj;       /* [n_steps][n_envs][3] or NULL */
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    long lo = j->n_envs * j->tid / j->n_threads, hi = j->n_envs * (j->tid + 1) / j->n_threads;
    int hf = j->history_len * j->n_features;
    double *obs = (double *)malloc(sizeof(double) * (size_t)hf);
    for (long i = lo; i < hi; i++) {
        pcco_env *e = pcco_create(j->history_len, j->feature_ids, j->n_features);
        pcco_seed_philox(e, j->seeds[i]);
        pcco_reset(e, j->bw[i], j->lat[i], j->queue[i], j->loss[i], j->start_rate[i]);
        double rs = 0.0;
        long cs[3] = {0, 0, 0};
        for (long t = 0; t < j->n_steps; t++) {
            double r; int d; long c[3];
            double a = j->actions ? j->actions[t * j->n_envs + i] : 0.0;
            pcco_step(e, a, obs, &r, &d, c, NULL);
            rs += r; cs[0] += c[0]; cs[1] += c[1]; cs[2] += c[2];
            if (j->reward_traj) j->reward_traj[t * j->n_envs + i] = r;
            if (j->count_traj) for (int k = 0; k < 3; k++) j->count_traj[(t * j->n_envs + i) * 3 + k] = (int)c[k];
        }
        j->reward_sum[i] = rs;
        for (int k = 0; k < 3; k++) j->count_sum[3 * i + k] = cs[k];
        if (j->obs_out) for (int k = 0; k < hf; k++) j->obs_out[i * hf + k] = obs[k];
        pcco_destroy(e);
    }
    free(obs);
    return NULL;
}

/* Each env: reset with its parameters (2 warm-up MIs), then n_steps steps.  Returns the wall
 * time in seconds (CLOCK_MONOTONIC) of the threaded region. */
double pcco_batch_run(long n_envs, long n_steps, int n_threads, int history_len,
                      const int *feature_ids, int n_features,
                      const double *bw, const double *lat, const long *queue, const double *loss,
                      const double *start_rate, const uint64_t *seeds, const double *actions,
                      double *obs_out, double *reward_sum, long *count_sum,
                      double *reward_traj, int *count_traj)
{
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        job_t j = {t, n_threads, n_envs, n_steps, history_len, n_features, feature_ids,
                   bw, lat, loss, start_rate, queue, seeds, actions, obs_out, reward_sum, count_sum,
                   reward_traj, count_traj};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------ */
/* Persistent batch: the CPU arm of bench.py steps the same env batch the GPU arm does,   */
/* one pcco_batch_step per "step", envs statically partitioned over host threads.         */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    long n;
    int hf;
    pcco_env **envs;
} pcco_batch;

pcco_batch *pcco_batch_create(long n, int history_len, const int *feature_ids, int n_features,
                              const uint64_t *seeds)
{
    pcco_batch *b = (pcco_batch *)calloc(1, sizeof(pcco_batch));
    b->n = n;
    b->hf = history_len * n_features;
    b->envs = (pcco_env **)calloc((size_t)n, sizeof(pcco_env *));
    for (long i = 0; i < n; i++) {
        b->envs[i] = pcco_create(history_len, feature_ids, n_features);
        pcco_seed_philox(b->envs[i], seeds[i]);
    }
    return b;
}

void pcco_batch_destroy(pcco_batch *b)
{
    if (!b) return;
    for (long i = 0; i < b->n; i++) pcco_destroy(b->envs[i]);
    free(b->envs);
    free(b);
}

typedef struct {
    pcco_batch *b;
    int tid, n_threads, op; /* op 0 = reset, 1 = step */
    const unsigned char *mask;
    const double *bw, *lat, *loss, *start_rate, *actions;
    const long *queue;
    double *obs, *reward;
    unsigned char *done;
    int *counts;
} bjob_t;

void pcco_get_obs(const pcco_env *e, double *obs);

static void *bworker(void *arg)
{
    bjob_t *j = (bjob_t *)arg;
    pcco_batch *b = j->b;
    long lo = b->n * j->tid / j->n_threads, hi = b->n * (j->tid + 1) / j->n_threads;
    for (long i = lo; i < hi; i++) {
        if (j->op == 0) {
            if (j->mask && !j->mask[i]) continue;
            pcco_reset(b->envs[i], j->bw[i], j->lat[i], j->queue[i], j->loss[i], j->start_rate[i]);
            if (j->obs) pcco_get_obs(b->envs[i], j->obs + i * b->hf);
        } else {
            double r; int d; long c[3];
            pcco_step(b->envs[i], j->actions[i], j->obs + i * b->hf, &r, &d, c, NULL);
            j->reward[i] = r;
            j->done[i] = (unsigned char)d;
            if (j->counts) for (int k = 0; k < 3; k++) j->counts[3 * i + k] = (int)c[k];
        }
    }
    return NULL;
}

static void brun(bjob_t proto, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    bjob_t *jobs = (bjob_t *)malloc(sizeof(bjob_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = proto; jobs[t].tid = t; jobs[t].n_threads = n_threads;
        pthread_create(&th[t], NULL, bworker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

void pcco_batch_reset(pcco_batch *b, const unsigned char *mask, const double *bw, const double *lat,
                      const long *queue, const double *loss, const double *start_rate, double *obs,
                      int n_threads)
{
    bjob_t j; memset(&j, 0, sizeof(j));
    j.b = b; j.op = 0; j.mask = mask; j.bw = bw; j.lat = lat; j.queue = queue; j.loss = loss;
    j.start_rate = start_rate; j.obs = obs;
    brun(j, n_threads);
}

void pcco_batch_step(pcco_batch *b, const double *actions, double *obs, double *reward,
                     unsigned char *done, int *counts, int n_threads)
{
    bjob_t j; memset(&j, 0, sizeof(j));
    j.b = b; j.op = 1; j.actions = actions; j.obs = obs; j.reward = reward; j.done = done; j.counts = counts;
    brun(j, n_threads);
}
