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
                      double *obs_out, double *reward_sum, long *count_sum)
{
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        job_t j = {t, n_threads, n_envs, n_steps, history_len, n_features, feature_ids,
                   bw, lat, loss, start_rate, queue, seeds, actions, obs_out, reward_sum, count_sum};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
