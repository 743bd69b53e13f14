/* oracle/oracle_batch.c -- TEST INFRASTRUCTURE: multi-threaded driver that steps many oracle worlds for the CPU
 * baseline timing in bench.py (cpu_baseline / --impl reference).  One world per env, pthreads over env ranges. */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <unistd.h>
typedef struct DgoWorld DgoWorld;
void dgo_env_step(DgoWorld* W, const double* act, double* obs, double* rew, uint8_t* term);
void dgo_env_reset(DgoWorld* W);

typedef struct {
  DgoWorld** worlds; int e0, e1, nworlds, nsteps, n_act, n_obs, n_rew, n_term;
  const double* act; double *obs, *rew; uint8_t* term; int reset_only;
} Job;

static void* run(void* arg) {
  Job* j = (Job*)arg;
  for (int e = j->e0; e < j->e1; e++) {
    if (j->reset_only) { dgo_env_reset(j->worlds[e]); continue; }
    for (int s = 0; s < j->nsteps; s++)
      dgo_env_step(j->worlds[e], j->act + ((size_t)s * j->nworlds + e) * j->n_act, j->obs + (size_t)e * j->n_obs,
                   j->rew + (size_t)e * j->n_rew, j->term + (size_t)e * j->n_term);
  }
  return NULL;
}
int dgo_batch_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void launch(Job base, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > base.nworlds) nthreads = base.nworlds > 0 ? base.nworlds : 1;
  pthread_t th[256]; Job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = base; jobs[t].e0 = (int)((long)base.nworlds * t / nthreads); jobs[t].e1 = (int)((long)base.nworlds * (t + 1) / nthreads);
    pthread_create(&th[t], NULL, run, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}
/* steps every world `nsteps` times; actions [nsteps][nworlds][n_act]; outputs hold the last step */
void dgo_batch_step(DgoWorld** worlds, int nworlds, int nsteps, const double* act, int n_act, double* obs, int n_obs, double* rew,
                    int n_rew, uint8_t* term, int n_term, int nthreads) {
  Job j = {worlds, 0, 0, nworlds, nsteps, n_act, n_obs, n_rew, n_term, act, obs, rew, term, 0};
  launch(j, nthreads);
}
void dgo_batch_reset(DgoWorld** worlds, int nworlds, int nthreads) {
  Job j = {worlds, 0, 0, nworlds, 0, 0, 0, 0, 0, NULL, NULL, NULL, NULL, 1};
  launch(j, nthreads);
}
