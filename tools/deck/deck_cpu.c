/* tools/deck/deck_cpu.c -- CPU build of the synthetic deck generator (test / bench support). */
#include "deck_gen.h"

#include <pthread.h>
#include <stdlib.h>

typedef struct {
  uint64_t seed;
  uint32_t first;
  int lo, hi, W, H;
  double jitter;
  uint8_t *out;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  int k, x, y;
  for (k = j->lo; k < j->hi; k++) {
    deck_params p;
    uint8_t *dst = j->out + (size_t)k * j->W * j->H;
    deck_frame_params(j->seed, j->first + (uint32_t)k, j->W, j->H, j->jitter, &p);
    for (y = 0; y < j->H; y++)
      for (x = 0; x < j->W; x++) dst[(size_t)y * j->W + x] = deck_pixel(&p, x, y);
  }
  return NULL;
}

/* Render frames [first, first+n) of deck `seed` into out (n dense W*H planes). */
void deck_render_cpu(uint64_t seed, uint32_t first, int n, int W, int H, double jitter, uint8_t *out, int nthreads) {
  pthread_t th[64];
  job_t jobs[64];
  int t;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  for (t = 0; t < nthreads; t++) {
    job_t j = {seed, first, (int)((long)n * t / nthreads), (int)((long)n * (t + 1) / nthreads), W, H, jitter, out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* Ground truth of one frame: digits[16], n_digits, quad[8] (tl,tr,bl,br). */
void deck_truth(uint64_t seed, uint32_t frame, int W, int H, double jitter, uint8_t *digits, int32_t *n_digits, double *quad) {
  deck_params p;
  int i;
  deck_frame_params(seed, frame, W, H, jitter, &p);
  for (i = 0; i < 16; i++) digits[i] = p.digits[i];
  *n_digits = p.n_digits;
  for (i = 0; i < 8; i++) quad[i] = p.quad[i];
}
