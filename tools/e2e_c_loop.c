/* tools/e2e_c_loop.c -- C-level latency of armsim_step_host (no Python in the loop): where the end-to-end time goes.
 *   gcc -O2 -Iinclude tools/e2e_c_loop.c -o gpurun_out/e2e_c_loop -ldl && gpurun_out/e2e_c_loop drl-on-robot-arm_b200/libarmsim.so 4096 2000 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "armsim.h"

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "drl-on-robot-arm_b200/libarmsim.so";
  int n = argc > 2 ? atoi(argv[2]) : 4096, steps = argc > 3 ? atoi(argv[3]) : 2000;
  void* so = dlopen(path, RTLD_NOW);
  if (!so) { fprintf(stderr, "%s\n", dlerror()); return 1; }
  int (*defcfg)(int32_t, ArmsimConfig*) = dlsym(so, "armsim_default_config");
  int (*create)(const ArmsimConfig*, ArmSim**) = dlsym(so, "armsim_create");
  int (*hostbuf)(ArmSim*, float**, float**, float**, uint8_t**, uint8_t**) = dlsym(so, "armsim_host_buffers");
  int (*step)(ArmSim*, const float*, float*, float*, uint8_t*, uint8_t*) = dlsym(so, "armsim_step_host");
  void (*destroy)(ArmSim*) = dlsym(so, "armsim_destroy");
  const char* (*lasterr)(void) = dlsym(so, "armsim_last_error");
  ArmsimConfig cfg;
  defcfg(ARMSIM_TASK_REACH, &cfg);
  cfg.n_envs = n; cfg.auto_reset = 1;
  ArmSim* sim = NULL;
  if (create(&cfg, &sim)) { fprintf(stderr, "create: %s\n", lasterr()); return 1; }
  float *a, *o, *r; uint8_t *d, *s;
  hostbuf(sim, &a, &o, &r, &d, &s);
  for (int i = 0; i < n * 3; ++i) a[i] = 1.4f * (float)rand() / RAND_MAX - 0.7f;
  for (int k = 0; k < 20; ++k) step(sim, a, o, r, d, s);
  double t0 = now(), acc = 0;
  for (int k = 0; k < steps; ++k) { step(sim, a, o, r, d, s); acc += r[0]; }
  double dt = now() - t0;
  printf("{\"n_envs\": %d, \"steps\": %d, \"us_per_step\": %.3f, \"env_steps_per_s\": %.4g, \"check\": %.3f}\n", n, steps,
         1e6 * dt / steps, n * (double)steps / dt, acc);
  /* with host "think time" between steps (a policy running on the CPU): time spent INSIDE the step call only */
  for (int think_us = 5; think_us <= 80; think_us *= 2) {
    double in_call = 0;
    for (int k = 0; k < steps; ++k) {
      double t1 = now();
      step(sim, a, o, r, d, s);
      double t2 = now();
      in_call += t2 - t1;
      while (now() - t2 < 1e-6 * think_us) { }
    }
    printf("{\"think_us\": %d, \"us_in_step_call\": %.3f}\n", think_us, 1e6 * in_call / steps);
  }
  /* staged variant: ordinary (pageable) buffers */
  float* a2 = malloc(n * 3 * 4); float* o2 = malloc(n * 6 * 4); float* r2 = malloc(n * 4); uint8_t* d2 = malloc(n); uint8_t* s2 = malloc(n);
  memcpy(a2, a, n * 3 * 4);
  t0 = now();
  for (int k = 0; k < steps; ++k) step(sim, a2, o2, r2, d2, s2);
  dt = now() - t0;
  printf("{\"variant\": \"pageable buffers (memcpy staging)\", \"us_per_step\": %.3f}\n", 1e6 * dt / steps);
  destroy(sim);
  return 0;
}
