// cube_host_check.cpp -- TEST INFRASTRUCTURE.  The device cube / contact model (csrc/cube_model.cuh, fp32, look-ahead
// Gauss-Seidel) compiled FOR THE HOST by g++ from the same source, against the oracle's plain restatement
// (oracle/cube_model.h, fp64), one sim step at a time from identical (f32-rounded) states.  Runs without a GPU:
//   g++ -O2 -I drl-on-robot-arm_b200/csrc -I oracle tests/host/cube_host_check.cpp -o /tmp/cube_host_check -lm
// Prints "<steps> <worst pos err> <worst vel err> <outliers> <touching> <squeezed>"; exit code 0 = within tolerance.
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "cube_model.cuh"
extern "C" {
#include "cube_model.h"
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (double)(rng_state >> 11) * (1.0 / 9007199254740992.0);
}
static double uni(double a, double b) { return a + (b - a) * urand(); }

template <bool PICK>
static int run(int episodes, int steps, double* worst_pos, double* worst_vel, long* n_steps, long* n_out, long* n_touch, long* n_squeezed) {
  for (int ep = 0; ep < episodes; ++ep) {
    CubeState oc;
    cube_init(&oc, uni(0.3, 0.7), uni(-0.2, 0.2), PICK && ep % 4 == 3 ? uni(0.0, 0.1) : 0.01 - 0.015 * (ep % 3 == 0), 1.57 + 3.1415925438 * urand());
    if (ep % 5 == 4) {   // tumbling start
      for (int i = 0; i < 3; ++i) { oc.v[i] = uni(-0.5, 0.5); oc.w[i] = uni(-10, 10); }
    }
    // tool pointing down with a small tilt; EE starts near the cube (above / beside / far)
    double tilt = uni(-0.05, 0.05), yaw = uni(-3.14, 3.14);
    double Ree[9] = {cos(yaw), sin(yaw), tilt, sin(yaw), -cos(yaw), -tilt, tilt, tilt, -1.0};
    double ee[3] = {oc.pos[0] + uni(-0.09, 0.09), oc.pos[1] + uni(-0.09, 0.09), 0.0};
    const int mode = ep % 4;
    if (PICK) ee[2] = oc.pos[2] + 0.257 + uni(-0.02, 0.03);
    else ee[2] = mode == 0 ? uni(0.0, 0.03) : uni(0.0, 0.1);
    if (mode == 1) { ee[0] = oc.pos[0] + uni(-0.01, 0.01); ee[1] = oc.pos[1] + uni(-0.01, 0.01); }   // press down on it
    double grip = PICK && (ep & 8) ? 1.0 : 0.0;
    double vel[3] = {uni(-0.006, 0.006), uni(-0.006, 0.006), uni(-0.004, 0.002)};
    for (int k = 0; k < steps; ++k) {
      // teacher forcing on f32-rounded state
      cube::State dc;
      for (int i = 0; i < 3; ++i) { dc.pos[i] = (float)oc.pos[i]; oc.pos[i] = dc.pos[i]; dc.v[i] = (float)oc.v[i]; oc.v[i] = dc.v[i]; dc.w[i] = (float)oc.w[i]; oc.w[i] = dc.w[i]; }
      for (int i = 0; i < 4; ++i) { dc.quat[i] = (float)oc.quat[i]; oc.quat[i] = dc.quat[i]; }
      float eef[3], Rf[9];
      double eed[3], Rd[9];
      for (int i = 0; i < 3; ++i) { eef[i] = (float)ee[i]; eed[i] = eef[i]; }
      for (int i = 0; i < 9; ++i) { Rf[i] = (float)Ree[i]; Rd[i] = Rf[i]; }
      const double v_before = fabs(oc.v[0]) + fabs(oc.v[1]);
      const int sweeps = cube_step(&oc, eed, Rd, PICK ? 1 : 0, grip);
      cube::step<PICK>(dc, eef, Rf, (float)grip);
      double pe = 0, ve = 0;
      for (int i = 0; i < 3; ++i) { pe = fmax(pe, fabs(dc.pos[i] - oc.pos[i])); ve = fmax(ve, fabs(dc.v[i] - oc.v[i])); }
      for (int i = 0; i < 4; ++i) pe = fmax(pe, 0.02 * fabs(dc.quat[i] - oc.quat[i]));
      *n_steps += 1;
      if (pe > 5e-5 || ve > 5e-3) *n_out += 1;
      else { *worst_pos = fmax(*worst_pos, pe); *worst_vel = fmax(*worst_vel, ve); }
      if (fabs(fabs(oc.v[0]) + fabs(oc.v[1]) - v_before) > 0.02) *n_touch += 1;
      if (sweeps == CUBE_PGS_ITERS) *n_squeezed += 1;
      if (!(pe < 1.0)) { fprintf(stderr, "diverged ep %d step %d pe %g\n", ep, k, pe); return 2; }
      for (int i = 0; i < 3; ++i) ee[i] += vel[i];
      if (ee[2] < 0.0) ee[2] = 0.0;
      if (PICK && ee[2] < 0.23) ee[2] = 0.23;
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  const int episodes = argc > 1 ? atoi(argv[1]) : 400;
  for (int pick = 0; pick < 2; ++pick) {
    double wp = 0, wv = 0;
    long n = 0, out = 0, touch = 0, sq = 0;
    const int rc = pick ? run<true>(episodes, 40, &wp, &wv, &n, &out, &touch, &sq) : run<false>(episodes, 40, &wp, &wv, &n, &out, &touch, &sq);
    printf("%s steps %ld worst_pos %.3g worst_vel %.3g outliers %ld touching %ld squeezed %ld\n", pick ? "pick" : "push", n, wp, wv, out, touch, sq);
    if (rc) return rc;
    if (out > n / 200 || touch < n / 100) return 1;
  }
  return 0;
}
