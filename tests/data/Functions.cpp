// Compressible Euler (gamma = 1.4) behind the reference's user-function interface, for tests of CPPPrinter output.
#include "Functions.h"
#include <cmath>
#ifndef DIMENSIONS
#define DIMENSIONS 2
#endif
namespace { constexpr double GAMMA = 1.4; }

static double pressure(const double* Q, double irho) {
  double ke = Q[1] * Q[1] + Q[2] * Q[2];
#if DIMENSIONS == 3
  ke = ke + Q[3] * Q[3];
#endif
  return (GAMMA - 1) * (Q[DIMENSIONS + 1] - 0.5 * irho * ke);
}

void Flux(const double* __restrict__ Q, int normal, double* __restrict__ F) {
  const double irho = 1.0 / Q[0];
  const double p = pressure(Q, irho);
  const double coeff = irho * Q[normal + 1];
  for (int v = 0; v <= DIMENSIONS; v++) F[v] = coeff * Q[v];
  F[DIMENSIONS + 1] = coeff * Q[DIMENSIONS + 1] + coeff * p;
  F[normal + 1] += p;
}

double maxEigenvalue(const double* __restrict__ Q, int normal) {
  const double irho = 1.0 / std::fabs(Q[0]);
  const double p = pressure(Q, irho);
  const double c = std::sqrt(GAMMA * std::fabs(p) * irho);
  const double u_n = Q[normal + 1] * irho;
  const double a = std::fabs(u_n - c), b = std::fabs(u_n + c);
  return a < b ? b : a;
}

double max(double* a, double* b) { return *a < *b ? *b : *a; }
