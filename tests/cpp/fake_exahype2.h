// Test infrastructure: the few ExaHyPE2 / Peano names that the C++ generated from examples/kernel-generator.py refers to,
// so that CPPPrinter's output for that declaration compiles and runs here and serves as the CPU side of the parity test
// of the CUDA back-end (the reference's own harness fakes Peano's solver classes the same way,
// "Unit test/correctness_test.cpp":14-100; Peano itself is not vendored).  Member layout of CellData follows ExaHyPE2's
// (one entry per patch in every member).  Build with -DDimensions=2|3 -ffp-contract=off.
//
// Volume geometry -- the same definitions, in the same evaluation order, as FvCellCtx in csrc/fv_patch_kernel.cuh:
//   getVolumeSize(h, n)[d]           = h[d] / n
//   getVolumeCentre(x, h, n, idx)[d] = (x[d] - 0.5 * h[d]) + (idx[d] + 0.5) * (h[d] / n)
//   getVolumeCentre(x, h, n)         = x                       (no index: the patch centre, kernel-generator.py:39)
#pragma once
#include <cmath>
#include <initializer_list>

namespace tarch {
namespace la {
template <int D, typename T>
struct Vector {
  T v[D];
  Vector() : v{} {}
  Vector(std::initializer_list<T> l) : v{} {
    int i = 0;
    for (T x : l) if (i < D) v[i++] = x;
  }
  T& operator()(int i) { return v[i]; }
  const T& operator()(int i) const { return v[i]; }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
}  // namespace la
namespace timing {
struct Measurement {};
}  // namespace timing
}  // namespace tarch

namespace exahype2 {
using Vec = tarch::la::Vector<Dimensions, double>;
struct CellData {
  double** QIn;
  Vec* cellCentre;
  Vec* cellSize;
  double* t;
  double* dt;
  double** QOut;
  double* maxEigenvalue;
  int numberOfCells;
};
namespace fv {
inline Vec getVolumeSize(const Vec& h, int n) {
  Vec r;
  for (int d = 0; d < Dimensions; ++d) r(d) = h(d) / double(n);
  return r;
}
inline Vec getVolumeCentre(const Vec& x, const Vec& h, int n, const tarch::la::Vector<Dimensions, int>& index) {
  Vec r;
  for (int d = 0; d < Dimensions; ++d) r(d) = (x(d) - 0.5 * h(d)) + (double(index(d)) + 0.5) * (h(d) / double(n));
  return r;
}
inline Vec getVolumeCentre(const Vec& x, const Vec&, int) { return x; }
}  // namespace fv
}  // namespace exahype2

// The solver instance the declaration names.  Position- and time-dependent on purpose: every argument of the ExaHyPE2
// signature has to reach the functor for the parity test to pass.  The device body of the test (tests/test_cell_data_*.py)
// states the same formulas in the same order.  The declaration calls `flux` for the eigenvalue too
// (kernel-generator.py:39), hence the 6-argument overload.
namespace benchmarks { namespace exahype2 { namespace kernelbenchmarks { namespace repositories {
struct FVRusanovSolver {
  void flux(const double* Q, const ::exahype2::Vec& x, const ::exahype2::Vec& h, double t, double dt, int normal,
            double* F) const {
    const double irho = 1.0 / Q[0];
#if Dimensions == 2
    const double p = (1.4 - 1) * (Q[3] - 0.5 * irho * (Q[1] * Q[1] + Q[2] * Q[2]));
#else
    const double p = (1.4 - 1) * (Q[4] - 0.5 * irho * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]));
#endif
    const double coeff = irho * Q[normal + 1];
    const double w = 1.0 + 0.01 * x(normal) + 0.1 * h(0) + 0.001 * t + 0.5 * dt;
    for (int v = 0; v <= Dimensions; ++v) F[v] = coeff * Q[v] * w;
    F[Dimensions + 1] = (coeff * Q[Dimensions + 1] + coeff * p) * w;
    F[normal + 1] += p;
  }
  double flux(const double* Q, const ::exahype2::Vec& x, const ::exahype2::Vec& h, double t, double dt, int normal) const {
    const double irho = 1.0 / Q[0];
    return std::fabs(Q[normal + 1] * irho) + 0.01 * x(0) + h(1) + t + dt;
  }
};
static const FVRusanovSolver instanceOfFVRusanovSolver;
}}}}  // namespace benchmarks::exahype2::kernelbenchmarks::repositories

inline double max(double* a, double* b) { return (*a < *b) ? *b : *a; }   // Functions.cpp:64-66
