// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// Thin C-ABI wrapper linked with the reference's OWN, UNMODIFIED sources
//   /root/reference/Unit test/test.cpp       (generated kernel, time_step)
//   /root/reference/Unit test/Functions.cpp  (Flux / maxEigenvalue / max)
// compiled from where they lie by oracle/build_ref.sh into oracle/_ref/libexahype_ref.so.
// No reference source is copied into this repository.
//
// The committed kernel reads temporaries it never wrote (test.cpp:22,42 vs :64,82).
// To make its output a deterministic golden without touching the file, this translation unit
// replaces the array forms of operator new/delete so that `new double[N]` is value-initialised;
// the library is linked -Bsymbolic so the replacement binds inside the library only.
#include <cstdlib>
#include <new>

void* operator new[](std::size_t n) {
  void* p = std::calloc(n ? n : 1, 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

// Unit test/test.h:3 and Unit test/Functions.h:2-4
void time_step(double* Q, double dt);
void Flux(const double* __restrict__ Q, int normal, double* __restrict__ F);
double maxEigenvalue(const double* __restrict__ Q, int normal);
double max(double* a, double* b);

extern "C" {
// fixed configuration of the committed kernel: dim 2, patch 4, halo 1, 5 + 5 variables, 1 patch
void ref_time_step(double* Q, double dt) { time_step(Q, dt); }
void ref_flux(const double* Q, int normal, double* F) { Flux(Q, normal, F); }
double ref_max_eigenvalue(const double* Q, int normal) { return maxEigenvalue(Q, normal); }
double ref_max(double* a, double* b) { return max(a, b); }
int ref_n_values(void) { return 360; }
// `reps` calls of the committed kernel on a fresh copy of Q0 each (360 doubles; the copy is part of what a caller of
// time_step(Q, dt) does per patch anyway): what bench.py times as cpu_baseline.reference_compiled
void ref_time_step_repeat(const double* Q0, double* Q, double dt, long reps) {
  for (long r = 0; r < reps; ++r) {
    for (int i = 0; i < 360; ++i) Q[i] = Q0[i];
    time_step(Q, dt);
  }
}
}
