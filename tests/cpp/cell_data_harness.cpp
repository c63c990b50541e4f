// Test infrastructure: runs the C++ that CPPPrinter generates from examples/kernel-generator.py (included as
// GENERATED_KERNEL) over a batch, one single-patch CellData per call -- the declaration has n_patches = 1.
//   q        n haloed patches, in/out (the declaration's `QOut`: copied in, interior copied back)
//   scratch  one haloed patch (the declaration's `QIn`, the working copy)
#ifndef FAKE_HEADER
#define FAKE_HEADER "fake_exahype2.h"
#endif
#include FAKE_HEADER

#include GENERATED_KERNEL

extern "C" void run_cell_data(int n, long long patch_elems, double* q, double* scratch, const double* centre,
                              const double* size, const double* t, const double* dt) {
  for (int p = 0; p < n; ++p) {
    double* qin = scratch;
    double* qout = q + p * patch_elems;
    exahype2::Vec c, s;
    for (int d = 0; d < Dimensions; ++d) { c(d) = centre[p * Dimensions + d]; s(d) = size[p * Dimensions + d]; }
    double tt = t[p], dd = dt[p], lam = 0.0;
    exahype2::CellData cd{&qin, &c, &s, &tt, &dd, &qout, &lam, 1};
    tarch::timing::Measurement m;
    time_step(cd, m);
  }
}
