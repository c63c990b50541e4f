// The reference's parity harness, revived against the B200 drop-in.
//
// Reference: "Unit test/correctness_test.cpp":176-221 -- sin-initialise two copies of one 4x4+1 patch with 5+5
// variables (:102-106), run the DSL kernel `time_step(Q1, 1)` on one (:195) and a comparator on the other (:196, whose
// body is commented out in the reference), compare with `!=` (:199-204).  Here
//   time_step      = the drop-in of include/exahype_cuda.h (host call shape, runs on the GPU), and
//   old_time_step  = the reference's OWN generated kernel compiled from its own sources (oracle/_ref/libexahype_ref.so,
//                    loaded with dlopen when present) and the CPU oracle (oracle/libfv_oracle.so).
// The reference's committed kernel reads rows of its temporaries that it never writes, so against it
// only the cells those rows cannot reach -- (2..3, 2..3) -- are compared; against the oracle (corrected ranges) all
// 360 values are.  Exit code = number of differing values.
//
//   g++ -std=c++17 -Iinclude -Ioracle tests/cpp/correctness_test_b200.cpp -Lexahype_b200 -lexahype_cuda -ldl
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>

#include "exahype_cuda.h"
#include "fv_rusanov_oracle.h"

namespace {
constexpr int kDim = 2, kPatch = 4, kHalo = 1, kReal = 5, kAux = 5;

// drop-in for the reference's generated `void time_step(double* Q, double dt)` ("Unit test/test.h":3)
void time_step(double* Q, double dt) {
  exahype_fv_config cfg = {EXAHYPE_MODEL_EULER, EXAHYPE_DTYPE_F64, kDim, kPatch, kHalo, kReal, kAux, 0u};
  if (exahype_cuda_time_step_host(&cfg, Q, Q, 1, dt, nullptr, nullptr) != EXAHYPE_OK) {
    std::fprintf(stderr, "time_step: %s\n", exahype_cuda_last_error());
    std::exit(100);
  }
}

void init_input(std::vector<double>& Q) {           // correctness_test.cpp:102-106
  const int n = (int)Q.size();
  for (int i = 0; i < n; ++i) Q[i] = std::sin(3.141 * i / n);
}

void show(const std::vector<double>& Q, int stride_a, int stride_b) {   // first variable of every cell, like :108-116
  for (int i = 0; i < (int)Q.size(); i += stride_a) {
    if (i % (stride_a * stride_b) == 0) std::printf("\n");
    std::printf("%.2f\t", std::ceil(Q[i] * 100.0) / 100.0);
  }
  std::printf("\n");
}
}  // namespace

int main(int argc, char** argv) {
  const int S = kPatch + 2 * kHalo, nv = kReal + kAux;
  const int n = nv * S * S;
  std::vector<double> Q1(n), Q2(n), Q3(n);
  init_input(Q1); init_input(Q2); init_input(Q3);

  time_step(Q1.data(), 1.0);                                            // the B200 path

  int bad = 0;
  // comparator 1: the CPU oracle with the corrected loop ranges -- every value
  typedef int (*step_fn)(const fvo_config*, double*, int64_t, double, double*, double*, int);
  const char* oracle_path = argc > 1 ? argv[1] : "oracle/libfv_oracle.so";
  void* ho = dlopen(oracle_path, RTLD_NOW);
  if (!ho) { std::fprintf(stderr, "cannot load %s: %s\n", oracle_path, dlerror()); return 101; }
  step_fn oracle_step = (step_fn)dlsym(ho, "fvo_step_f64");
  fvo_config ocfg = {kDim, kPatch, kHalo, kReal, kAux, FVO_MODEL_EULER, FVO_RANGES_HEAD, FVO_DISS_VAR0};
  oracle_step(&ocfg, Q2.data(), 1, 1.0, nullptr, nullptr, 1);
  int bad_oracle = 0;
  for (int i = 0; i < n; ++i) bad_oracle += (Q1[i] != Q2[i]);
  std::printf("vs CPU oracle (all %d values): %d differences\n", n, bad_oracle);
  bad += bad_oracle;

  // comparator 2: the reference's own compiled kernel -- the cells its uninitialised rows cannot reach
  const char* ref_path = argc > 2 ? argv[2] : "oracle/_ref/libexahype_ref.so";
  if (void* hr = dlopen(ref_path, RTLD_NOW)) {
    typedef void (*ref_fn)(double*, double);
    ref_fn old_time_step = (ref_fn)dlsym(hr, "ref_time_step");
    old_time_step(Q3.data(), 1.0);
    int bad_ref = 0, compared = 0;
    for (int i = 2; i <= 3; ++i)
      for (int j = 2; j <= 3; ++j)
        for (int v = 0; v < nv; ++v, ++compared) bad_ref += (Q1[(i * S + j) * nv + v] != Q3[(i * S + j) * nv + v]);
    std::printf("vs the reference's compiled kernel (%d values of the inner cells): %d differences\n", compared, bad_ref);
    bad += bad_ref;
  } else {
    std::printf("reference kernel library not present (%s): comparator skipped\n", ref_path);
  }

  if (bad > 0) std::printf("there are %d differences between the outputs\n", bad);
  else std::printf("no differences! :)\n");
  show(Q1, nv, S);
  show(Q2, nv, S);
  return bad > 255 ? 255 : bad;
}
