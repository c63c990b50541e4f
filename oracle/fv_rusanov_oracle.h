/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
 *
 * CPU oracle for ExaHyPE's batched stateless finite-volume Rusanov patch update.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the CUDA product path never calls into it.
 *
 * It restates, parametrised over (dim, patch_size, halo, n_real, n_aux, n_patches),
 * the statement sequence of the reference's generated kernel
 *   /root/reference/Unit test/test.cpp:11-104
 * with the user physics of
 *   /root/reference/Unit test/Functions.cpp:9-66
 * and the loop-range rule of
 *   /root/reference/exahype/printers/CPPPrinter.py:116-137.
 *
 * Parity status: PINNED.  `ranges = FVO_RANGES_COMMITTED` reproduces bit-for-bit the
 * output of the reference's own committed sources compiled from where they lie
 * (oracle/build_ref.sh -> oracle/_ref/libexahype_ref.so) and the golden hash G0 of
 * SURVEY.md section 8c; `FVO_RANGES_HEAD` reproduces goldens G0h/G1/G3s/G3r/G2r.
 */
#ifndef FV_RUSANOV_ORACLE_H
#define FV_RUSANOV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SWE_SOURCE: shallow water with the bathymetry source term (no reference counterpart beyond the
 * signature sourceTerm(Q, x, h, t, dt, S) of "Unit test/correctness_test.cpp":16-23): q = (h, hu, hv | b, db/dx, db/dy),
 * S = (0, -g h db/dx, -g h db/dy) evaluated on the ORIGINAL state; after the dissipation statements, interior cells,
 * v < n_real:  Q_copy = Q_copy + dt*S. */
enum { FVO_MODEL_EULER = 0, FVO_MODEL_SWE = 1, FVO_MODEL_SWE_SOURCE = 2 };
/* HEAD: flux/eigen on cells {full along n, interior across} (CPPPrinter.py:132-137).
 * COMMITTED: the transposed ranges found in Unit test/test.cpp:22-23,32-33. */
enum { FVO_RANGES_HEAD = 0, FVO_RANGES_COMMITTED = 1 };
/* VAR0: dissipation reaches variable 0 only, as emitted (test.cpp:81,90).
 * ALL : dissipation on all n_real unknowns (what struct=True intended). */
enum { FVO_DISS_VAR0 = 0, FVO_DISS_ALL = 1 };

typedef struct {
  int dim;        /* 2 or 3 */
  int patch_size; /* P >= 1 */
  int halo;       /* h >= 1 */
  int n_real;
  int n_aux;
  int model;      /* FVO_MODEL_* */
  int ranges;     /* FVO_RANGES_* */
  int diss;       /* FVO_DISS_* */
} fvo_config;

/* In-place update of Q[n_patches][S]^dim[n_real+n_aux] (AoS, S = P+2h).
 * lambda_patch (nullable) receives n_patches values, lambda_max (nullable) one.
 * nthreads <= 1 runs the serial loop the reference runs; > 1 uses OpenMP over patches.
 * Returns 0, or a negative code for an invalid configuration. */
int fvo_step_f64(const fvo_config* cfg, double* Q, int64_t n_patches, double dt,
                 double* lambda_patch, double* lambda_max, int nthreads);
int fvo_step_f32(const fvo_config* cfg, float* Q, int64_t n_patches, float dt,
                 float* lambda_patch, float* lambda_max, int nthreads);

/* Q[i] = sin(3.141*i/n)  (correctness_test.cpp:102-106) */
void fvo_fill_sin_f64(double* Q, int64_t n);
void fvo_fill_sin_f32(float* Q, int64_t n);

/* Counter-based admissible synthetic state: slot index = first_cell*nv + ...,
 * so every shard generates identical bits without communication. */
void fvo_fill_synthetic_f64(const fvo_config* cfg, double* Q, int64_t first_cell, int64_t n_cells,
                            uint64_t seed);
void fvo_fill_synthetic_f32(const fvo_config* cfg, float* Q, int64_t first_cell, int64_t n_cells,
                            uint64_t seed);

/* FNV-1a-64 over 8-byte words (xor then multiply). */
uint64_t fvo_fnv1a64_words(const void* data, int64_t n_words);

int fvo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
