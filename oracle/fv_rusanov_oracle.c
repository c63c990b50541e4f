/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.  See fv_rusanov_oracle.h.
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC.
 * -ffp-contract=off matters: the reference is built without FMA contraction
 * (Unit test/correctness_test.sbatch:24, plain g++ on x86-64), and the parity contract
 * is bit-exactness against that evaluation.
 */
#include "fv_rusanov_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FVO_MAX_DIM 3
#define FVO_MAX_NR 8

typedef struct {
  int dim, P, h, S, nr, na, nv;
  int ncell;              /* S^dim */
  int stride[FVO_MAX_DIM]; /* cell stride of axis m; axis 0 is slowest (CPPPrinter.py:247-261) */
} fvo_geom;

static int fvo_make_geom(const fvo_config* c, fvo_geom* g) {
  if (!c) return -1;
  if (c->dim != 2 && c->dim != 3) return -2;          /* KernelBuilder.py:41-48 viable() */
  if (c->patch_size < 1) return -3;
  if (c->halo < 1) return -4;                          /* the +-1 stencil needs one halo layer */
  if (c->n_real < 1 || c->n_real > FVO_MAX_NR || c->n_aux < 0) return -5;
  if (c->model == FVO_MODEL_EULER && c->n_real < c->dim + 2) return -6;
  if ((c->model == FVO_MODEL_SWE || c->model == FVO_MODEL_SWE_SOURCE) && (c->dim != 2 || c->n_real < 3)) return -6;
  if (c->model == FVO_MODEL_SWE_SOURCE && c->n_aux < 3) return -6;     /* b, db/dx, db/dy */
  if (c->model != FVO_MODEL_EULER && c->model != FVO_MODEL_SWE && c->model != FVO_MODEL_SWE_SOURCE) return -7;
  g->dim = c->dim; g->P = c->patch_size; g->h = c->halo; g->S = g->P + 2 * g->h;
  g->nr = c->n_real; g->na = c->n_aux; g->nv = g->nr + g->na;
  g->ncell = 1;
  for (int m = g->dim - 1; m >= 0; --m) { g->stride[m] = g->ncell; g->ncell *= g->S; }
  return 0;
}

uint64_t fvo_fnv1a64_words(const void* data, int64_t n_words) {
  const uint64_t* w = (const uint64_t*)data;
  uint64_t hsh = 1469598103934665603ULL;
  for (int64_t i = 0; i < n_words; ++i) { hsh ^= w[i]; hsh *= 1099511628211ULL; }
  return hsh;
}

int fvo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline double fvo_u01(uint64_t idx, uint64_t seed) {
  uint64_t z = (idx + seed) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

/* one cell of the synthetic admissible state, always evaluated in double */
static void fvo_synth_cell(const fvo_config* c, int nv, int64_t cell, uint64_t seed, double* out) {
  double u[16];
  for (int v = 0; v < nv && v < 16; ++v) u[v] = fvo_u01((uint64_t)cell * (uint64_t)nv + (uint64_t)v, seed);
  if (c->model == FVO_MODEL_EULER) {
    const int d = c->dim;
    const double rho = 1.0 + u[0];
    double ke = 0.0;
    for (int k = 0; k < d; ++k) { const double vel = u[1 + k] - 0.5; out[1 + k] = rho * vel; ke = ke + vel * vel; }
    const double p = 1.0 + u[d + 1];
    out[0] = rho;
    out[d + 1] = p / (1.4 - 1.0) + 0.5 * rho * ke;
    for (int v = d + 2; v < nv; ++v) out[v] = u[v];
  } else {
    const double hgt = 1.0 + u[0];
    out[0] = hgt;
    out[1] = hgt * (0.2 * (u[1] - 0.5));
    out[2] = hgt * (0.2 * (u[2] - 0.5));
    for (int v = 3; v < nv; ++v) out[v] = 0.1 * u[v];
  }
}


/* Cell lists, built once per call: the loop ranges of the generated kernel as index sets. */
typedef struct {
  int* sweep[FVO_MAX_DIM];   /* cells of the flux/eigen sweep of axis n, in loop order */
  int n_sweep[FVO_MAX_DIM];
  int* interior;             /* interior cells, in loop order */
  int n_interior;
} fvo_cells;

static int fvo_is_interior(const fvo_geom* g, int cell, int skip_axis) {
  for (int m = 0; m < g->dim; ++m) {
    if (m == skip_axis) continue;
    const int cm = (cell / g->stride[m]) % g->S;
    if (cm < g->h || cm >= g->P + g->h) return 0;
  }
  return 1;
}

/* HEAD: full along n, interior across (CPPPrinter.py:132-137); COMMITTED: transposed (test.cpp:22-23) */
static int fvo_in_sweep(const fvo_geom* g, int ranges, int cell, int n) {
  if (ranges == FVO_RANGES_HEAD) return fvo_is_interior(g, cell, n);
  const int cn = (cell / g->stride[n]) % g->S;
  return cn >= g->h && cn < g->P + g->h;
}

static void fvo_free_cells(fvo_cells* c) {
  for (int n = 0; n < FVO_MAX_DIM; ++n) free(c->sweep[n]);
  free(c->interior);
}

static int fvo_make_cells(const fvo_geom* g, int ranges, fvo_cells* c) {
  memset(c, 0, sizeof *c);
  c->interior = (int*)malloc(sizeof(int) * (size_t)g->ncell);
  if (!c->interior) return -1;
  for (int cell = 0; cell < g->ncell; ++cell)
    if (fvo_is_interior(g, cell, -1)) c->interior[c->n_interior++] = cell;
  for (int n = 0; n < g->dim; ++n) {
    c->sweep[n] = (int*)malloc(sizeof(int) * (size_t)g->ncell);
    if (!c->sweep[n]) { fvo_free_cells(c); return -1; }
    for (int cell = 0; cell < g->ncell; ++cell)
      if (fvo_in_sweep(g, ranges, cell, n)) c->sweep[n][c->n_sweep[n]++] = cell;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
#define FVO_DEFINE(T, SFX, SQRT, FABS)                                                          \
                                                                                                \
  /* Functions.cpp:9-37.  2-D: F[3] = energy flux, entries >= 4 are never written (the        \
   * committed kernel runs with n_real = 5).  3-D: the `#endif` at Functions.cpp:34 precedes  \
   * an unconditional F[3] overwrite; the corrected form is used.  */         \
  static void euler_flux_##SFX(int dim, const T* Q, int normal, T* F) {                        \
    const T GAMMA = (T)1.4;                                                                     \
    const T rho = Q[0], u = Q[1], v = Q[2];                                                     \
    const T irho = (T)1.0 / rho;                                                                \
    if (dim == 3) {                                                                             \
      const T w = Q[3], e = Q[4];                                                               \
      const T p = (GAMMA - 1) * (e - (T)0.5 * irho * (u * u + v * v + w * w));                  \
      const T coeff = irho * Q[normal + 1];                                                     \
      F[0] = coeff * rho; F[1] = coeff * u; F[2] = coeff * v; F[3] = coeff * w;                 \
      F[4] = coeff * e + coeff * p;                                                             \
      F[normal + 1] += p;                                                                       \
    } else {                                                                                    \
      const T e = Q[3];                                                                         \
      const T p = (GAMMA - 1) * (e - (T)0.5 * irho * (u * u + v * v));                          \
      const T coeff = irho * Q[normal + 1];                                                     \
      F[0] = coeff * rho; F[1] = coeff * u; F[2] = coeff * v;                                   \
      F[3] = coeff * e + coeff * p;                                                             \
      F[normal + 1] += p;                                                                       \
    }                                                                                           \
  }                                                                                             \
                                                                                                \
  /* Functions.cpp:39-62 */                                                                     \
  static T euler_eig_##SFX(int dim, const T* Q, int normal) {                                  \
    const T GAMMA = (T)1.4;                                                                     \
    const T rho = Q[0], u = Q[1], v = Q[2];                                                     \
    const T irho = (T)1.0 / FABS(rho);                                                          \
    T p;                                                                                        \
    if (dim == 3) {                                                                             \
      const T w = Q[3], e = Q[4];                                                               \
      p = (GAMMA - 1) * (e - (T)0.5 * irho * (u * u + v * v + w * w));                          \
    } else {                                                                                    \
      const T e = Q[3];                                                                         \
      p = (GAMMA - 1) * (e - (T)0.5 * irho * (u * u + v * v));                                  \
    }                                                                                           \
    const T c = SQRT(GAMMA * FABS(p) * irho);                                                   \
    const T u_n = Q[normal + 1] * irho;                                                         \
    const T a = FABS(u_n - c), b = FABS(u_n + c);                                               \
    return (a < b) ? b : a; /* std::max */                                                      \
  }                                                                                             \
                                                                                                \
  /* Shallow water, this repo's definition in the style of Functions.cpp:      \
   * q = (h, hu, hv | b), g = 9.81; bathymetry is aux and does not enter the flux. */          \
  static void swe_flux_##SFX(const T* Q, int normal, T* F) {                                   \
    const T G = (T)9.81;                                                                        \
    const T ih = (T)1.0 / Q[0];                                                                 \
    const T un = ih * Q[normal + 1];                                                            \
    F[0] = un * Q[0]; F[1] = un * Q[1]; F[2] = un * Q[2];                                       \
    F[normal + 1] += (T)0.5 * G * Q[0] * Q[0];                                                  \
  }                                                                                             \
  static T swe_eig_##SFX(const T* Q, int normal) {                                             \
    const T G = (T)9.81;                                                                        \
    const T ih = (T)1.0 / FABS(Q[0]);                                                           \
    const T un = Q[normal + 1] * ih;                                                            \
    const T c = SQRT(G * FABS(Q[0]));                                                           \
    const T a = FABS(un - c), b = FABS(un + c);                                                 \
    return (a < b) ? b : a;                                                                     \
  }                                                                                             \
                                                                                                \
  /* bathymetry source in the style of sourceTerm(Q, x, h, t, dt, S), correctness_test.cpp:16-23: aux = (b, bx, by) */ \
  static void swe_source_##SFX(const T* Q, int nr, T* S) {                                      \
    const T G = (T)9.81;                                                                        \
    const T gh = G * Q[0];                                                                      \
    for (int v = 0; v < nr; ++v) S[v] = (T)0;                                                   \
    S[1] = -gh * Q[nr + 1];                                                                     \
    S[2] = -gh * Q[nr + 2];                                                                     \
  }                                                                                             \
                                                                                                \
  static inline T maxp_##SFX(const T* a, const T* b) { /* Functions.cpp:64-66 */               \
    return (*a < *b) ? *b : *a;                                                                 \
  }                                                                                             \
                                                                                                \
  /* One patch.  Scratch: Qc[ncell*nv], F[dim][ncell*nr], L[dim][ncell]. */                     \
  static T patch_step_##SFX(const fvo_config* cfg, const fvo_geom* g, const fvo_cells* cl, T* Q, \
                            T dt, T* Qc, T* F, T* L) {                                          \
    const int nv = g->nv, nr = g->nr, nc = g->ncell, dim = g->dim;                              \
    /* test.cpp:11-19 : copy every haloed cell, every variable */                               \
    memcpy(Qc, Q, sizeof(T) * (size_t)nc * nv);                                                 \
    /* test.cpp:4-8 allocates these uninitialised; value-initialised here */    \
    memset(F, 0, sizeof(T) * (size_t)dim * nc * nr);                                            \
    memset(L, 0, sizeof(T) * (size_t)dim * nc);                                                 \
    /* test.cpp:20-39 */                                                                        \
    for (int n = 0; n < dim; ++n)                                                               \
      for (int k = 0; k < cl->n_sweep[n]; ++k) {                                                \
        const int c = cl->sweep[n][k];                                                          \
        T* f = F + ((size_t)n * nc + c) * nr;                                                   \
        if (cfg->model == FVO_MODEL_EULER) euler_flux_##SFX(dim, Qc + (size_t)c * nv, n, f);   \
        else swe_flux_##SFX(Qc + (size_t)c * nv, n, f);                                         \
      }                                                                                         \
    /* test.cpp:40-59 */                                                                        \
    for (int n = 0; n < dim; ++n)                                                               \
      for (int k = 0; k < cl->n_sweep[n]; ++k) {                                                \
        const int c = cl->sweep[n][k];                                                          \
        L[(size_t)n * nc + c] = (cfg->model == FVO_MODEL_EULER)                                 \
                                    ? euler_eig_##SFX(dim, Qc + (size_t)c * nv, n)              \
                                    : swe_eig_##SFX(Qc + (size_t)c * nv, n);                    \
      }                                                                                         \
    /* test.cpp:60-77 : central flux difference, axis by axis in order */                       \
    for (int n = 0; n < dim; ++n) {                                                             \
      const int e = g->stride[n];                                                               \
      const T* f = F + (size_t)n * nc * nr;                                                     \
      for (int k = 0; k < cl->n_interior; ++k) {                                                \
        const int c = cl->interior[k];                                                          \
        for (int v = 0; v < nr; ++v)                                                            \
          Qc[(size_t)c * nv + v] = Qc[(size_t)c * nv + v] - (T)0.5 * f[(size_t)(c + e) * nr + v] + \
                                   (T)0.5 * f[(size_t)(c - e) * nr + v];                        \
      }                                                                                         \
    }                                                                                           \
    /* test.cpp:78-95 : Rusanov dissipation from the ORIGINAL Q */                              \
    const int dv = (cfg->diss == FVO_DISS_ALL) ? nr : 1;                                        \
    for (int n = 0; n < dim; ++n) {                                                             \
      const int e = g->stride[n];                                                               \
      const T* l = L + (size_t)n * nc;                                                          \
      for (int k = 0; k < cl->n_interior; ++k) {                                                \
        const int c = cl->interior[k];                                                          \
          for (int v = 0; v < dv; ++v)                                                          \
            Qc[(size_t)c * nv + v] =                                                            \
                (T)0.5 * dt *                                                                   \
                    ((-Q[(size_t)(c + e) * nv + v] + Q[(size_t)c * nv + v]) *                   \
                         maxp_##SFX(&l[c + e], &l[c]) +                                         \
                     (Q[(size_t)(c - e) * nv + v] - Q[(size_t)c * nv + v]) *                    \
                         maxp_##SFX(&l[c - e], &l[c])) +                                        \
                Qc[(size_t)c * nv + v];                                                         \
      }                                                                                         \
    }                                                                                           \
    /* source term (8f-3): S of the ORIGINAL state, Q_copy = Q_copy + dt*S, interior cells */    \
    if (cfg->model == FVO_MODEL_SWE_SOURCE)                                                     \
      for (int k = 0; k < cl->n_interior; ++k) {                                                \
        const int c = cl->interior[k];                                                          \
        T S[FVO_MAX_NR];                                                                        \
        swe_source_##SFX(Q + (size_t)c * nv, nr, S);                                            \
        for (int v = 0; v < nr; ++v) Qc[(size_t)c * nv + v] = Qc[(size_t)c * nv + v] + dt * S[v]; \
      }                                                                                         \
    /* test.cpp:96-104 : interior copy-back, all variables; plus the patch's max eigenvalue     \
     * over interior cells of the INPUT state (not in the reference) */         \
    T lam = (T)0;                                                                               \
    for (int k = 0; k < cl->n_interior; ++k) {                                                  \
      {                                                                                         \
        const int c = cl->interior[k];                                                          \
        for (int v = 0; v < nv; ++v) Q[(size_t)c * nv + v] = Qc[(size_t)c * nv + v];            \
        for (int n = 0; n < dim; ++n) { /* interior cells are swept under both range rules */  \
          const T ln = L[(size_t)n * nc + c];                                                   \
          if (lam < ln) lam = ln;                                                               \
        }                                                                                       \
      }                                                                                         \
    }                                                                                           \
    return lam;                                                                                 \
  }                                                                                             \
                                                                                                \
  int fvo_step_##SFX(const fvo_config* cfg, T* Q, int64_t n_patches, T dt, T* lambda_patch,    \
                     T* lambda_max, int nthreads) {                                             \
    fvo_geom g;                                                                                 \
    const int rc = fvo_make_geom(cfg, &g);                                                      \
    if (rc) return rc;                                                                          \
    if (n_patches < 0 || (!Q && n_patches > 0)) return -8;                                      \
    fvo_cells cl;                                                                               \
    if (fvo_make_cells(&g, cfg->ranges, &cl)) return -9;                                        \
    const size_t per = (size_t)g.ncell * g.nv;                                                  \
    const size_t scratch = per + (size_t)g.dim * g.ncell * (g.nr + 1);                          \
    T gmax = (T)0;                                                                              \
    int fail = 0;                                                                               \
    if (nthreads < 1) nthreads = 1;                                                             \
    _Pragma("omp parallel num_threads(nthreads) if (nthreads > 1)")                             \
    {                                                                                           \
      T* buf = (T*)malloc(sizeof(T) * scratch);                                                 \
      T lmax = (T)0;                                                                            \
      if (!buf) {                                                                               \
        _Pragma("omp atomic write") fail = 1;                                                   \
      } else {                                                                                  \
        T* Qc = buf; T* F = Qc + per; T* L = F + (size_t)g.dim * g.ncell * g.nr;                \
        _Pragma("omp for schedule(static)")                                                     \
        for (int64_t b = 0; b < n_patches; ++b) {                                               \
          const T lam = patch_step_##SFX(cfg, &g, &cl, Q + (size_t)b * per, dt, Qc, F, L);           \
          if (lambda_patch) lambda_patch[b] = lam;                                              \
          if (lmax < lam) lmax = lam;                                                           \
        }                                                                                       \
        free(buf);                                                                              \
      }                                                                                         \
      _Pragma("omp critical") { if (gmax < lmax) gmax = lmax; }                                 \
    }                                                                                           \
    fvo_free_cells(&cl);                                                                        \
    if (fail) return -9;                                                                        \
    if (lambda_max) *lambda_max = gmax;                                                         \
    return 0;                                                                                   \
  }                                                                                             \
                                                                                                \
  void fvo_fill_sin_##SFX(T* Q, int64_t n) {                                                    \
    for (int64_t i = 0; i < n; ++i) Q[i] = (T)sin(3.141 * (double)i / (double)n);               \
  }                                                                                             \
                                                                                                \
  void fvo_fill_synthetic_##SFX(const fvo_config* cfg, T* Q, int64_t first_cell,               \
                                int64_t n_cells, uint64_t seed) {                               \
    const int nv = cfg->n_real + cfg->n_aux;                                                    \
    _Pragma("omp parallel for schedule(static)")                                                \
    for (int64_t c = 0; c < n_cells; ++c) {                                                     \
      double cell[16];                                                                          \
      fvo_synth_cell(cfg, nv, first_cell + c, seed, cell);                                      \
      for (int v = 0; v < nv; ++v) Q[(size_t)c * nv + v] = (T)cell[v];                          \
    }                                                                                           \
  }

FVO_DEFINE(double, f64, sqrt, fabs)
FVO_DEFINE(float, f32, sqrtf, fabsf)
