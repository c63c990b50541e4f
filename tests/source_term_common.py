"""Shared by the source-term tests.  The reference has no kernel statement for a source term --
only the solver signature sourceTerm(Q, x, h, t, dt, S) ("Unit test/correctness_test.cpp":16-23) -- so the declaration
below is this repository's: two statements behind the dissipation, written with the reference's own DSL,

    kernel.single(sourceTerm(Q[0], tmp_source[0]))                        # S of the ORIGINAL state
    kernel.single(Q_copy[0], Q_copy[0] + dt*tmp_source[0], struct=True)   # Q_copy += dt*S on the unknowns

The model: shallow water with bathymetry slopes as auxiliary variables, q = (h, hu, hv | b, db/dx, db/dy),
S = (0, -g h db/dx, -g h db/dy)  (oracle: FVO_MODEL_SWE_SOURCE; device: SweSourcePhysics in csrc/physics.cuh)."""

# the user's functions as host C++ (the reference's Functions.h style) ...
HOST_FUNCTIONS = """
#include <cmath>
void Flux(const double* __restrict__ Q, int normal, double* __restrict__ F) {
  const double ih = 1.0 / Q[0];
  const double un = ih * Q[normal + 1];
  F[0] = un * Q[0]; F[1] = un * Q[1]; F[2] = un * Q[2];
  F[normal + 1] += 0.5 * 9.81 * Q[0] * Q[0];
}
double maxEigenvalue(const double* __restrict__ Q, int normal) {
  const double ih = 1.0 / std::fabs(Q[0]);
  const double un = Q[normal + 1] * ih;
  const double c = std::sqrt(9.81 * std::fabs(Q[0]));
  return std::fmax(std::fabs(un - c), std::fabs(un + c));
}
double max(double* a, double* b) { return (*a < *b) ? *b : *a; }
void sourceTerm(const double* __restrict__ Q, double* __restrict__ S) {
  const double gh = 9.81 * Q[0];
  S[0] = 0.0; S[1] = -gh * Q[4]; S[2] = -gh * Q[5];
}
"""
HOST_HEADER = """
void Flux(const double* __restrict__ Q, int normal, double* __restrict__ F);
double maxEigenvalue(const double* __restrict__ Q, int normal);
double max(double* a, double* b);
void sourceTerm(const double* __restrict__ Q, double* __restrict__ S);
"""
# ... and as device source, same formulas in the same order
DEVICE_FUNCTIONS = """
template <class T> __device__ void Flux(const T* Q, int normal, T* F) {
  const T ih = T(1.0) / Q[0];
  const T un = ih * Q[normal + 1];
  F[0] = un * Q[0]; F[1] = un * Q[1]; F[2] = un * Q[2];
  F[normal + 1] += T(0.5) * T(9.81) * Q[0] * Q[0];
}
template <class T> __device__ T maxEigenvalue(const T* Q, int normal) {
  const T ih = T(1.0) / fabs(Q[0]);
  const T un = Q[normal + 1] * ih;
  const T c = sqrt(T(9.81) * fabs(Q[0]));
  return fmax(fabs(un - c), fabs(un + c));
}
template <class T> __device__ void sourceTerm(const T* Q, T* S) {
  const T gh = T(9.81) * Q[0];
  S[0] = T(0); S[1] = -gh * Q[4]; S[2] = -gh * Q[5];
}
"""


def declare(patch_size=8, source=True, device_source=False):
    from exahype import KernelBuilder
    k = KernelBuilder(dim=2, patch_size=patch_size, halo_size=1, n_real=3, n_aux=3)
    Q, Qc = k.item('Q'), k.item('Q_copy')
    F, L = k.directional_item('tmp_flux'), k.directional_item('tmp_eigen', struct=False)
    S = k.item('tmp_source')
    dt = k.const('dt')
    normal = k.directional_const('normal', [0, 1])
    body = DEVICE_FUNCTIONS if device_source else None
    Flux, Eig = k.function('Flux', body=body), k.function('maxEigenvalue', body=body)
    Max, Src = k.function('max'), k.function('sourceTerm', body=body)
    k.single(Qc[0], Q[0])
    k.directional(Flux(Qc[0], normal, F[0]))
    k.directional(L[0], Eig(Qc[0], normal))
    k.directional(Qc[0], Qc[0] + 0.5 * (F[-1] - F[1]))
    left = -Max(L[-1], L[0]) * (Q[0] - Q[-1])
    right = -Max(L[1], L[0]) * (Q[0] - Q[1])
    k.directional(Qc[0], Qc[0] + 0.5 * dt * (left - right), struct=True)
    if source:
        k.single(Src(Q[0], S[0]))
        k.single(Qc[0], Qc[0] + dt * S[0], struct=True)
    k.single(Q[0], Qc[0])
    return k
