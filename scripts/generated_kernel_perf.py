"""What a GENERATED kernel costs next to the committed hand-written family: the same declaration
(the 14 / 20 statements of examples/Batched_stateless.py) compiled by CUDAPrinter three ways -- committed functor family
(per-cell primitive cache `Prims`), SymPy-bodied functors, the user's device source with the Functions.h signatures (both
with an empty `Prims`: 1/rho, p, c are evaluated by every flux / eigenvalue call and left to the compiler's CSE) -- on the
BASELINE batch of C3 and C2, kernel-only, burst.  Run on a B200:  python scripts/generated_kernel_perf.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import sympy
from make_golden import batched_stateless
from exahype import KernelBuilder
from exahype.printers import CUDAPrinter
from exahype_b200 import runtime

GAMMA = 1.4


def euler_bodies(dim):
    def flux(q, n):
        irho = 1 / q[0]
        ke = sum(q[1 + a] * q[1 + a] for a in range(dim))
        p = (GAMMA - 1) * (q[dim + 1] - sympy.Rational(1, 2) * irho * ke)
        coeff = irho * q[n + 1]
        f = [coeff * q[v] for v in range(dim + 1)] + [coeff * q[dim + 1] + coeff * p]
        f[n + 1] = f[n + 1] + p
        return f

    def eig(q, n):
        irho = 1 / sympy.Abs(q[0])
        ke = sum(q[1 + a] * q[1 + a] for a in range(dim))
        p = (GAMMA - 1) * (q[dim + 1] - sympy.Rational(1, 2) * irho * ke)
        c = sympy.sqrt(GAMMA * sympy.Abs(p) * irho)
        u = q[n + 1] * irho
        return sympy.Max(sympy.Abs(u - c), sympy.Abs(u + c))
    return flux, eig


def user_source(dim):
    e = dim + 1
    ke = " + ".join(f"Q[{1 + a}] * Q[{1 + a}]" for a in range(dim))
    comps = " ".join(f"F[{v}] = coeff * Q[{v}];" for v in range(dim + 1))
    return f"""
template <class T> __device__ void Flux(const T* Q, int normal, T* F) {{
  const T irho = T(1.0) / Q[0];
  const T p = (T(1.4) - 1) * (Q[{e}] - T(0.5) * irho * ({ke}));
  const T coeff = irho * Q[normal + 1];
  {comps} F[{e}] = coeff * Q[{e}] + coeff * p;
  F[normal + 1] += p;
}}
template <class T> __device__ T maxEigenvalue(const T* Q, int normal) {{
  const T irho = T(1.0) / fabs(Q[0]);
  const T p = (T(1.4) - 1) * (Q[{e}] - T(0.5) * irho * ({ke}));
  const T c = sqrt(T(1.4) * fabs(p) * irho);
  const T u = Q[normal + 1] * irho;
  return fmax(fabs(u - c), fabs(u + c));
}}
"""


BUILD_DIR = os.path.join(ROOT, "exahype_b200", "variants", "generated")   # in-tree: built here, travels to the GPU box


def timed(step, steps=20):
    import torch
    for _ in range(3):
        step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(steps):
        step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    build_only = "--build-only" in sys.argv          # cross-compile the units (no GPU needed), then stop
    if not build_only:
        import torch
    tmp = BUILD_DIR
    os.makedirs(tmp, exist_ok=True)
    peak = 6547.8
    for name, dim, P, batch in (("C3", 3, 8, 32768), ("C2", 2, 16, 65536)):
        nr = dim + 2
        upd = runtime.PatchUpdate("euler", dim, P, 1, nr, 0, output="unhaloed")
        if build_only:
            q = out = lam = None
            rows = []
        else:
          q = upd.fill_synthetic(torch.empty(upd.in_shape(batch), dtype=torch.float64, device="cuda"), 0)
          out = torch.empty(upd.out_shape(batch), dtype=torch.float64, device="cuda")
          lam = torch.zeros(1, dtype=torch.float64, device="cuda")
          rows = [("committed instantiation (libexahype_cuda.so)", timed(lambda: upd.step(q, out, 0.01, None, lam)))]
        gb = upd.algorithmic_bytes_per_patch * batch / 1e9
        flux, eig = euler_bodies(dim)
        for label, kw in (("generated, committed functor family (model='euler')", dict(model="euler")),
                          ("generated, SymPy-bodied functors (Prims derived by CSE + strength reduction)", dict(bodies=(flux, eig))),
                          ("generated, user device source, Functions.h signatures (empty Prims)", dict(source=user_source(dim)))):
            k = batched_stateless(KernelBuilder, dim, P, 1, nr, 0)
            if "bodies" in kw:
                k.all_items["Flux"].deviceBody(kw["bodies"][0]); k.all_items["maxEigenvalue"].deviceBody(kw["bodies"][1])
                pr = CUDAPrinter(k, function_name=f"gen_{name}_sympy")
            elif "source" in kw:
                k.all_items["Flux"].deviceBody(kw["source"]); k.all_items["maxEigenvalue"].deviceBody(kw["source"])
                pr = CUDAPrinter(k, function_name=f"gen_{name}_user")
            else:
                pr = CUDAPrinter(k, model="euler", function_name=f"gen_{name}_family")
            gk = pr.build(directory=tmp)
            if build_only:
                print("built", gk.lib_path)
                continue
            rows.append((f"{label} [{pr.template}]", timed(lambda: gk.step(q, out, 0.01, None, lam, unhaloed=True))))
        if build_only:
            continue
        print(f"{name}: {batch} patches, un-haloed output, var0 dissipation, burst (3 warm-up + 20 launches, CUDA events)")
        for label, ms in rows:
            print(f"   {ms:7.4f} ms  {gb / (ms * 1e-3):6.0f} GB/s  {gb / (ms * 1e-3) / peak:5.3f} of the copy peak   {label}")
        del q, out


if __name__ == "__main__":
    main()
