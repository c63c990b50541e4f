"""Back-end printers: CPPPrinter output compiles and equals the oracle bit for bit (CPU); CUDAPrinter recognises the
program, emits the expected functors and its unit cross-compiles for sm_100a (no GPU needed for that)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import batched_stateless  # noqa: E402

from exahype import KernelBuilder  # noqa: E402
from exahype.printers import CPPPrinter, CUDAPrinter, MLIRPrinter  # noqa: E402
from exahype_b200.printers import UnsupportedKernel, analyse  # noqa: E402

GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def compile_cpp(tmp_path, kernel, dims):
    src = tmp_path / "generated.cpp"
    CPPPrinter(kernel).file(str(src), header_file_name="Functions.h")
    lib = tmp_path / "libgenerated.so"
    subprocess.run([GXX, "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", f"-DDIMENSIONS={dims}",
                    "-I", os.path.join(HERE, "data"), str(src), os.path.join(HERE, "data", "Functions.cpp"),
                    "-o", str(lib)], check=True)
    fn = getattr(ctypes.CDLL(str(lib)), "_Z9time_stepPdd")     # C++ linkage, like the reference's test.h:3
    fn.argtypes = [ctypes.c_void_p, ctypes.c_double]
    fn.restype = None
    return fn


@pytest.mark.parametrize("shape", [(2, 3, 1, 4, 0, 7), (2, 4, 1, 4, 2, 3), (3, 4, 1, 5, 0, 2), (2, 5, 2, 4, 0, 2)])
def test_cpp_printer_output_equals_oracle(tmp_path, oracle, shape):
    dim, P, h, nr, na, B = shape
    time_step = compile_cpp(tmp_path, batched_stateless(KernelBuilder, dim, P, h, nr, na, B), dim)
    cfg = oracle.OracleConfig(dim=dim, patch_size=P, halo=h, n_real=nr, n_aux=na)
    q = oracle.fill_synthetic(cfg, B)
    want = q.copy()
    oracle.step(cfg, want, 0.01)
    time_step(q.ctypes.data, 0.01)
    assert np.array_equal(q, want)


def test_cpp_printer_signature_and_ranges():
    code = CPPPrinter(batched_stateless(KernelBuilder, 2, 4, 1, 5, 5)).code
    assert code.startswith("void time_step(double* Q, double dt) {")
    assert "new double[360]()" in code and "new double[36]()" in code
    assert "Flux(&Q_copy[360*patch + 60*i + 10*j], normal, &tmp_flux_x[180*patch + 30*i + 5*j]);" in code
    # flux sweep along i: full range on i, interior on j (reference CPPPrinter.py:132-137; committed file has it transposed)
    flux_x = code[code.index("normal = 0;"):code.index("normal = 1;")]
    assert "int i = 0; i < 6" in flux_x and "int j = 1; j < 5" in flux_x
    assert "&&" not in code and "None" not in code and "patch - 1" not in code
    with pytest.raises(NotImplementedError):
        MLIRPrinter(batched_stateless(KernelBuilder, 2, 4, 1, 5, 5))


def test_cuda_printer_recognises_the_program():
    k = batched_stateless(KernelBuilder, 3, 8, 1, 5, 0, 4)
    prog = analyse(k)
    assert (prog.q_in, prog.q_work, prog.flux_tmp, prog.eigen_tmp) == ("Q", "Q_copy", "tmp_flux", "tmp_eigen")
    assert (prog.flux_fn, prog.eigen_fn, prog.max_fn, prog.dt) == ("Flux", "maxEigenvalue", "max", "dt")
    assert prog.normals == [0, 1, 2]
    assert prog.flux_update == "qc - T(0.5)*f_plus + T(0.5)*f_minus"
    assert prog.dissipation == ("T(0.5)*dt*((-q_plus + q0)*::exahype::fv_max(l_plus, l0) + "
                                "(q_minus - q0)*::exahype::fv_max(l_minus, l0)) + qc")
    assert prog.dissipation_all is False            # tmp_eigen silences the var loop, as in the reference's output
    p = CUDAPrinter(k, model="euler")
    assert "using Physics = ::exahype::EulerPhysics<3, 5, 0>;" in p.code
    assert 'extern "C"' in p.code and "int time_step(const void* q_in" in p.code
    assert '#include "fv3d_pair_kernel.cuh"' in p.code
    assert "::exahype::Fv3dPairAuto<Physics, Update, double, 8, 1, false, true>::launch" in p.code
    grp = CUDAPrinter(k, model="euler", template="march")
    assert '#include "fv3d_march_kernel.cuh"' in grp.code
    assert "::exahype::Fv3dMarchAuto<Physics, Update, double, 8, 1, false, true>::launch" in grp.code
    assert "Fv3dMarchAuto<Physics, Update, double, 4, 1" in CUDAPrinter(
        batched_stateless(KernelBuilder, 3, 4, 1, 5, 0), model="euler").code
    cell = CUDAPrinter(k, model="euler", template="cell")
    assert "FvKernelConfig<Physics, Update, double, 3, 8, 1, 1, 512, 1, false, true, false>" in cell.code
    assert "Fv2dMarchAuto<Physics, Update, double, 16, 1, false, false>" in CUDAPrinter(
        batched_stateless(KernelBuilder, 2, 16, 1, 4, 0), model="euler").code
    assert "FvKernelConfig" in CUDAPrinter(batched_stateless(KernelBuilder, 2, 3, 1, 4, 0), model="euler").code
    assert CUDAPrinter(k, function_name="step32", dtype="f32", dissipation="all", model="euler").code.count("float") >= 2


def test_cuda_printer_rejects_other_programs():
    k = KernelBuilder(2, 4, 1, 4, 0)
    q, qc = k.item("Q"), k.item("Q_copy")
    k.directional_item("tmp_flux"); k.directional_item("tmp_eigen", struct=False)
    k.single(qc[0], q[0]); k.single(q[0], qc[0])
    with pytest.raises(UnsupportedKernel):
        CUDAPrinter(k)
    with pytest.raises(UnsupportedKernel):
        CUDAPrinter(batched_stateless(KernelBuilder, 2, 4, 0, 4, 0))   # no halo layer to read


def _swe_kernel():
    """Shallow water declared with SymPy bodies: the CUDA functors are generated, not hand-written."""
    import sympy
    from sympy.codegen.ast import real, integer, none
    k = KernelBuilder(dim=2, patch_size=16, halo_size=1, n_real=3, n_aux=1)
    Q, Qc = k.item('Q'), k.item('Q_copy')
    F, L = k.directional_item('tmp_flux'), k.directional_item('tmp_eigen', struct=False)
    dt = k.const('dt'); normal = k.directional_const('normal', [0, 1])
    g = 9.81

    def flux(q, n):
        un = q[n + 1] / q[0]
        f = [un * q[0], un * q[1], un * q[2]]
        f[n + 1] = f[n + 1] + 0.5 * g * q[0] * q[0]
        return f

    def eig(q, n):
        un, c = q[n + 1] / sympy.Abs(q[0]), sympy.sqrt(g * sympy.Abs(q[0]))
        return sympy.Max(sympy.Abs(un - c), sympy.Abs(un + c))
    Flux = k.function('Flux', parameter_types=[Q, real, Q], return_type=integer, body=flux)
    Eig = k.function('maxEigenvalue', parameter_types=[Q, real], return_type=real, body=eig)
    Max = k.function('max', parameter_types=[Q, Q], return_type=none)
    k.single(Qc[0], Q[0])
    k.directional(Flux(Qc[0], normal, F[0]))
    k.directional(L[0], Eig(Qc[0], normal))
    k.directional(Qc[0], Qc[0] + 0.5 * (F[-1] - F[1]))
    left = -Max(L[-1], L[0]) * (Q[0] - Q[-1]); right = -Max(L[1], L[0]) * (Q[0] - Q[1])
    k.directional(Qc[0], Qc[0] + 0.5 * dt * (left - right), struct=True)
    k.single(Q[0], Qc[0])
    return k


def test_generated_units_cross_compile_for_sm100a(tmp_path):
    """nvcc -gencode arch=compute_100a,code=sm_100a on generated units: hand-written family, SymPy bodies, and the
    user's own device source with the reference's Functions.h signatures."""
    lib = CUDAPrinter(_swe_kernel(), function_name="swe_step").build(directory=str(tmp_path))
    assert os.path.exists(lib.lib_path) and lib.in_shape(2) == (2, 18, 18, 4)
    assert "F[1] = " in CUDAPrinter(_swe_kernel()).code

    k = batched_stateless(KernelBuilder, 2, 8, 1, 4, 0)
    hdr = tmp_path / "Functions.cuh"
    hdr.write_text('''
template <class T> __device__ void Flux(const T* Q, int normal, T* F) {
  const T irho = T(1.0) / Q[0];
  const T p = (T(1.4) - 1) * (Q[3] - T(0.5) * irho * (Q[1] * Q[1] + Q[2] * Q[2]));
  const T coeff = irho * Q[normal + 1];
  F[0] = coeff * Q[0]; F[1] = coeff * Q[1]; F[2] = coeff * Q[2]; F[3] = coeff * Q[3] + coeff * p;
  F[normal + 1] += p;
}
template <class T> __device__ T maxEigenvalue(const T* Q, int normal) {
  const T irho = T(1.0) / fabs(Q[0]);
  const T p = (T(1.4) - 1) * (Q[3] - T(0.5) * irho * (Q[1] * Q[1] + Q[2] * Q[2]));
  const T c = sqrt(T(1.4) * fabs(p) * irho);
  const T u = Q[normal + 1] * irho;
  return fmax(fabs(u - c), fabs(u + c));
}
''')
    printer = CUDAPrinter(k, function_name="user_step")
    src = tmp_path / "user_step.cu"
    printer.file(str(src), header_file_name="Functions.cuh")
    assert src.read_text().startswith('#include "Functions.cuh"')
    built = printer.build(directory=str(tmp_path), include_dirs=[str(tmp_path)])
    assert os.path.exists(built.lib_path)


def test_generated_3d_unit_instantiates_the_warp_per_patch_template(tmp_path):
    """8x8x8 patches: `template='auto'` picks csrc/fv3d_pair_kernel.cuh (Fv3dPairAuto) and the unit cross-compiles for
    sm_100a, fp64 and fp32, exporting the drop-in entry."""
    import ctypes
    k = batched_stateless(KernelBuilder, 3, 8, 1, 5, 0)
    for dtype in ("f64", "f32"):
        p = CUDAPrinter(k, model="euler", dtype=dtype, function_name=f"step3d_{dtype}")
        assert p.template == "pair" and "Fv3dPairAuto<Physics, Update" in p.code
        built = p.build(directory=str(tmp_path))
        assert os.path.exists(built.lib_path)
        assert hasattr(ctypes.CDLL(built.lib_path), f"step3d_{dtype}")
    with pytest.raises(Exception):
        CUDAPrinter(batched_stateless(KernelBuilder, 3, 4, 1, 5, 0), model="euler", template="pair")


def test_warp_per_patch_tuning_switches_keep_compiling(tmp_path):
    """The A/B switches of csrc/fv3d_pair_kernel.cuh at their NON-default values -- ring slot released a step later
    (EXAHYPE_3D_EARLY2=0), the two window-filling steps of a patch merged, CTA-major patch order -- and the cache-less
    functor family (registers carry F_1, F_2, L_1, L_2 for one step) still cross-compile without spills."""
    import importlib
    from exahype_b200 import build as B
    cp = importlib.import_module("exahype_b200.printers.CUDAPrinter")     # the module (CSRC / INCLUDE), not the class
    k = batched_stateless(KernelBuilder, 3, 8, 1, 5, 0)
    p = CUDAPrinter(k, model="euler", function_name="step3d_switches")
    src = tmp_path / "step3d_switches.cu"
    src.write_text(p.code)
    cmd = [B.nvcc()] + B.NVCC_FLAGS + B._host_compiler_args() + ["-I", cp.CSRC, "-I", cp.INCLUDE, "-Xptxas", "-v",
           "-DEXAHYPE_3D_EARLY2=0", "-DEXAHYPE_3D_MERGED_PRE=1", "-DEXAHYPE_3D_INTERLEAVE=0",
           "-c", str(src), "-o", str(tmp_path / "step3d_switches.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    def spills(log):
        return [ln for ln in log.splitlines() if "spill stores" in ln and "0 bytes stack frame, 0 bytes spill stores" not in ln]
    assert not spills(r.stderr), spills(r.stderr)

    # cache-less family: the user's device source with the Functions.h signatures, default switches
    hdr = tmp_path / "Functions3d.cuh"
    hdr.write_text("""
template <class T> __device__ void Flux(const T* Q, int normal, T* F) {
  const T irho = T(1.0) / Q[0];
  const T p = (T(1.4) - 1) * (Q[4] - T(0.5) * irho * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]));
  const T coeff = irho * Q[normal + 1];
  F[0] = coeff * Q[0]; F[1] = coeff * Q[1]; F[2] = coeff * Q[2]; F[3] = coeff * Q[3]; F[4] = coeff * Q[4] + coeff * p;
  F[normal + 1] += p;
}
template <class T> __device__ T maxEigenvalue(const T* Q, int normal) {
  const T irho = T(1.0) / fabs(Q[0]);
  const T p = (T(1.4) - 1) * (Q[4] - T(0.5) * irho * (Q[1] * Q[1] + Q[2] * Q[2] + Q[3] * Q[3]));
  const T c = sqrt(T(1.4) * fabs(p) * irho);
  const T u = Q[normal + 1] * irho;
  return fmax(fabs(u - c), fabs(u + c));
}
""")
    pu = CUDAPrinter(batched_stateless(KernelBuilder, 3, 8, 1, 5, 0), function_name="user_step3d_stash")
    assert pu.template == "pair"
    usrc = tmp_path / "user_step3d_stash.cu"
    pu.file(str(usrc), header_file_name="Functions3d.cuh")
    cmd = [B.nvcc()] + B.NVCC_FLAGS + B._host_compiler_args() + ["-I", cp.CSRC, "-I", cp.INCLUDE, "-I", str(tmp_path),
           "-Xptxas", "-v", "-c", str(usrc), "-o", str(tmp_path / "user_step3d_stash.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not spills(r.stderr), spills(r.stderr)


def _euler_sympy_bodies(dim, gamma=1.4):
    import sympy

    def flux(q, n):
        irho = 1 / q[0]
        ke = sum(q[1 + a] * q[1 + a] for a in range(dim))
        p = (gamma - 1) * (q[dim + 1] - sympy.Rational(1, 2) * irho * ke)
        coeff = irho * q[n + 1]
        f = [coeff * q[v] for v in range(dim + 1)] + [coeff * q[dim + 1] + coeff * p]
        f[n + 1] = f[n + 1] + p
        return f

    def eig(q, n):
        irho = 1 / sympy.Abs(q[0])
        ke = sum(q[1 + a] * q[1 + a] for a in range(dim))
        p = (gamma - 1) * (q[dim + 1] - sympy.Rational(1, 2) * irho * ke)
        c = sympy.sqrt(gamma * sympy.Abs(p) * irho)
        u = q[n + 1] * irho
        return sympy.Max(sympy.Abs(u - c), sympy.Abs(u + c))
    return flux, eig


def test_strength_reduction_of_sympy_functors_keeps_the_values():
    """CUDAPrinter.strength_reduce gathers a product's half-power factors under one root and turns 1/|x| into |1/x|
    (opaque functions, so that SymPy does not split them again): same values to rounding, one reciprocal and one root
    per cell in the generated per-cell cache."""
    import sympy
    from exahype_b200.printers.CUDAPrinter import strength_reduce
    q = sympy.symbols("q0:5", real=True)
    flux, eig = _euler_sympy_bodies(3)
    rng = np.random.default_rng(5)
    states = np.column_stack([rng.uniform(0.5, 2.0, 200) * rng.choice([-1.0, 1.0], 200),     # either sign of the density
                              rng.uniform(-1, 1, (200, 3)), rng.uniform(2.0, 4.0, 200)])
    custom = {"exahype_recip": lambda x: 1.0 / x, "exahype_sqrt": np.sqrt}
    for n in range(3):
        for expr in list(flux(list(q), n)) + [eig(list(q), n)]:
            reduced = strength_reduce(expr)
            f0 = sympy.lambdify(q, expr, "numpy")
            f1 = sympy.lambdify(q, reduced, [custom, "numpy"])
            a, b = f0(*states.T) * np.ones(200), f1(*states.T) * np.ones(200)
            np.testing.assert_allclose(b, a, rtol=2e-13, atol=2e-13)      # a/b against a*(1/b), and cancellation in F
    # the eigenvalue: SymPy's own form has two roots and a reciprocal of |rho|, the reduced one a single root
    from sympy.core.function import AppliedUndef
    roots = {a for a in strength_reduce(eig(list(q), 0)).atoms(AppliedUndef) if a.func.__name__ == "exahype_sqrt"}
    assert len(roots) == 1 and len(eig(list(q), 0).atoms(sympy.Pow)) > 2
    k = batched_stateless(KernelBuilder, 3, 8, 1, 5, 0)
    k.all_items["Flux"].deviceBody(flux)
    k.all_items["maxEigenvalue"].deviceBody(eig)
    code = CUDAPrinter(k, function_name="sympy_euler3d").code
    prims = code[code.index("Prims<T> prims("):code.index("return pr;")]
    assert prims.count("T(1.0)/(") == 1 and prims.count("fv_sqrt<T>(") == 1, prims
    body = code[code.index("return pr;"):code.index("using Update")]
    assert "T(1.0)/(" not in body and "fv_sqrt" not in body       # flux / eigenvalue calls only combine cached values
