"""The ExaHyPE2 CellData boundary on the CPU: the reference's own user script
examples/kernel-generator.py, run verbatim, drives CPPPrinter and CUDAPrinter; the C++ it generates compiles against a
minimal fake of ExaHyPE2's types and -- with a solver that ignores position and time -- reproduces the oracle bit for
bit; the CUDA unit cross-compiles for sm_100a."""
import os

import numpy as np
import pytest

import cell_data_common as C

REFERENCE_SCRIPT = "/root/reference/examples/kernel-generator.py"


def _run_reference_script(tmp_path):
    with open(REFERENCE_SCRIPT) as f:
        source = f.read()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        env = {"__name__": "__main__"}
        exec(compile(source, REFERENCE_SCRIPT, "exec"), env)
    finally:
        os.chdir(cwd)
    return env


@pytest.mark.skipif(not os.path.exists(REFERENCE_SCRIPT), reason="the reference tree is not present on this machine")
def test_restated_declaration_equals_the_reference_script(tmp_path):
    """tests/cell_data_common.declare() is what the GPU tests use (no /root/reference there): same tables, same statements,
    same generated C++ and CUDA as the script itself."""
    from exahype.printers import CPPPrinter, CUDAPrinter
    ref = _run_reference_script(tmp_path)["kernel"]
    mine = C.declare()
    for attr in ("dim", "patch_size", "halo_size", "n_real", "n_aux", "n_patches", "items", "directional_items",
                 "functions", "inputs", "input_types", "parents", "item_struct", "directions", "struct_inclusion"):
        assert getattr(ref, attr) == getattr(mine, attr), attr
    assert [str(x) for x in ref.LHS] == [str(x) for x in mine.LHS]
    assert [str(x) for x in ref.RHS] == [str(x) for x in mine.RHS]
    assert CPPPrinter(ref).code == CPPPrinter(mine).code
    assert (tmp_path / "generated_kernel.cpp").read_text().endswith(CPPPrinter(mine).code)
    for k in (ref, mine):
        k.all_items["flux"].deviceBody(C.device_solver(2))
    assert CUDAPrinter(ref).code == CUDAPrinter(mine).code


def test_cuda_printer_accepts_the_cell_data_declaration():
    """The declaration CUDAPrinter used to reject ("symbol dt is not available inside the kernel"): CellData members as
    kernel-side values, the solver signature flux(Q, x, h, t, dt, normal, F), one name for flux and eigenvalue."""
    from exahype.printers import CUDAPrinter
    from exahype_b200.printers.CUDAPrinter import analyse
    k = C.declare()
    p = analyse(k)
    assert (p.q_in, p.q_work, p.flux_fn, p.eigen_fn, p.max_fn, p.dt) == ("QOut", "QIn", "flux", "flux", "max", "dt")
    assert p.flux_args == ["Q", "x", "h", "t", "dt", "normal", "F"] and p.eigen_args == ["Q", "X", "h", "t", "dt", "normal"]
    assert p.context and not p.dissipation_all       # `struct=True` is silenced by tmp_eigen, as in the reference's printer
    k.all_items["flux"].deviceBody(C.device_solver(2))
    cu = CUDAPrinter(k)
    assert cu.template == "cell" and cu.context
    assert "static constexpr bool NEEDS_CONTEXT = true;" in cu.code
    assert "user::flux(q, c.x, c.h, c.t, c.dt, N, F);" in cu.code and "return user::flux(q, c.X, c.h, c.t, c.dt, N);" in cu.code
    assert "int time_step_cell_data(const exahype_cell_data* cells" in cu.code
    # the committed Euler family ignores position and time: every kernel template stays available
    cu2 = CUDAPrinter(C.declare(dim=3, patch_size=8, n_real=5), model="euler")
    assert cu2.template == "pair" and not cu2.context and "NEEDS_CONTEXT" not in cu2.code
    assert "Fv3dPairAutoConfig<Physics, Update, double, 8, 1, false, true>::gather_type" in cu2.code


@pytest.mark.parametrize("dim,P", [(2, 4), (3, 4)])
def test_generated_units_cross_compile(dim, P):
    from exahype.printers import CUDAPrinter
    k = C.declare(dim=dim, patch_size=P, n_real=dim + 2)
    k.all_items["flux"].deviceBody(C.device_solver(dim))
    gk = CUDAPrinter(k).build()                   # nvcc -gencode arch=compute_100a,code=sm_100a, no GPU needed
    assert os.path.exists(gk.lib_path)


def test_generated_cpp_with_a_context_free_solver_equals_the_oracle(tmp_path, oracle):
    """CPPPrinter's output for the CellData declaration (per-patch members, `patchData.QIn[patch][...]`) against the pinned
    oracle: with a solver whose flux / eigenvalue are the reference's Functions.cpp formulas, the generated C++ must
    reproduce the oracle bit for bit -- this pins the CPU side of the GPU parity test."""
    from exahype.printers import CPPPrinter
    k = C.declare()
    CPPPrinter(k).file(file_name=str(tmp_path / "generated_kernel.cpp"))
    # a fake header variant whose solver is plain Euler (Functions.cpp:9-62), ignoring x / h / t / dt
    fake = open(os.path.join(C.HERE, "cpp", "fake_exahype2.h")).read()
    plain = fake.replace("const double w = 1.0 + 0.01 * x(normal) + 0.1 * h(0) + 0.001 * t + 0.5 * dt;", "const double w = 1.0;") \
                .replace("return std::fabs(Q[normal + 1] * irho) + 0.01 * x(0) + h(1) + t + dt;",
                         "const double ir = 1.0 / std::fabs(Q[0]);\n"
                         "    const double pp = (1.4 - 1) * (Q[3] - 0.5 * ir * (Q[1] * Q[1] + Q[2] * Q[2]));\n"
                         "    const double c = std::sqrt(1.4 * std::fabs(pp) * ir);\n"
                         "    const double u = Q[normal + 1] * ir;\n"
                         "    return std::fmax(std::fabs(u - c), std::fabs(u + c));")
    assert plain != fake
    (tmp_path / "plain_exahype2.h").write_text(plain)
    import subprocess, ctypes
    lib = str(tmp_path / "libplain.so")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-DDimensions=2",
                        '-DGENERATED_KERNEL="generated_kernel.cpp"', '-DFAKE_HEADER="plain_exahype2.h"', "-I", str(tmp_path),
                        os.path.join(C.HERE, "cpp", "cell_data_harness.cpp"), "-o", lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    fn = ctypes.CDLL(lib).run_cell_data
    vp = ctypes.c_void_p
    fn.argtypes = [ctypes.c_int, ctypes.c_longlong, vp, vp, vp, vp, vp, vp]
    cfg = oracle.OracleConfig(dim=2, patch_size=4, halo=1, n_real=4, n_aux=0)
    n = 9
    q0 = oracle.fill_synthetic(cfg, n)
    want = q0.copy()
    centre, size, t, dt = C.patch_geometry(n, 2)
    for p in range(n):                            # the oracle takes one dt per call
        oracle.step(cfg, want[p:p + 1], float(dt[p]))
    got = q0.copy()
    scratch = np.zeros(got[0].size)
    fn(n, got[0].size, got.ctypes.data, scratch.ctypes.data, centre.ctypes.data, size.ctypes.data, t.ctypes.data, dt.ctypes.data)
    assert np.array_equal(got, want)
