#!/usr/bin/env python
"""Regenerates the committed golden fixtures.  Runs ONLY in the dev container (needs /root/reference).

  python tests/golden/make_golden.py

Two kinds of fixture, both produced by the reference itself:

* ``g0_reference_kernel.json`` -- input and output (IEEE-754 bit patterns, hex) of the reference's committed
  generated kernel ``Unit test/test.cpp`` + ``Functions.cpp`` compiled unmodified by ``oracle/build_ref.sh``
  (temporaries value-initialised through a replaced ``operator new[]``), on the reference's own test input
  ``Q[i] = sin(3.141*i/360)``, ``dt = 1`` (``Unit test/correctness_test.cpp:102-106,191-196``); plus samples of the
  reference's ``Flux`` / ``maxEigenvalue``.
* ``statements_*.json`` -- the statement list (LHS/RHS/directions/struct_inclusion and the declaration tables) that
  the reference's Python ``KernelBuilder`` (``/root/reference/exahype/KernelBuilder.py``) records for the
  ``examples/Batched_stateless.py`` kernel, imported with ``xdsl`` stubbed out (it is not installed and is not on
  this path).
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REFERENCE = os.environ.get("EXAHYPE_REFERENCE_DIR", "/root/reference")


def hexbits(a):
    return [format(int(x), "016x") for x in np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).ravel()]


def golden_kernel():
    import oracle as O
    O.build()
    cfg = O.REFERENCE_CONFIG
    q_in = O.fill_sin(cfg, 1)
    q_out = q_in.copy()
    O.reference_time_step(q_out.reshape(-1), 1.0)
    rng = np.random.default_rng(7)
    samples = []
    for _ in range(16):
        q = np.concatenate([[1.0 + rng.random()], rng.random(2) - 0.5, [2.0 + rng.random()], rng.random(1)])
        for normal in (0, 1):
            samples.append({"q": hexbits(q), "normal": normal,
                            "F": hexbits(O.reference_flux(q, normal)),  # F[4] is never written in 2-D
                            "lambda": hexbits([O.reference_max_eigenvalue(q, normal)])[0]})
    out = {"config": {"dim": 2, "patch_size": 4, "halo": 1, "n_real": 5, "n_aux": 5, "n_patches": 1},
           "dt": 1.0, "input": hexbits(q_in), "output": hexbits(q_out), "fnv1a64": O.fnv1a64(q_out),
           "physics_samples": samples}
    with open(os.path.join(HERE, "g0_reference_kernel.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("g0", out["fnv1a64"])


def import_reference_exahype():
    """Import /root/reference/exahype with xdsl stubbed (exahype/__init__.py:3 imports SymPyToMLIR eagerly)."""
    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (), {})
    for mod in ["xdsl", "xdsl.dialects", "xdsl.dialects.builtin", "xdsl.dialects.experimental", "xdsl.ir",
                "xdsl.irdl", "xdsl.utils", "xdsl.utils.hints", "xdsl.utils.exceptions", "xdsl.traits",
                "xdsl.dialects.func", "xdsl.dialects.arith", "xdsl.dialects.scf", "xdsl.dialects.memref",
                "xdsl.dialects.llvm", "xdsl.dialects.math", "xdsl.dialects.experimental.math", "xdsl.printer",
                "xdsl.builder", "xdsl.utils.test_value"]:
        sys.modules.setdefault(mod, _Any(mod))
    sys.path.insert(0, REFERENCE)
    for k in [k for k in sys.modules if k == "exahype" or k.startswith("exahype.")]:
        del sys.modules[k]
    import exahype  # noqa: the reference's package
    assert exahype.__file__.startswith(REFERENCE), exahype.__file__
    return exahype


def batched_stateless(KernelBuilder, dim, patch_size, halo_size, n_real, n_aux, n_patches=1):
    """The declaration of /root/reference/examples/Batched_stateless.py:9-35, parametrised."""
    from sympy.codegen.ast import integer, real, none
    kernel = KernelBuilder(dim=dim, patch_size=patch_size, halo_size=halo_size, n_real=n_real, n_aux=n_aux,
                           n_patches=n_patches)
    Q = kernel.item('Q')
    Q_copy = kernel.item('Q_copy')
    tmp_flux = kernel.directional_item('tmp_flux')
    tmp_eig = kernel.directional_item('tmp_eigen', struct=False)
    dt = kernel.const('dt')
    normal = kernel.directional_const('normal', list(range(dim)))
    Flux = kernel.function('Flux', parameter_types=[Q, real, Q], return_type=integer)
    Eigen = kernel.function('maxEigenvalue', parameter_types=[Q, real], return_type=real)
    Max = kernel.function('max', parameter_types=[Q, Q], return_type=none)
    kernel.single(Q_copy[0], Q[0])
    kernel.directional(Flux(Q_copy[0], normal, tmp_flux[0]))
    kernel.directional(tmp_eig[0], Eigen(Q_copy[0], normal))
    kernel.directional(Q_copy[0], Q_copy[0] + 0.5 * (tmp_flux[-1] - tmp_flux[1]))
    left = -Max(tmp_eig[-1], tmp_eig[0]) * (Q[0] - Q[-1])
    right = -Max(tmp_eig[1], tmp_eig[0]) * (Q[0] - Q[1])
    kernel.directional(Q_copy[0], Q_copy[0] + 0.5 * dt * (left - right), struct=True)
    kernel.single(Q[0], Q_copy[0])
    return kernel


def dump_kernel(k):
    return {
        "dim": k.dim, "patch_size": k.patch_size, "halo_size": k.halo_size, "n_patches": k.n_patches,
        "n_real": k.n_real, "n_aux": k.n_aux,
        "indexes": [str(i) for i in k.indexes], "literals": list(k.literals), "parents": dict(k.parents),
        "inputs": list(k.inputs), "input_types": list(k.input_types), "items": list(k.items),
        "directional_items": list(k.directional_items),
        "directional_consts": {a: list(b) for a, b in k.directional_consts.items()},
        "functions": [str(f) for f in k.functions], "item_struct": dict(k.item_struct),
        "all_items": sorted(k.all_items.keys()),
        "LHS": [str(x) for x in k.LHS], "RHS": [str(x) for x in k.RHS],
        "directions": list(k.directions), "struct_inclusion": list(k.struct_inclusion),
    }


def golden_statements():
    exahype = import_reference_exahype()
    for name, args in {"2d_p4_h1_r5_a5": (2, 4, 1, 5, 5, 1), "2d_p3_h1_r4_a0_b1000": (2, 3, 1, 4, 0, 1000),
                       "3d_p8_h1_r5_a0_b4": (3, 8, 1, 5, 0, 4)}.items():
        k = batched_stateless(exahype.KernelBuilder, *args)
        with open(os.path.join(HERE, f"statements_{name}.json"), "w") as f:
            json.dump(dump_kernel(k), f, indent=1)
        print(name, len(k.LHS), "statements")


if __name__ == "__main__":
    golden_kernel()
    golden_statements()
