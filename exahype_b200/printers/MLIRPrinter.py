"""Placeholder for the reference's MLIR back-end (``exahype/printers/MLIRPrinter.py``, ``exahype/SymPyToMLIR.py``).

Out of scope for this repository: the path needs xDSL, which is not installed and
is not on the accelerated hot path.  The name is kept importable so that user scripts written against the reference
(``from exahype.printers import CPPPrinter, MLIRPrinter``, ``examples/Batched_stateless.py:6``) still import.
"""
from __future__ import annotations

from .CodePrinter import CodePrinter


class MLIRPrinter(CodePrinter):
    def __init__(self, kernel, function_name: str = "time_step"):
        raise NotImplementedError(
            "MLIRPrinter is not part of exahype_b200: the SymPy->MLIR path needs xDSL and is out of scope; "
            "use CUDAPrinter (GPU) or CPPPrinter (CPU)")

    def loop(self, expr, direction, below, struct_inclusion):  # pragma: no cover
        raise NotImplementedError
