"""``CPPPrinter`` -- serial C++ from the same statement list the CUDA back-end consumes.

Counterpart of the reference's ``exahype/printers/CPPPrinter.py``: one loop nest per
statement, temporaries allocated per call, user functions (``Flux`` / ``maxEigenvalue`` / ``max``) linked from a
header such as the reference's ``Unit test/Functions.h``.  It exists so the CPU and CUDA back-ends can be generated
from one declaration and cross-checked.  Written from the rules, not from the reference's string surgery, and it
emits code that compiles -- the reference at HEAD does not (``&&Q_copy[..]``, ``Flux(..) = None;``, missing ``Q``
parameter):

* loop ranges (reference ``CPPPrinter.py:116-137``): ``patch`` over the batch; a spatial axis runs over the whole
  haloed side only when it is the statement's sweep axis and the statement has no shifted access (flux / eigenvalue
  sweeps), otherwise over the interior; the copy-in runs over every haloed cell (``Unit test/test.cpp:11-19``);
* ``var`` extent (``:118-126``): min of the statement's struct code and of ``item_struct`` over every declared name
  that is a substring of the statement -- ``0`` means no ``var`` loop;
* linearisation (``:247-261``): AoS, ``X[patch][i][j]([k])[var]`` with the item's own variable count;
* temporaries are value-initialised (the reference's are not, and are read before being written; section 0.2).
"""
from __future__ import annotations

from typing import List, Optional

import sympy
from sympy import Indexed
from sympy.printing.str import StrPrinter

from ..KernelBuilder import KernelBuilder
from .CodePrinter import CodePrinter


class _CxxExpr(StrPrinter):
    printmethod = "_exahype_device_str"   # Indexed._sympystr would otherwise bypass _print_Indexed

    def __init__(self, printer: "CPPPrinter", with_var: bool):
        super().__init__()
        self._p = printer
        self._with_var = with_var
        self._address_of = 0

    def _print_Indexed(self, expr):
        k = self._p.kernel()
        name = str(expr.base)
        width = {0: 1, 1: k.n_real, 2: k.n_real + k.n_aux}[k.item_struct[name]]
        side = k.patch_size if name in k.unhaloed_items else k.patch_size + 2 * k.halo_size
        spatial = list(expr.indices[: 1 + k.dim])
        terms = []
        stride = width * side ** k.dim
        for idx in spatial:
            text = self._print(idx)
            text = text if isinstance(idx, sympy.Symbol) or idx.is_Atom else f"({text})"
            terms.append(f"{stride}*{text}")
            stride //= side
        if len(expr.indices) > 1 + k.dim and self._with_var and not self._address_of:
            terms.append(self._print(expr.indices[-1]))
        if self._p.per_patch_member(name):
            # member of an ExaHyPE2 CellData: one array per patch, `patchData.QIn[patch][...]` -- what the reference's
            # CPPPrinter.parse (CPPPrinter.py:278-316) rewrites `patchData.QIn[stride*patch + ...]` into
            ref = f"{self._p.qualified(name)}[{self._print(spatial[0])}][{' + '.join(terms[1:])}]"
        else:
            ref = f"{self._p.qualified(name)}[{' + '.join(terms)}]"
        return f"&{ref}" if self._address_of else ref

    def _print_Symbol(self, expr):
        name = str(expr)
        if self._p.per_patch_member(name):      # `patchData.dt` -> `patchData.dt[patch]` (CPPPrinter.py:310-311)
            return f"{self._p.qualified(name)}[{self._p.kernel().indexes[0]}]"
        return self._p.qualified(name) if name in self._p.kernel().parents else super()._print_Symbol(expr)

    def _print_Function(self, expr):
        k = self._p.kernel()
        name = expr.func.__name__
        if name not in k.functions:
            return super()._print_Function(expr)
        args = []
        for a in expr.args:
            if isinstance(a, Indexed):           # arrays go by pointer to the cell (Functions.h:2-4)
                self._address_of += 1
                args.append(self._print(a))
                self._address_of -= 1
            else:
                args.append(self._print(a))
        return f"{self._p.qualified(name)}({', '.join(args)})"

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Idx(self, expr):
        return str(expr.label)


class CPPPrinter(CodePrinter):
    def __init__(self, kernel: KernelBuilder, function_name: str = "time_step"):
        super().__init__(kernel, function_name=function_name)
        self._lines: List[str] = []
        self._depth = 1
        k = kernel
        if not k.items:
            raise Exception("declare at least one item")
        params = [f"{k.input_types[0]} {k.items[0]}"]
        params += [f"{t} {n}" for n, t in zip(k.inputs, k.input_types[1:])]
        self._emit_raw(f"void {function_name}({', '.join(params)}) {{")
        for lit in k.literals:
            self._emit(lit.replace("int ", "const int ", 1))
        self._temporaries = [str(v) for n, v in k.all_items.items()
                             if isinstance(v, sympy.IndexedBase) and n != k.items[0] and n not in k.parents]
        side = k.patch_size + 2 * k.halo_size
        for name in self._temporaries:
            width = {0: 1, 1: k.n_real, 2: k.n_real + k.n_aux}[k.item_struct[name]]
            s = k.patch_size if name in k.unhaloed_items else side
            self._emit(f"double* {name} = new double[{k.n_patches * s ** k.dim * width}]();")
        for name in k.directional_consts:
            self._emit(f"int {name} = 0;")
        self._emit("(void)dim; (void)patch_size; (void)halo_size; (void)n_real; (void)n_aux;")
        for lhs, rhs, direction, struct in zip(k.LHS, k.RHS, k.directions, k.struct_inclusion):
            if str(lhs) in k.directional_consts:
                self._emit(f"{lhs} = {rhs};")
            else:
                self.loop([lhs, rhs], direction, k.dim + 1, struct)
        for name in self._temporaries:
            self._emit(f"delete[] {name};")
        self._emit_raw("}")
        self.code = "\n".join(self._lines) + "\n"

    # ------------------------------------------------------------------
    def qualified(self, name: str) -> str:
        parent = self.kernel().parents.get(name)
        if parent is None:
            return name
        return f"{parent}{name}" if parent.endswith(":") else f"{parent}.{name}"

    def per_patch_member(self, name: str) -> bool:
        """``name`` is a member of the first declared item and that item is an object, not an array (``in_type`` such as
        ``::exahype2::CellData&``, reference ``examples/kernel-generator.py:8``): ExaHyPE2's ``CellData`` keeps one entry
        per patch in every member -- ``QIn[patch]``, ``dt[patch]``, ``cellCentre[patch]``."""
        k = self.kernel()
        return bool(k.items) and k.parents.get(name) == k.items[0] and not k.input_types[0].rstrip().endswith("*")

    def _emit_raw(self, text: str):
        self._lines.append(text)

    def _emit(self, text: str):
        self._lines.append("\t" * self._depth + text)

    def _range(self, lhs, rhs, direction: int, level: int, first: bool, last: bool):
        k = self.kernel()
        lo, hi = k.halo_size, k.patch_size + k.halo_size
        if level == 0:
            return 0, k.n_patches
        if first and not last:
            return 0, k.patch_size + 2 * k.halo_size
        text = str(lhs) + " " + str(rhs)
        shifted = any(f"{a} {s}" in text for a in "ijk" for s in "+-")
        if direction == level and direction >= 1 and not shifted and not last:
            return 0, k.patch_size + 2 * k.halo_size
        return lo, hi

    def loop(self, expr: list, direction: int, below: int, struct_inclusion: int):
        """Emit the loop nest for one statement.  ``below`` counts the index levels still to open
        (``dim + 1`` at the outermost call), as in the reference."""
        k = self.kernel()
        lhs, rhs = expr
        level = k.dim + 1 - below
        first = lhs is k.LHS[0] and rhs is k.RHS[0]
        last = lhs is k.LHS[-1] and rhs is k.RHS[-1]
        if below > 0:
            idx = k.indexes[level]
            lo, hi = self._range(lhs, rhs, direction, level, first, last)
            self._emit(f"for (int {idx} = {lo}; {idx} < {hi}; {idx}++) {{")
            self._depth += 1
            self.loop(expr, direction, below - 1, struct_inclusion)
            self._depth -= 1
            self._emit("}")
            return
        widths = [w for name, w in k.item_struct.items() if name in str(expr)] + [struct_inclusion]
        extent = {0: 1, 1: k.n_real, 2: k.n_real + k.n_aux}[min(widths)]
        pr = _CxxExpr(self, with_var=extent > 1)
        body = pr.doprint(lhs) if rhs is None or rhs == '' else f"{pr.doprint(lhs)} = {pr.doprint(rhs)}"
        if extent > 1:
            self._emit(f"for (int var = 0; var < {extent}; var++) {{")
            self._emit("\t" + body + ";")
            self._emit("}")
        else:
            self._emit(body + ";")

    def file(self, file_name: str = 'test.cpp', header_file_name: Optional[str] = None):
        """Unlike the reference (``CPPPrinter.py:320-354``) no Peano headers are prepended: the unit is self-contained
        apart from ``header_file_name``, which declares the user functions."""
        head = f'#include "{header_file_name}"\n\n' if header_file_name is not None else ""
        if not self.code.startswith(head) or not head:
            self.code = head + self.code if head else self.code
        super().file(file_name, header_file_name)
