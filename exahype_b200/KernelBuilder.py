"""``KernelBuilder`` -- records a patch kernel as parallel lists of SymPy statements.

Same constructor, methods and public attributes as the reference's ``exahype/KernelBuilder.py``
(ctor ``:51-90``; ``const :92``, ``directional_const :105``, ``item :112``, ``directional_item :122``,
``function :134``, ``single :144``, ``directional :165``, ``index :175``), because the printers read the
attribute lists directly (``exahype/printers/CPPPrinter.py:53-101``).  The implementation is new (regular-expression driven
rewriting of the relative accesses; table-driven declarations).

Relative-index DSL: inside a statement ``X[n]`` (one integer index) means "the cell at offset ``n`` along the
current sweep axis"; ``index()`` expands it to ``X[patch, i(+n), j, (k,) var]``.  Sweep axis 1 is ``i`` -- the
slowest spatial index -- and pairs with ``normal = 0``.

Statement codes (reference ``KernelBuilder.py:144-163``):

``directions``        ``-2`` the LHS is a declared input; ``-1`` not directional; ``1..3`` sweep axis.
``struct_inclusion``  ``-1`` scalar pseudo-statement ``normal = v``; ``0`` no loop over ``var``;
                      ``1`` loop ``var < n_real``; ``2`` loop ``var < n_real + n_aux``.

Differences from the reference at HEAD, both deliberate:

* HEAD subtracts 1 from every spatial index (and from ``patch``) of the *second declared item*
  (``KernelBuilder.py:217-218``) -- a half-finished move towards ExaHyPE2's un-haloed ``QOut`` that corrupts the
  ``Batched_stateless.py`` kernel (``Q_copy[patch - 1, i - 1, ...]``).  Here an item is shifted only when it was
  declared ``item(..., haloed=False)``, by ``halo_size`` on the spatial indices and never on ``patch``.  Set
  ``KernelBuilder.reference_head_quirks = True`` to get HEAD's behaviour verbatim (used by the golden tests).
* ``function()`` accepts ``body=`` so a CUDA functor can be emitted (the reference keeps bodies in external C++).
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional

from sympy import Idx, IndexedBase, symbols, sympify
from sympy.codegen.ast import none
from sympy.core.basic import Basic

from .TypedFunction import TypedFunction

_AXIS_SUFFIX = ("_patch", "_x", "_y", "_z")  # indexed by sweep axis; axis 0 is the patch index
_RELATIVE_ACCESS = re.compile(r"([A-Za-z_]\w*)\[\s*(-?\d+)\s*\]")


def viable(dim: int, patch_size: int, halo_size: int) -> bool:
    """Accepts exactly what the reference accepts (``KernelBuilder.py:41-48``)."""
    return dim in (2, 3) and patch_size >= 1 and halo_size >= 0


class KernelBuilder:
    #: reproduce HEAD's "second declared item is shifted by one" rewriting (see module docstring)
    reference_head_quirks = False

    def __init__(self, dim: int, patch_size: int, halo_size: int, n_real: int, n_aux: int,
                 n_patches: int = 1):
        if not viable(dim, patch_size, halo_size):
            raise Exception('check viability of inputs')
        self.dim, self.patch_size, self.halo_size = dim, patch_size, halo_size
        self.n_patches, self.n_real, self.n_aux = n_patches, n_real, n_aux

        side = (0, patch_size + 2 * halo_size)
        spatial = {name: Idx(name, side) for name in ("i", "j", "k")}
        self.all_items: Dict[str, object] = dict(spatial)
        self.all_items["patch"] = Idx("patch", (0, n_patches))
        self.all_items["var"] = Idx("var", (0, n_real + n_aux))
        axes = ["i", "j"] + (["k"] if dim == 3 else [])
        # printers iterate this list level by level: patch, i, j, (k), var
        self.indexes = [Idx(n) for n in ["patch"] + axes + ["var"]]
        self.default_shape = [n_patches] + [side] * dim

        self.literals: List[str] = []           # C++ lines such as 'int dim = 2;'
        self.parents: Dict[str, str] = {}       # name -> owning object (ExaHyPE2 CellData members)
        self.inputs: List[str] = []
        self.input_types: List[str] = []
        self.items: List[str] = []
        self.directional_items: List[str] = []
        self.directional_consts: Dict[str, list] = {}
        self.functions: List[str] = []
        self.item_struct: Dict[str, int] = {}   # 0 scalar per cell, 1 n_real, 2 n_real + n_aux
        self.unhaloed_items: set = set()

        self.LHS: list = []
        self.RHS: list = []
        self.directions: List[int] = []
        self.struct_inclusion: List[int] = []

        for name, value in (("dim", dim), ("patch_size", patch_size), ("halo_size", halo_size),
                            ("n_real", n_real), ("n_aux", n_aux)):
            self.const(name, define=f"int {name} = {value};")

    # ------------------------------------------------------------------ declarations
    def const(self, expr: str, in_type: str = "double", parent: Optional[Basic] = None, define=None):
        self.all_items[expr] = symbols(expr)
        if parent is not None:
            self.parents[expr] = str(parent)
        elif define is not None:
            self.literals.append(define)
        else:
            self.inputs.append(expr)
            self.input_types.append(in_type)
            return symbols(expr, real=True)
        return symbols(expr)

    def directional_const(self, expr: str, vals):
        if len(vals) != self.dim:
            raise Exception("directional constant must have values for each direction")
        self.directional_consts[expr] = vals
        sym = symbols(expr, real=True)
        self.all_items[expr] = sym
        return sym

    def item(self, expr: str, struct: bool = True, in_type: str = "double*", parent=None,
             haloed: bool = True):
        self.items.append(expr)
        base = IndexedBase(expr, real=True)
        self.all_items[expr] = base
        if len(self.items) == 1:      # only the first item contributes to the signature (reference :115-116)
            self.input_types.append(in_type)
        self.item_struct[expr] = 2 if struct else 0
        if parent is not None:
            self.parents[expr] = str(parent)
        if not haloed:
            self.unhaloed_items.add(expr)
        return base

    def directional_item(self, expr: str, struct: bool = True):
        self.directional_items.append(expr)
        width = 1 if struct else 0
        self.item_struct[expr] = width
        for suffix in _AXIS_SUFFIX[1:1 + self.dim]:
            self.all_items[expr + suffix] = IndexedBase(expr + suffix, real=True)
            self.item_struct[expr + suffix] = width
        return IndexedBase(expr, real=True)

    def function(self, expr: str, parent: Optional[Basic] = None, parameter_types: Optional[list] = None,
                 return_type=none, body=None):
        if parent is not None:
            self.parents[expr] = str(parent)
        self.functions.append(expr)
        func = TypedFunction(expr)
        func.returnType(return_type)
        func.parameterTypes(list(parameter_types) if parameter_types is not None else [])
        if body is not None:
            func.deviceBody(body)
        self.all_items[expr] = func
        return func

    # ------------------------------------------------------------------ statements
    def _var_extent_code(self, LHS, RHS, struct: bool) -> int:
        lhs_base = str(LHS).partition('[')[0]
        if struct:
            return 1
        if str(type(LHS)) in self.functions or str(type(RHS)) in self.functions:
            return 0                      # a call handles its own variables
        if lhs_base in self.inputs:
            return 2
        text = str(LHS) + str(RHS)        # substring match, as the reference does: this is how a scalar-per-cell
        return min(w for name, w in self.item_struct.items() if name in text)   # item silences the var loop

    def single(self, LHS: Basic, RHS: Optional[Basic] = None, direction: int = -1, struct: bool = False):
        self.struct_inclusion.append(self._var_extent_code(LHS, RHS, struct))
        self.directions.append(-2 if str(LHS).partition('[')[0] in self.inputs else direction)
        self.LHS.append(self.index(LHS, direction))
        self.RHS.append(self.index(RHS, direction))

    def directional(self, LHS: Basic, RHS: Optional[Basic] = None, struct: bool = False):
        text = str(LHS) + "\0" + str(RHS)
        for axis in range(1, self.dim + 1):
            for name, vals in self.directional_consts.items():
                if name in text:          # 'normal = v' pseudo-statement ahead of the sweep
                    self.LHS.append(self.all_items[name])
                    self.RHS.append(vals[axis - 1])
                    self.struct_inclusion.append(-1)
                    self.directions.append(-1)
            self.single(LHS, RHS, axis, struct)

    # ------------------------------------------------------------------ relative -> absolute indices
    def _absolute_text(self, name: str, offset: int, direction: int) -> str:
        """Text of the fully indexed access for the relative access ``name[offset]`` on sweep axis ``direction``."""
        if name in self.directional_items:
            if direction < 0:
                raise Exception(f"directional item '{name}' used in a non-directional statement")
            name += _AXIS_SUFFIX[direction]
        if name not in self.all_items:
            raise Exception(f"'{name}' was not declared with item()/directional_item()")
        if self.reference_head_quirks:
            shift_spatial = shift_patch = (1 if len(self.items) > 1 and name == self.items[1] else 0)
        else:
            shift_spatial, shift_patch = (self.halo_size if name in self.unhaloed_items else 0), 0
        full = []
        for level, idx in enumerate(self.indexes):
            term = str(idx)
            if term == "var":
                if self.item_struct.get(name, 2) != 0:
                    full.append(term)
                continue
            if term == "patch":
                term += f"-{shift_patch}" if shift_patch else ""
            elif self.reference_head_quirks:
                # HEAD drops the shift on the sweep axis whenever the access is offset (KernelBuilder.py:204-218)
                if level == direction and offset != 0:
                    term += f"{offset:+d}"
                elif shift_spatial:
                    term += f"-{shift_spatial}"
            else:
                total = (offset if level == direction else 0) - shift_spatial
                term += f"{total:+d}" if total else ""
            full.append(term)
        return f"{name}[{','.join(full)}]"

    def index(self, expr_in, direction: int = -1):
        """Expand every relative ``X[n]`` in ``expr_in`` for sweep axis ``direction`` (-1: no sweep).

        Like the reference (``KernelBuilder.py:175-227``) the expansion is done on the statement's text and parsed
        back with ``sympify``.  That round trip is part of the semantics, not an implementation detail: it
        re-canonicalises the expression (``-(-Q[i-1] + Q[i])*max(..)`` becomes ``(Q[i-1] - Q[i])*max(..)``), and since
        printers emit ``str(statement)`` the canonical form fixes the floating-point evaluation order.
        """
        if expr_in is None:
            return None
        text = str(expr_in)
        if text == '':
            return ''
        expanded = _RELATIVE_ACCESS.sub(
            lambda m: self._absolute_text(m.group(1), int(m.group(2)), direction), text)
        return sympify(expanded, locals=self.all_items)
