"""``CUDAPrinter`` -- turns a :class:`KernelBuilder` declaration into a CUDA translation unit for sm_100a.

Where the reference's ``CPPPrinter`` (``exahype/printers/CPPPrinter.py:48-102``) prints one serial loop nest per
statement plus five heap temporaries, this back-end recognises the *program* those statements form -- the batched
stateless finite-volume Rusanov update of ``examples/Batched_stateless.py:25-35`` -- and emits

* ``__device__`` functors for the declared flux / eigenvalue functions (hand-written family, SymPy expressions, or
  the user's own device source with the reference's ``Functions.h`` signatures),
* an update functor whose two expressions are printed from the declaration's own statements in SymPy's ``str``
  order, i.e. the evaluation order the reference's generated C++ has, and
* an ``extern "C"`` entry that instantiates a hand-written kernel template for the declared geometry: row marching
  (``csrc/fv2d_march_kernel.cuh``) for 2-D patches whose side divides the warp, warp-per-patch plane marching
  (``csrc/fv3d_pair_kernel.cuh``) for 8x8x8 patches with one halo layer, plane marching by warp groups
  (``csrc/fv3d_march_kernel.cuh``) for the other 3-D patches, thread per cell (``csrc/fv_patch_kernel.cuh``) otherwise.

``.code`` / ``.file()`` / ``.here()`` / ``.loop()`` follow ``CodePrinter`` (reference ``CodePrinter.py:46-71``);
``.build()`` compiles the unit with nvcc and returns a callable bound through ctypes.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import subprocess
import tempfile
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import sympy
from sympy import Idx, Indexed, Symbol
from sympy.printing.c import C99CodePrinter
from sympy.printing.str import StrPrinter

from ..KernelBuilder import KernelBuilder
from ..TypedFunction import DeviceBody
from .CodePrinter import CodePrinter

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc")
INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "include")

_SMEM_LIMIT = 227 * 1024


class UnsupportedKernel(NotImplementedError):
    """The declaration is not the batched stateless Rusanov patch update this back-end accelerates."""


# ---------------------------------------------------------------------------------------------------------------
@dataclass
class RusanovProgram:
    """What the statement list means, after pattern recognition."""
    q_in: str                     # declared input / output item (reference: 'Q')
    q_work: str                   # working copy (reference: 'Q_copy')
    flux_tmp: str                 # directional item receiving the flux
    eigen_tmp: str                # directional item receiving the eigenvalue
    flux_fn: str
    eigen_fn: str
    max_fn: Optional[str]
    normal: str                   # directional const passed to the functions
    normals: List[int]            # its value per axis
    dt: Optional[str]
    flux_update: str              # device expression in (qc, f_plus, f_minus)
    dissipation: str              # device expression in (qc, q0, q_plus, q_minus, l0, l_plus, l_minus, dt)
    dissipation_all: bool         # False: variable 0 only, as the reference emits
    roles: List[str] = field(default_factory=list)   # one label per statement, for the listing
    # what each argument of the flux / eigenvalue call is, in call order: 'Q' the cell's variables, 'F' the flux output,
    # 'normal', and -- for ExaHyPE2's solver signature flux(Q, x, h, t, dt, normal, F) (kernel-generator.py:37-39) -- the
    # cell context: 'x' volume centre, 'h' volume size, 'X' patch centre, 'H' patch size, 't', 'dt'
    flux_args: List[str] = field(default_factory=lambda: ["Q", "normal", "F"])
    eigen_args: List[str] = field(default_factory=lambda: ["Q", "normal"])
    # optional source term: `sourceTerm(Q[0], S[0])` on the ORIGINAL state followed by
    # `Q_copy[0] = Q_copy[0] + dt*S[0]`, after the dissipation statements
    source_fn: Optional[str] = None
    source_tmp: Optional[str] = None
    source_update: Optional[str] = None       # device expression in (qc, s, dt)

    @property
    def context(self) -> bool:
        return any(a not in ("Q", "normal", "F") for a in self.flux_args + self.eigen_args)


class _StatementPrinter(StrPrinter):
    """Prints a statement's RHS as a C++ expression in SymPy ``str`` order -- the order the reference's CPPPrinter
    emits (it prints ``str(expr)``) -- with array accesses replaced by the kernel template's local names."""

    printmethod = "_exahype_device_str"   # Indexed._sympystr would otherwise bypass _print_Indexed

    def __init__(self, names: Dict[Indexed, str], max_fn: Optional[str]):
        super().__init__()
        self._names = names
        self._max_fn = max_fn

    def _print_Indexed(self, expr):
        if expr not in self._names:
            raise UnsupportedKernel(f"unexpected array access {expr} in an update statement")
        return self._names[expr]

    def _print_Float(self, expr):
        return f"T({repr(float(expr))})"

    def _print_Integer(self, expr):
        return str(int(expr))

    def _print_Idx(self, expr):
        return str(expr.label)

    def _print_Rational(self, expr):
        return f"(T({int(expr.p)})/T({int(expr.q)}))"

    def _print_Pow(self, expr, rational=False):
        raise UnsupportedKernel("powers are not supported in update statements (the reference emits str(expr))")

    def _print_Function(self, expr):
        name = expr.func.__name__
        args = ", ".join(self._print(a) for a in expr.args)
        if name == self._max_fn and len(expr.args) == 2:
            return f"::exahype::fv_max({args})"          # reference: double max(double*, double*), Functions.cpp:64-66
        raise UnsupportedKernel(f"call to {name} inside an update statement")


def _offset_on_axis(access: Indexed, kernel: KernelBuilder, axis: int) -> int:
    """Offset of ``access`` along sweep axis ``axis`` (1-based); every other index must be un-shifted."""
    off = 0
    for level, idx in enumerate(access.indices):
        if level >= 1 + kernel.dim:
            break
        base = kernel.all_items[str(kernel.indexes[level])]
        delta = sympy.simplify(idx - base)
        if not delta.is_Integer:
            raise UnsupportedKernel(f"index {idx} of {access} is not a constant offset")
        if int(delta) != 0:
            if level != axis:
                raise UnsupportedKernel(f"{access} is shifted off the sweep axis")
            off = int(delta)
    return off


def analyse(kernel: KernelBuilder) -> RusanovProgram:
    """Recognise the Rusanov patch-update program in ``kernel``'s statement list, or raise."""
    k = kernel
    if len(k.items) < 2 or len(k.directional_items) < 2:
        raise UnsupportedKernel("expected two items (input, working copy) and two directional items (flux, eigenvalue)")
    stmts = list(zip(k.LHS, k.RHS, k.directions, k.struct_inclusion))
    roles: List[str] = []
    dim = k.dim

    def base_of(e) -> str:
        return str(e.base) if isinstance(e, Indexed) else ""

    def fn_name(e) -> str:
        return e.func.__name__ if isinstance(e, sympy.Function) and e.func.__name__ in k.functions else ""

    # --- copy-in / copy-out ------------------------------------------------------------------------------------
    first, last = stmts[0], stmts[-1]
    if not (isinstance(first[0], Indexed) and isinstance(first[1], Indexed) and first[2] in (-1, -2)):
        raise UnsupportedKernel("first statement must copy the input into the working item")
    q_work, q_in = base_of(first[0]), base_of(first[1])
    if not (isinstance(last[0], Indexed) and base_of(last[0]) == q_in and base_of(last[1]) == q_work):
        raise UnsupportedKernel("last statement must copy the working item back into the input")

    flux_calls, eig_calls, flux_upd, diss_upd = {}, {}, {}, {}
    source_call: Dict[str, object] = {}
    normal_name, normals = None, {}
    pending_normal = None
    # scalar members of the CellData object (kernel-generator.py:15-19: dt, t, cellCentre, cellSize with parent=Data)
    members = {n for n in k.parents if isinstance(k.all_items.get(n), Symbol)}
    flux_arg_tags, eigen_arg_tags = {}, {}

    def classify(arg, is_flux_call: bool) -> str:
        """What one argument of a flux / eigenvalue call is (see RusanovProgram.flux_args)."""
        if isinstance(arg, Indexed):
            b = base_of(arg)
            if b == q_work:
                return "Q"
            if is_flux_call and any(b == d + sfx for d in k.directional_items for sfx in ("_x", "_y", "_z")):
                return "F"
            raise UnsupportedKernel(f"unexpected array argument {arg} in a flux / eigenvalue call")
        if isinstance(arg, Symbol):
            name = str(arg)
            if name in k.directional_consts:
                return "normal"
            if name in members or name in k.inputs:
                tag = {"dt": "dt", "t": "t", "cellCentre": "X", "cellSize": "H"}.get(name)
                if tag:
                    return tag
            raise UnsupportedKernel(f"argument {arg} of a flux / eigenvalue call is not available inside the kernel")
        if isinstance(arg, sympy.Function):
            name = arg.func.__name__
            # ExaHyPE2's exahype2::fv::getVolumeCentre(x, h, patch_size, index) / getVolumeSize(h, patch_size); without
            # the index the declaration passes the patch centre (kernel-generator.py:39)
            if name == "getVolumeCentre":
                return "x" if any(isinstance(a, (sympy.Set, sympy.Tuple, set, frozenset)) for a in arg.args) else "X"
            if name == "getVolumeSize":
                return "h"
        raise UnsupportedKernel(f"argument {arg} of a flux / eigenvalue call is not supported")

    for pos, (lhs, rhs, direction, struct) in enumerate(stmts):
        if pos == 0:
            roles.append("copy-in"); continue
        if pos == len(stmts) - 1:
            roles.append("copy-out"); continue
        if struct == -1 and isinstance(lhs, Symbol) and str(lhs) in k.directional_consts:
            normal_name = str(lhs); pending_normal = int(rhs); roles.append(f"{lhs} = {rhs}"); continue
        if fn_name(lhs) and rhs is None and direction >= 1:                # Flux(Qc[c], normal, F_d[c])
            args = lhs.args
            if len(args) < 3 or base_of(args[0]) != q_work or not isinstance(args[-1], Indexed):
                raise UnsupportedKernel(f"flux call {lhs} must be f({q_work}[c], {normal_name}, tmp[c])")
            flux_calls[direction] = (fn_name(lhs), base_of(args[-1]), pending_normal)
            flux_arg_tags[direction] = [classify(a, True) for a in args]
            roles.append(f"flux axis {direction}"); continue
        if isinstance(lhs, Indexed) and fn_name(rhs):                      # L_d[c] = maxEigenvalue(Qc[c], normal)
            if base_of(rhs.args[0]) != q_work:
                raise UnsupportedKernel(f"eigenvalue call {rhs} must read {q_work}[c]")
            eig_calls[direction] = (fn_name(rhs), base_of(lhs), pending_normal)
            eigen_arg_tags[direction] = [classify(a, False) for a in rhs.args]
            roles.append(f"eigenvalue axis {direction}"); continue
        if isinstance(lhs, Indexed) and base_of(lhs) == q_work and isinstance(rhs, sympy.Expr):
            used = {base_of(a) for a in rhs.atoms(Indexed)}
            if any(b.startswith(tuple(k.directional_items)) for b in used) and q_in not in used:
                flux_upd[direction] = (lhs, rhs, struct); roles.append(f"flux update axis {direction}"); continue
            if q_in in used:
                diss_upd[direction] = (lhs, rhs, struct); roles.append(f"dissipation axis {direction}"); continue
        if fn_name(lhs) and rhs is None and direction < 0:                  # sourceTerm(Q[c], S[c])
            args = lhs.args
            if len(args) != 2 or base_of(args[0]) != q_in or not isinstance(args[1], Indexed) or base_of(args[1]) in (q_in, q_work):
                raise UnsupportedKernel(f"source call {lhs} must be f({q_in}[c], tmp[c]): the source is evaluated on the original state")
            if source_call:
                raise UnsupportedKernel("more than one source call")
            source_call.update(fn=fn_name(lhs), tmp=base_of(args[1]))
            roles.append("source call"); continue
        if isinstance(lhs, Indexed) and base_of(lhs) == q_work and isinstance(rhs, sympy.Expr) and source_call \
                and source_call["tmp"] in {base_of(a) for a in rhs.atoms(Indexed)} and direction < 0:
            source_call["update"] = (lhs, rhs, struct)
            roles.append("source update"); continue
        raise UnsupportedKernel(f"statement {pos} ({lhs} = {rhs}) is not part of the Rusanov patch update")

    axes = list(range(1, dim + 1))
    for table, what in ((flux_calls, "flux call"), (eig_calls, "eigenvalue call"), (flux_upd, "flux update"),
                        (diss_upd, "dissipation update")):
        if sorted(table) != axes:
            raise UnsupportedKernel(f"need exactly one {what} per axis, got axes {sorted(table)}")
    flux_fn = {v[0] for v in flux_calls.values()}
    eigen_fn = {v[0] for v in eig_calls.values()}
    if len(flux_fn) != 1 or len(eigen_fn) != 1:
        raise UnsupportedKernel("one flux and one eigenvalue function expected")
    suffix = ("", "_x", "_y", "_z")
    flux_tmp = flux_calls[1][1][: -len(suffix[1])]
    eigen_tmp = eig_calls[1][1][: -len(suffix[1])]
    for d in axes:
        if flux_calls[d][1] != flux_tmp + suffix[d] or eig_calls[d][1] != eigen_tmp + suffix[d]:
            raise UnsupportedKernel("directional temporaries are not used consistently across axes")
        normals[d] = flux_calls[d][2]
        if normals[d] is None or eig_calls[d][2] != normals[d]:
            raise UnsupportedKernel("the directional constant must be set ahead of each sweep")
    if [normals[d] for d in axes] != list(range(dim)):
        raise UnsupportedKernel("kernel template pairs sweep axis d with normal d-1 (reference Batched_stateless.py:17)")
    for tags, what in ((flux_arg_tags, "flux"), (eigen_arg_tags, "eigenvalue")):
        if any(tags[d] != tags[1] for d in axes):
            raise UnsupportedKernel(f"the {what} call takes different arguments on different axes")
        if tags[1].count("Q") != 1 or tags[1].count("normal") != 1 or (what == "flux" and tags[1][-1] != "F"):
            raise UnsupportedKernel(f"the {what} call needs the cell's variables and the normal once each"
                                    + (", and the flux output last" if what == "flux" else ""))

    # --- update expressions, printed per axis and required to be the same program on every axis -----------------
    max_candidates = [f for f in k.functions if f not in (next(iter(flux_fn)), next(iter(eigen_fn)))]
    dt_names = [n for n in k.inputs] + sorted(members)

    def names_for(rhs, d):
        names = {}
        for a in rhs.atoms(Indexed):
            b, off = base_of(a), _offset_on_axis(a, k, d)
            tag = {0: "0", 1: "plus", -1: "minus"}.get(off)
            if tag is None:
                raise UnsupportedKernel(f"{a}: only offsets -1, 0, +1 are supported (one halo layer)")
            if b == q_work and off == 0:
                names[a] = "qc"
            elif b == q_in:
                names[a] = "q0" if off == 0 else f"q_{tag}"
            elif b == flux_tmp + suffix[d] and off != 0:
                names[a] = f"f_{tag}"
            elif b == eigen_tmp + suffix[d]:
                names[a] = "l0" if off == 0 else f"l_{tag}"
            else:
                raise UnsupportedKernel(f"unexpected access {a} in the update along axis {d}")
        return names

    def printed(table, allowed_syms):
        texts = set()
        max_fn = None
        for d in axes:
            lhs, rhs, _ = table[d]
            if _offset_on_axis(lhs, k, d) != 0:
                raise UnsupportedKernel("updates must write the centre cell")
            used_fns = {f.func.__name__ for f in rhs.atoms(sympy.Function) if f.func.__name__ in k.functions}
            if len(used_fns) > 1 or not used_fns <= set(max_candidates):
                raise UnsupportedKernel(f"unexpected function calls {used_fns} in an update")
            max_fn = next(iter(used_fns)) if used_fns else max_fn
            inside_accesses = set().union(*[a.free_symbols for a in rhs.atoms(Indexed)]) if rhs.atoms(Indexed) else set()
            for sym in rhs.free_symbols - inside_accesses:
                if str(sym) not in allowed_syms:
                    raise UnsupportedKernel(f"symbol {sym} is not available inside the kernel")
            texts.add(_StatementPrinter(names_for(rhs, d), max_fn).doprint(rhs))
        if len(texts) != 1:
            raise UnsupportedKernel(f"the update differs between axes: {sorted(texts)}")
        return texts.pop(), max_fn

    flux_text, _ = printed(flux_upd, set())
    diss_text, max_fn = printed(diss_upd, set(dt_names))
    dt = next((n for n in dt_names if Symbol(n) in diss_upd[1][1].free_symbols
               or any(str(s) == n for s in diss_upd[1][1].free_symbols)), None)
    if dt is not None and dt != "dt":
        diss_text = diss_text.replace(dt, "dt")

    # var extent of the dissipation exactly as the reference's printer decides it (CPPPrinter.py:118-126):
    # min over the statement's struct code and every item_struct name that is a substring of the statement
    lhs, rhs, struct = diss_upd[1]
    widths = [w for name, w in k.item_struct.items() if name in str([lhs, rhs])] + [struct]
    dissipation_all = min(widths) >= 1

    source_text = None
    if source_call:
        if "update" not in source_call:
            raise UnsupportedKernel("a source call needs the statement that adds dt*S to the working item")
        if roles.index("source call") < max(i for i, r in enumerate(roles) if r.startswith("dissipation")):
            raise UnsupportedKernel("the source statements follow the dissipation statements")
        lhs, rhs, struct = source_call["update"]
        names = {}
        for a in rhs.atoms(Indexed):
            if any(sympy.simplify(idx - k.all_items[str(k.indexes[level])]) != 0
                   for level, idx in enumerate(a.indices[: 1 + dim])):
                raise UnsupportedKernel(f"{a}: the source update reads the centre cell only")
            names[a] = {q_work: "qc", source_call["tmp"]: "s"}.get(base_of(a))
            if names[a] is None:
                raise UnsupportedKernel(f"unexpected access {a} in the source update")
        for sym in rhs.free_symbols - set().union(*[a.free_symbols for a in rhs.atoms(Indexed)]):
            if str(sym) not in dt_names:
                raise UnsupportedKernel(f"symbol {sym} is not available inside the kernel")
        source_text = _StatementPrinter(names, None).doprint(rhs)
        if dt is not None and dt != "dt":
            source_text = source_text.replace(dt, "dt")
        if min([w for name, w in k.item_struct.items() if name in str([lhs, rhs])] + [struct]) < 1:
            raise UnsupportedKernel("the source update must run over the unknowns (struct=True)")

    return RusanovProgram(q_in=q_in, q_work=q_work, flux_tmp=flux_tmp, eigen_tmp=eigen_tmp,
                          flux_fn=next(iter(flux_fn)), eigen_fn=next(iter(eigen_fn)), max_fn=max_fn,
                          normal=normal_name or "normal", normals=[normals[d] for d in axes], dt=dt,
                          flux_update=flux_text, dissipation=diss_text, dissipation_all=dissipation_all, roles=roles,
                          flux_args=flux_arg_tags[1], eigen_args=eigen_arg_tags[1],
                          source_fn=source_call.get("fn"), source_tmp=source_call.get("tmp"), source_update=source_text)


# ---------------------------------------------------------------------------------------------------------------
def smem_bytes(dim, P, H, nr, na, G, elem, diss_all, unhaloed_unused=False) -> int:
    """Python mirror of ``FvKernelConfig::SMEM_BYTES`` (csrc/fv_patch_kernel.cuh)."""
    up = lambda x, a: (x + a - 1) // a * a
    nv, S = nr + na, P + 2 * H
    ncell, pd, pf = S ** dim, P ** dim, P ** (dim - 1)
    patch_bytes = ncell * nv * elem
    qbuf = 2 if patch_bytes % 16 == 0 else 1
    slots = (P + 2) * pf
    dv = nr if diss_all else 1
    off_f = up(qbuf * G * patch_bytes, 16)
    off_l = up(off_f + dim * nr * G * slots * elem, 16)
    off_r = up(off_l + dim * G * slots * elem, 16)
    off_out = up(off_r + (dv * G * ncell * elem if nv % 2 == 0 else 0), 128)
    off_lam = up(off_out + G * pd * nv * elem, 16)
    off_bar = up(off_lam + 2 * G * 8, 16)
    return off_bar + 16


def pick_geometry(dim, P, H, nr, na, elem):
    """(patches per tile G, threads per CTA, min CTAs/SM): about one thread per interior cell, 128..512 threads."""
    pd = P ** dim
    if pd >= 1024:
        G, nt = 1, 512
    elif pd >= 128:
        G, nt = 1, (pd + 31) // 32 * 32
    else:
        G = max(1, 256 // pd)
        nt = min(256, (G * pd + 31) // 32 * 32)
    while G > 1 and smem_bytes(dim, P, H, nr, na, G, elem, True) > _SMEM_LIMIT // 2:
        G -= 1
    if smem_bytes(dim, P, H, nr, na, G, elem, True) > _SMEM_LIMIT:
        raise UnsupportedKernel(f"a {P}^{dim} patch with {nr + na} variables does not fit 227 KB of shared memory")
    minb = max(1, min(4, _SMEM_LIMIT // smem_bytes(dim, P, H, nr, na, G, elem, True), 65536 // (nt * 128)))
    return G, nt, minb


class _DevicePrinter(C99CodePrinter):
    """SymPy expression -> C++ over the template type ``T``."""

    def _print_Float(self, expr):
        return f"T({repr(float(expr))})"

    def _print_Rational(self, expr):
        return f"(T({int(expr.p)})/T({int(expr.q)}))"

    def _print_Abs(self, expr):
        return f"::exahype::fv_abs<T>({self._print(expr.args[0])})"

    def _print_Pow(self, expr):
        b, e = expr.args
        if e == sympy.Rational(1, 2):
            return f"::exahype::fv_sqrt<T>({self._print(b)})"
        if e == -1:
            return f"(T(1.0)/({self._print(b)}))"
        if e == sympy.Rational(-1, 2):
            return f"(T(1.0)/::exahype::fv_sqrt<T>({self._print(b)}))"
        if e.is_Integer and 2 <= int(e) <= 4:
            return "(" + "*".join([f"({self._print(b)})"] * int(e)) + ")"
        return f"pow({self._print(b)}, {self._print(e)})"

    def _print_Max(self, expr):
        args = [self._print(a) for a in expr.args]
        out = args[0]
        for a in args[1:]:
            out = f"::exahype::fv_max<T>({out}, {a})"
        return out

    # the two opaque operations of strength_reduce() (SymPy routes every undefined function through this method)
    def _print_AppliedUndef(self, expr):
        name = expr.func.__name__
        if name == "exahype_recip":
            return f"(T(1.0)/({self._print(expr.args[0])}))"
        if name == "exahype_sqrt":
            return f"::exahype::fv_sqrt<T>({self._print(expr.args[0])})"
        return super()._print_Function(expr)


_RECIP = sympy.Function("exahype_recip")
_SQRT = sympy.Function("exahype_sqrt")
_HALF = sympy.Rational(1, 2)


def strength_reduce(expr):
    """Fewer divisions and roots in a SymPy-declared functor, within the 1e-12 contract of generated kernels.

    SymPy canonicalises ``sqrt(g*|p|/|rho|)`` to ``sqrt(g)*sqrt(|p|)/sqrt(|rho|)`` and keeps ``1/rho`` and ``1/|rho|`` apart: written
    out, a cell pays two reciprocals, two roots and a division where the hand-written family pays one reciprocal and one
    root.  Here every product's half-power factors are gathered under ONE root again (valid because SymPy only splits
    non-negative radicands) and ``1/|x|`` becomes ``|1/x|`` (exact: the IEEE reciprocal is sign-symmetric), both as opaque
    functions so that SymPy does not undo them and the common-subexpression pass shares ``1/rho`` between flux and
    eigenvalue."""
    def roots(e):
        if isinstance(e, sympy.Pow) and e.exp == _HALF:
            return _SQRT(e.base)
        if isinstance(e, sympy.Pow) and e.exp == -_HALF:
            return _SQRT(sympy.Pow(e.base, -1))
        if isinstance(e, sympy.Mul):
            rs = [a for a in e.args if isinstance(a, _SQRT)]
            if len(rs) > 1:
                rest = [a for a in e.args if not isinstance(a, _SQRT)]
                return sympy.Mul(*rest) * _SQRT(sympy.Mul(*[a.args[0] for a in rs]))
        return e
    expr = sympy.sympify(expr).replace(lambda e: isinstance(e, (sympy.Pow, sympy.Mul)), roots)

    def recip(e):
        if isinstance(e.base, sympy.Abs):
            return sympy.Abs(_RECIP(e.base.args[0]))
        return _RECIP(e.base)
    return expr.replace(lambda e: isinstance(e, sympy.Pow) and e.exp == -1, recip)


def plan_symbolic_functors(groups: "Dict[tuple, list]"):
    """Common-subexpression plan for SymPy-bodied functors.

    ``groups`` maps ``('F', n)`` / ``('L', n)`` to the expressions of the flux components / the eigenvalue along axis ``n``.
    A cell evaluates all of them (2*dim functor calls), and written out one by one they repeat the reciprocal of the
    density, the pressure and the sound speed in every call -- a SymPy-declared 3-D Euler kernel ran at 0.24 of the HBM
    peak where the hand-written family's per-cell cache (`Prims`) reaches 0.88.  This is that cache, derived: `sympy.cse`
    over all groups; the EXPENSIVE temporaries (a division, a root, or more than two operations) that at least two groups
    need go into `Prims` and are computed once per cell; cheap ones (one product or sum) are re-emitted where they are
    used, so that the cache stays a handful of values (the marching kernels carry it in registers from plane to plane).

    Returns ``(prims_lines, stored, bodies)``: assignments computing the stored temporaries (with everything they depend
    on), the stored symbols in struct order, and per group ``(local assignments, reduced expressions)``."""
    keys = list(groups)
    flat = [strength_reduce(e) for k in keys for e in groups[k]]
    repl, reduced = sympy.cse(flat, symbols=sympy.numbered_symbols("cse_t"), order="none")
    defs = dict(repl)
    order = [sym for sym, _ in repl]
    temps = set(order)

    def expensive(expr) -> bool:
        if any((isinstance(a, sympy.Pow) and not (a.exp.is_Integer and a.exp > 0)) or isinstance(a, (_RECIP, _SQRT))
               for a in sympy.preorder_traversal(expr)):
            return True
        return bool(sympy.count_ops(expr) > 2)
    costly = {t for t in order if expensive(defs[t])}

    def closure(exprs, stop):
        need, stack = set(), [x for e in exprs for x in e.free_symbols if x in temps]
        while stack:
            t = stack.pop()
            if t in need:
                continue
            need.add(t)
            if t not in stop:
                stack.extend(x for x in defs[t].free_symbols if x in temps)
        return need
    per_group, i = {}, 0
    for k in keys:
        per_group[k] = reduced[i:i + len(groups[k])]
        i += len(groups[k])
    full = {k: closure(per_group[k], set()) for k in keys}
    candidates = {t for t in costly if sum(t in full[k] for k in keys) >= 2}
    reach = {k: closure(per_group[k], candidates) for k in keys}
    stored = [t for t in order if t in candidates and any(t in reach[k] for k in keys)]
    stored_set = set(stored)
    prims_need = closure([sympy.Add(*stored)] if stored else [], set()) if stored else set()
    prims_lines = [(t, defs[t]) for t in order if t in prims_need]
    bodies = {}
    for k in keys:
        local = closure(per_group[k], stored_set)
        bodies[k] = ([(t, defs[t]) for t in order if t in local and t not in stored_set],
                     [t for t in stored if t in local], per_group[k])
    return prims_lines, stored, bodies


class CUDAPrinter(CodePrinter):
    """``CUDAPrinter(kernel, function_name="time_step", dtype="f64")``.

    dtype        'f64' | 'f32' -- arithmetic type of the generated entry.
    dissipation  None: as the reference's printer would emit it (variable 0 only for the reference declaration,
                 CPPPrinter.py:118-126); 'all' / 'var0' to force.
    model        'euler' | 'swe': use a committed hand-written functor family for functions declared without a body.
    template     'auto' | 'pair' | 'march' | 'cell': which hand-written kernel template the entry instantiates (see
                 module doc; 'pair' is the warp-per-patch kernel for 3-D patches of side 8, halo 1).
    """

    def __init__(self, kernel: KernelBuilder, function_name: str = "time_step", dtype: str = "f64",
                 dissipation: Optional[str] = None, model: Optional[str] = None,
                 patches_per_tile: Optional[int] = None, threads: Optional[int] = None, template: str = "auto"):
        super().__init__(kernel, function_name=function_name)
        if dtype not in ("f64", "f32"):
            raise ValueError("dtype must be 'f64' or 'f32'")
        if kernel.halo_size < 1:
            raise UnsupportedKernel("the Rusanov update reads one halo layer: halo_size must be >= 1")
        self.dtype = dtype
        self.ctype = "double" if dtype == "f64" else "float"
        self.program = analyse(kernel)
        if dissipation not in (None, "all", "var0"):
            raise ValueError("dissipation must be None, 'all' or 'var0'")
        self.dissipation_all = self.program.dissipation_all if dissipation is None else dissipation == "all"
        self.model = model
        k = kernel
        if template not in ("auto", "pair", "march", "cell"):
            raise ValueError("template must be 'auto', 'pair', 'march' or 'cell'")
        can_march = (k.dim == 2 and k.patch_size in (8, 16, 32)) or (k.dim == 3 and k.patch_size in (4, 8))
        if template == "march" and not can_march:
            raise UnsupportedKernel("the marching templates serve 2-D patches of side 8/16/32 and 3-D patches of side 4/8")
        can_pair = k.dim == 3 and k.patch_size == 8 and k.halo_size == 1
        if template == "pair" and not can_pair:
            raise UnsupportedKernel("the warp-per-patch template serves 3-D patches of side 8 with one halo layer")
        small = k.n_real + k.n_aux <= 6
        self.context = self.program.context and not self._all_builtin()
        if self.context:
            # functors that see the cell's position / time (ExaHyPE2's solver signature) are served by the thread-per-cell
            # kernel, which knows every cell's index; the committed families ignore the context and keep every template
            if template not in ("auto", "cell"):
                raise UnsupportedKernel("functors taking x / h / t / dt run in the thread-per-cell template only")
            can_pair = can_march = False
        if can_pair and (template == "pair" or (template == "auto" and small)):
            self.template = "pair"
        elif can_march and (template == "march" or (template == "auto" and small)):
            self.template = "march"
        else:
            self.template = "cell"
        elem = 8 if dtype == "f64" else 4
        # tile geometry of the thread-per-cell template (the marching templates pick theirs at compile time)
        G, nt, minb = pick_geometry(k.dim, k.patch_size, k.halo_size, k.n_real, k.n_aux, elem) if self.template == "cell" \
            else (1, 256, 1)
        self.patches_per_tile = patches_per_tile or G
        self.threads = threads or nt
        self.min_ctas = minb
        self.header_file_name: Optional[str] = None
        self._statements: List[str] = []
        self.code = self._emit()

    # ------------------------------------------------------------------ CodePrinter interface
    def loop(self, expr: list, direction: int, below: int = 0, struct_inclusion: int = 0):
        """One statement -> one line of the listing placed at the top of the unit: on the GPU a statement is not a
        loop nest but a phase of the fused kernel (see csrc/fv_patch_kernel.cuh)."""
        role = self.program.roles[len(self._statements)] if len(self._statements) < len(self.program.roles) else "?"
        lhs, rhs = expr
        text = f"{lhs};" if rhs is None or rhs == '' else f"{lhs} = {rhs};"
        phase = {"copy-in": "TMA bulk load of the tile into shared memory", "copy-out": "phase C: staged interior -> HBM"}.get(
            role, "phase A: per-cell functor" if role.startswith(("flux axis", "eigen")) else
            "phase B: register update" if "update" in role or "dissipation" in role else "compile-time axis")
        line = f"//   [{len(self._statements):2d}] dir={direction:2d} struct={struct_inclusion:2d}  {role:<22s} -> {phase}\n//        {text}\n"
        self._statements.append(line)
        return line

    def file(self, file_name: str = "time_step.cu", header_file_name: Optional[str] = None):
        """Write the unit; ``header_file_name`` is #included first and must define the declared functions as device
        code with the reference's signatures (``Unit test/Functions.h:2-4``) when they have no body attached."""
        if header_file_name != self.header_file_name:
            self.header_file_name = header_file_name
            self._statements = []
            self.code = self._emit()
        super().file(file_name, header_file_name)

    # ------------------------------------------------------------------ emission
    def _all_builtin(self) -> bool:
        fb, eb = self._body_of(self.program.flux_fn), self._body_of(self.program.eigen_fn)
        return fb is not None and bool(fb.builtin) and (eb is None or eb.builtin == fb.builtin)

    @staticmethod
    def _call_args(tags: List[str]) -> str:
        """Argument list of the user's device function, in the order the declaration calls it."""
        return ", ".join({"Q": "q", "F": "F", "normal": "N", "x": "c.x", "h": "c.h", "X": "c.X", "H": "c.H", "t": "c.t",
                          "dt": "c.dt"}[t] for t in tags)

    def _body_of(self, name: str) -> Optional[DeviceBody]:
        fn = self.kernel().all_items.get(name)
        body = getattr(fn, "device_body", None)
        if body is None and self.model is not None:
            body = DeviceBody(builtin=self.model)
        return body

    def _physics(self) -> str:
        k, p = self.kernel(), self.program
        nr, na, nv, dim = k.n_real, k.n_aux, k.n_real + k.n_aux, k.dim
        fb, eb = self._body_of(p.flux_fn), self._body_of(p.eigen_fn)
        if fb is not None and fb.builtin and (eb is None or eb.builtin == fb.builtin):
            fam = fb.builtin
            if fam == "euler":
                return f"using Physics = ::exahype::EulerPhysics<{dim}, {nr}, {na}>;   // csrc/physics.cuh\n"
            if fam in ("swe", "swe_source"):
                if dim != 2:
                    raise UnsupportedKernel("the shallow-water family is 2-D")
                if (fam == "swe_source") != bool(p.source_fn):
                    raise UnsupportedKernel("model 'swe_source' goes with a declaration that has the source statements, 'swe' with one that has none")
                cls = "SweSourcePhysics" if fam == "swe_source" else "SwePhysics"
                return f"using Physics = ::exahype::{cls}<{nr}, {na}>;   // csrc/physics.cuh\n"
            raise UnsupportedKernel(f"unknown builtin physics '{fam}'")

        ctx = self.context
        q = sympy.symbols(f"q0:{nv}", real=True)
        prn = _DevicePrinter()

        def lower(expr):
            text = prn.doprint(sympy.sympify(expr))
            for v in reversed(range(nv)):
                text = re.sub(rf"\bq{v}\b", f"q[{v}]", text)
            return text

        # SymPy bodies: one common-subexpression plan over every functor call of a cell (plan_symbolic_functors)
        groups = {}
        if fb is not None and fb.expressions:
            for n in range(dim):
                comps = [sympy.sympify(c) for c in fb.expressions(list(q), n)]
                if len(comps) != nr:
                    raise ValueError(f"{p.flux_fn}: expected {nr} flux components, got {len(comps)}")
                groups[("F", n)] = comps
        if eb is not None and eb.expressions:
            for n in range(dim):
                groups[("L", n)] = [sympy.sympify(eb.expressions(list(q), n))]
        prims_lines, stored, bodies = plan_symbolic_functors(groups) if groups else ([], [], {})
        slot = {t: i for i, t in enumerate(stored)}

        def body(key, indent="      "):
            local, used, exprs = bodies[key]
            text = "".join(f"{indent}const T {t} = pr.t[{slot[t]}];\n" for t in used)
            text += "".join(f"{indent}const T {t} = {lower(e)};\n" for t, e in local)
            return text, exprs
        out = [f"struct Physics {{\n  static constexpr int NR = {nr}, NA = {na}, NV = {nv};\n"
               + ("  static constexpr bool NEEDS_CONTEXT = true;   // functors take the cell's x / h / t / dt (FvCellCtx)\n" if ctx else "")]
        if stored:
            out.append(f"  // per-cell cache shared by the {len(groups)} functor calls of a cell (derived by common-subexpression elimination)\n"
                       f"  template <typename T> struct Prims {{ T t[{len(stored)}]; }};\n"
                       "  template <typename T> static __device__ __forceinline__ Prims<T> prims(const T (&q)[NV]) {\n"
                       "    Prims<T> pr;\n"
                       + "".join(f"    const T {t} = {lower(e)};\n" for t, e in prims_lines)
                       + "".join(f"    pr.t[{i}] = {t};\n" for i, t in enumerate(stored))
                       + "    return pr;\n  }\n")
        else:
            out.append("  template <typename T> struct Prims {};\n"
                       "  template <typename T> static __device__ __forceinline__ Prims<T> prims(const T (&)[NV]) { return {}; }\n")
        ctx_tpl = ", class Ctx" if ctx else ""
        # a body given as device source lives in `namespace user` (below); functions that come from the header passed to
        # file() are global -- qualified either way, so that a declared name such as `flux` cannot hit the functor's members
        scope = lambda body: "user::" if (body is not None and body.source) else "::"
        ctx_arg = ", const Ctx& c" if ctx else ""

        # flux
        out.append(f"  template <int N, typename T{ctx_tpl}>\n"
                   f"  static __device__ __forceinline__ void flux(const T (&q)[NV], const Prims<T>& pr, T (&F)[NR]{ctx_arg}) {{\n")
        if ctx and ((fb is not None and fb.expressions) or (eb is not None and eb.expressions)):
            raise UnsupportedKernel("functors taking x / h / t / dt need a device-source body (DeviceBody(source=...))")
        if fb is not None and fb.expressions:
            for n in range(dim):
                pre_text, comps = body(("F", n))
                out.append(f"    if (N == {n}) {{\n" + pre_text + "".join(f"      F[{v}] = {lower(c)};\n" for v, c in enumerate(comps)) + "    }\n")
        else:   # user's device function with the declared signature: void Flux(const T* Q, int normal, T* F) in the
            # reference's Functions.h:2, flux(Q, x, h, t, dt, normal, F) for an ExaHyPE2 solver
            out.append(f"    {scope(fb)}{p.flux_fn}({self._call_args(p.flux_args)});\n")
        out.append("  }\n")
        # eigenvalue
        out.append(f"  template <int N, typename T{ctx_tpl}>\n"
                   f"  static __device__ __forceinline__ T eigen(const T (&q)[NV], const Prims<T>& pr{ctx_arg}) {{\n")
        if eb is not None and eb.expressions:
            for n in range(dim):
                pre_text, exprs = body(("L", n))
                out.append(f"    if (N == {n}) {{\n" + pre_text + f"      return {lower(exprs[0])};\n    }}\n")
            out.append("    return T(0);\n")
        else:
            out.append(f"    return {scope(eb)}{p.eigen_fn}({self._call_args(p.eigen_args)});\n")
        out.append("  }\n")
        if p.source_fn:           # void sourceTerm(const T* Q, T* S) in the style of Functions.h
            sb = self._body_of(p.source_fn)
            if sb is not None and (sb.expressions or sb.builtin):
                raise UnsupportedKernel("a generated source term needs a device-source body (or model='swe_source')")
            out.append("  static constexpr bool HAS_SOURCE = true;\n"
                       "  template <typename T>\n"
                       "  static __device__ __forceinline__ void source(const T (&q)[NV], T (&S)[NR]) {\n"
                       f"    {scope(sb)}{p.source_fn}(q, S);\n  }}\n")
        out.append("};\n")
        pre = ""
        for b in (fb, eb, self._body_of(p.source_fn) if p.source_fn else None):
            if b is not None and b.source and b.source not in pre:
                pre += b.source.rstrip() + "\n\n"
        # the user's functions live in their own namespace: a declared name such as `flux` must not collide with the
        # functor's members
        return "namespace user {\n" + pre + "}  // namespace user\n\n" + "".join(out)

    def _emit(self) -> str:
        k, p = self.kernel(), self.program
        self._statements = []
        for lhs, rhs, d, s in zip(k.LHS, k.RHS, k.directions, k.struct_inclusion):
            self.loop([lhs, rhs], d, k.dim + 1, s)
        T = self.ctype
        fname = self.functionName()
        b = lambda x: str(bool(x)).lower()
        geo = f"Physics, Update, {T}, {k.patch_size}, {k.halo_size}"
        if self.template == "cell":
            header = "fv_patch_kernel.cuh"
            launcher = lambda da, uh, gather=False: (
                f"::exahype::FvLauncher<::exahype::FvKernelConfig<Physics, Update, {T}, {k.dim}, "
                f"{k.patch_size}, {k.halo_size}, {self.patches_per_tile}, {self.threads}, "
                f"{self.min_ctas}, {b(da)}, {b(uh)}, {b(gather)}>>")
        elif k.dim == 2:
            header = "fv2d_march_kernel.cuh"
            launcher = lambda da, uh, gather=False: (
                f"::exahype::Fv2dMarchGather<{geo}, {b(da)}, {b(uh)}>" if gather else
                f"::exahype::Fv2dMarchAuto<{geo}, {b(da)}, {b(uh)}>")
        elif self.template == "pair":
            header = "fv3d_pair_kernel.cuh"
            launcher = lambda da, uh, gather=False: (
                f"::exahype::Fv3dPairLauncher<typename ::exahype::Fv3dPairAutoConfig<{geo}, {b(da)}, {b(uh)}>::gather_type>"
                if gather else f"::exahype::Fv3dPairAuto<{geo}, {b(da)}, {b(uh)}>")
        else:
            header = "fv3d_march_kernel.cuh"
            launcher = lambda da, uh, gather=False: (
                f"::exahype::Fv3dMarchLauncher<typename ::exahype::Fv3dMarchAutoConfig<{geo}, {b(da)}, {b(uh)}>::gather_type>"
                if gather else f"::exahype::Fv3dMarchAuto<{geo}, {b(da)}, {b(uh)}>")
        da = self.dissipation_all
        parts = []
        if self.header_file_name:
            parts.append(f'#include "{self.header_file_name}"\n')
        parts.append(
            "// Generated by exahype.printers.CUDAPrinter -- do not edit.\n"
            f"// Kernel: dim={k.dim} patch_size={k.patch_size} halo_size={k.halo_size} n_real={k.n_real} n_aux={k.n_aux}; "
            f"dtype={self.dtype}; dissipation={'all' if da else 'var0'}; kernel template: {header}\n"
            "// Statement list (KernelBuilder) and where each statement went in the fused kernel:\n"
            + "".join(self._statements) +
            f"#include <stdint.h>\n#include <cuda_runtime.h>\n#include \"{header}\"\n#include \"exahype_cuda.h\"\n\nnamespace {{\n\n")
        parts.append(self._physics())
        # The reference declaration's two statements, character for character: the hand-written functor states the same
        # arithmetic in its cheaper bit-identical form (single-rounding FMAs for the +-0.5*F terms, one max(L, L') per pair
        # of cells; csrc/physics.cuh RusanovUpdate) -- a generated kernel for the reference's own declaration then IS the
        # committed instantiation (C3: 0.320 -> 0.307 ms).  Any other text is emitted as written.
        reference_update = (p.flux_update == "qc - T(0.5)*f_plus + T(0.5)*f_minus" and p.dissipation ==
                            "T(0.5)*dt*((-q_plus + q0)*::exahype::fv_max(l_plus, l0) + (q_minus - q0)*::exahype::fv_max(l_minus, l0)) + qc"
                            and p.source_update in (None, "dt*s + qc"))
        parts.append(
            "\n// update statements: those of examples/Batched_stateless.py:29,31-33 -> the hand-written functor (same bits)\n"
            "using Update = ::exahype::RusanovUpdate;\n\n}  // namespace\n\n" if reference_update else
            "\n// update statements in the evaluation order of the declaration (SymPy str order == reference C++ order)\n"
            "struct Update {\n"
            "  template <typename T>\n"
            "  static __device__ __forceinline__ T flux(T qc, T f_plus, T f_minus) {\n"
            f"    return {p.flux_update};\n  }}\n"
            "  template <typename T>\n"
            "  static __device__ __forceinline__ T dissipation(T qc, T q0, T q_plus, T q_minus, T l0, T l_plus, T l_minus, T dt) {\n"
            f"    return {p.dissipation};\n  }}\n"
            + ("  template <typename T>\n"
               "  static __device__ __forceinline__ T source(T qc, T s, T dt) {\n"
               f"    return {p.source_update};\n  }}\n" if p.source_update else "")
            + "};\n\n}  // namespace\n\n")
        parts.append(
            f"// Drop-in for the reference's generated `void {fname}(double* Q, double dt)` over a device-resident batch.\n"
            "// flags: bit 1 = un-haloed output, bit 2 = accumulate into *lambda_max (include/exahype_cuda.h).\n"
            f'extern "C" __attribute__((visibility("default")))\n'
            f"int {fname}(const void* q_in, void* q_out, int64_t n_patches, double dt, void* lambda_patch, void* lambda_max,\n"
            f"    unsigned flags, void* stream) {{\n"
            "  cudaStream_t s = static_cast<cudaStream_t>(stream);\n"
            "  if (flags & ~6u) return -1;   // a generated unit knows the two output layouts and the accumulate bit, nothing else\n"
            f"  if (lambda_max && !(flags & 4u) && cudaMemsetAsync(lambda_max, 0, sizeof({T}), s) != cudaSuccess) return -3;\n"
            "  if (n_patches <= 0) return n_patches < 0 ? -1 : 0;\n"
            "  if (!q_in || !q_out || ((uintptr_t)q_in & 15) || ((uintptr_t)q_out & 15)) return -1;\n"
            "  if ((flags & 2u) && q_in == q_out) return -1;   // un-haloed output cannot alias the haloed input\n"
            "  cudaError_t err = (flags & 2u)\n"
            f"      ? {launcher(da, True)}::launch(q_in, q_out, n_patches, dt, lambda_patch, lambda_max, s)\n"
            f"      : {launcher(da, False)}::launch(q_in, q_out, n_patches, dt, lambda_patch, lambda_max, s);\n"
            "  return err == cudaSuccess ? 0 : -3;\n}\n\n"
            "// The ExaHyPE2 CellData form of the same step (include/exahype_cuda.h exahype_cell_data: per-patch QIn / QOut\n"
            "// pointers, dt, t, cellCentre, cellSize, maxEigenvalue -- reference examples/kernel-generator.py:8-19).\n"
            f'extern "C" __attribute__((visibility("default")))\n'
            f"int {fname}_cell_data(const exahype_cell_data* cells, double dt, void* lambda_max, unsigned flags, void* stream) {{\n"
            "  cudaStream_t s = static_cast<cudaStream_t>(stream);\n"
            "  if (!cells || cells->n_patches < 0 || (flags & ~6u)) return -1;\n"
            f"  if (lambda_max && !(flags & 4u) && cudaMemsetAsync(lambda_max, 0, sizeof({T}), s) != cudaSuccess) return -3;\n"
            "  if (cells->n_patches == 0) return 0;\n"
            "  if (!cells->q_in || !cells->q_out) return -1;\n"
            "  const ::exahype::FvGatherRaw g = {cells->q_in, cells->q_out, cells->dt, {}, cells->cell_centre, cells->cell_size, cells->t};\n"
            "  cudaError_t err = (flags & 2u)\n"
            f"      ? {launcher(da, True, True)}::launch(nullptr, nullptr, cells->n_patches, dt, cells->max_eigenvalue, lambda_max, s, &g)\n"
            f"      : {launcher(da, False, True)}::launch(nullptr, nullptr, cells->n_patches, dt, cells->max_eigenvalue, lambda_max, s, &g);\n"
            "  return err == cudaSuccess ? 0 : -3;\n}\n")
        return "".join(parts)

    # ------------------------------------------------------------------ compile + bind
    @staticmethod
    def cache_directory() -> str:
        """Per-user, mode 0700: ``$EXAHYPE_B200_CACHE`` or ``~/.cache/exahype_b200/generated``."""
        d = os.environ.get("EXAHYPE_B200_CACHE") or os.path.join(os.path.expanduser("~"), ".cache", "exahype_b200", "generated")
        try:
            os.makedirs(d, mode=0o700, exist_ok=True)
            if not os.access(d, os.W_OK):
                raise OSError("not writable")
        except OSError:                       # no usable home directory: a private directory for this process
            d = tempfile.mkdtemp(prefix="exahype_b200_generated_")
        return d

    def build_tag(self, extra=()) -> str:
        """Key of a built unit: the generated code AND everything else that decides the binary -- the kernel templates and
        the ABI header it includes, the nvcc flags, the compiler version."""
        from .. import build as _b
        h = hashlib.sha256()
        h.update(self.code.encode())
        for root in (CSRC, INCLUDE):
            for name in sorted(os.listdir(root)):
                if name.endswith((".cuh", ".h")):
                    with open(os.path.join(root, name), "rb") as f:
                        h.update(name.encode() + b"\0" + f.read())
        h.update(" ".join(_b.NVCC_FLAGS + list(extra)).encode())
        try:
            h.update(subprocess.run([_b.nvcc(), "--version"], capture_output=True, text=True).stdout.encode())
        except Exception:
            pass
        return h.hexdigest()[:20]

    def build(self, directory: Optional[str] = None, include_dirs=(), verbose: bool = False) -> "GeneratedKernel":
        """nvcc the unit for sm_100a into a shared library (cross-compiles without a GPU) and bind it."""
        from .. import build as _b
        directory = directory or self.cache_directory()
        os.makedirs(directory, exist_ok=True)
        tag = self.build_tag(include_dirs)
        lib = os.path.join(directory, f"lib{self.functionName()}_{tag}.so")
        if os.path.exists(lib) and hasattr(os, "getuid") and os.stat(lib).st_uid != os.getuid():
            raise RuntimeError(f"{lib} exists but belongs to another user: refusing to load it")
        if not os.path.exists(lib):
            # private names until the atomic rename: ranks of one job build the same unit at the same time
            fd, src = tempfile.mkstemp(prefix=f"{self.functionName()}_{tag}_", suffix=".cu", dir=directory)
            with os.fdopen(fd, "w") as f:
                f.write(self.code)
            tmp = src[:-3] + ".so.tmp"
            cmd = [_b.nvcc()] + _b.NVCC_FLAGS + _b._host_compiler_args() + ["-I", CSRC, "-I", INCLUDE]
            for d in include_dirs:
                cmd += ["-I", d]
            cmd += ["-shared", src, "-o", tmp, "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on the generated kernel ({src}):\n{r.stderr}")
            os.replace(tmp, lib)
            os.replace(src, os.path.join(directory, f"{self.functionName()}_{tag}.cu"))
        return GeneratedKernel(self, lib)


class _CellData(ctypes.Structure):
    """``exahype_cell_data`` of include/exahype_cuda.h."""
    _fields_ = [("n_patches", ctypes.c_int64), ("q_in", ctypes.c_void_p), ("q_out", ctypes.c_void_p),
                ("dt", ctypes.c_void_p), ("max_eigenvalue", ctypes.c_void_p), ("cell_centre", ctypes.c_void_p),
                ("cell_size", ctypes.c_void_p), ("t", ctypes.c_void_p)]


class GeneratedKernel:
    """A compiled ``CUDAPrinter`` unit: same ``step`` / ``step_cell_data`` calls as
    :class:`exahype_b200.runtime.PatchUpdate`, with the same argument checks."""

    def __init__(self, printer: CUDAPrinter, lib_path: str):
        k = printer.kernel()
        self.lib_path = lib_path
        self.dim, self.patch_size, self.halo_size = k.dim, k.patch_size, k.halo_size
        self.n_real, self.n_aux, self.dtype = k.n_real, k.n_aux, printer.dtype
        self._lib = ctypes.CDLL(lib_path)
        self._fn = getattr(self._lib, printer.functionName())
        vp = ctypes.c_void_p
        self._fn.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_double, vp, vp, ctypes.c_uint, vp]
        self._fn.restype = ctypes.c_int
        self._fn_cells = getattr(self._lib, printer.functionName() + "_cell_data")
        self._fn_cells.argtypes = [ctypes.POINTER(_CellData), ctypes.c_double, vp, ctypes.c_uint, vp]
        self._fn_cells.restype = ctypes.c_int

    def in_shape(self, n):
        return (n,) + (self.patch_size + 2 * self.halo_size,) * self.dim + (self.n_real + self.n_aux,)

    def out_shape(self, n, unhaloed=False):
        return (n,) + (self.patch_size,) * self.dim + (self.n_real + self.n_aux,) if unhaloed else self.in_shape(n)

    def _tdt(self):
        import torch
        return torch.float64 if self.dtype == "f64" else torch.float32

    def step(self, q_in, q_out=None, dt: float = 0.0, lambda_patch=None, lambda_max=None, stream=None,
             unhaloed: bool = False, accumulate_lambda: bool = False):
        import torch
        tdt = self._tdt()
        if q_out is None:
            if unhaloed:
                raise ValueError("un-haloed output needs an explicit q_out")
            q_out = q_in
        per = (self.patch_size + 2 * self.halo_size) ** self.dim * (self.n_real + self.n_aux)
        if not q_in.is_cuda or q_in.numel() % per:
            raise ValueError("q_in must be a CUDA tensor holding whole patches")
        for name, t in (("q_in", q_in), ("q_out", q_out), ("lambda_patch", lambda_patch), ("lambda_max", lambda_max)):
            if t is None:
                continue
            if t.dtype != tdt or not t.is_contiguous() or not t.is_cuda or t.device != q_in.device:
                raise ValueError(f"{name} must be a contiguous {tdt} CUDA tensor on {q_in.device}")
        n = q_in.numel() // per
        want = 1
        for d in self.out_shape(n, unhaloed):
            want *= d
        if q_out is not q_in and q_out.numel() != want:
            raise ValueError(f"q_out must hold {self.out_shape(n, unhaloed)}")
        if unhaloed and q_out.data_ptr() == q_in.data_ptr():
            raise ValueError("un-haloed output cannot alias the haloed input")
        if lambda_patch is not None and lambda_patch.numel() < n:
            raise ValueError("lambda_patch must hold one value per patch")
        if lambda_max is not None and lambda_max.numel() < 1:
            raise ValueError("lambda_max must hold one value")
        if stream is None:
            stream = torch.cuda.current_stream(q_in.device).cuda_stream
        flags = (2 if unhaloed else 0) | (4 if accumulate_lambda else 0)
        with torch.cuda.device(q_in.device):
            rc = self._fn(q_in.data_ptr(), q_out.data_ptr(), n, float(dt),
                          lambda_patch.data_ptr() if lambda_patch is not None else None,
                          lambda_max.data_ptr() if lambda_max is not None else None, flags, stream)
        if rc:
            raise RuntimeError(f"generated kernel failed with code {rc}")
        return q_out

    def step_cell_data(self, q_in_ptrs, q_out_ptrs, dt: float = 0.0, dt_patch=None, max_eigenvalue=None, lambda_max=None,
                       cell_centre=None, cell_size=None, t_patch=None, unhaloed: bool = False, stream=None):
        """The ``CellData`` form: int64 CUDA tensors of per-patch device pointers, optional per-patch ``dt`` / ``t``
        (``[n]``) and ``cellCentre`` / ``cellSize`` (``[n, dim]``) of this kernel's dtype, per-patch ``maxEigenvalue`` out."""
        import torch
        tdt = self._tdt()
        n = int(q_in_ptrs.numel())
        for name, ten, want, count in (("q_in_ptrs", q_in_ptrs, torch.int64, n), ("q_out_ptrs", q_out_ptrs, torch.int64, n),
                                       ("dt_patch", dt_patch, tdt, n), ("max_eigenvalue", max_eigenvalue, tdt, n),
                                       ("lambda_max", lambda_max, tdt, 1), ("cell_centre", cell_centre, tdt, n * self.dim),
                                       ("cell_size", cell_size, tdt, n * self.dim), ("t_patch", t_patch, tdt, n)):
            if ten is None:
                continue
            if not ten.is_cuda or not ten.is_contiguous() or ten.dtype != want or ten.numel() < count:
                raise ValueError(f"{name} must be a contiguous CUDA tensor of dtype {want} with at least {count} values")
        if stream is None:
            stream = torch.cuda.current_stream(q_in_ptrs.device).cuda_stream
        ptr = lambda x: x.data_ptr() if x is not None else None
        cells = _CellData(n, q_in_ptrs.data_ptr(), q_out_ptrs.data_ptr(), ptr(dt_patch), ptr(max_eigenvalue),
                          ptr(cell_centre), ptr(cell_size), ptr(t_patch))
        with torch.cuda.device(q_in_ptrs.device):
            rc = self._fn_cells(ctypes.byref(cells), float(dt), ptr(lambda_max), 2 if unhaloed else 0, stream)
        if rc:
            raise RuntimeError(f"generated kernel (cell data) failed with code {rc}")
