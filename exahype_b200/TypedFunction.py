"""``TypedFunction`` -- a SymPy undefined function that also carries a C-level signature.

Mirrors the public behaviour of the reference's ``exahype/TypedFunction.py:11-34``: calling
``TypedFunction(name)`` yields a SymPy function *class* (so ``Flux(Q[0], normal, F[0])`` builds an
applied-function node) with ``return_type`` / ``parameter_types`` attributes and the accessor methods
``returnType()`` / ``parameterTypes()``.

B200 extension (not in the reference, where bodies live in hand-written C++ such as
``Unit test/Functions.cpp``): ``deviceBody()`` attaches what the CUDA printer needs to emit a
``__device__`` functor for the function -- see :class:`DeviceBody`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Union

import sympy


@dataclass
class DeviceBody:
    """How a declared function becomes device code.

    Exactly one of the fields is used:

    * ``builtin``: name of a hand-written functor family shipped in ``csrc/physics.cuh``
      (``"euler"`` or ``"swe"``);
    * ``expressions``: a callable ``f(q, normal) -> sequence | expr`` evaluated on SymPy symbols
      ``q[0..n_var)`` with ``normal`` a Python int; a sequence gives the flux components, a single
      expression gives a scalar (eigenvalue);
    * ``source``: verbatim CUDA source of a ``__device__`` function with the reference's user-function
      signature (``Unit test/Functions.h:2-4``), e.g.
      ``template <class T> __device__ void Flux(const T* Q, int normal, T* F)``.
    """
    builtin: Optional[str] = None
    expressions: Optional[Callable] = None
    source: Optional[str] = None

    def __post_init__(self):
        given = [x is not None for x in (self.builtin, self.expressions, self.source)]
        if sum(given) != 1:
            raise ValueError("DeviceBody takes exactly one of builtin=, expressions=, source=")


def _return_type(func, value=None):
    if value is not None:
        func.return_type = value
    return func.return_type


def _parameter_types(func, value: Optional[Sequence] = None):
    if value is not None:
        func.parameter_types = value
    return func.parameter_types


def _device_body(func, body: Union[None, str, Callable, DeviceBody] = None, **kwargs):
    if body is not None or kwargs:
        if isinstance(body, DeviceBody):
            func.device_body = body
        elif callable(body):
            func.device_body = DeviceBody(expressions=body)
        elif isinstance(body, str):
            func.device_body = DeviceBody(source=body)
        else:
            func.device_body = DeviceBody(**kwargs)
    return func.device_body


class TypedFunction:
    """Factory: ``TypedFunction('Flux')`` returns a fresh SymPy function class named ``Flux``."""

    def __new__(cls, name: str, **options):
        func = sympy.Function(name, **options)
        func.return_type = None
        func.parameter_types = None
        func.device_body = None
        # bound per created class, so two functions never share a signature
        func.returnType = lambda value=None, _f=func: _return_type(_f, value)
        func.parameterTypes = lambda value=None, _f=func: _parameter_types(_f, value)
        func.deviceBody = lambda body=None, _f=func, **kw: _device_body(_f, body, **kw)
        return func
