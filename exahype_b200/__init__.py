"""exahype_b200 -- B200-native batched stateless finite-volume Rusanov patch update.

Front-end API of the reference kept (``KernelBuilder`` / ``TypedFunction``, reference
``exahype/__init__.py:1-2``); back-end is ``printers.CUDAPrinter`` + ``libexahype_cuda.so``
(``exahype_b200.runtime``).  Nothing here imports the test oracle, and importing this package
does not need xdsl.
"""
from .KernelBuilder import KernelBuilder, viable
from .TypedFunction import TypedFunction, DeviceBody

__all__ = ["KernelBuilder", "TypedFunction", "DeviceBody", "viable"]
__version__ = "0.1.0"
