"""Drop-in alias for the reference's package name: ``from exahype import KernelBuilder`` keeps working
(reference ``exahype/__init__.py:1-2``).  The implementation lives in :mod:`exahype_b200`; the reference's eager
``SymPyToMLIR`` import (``exahype/__init__.py:3``, needs xdsl) is intentionally absent."""
from exahype_b200 import KernelBuilder, TypedFunction, DeviceBody, viable  # noqa: F401
