"""``from exahype.printers import CUDAPrinter`` (new) next to the reference's names
(reference ``exahype/printers/__init__.py:1-2``)."""
from exahype_b200.printers import CodePrinter, CUDAPrinter, CPPPrinter, MLIRPrinter  # noqa: F401
