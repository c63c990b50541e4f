from .CodePrinter import CodePrinter
from .CUDAPrinter import CUDAPrinter, GeneratedKernel, UnsupportedKernel, analyse
from .CPPPrinter import CPPPrinter
from .MLIRPrinter import MLIRPrinter

__all__ = ["CodePrinter", "CUDAPrinter", "CPPPrinter", "MLIRPrinter", "GeneratedKernel", "UnsupportedKernel", "analyse"]
