"""Abstract back-end: holds a kernel declaration and a function name, produces ``.code``.

Same public surface as the reference's ``exahype/printers/CodePrinter.py:46-71``
(``kernel()``, ``functionName()``, ``file()``, ``here()``, abstract ``loop()``).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

from ..KernelBuilder import KernelBuilder


class CodePrinter(ABC):
    code: str = ""

    def __init__(self, kernel: KernelBuilder, function_name: str):
        self._kernel = kernel
        self._functionName = function_name

    def kernel(self, kernel: Optional[KernelBuilder] = None) -> KernelBuilder:
        """Getter, or setter when an argument is given."""
        if kernel is not None:
            self._kernel = kernel
        return self._kernel

    def functionName(self, function_name: Optional[str] = None) -> str:
        if function_name is not None:
            self._functionName = function_name
        return self._functionName

    def file(self, file_name: str, header_file_name: Optional[str] = None):
        """Write ``self.code`` to ``file_name``; subclasses decide what the header name means."""
        with open(file_name, "w") as out:
            out.write(self.code)

    def here(self):
        print(self.code)

    @abstractmethod
    def loop(self, expr: list, direction: int, below: int, struct_inclusion: int):
        """Emit the code for one statement ``[LHS, RHS]`` swept along ``direction``."""
