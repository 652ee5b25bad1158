"""Import the UNMODIFIED reference (TEST INFRASTRUCTURE; only works where /root/reference exists).

Nothing is copied: `compressai` is imported from the reference tree where it lies, its two
pybind11 extensions (`compressai._CXX`, `compressai.ans`) are resolved to the binaries that
`make -C oracle ref` compiled from the reference's own C++ sources into oracle/_ref/, and
the two third-party modules missing offline (`kornia`, `range_coder`) come from oracle/shims.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import sysconfig
from pathlib import Path

ORACLE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("MASIC_REFERENCE", "/root/reference"))
_EXT = sysconfig.get_config_var("EXT_SUFFIX")


def ref_ext_path(name: str) -> Path:
    return ORACLE / "_ref" / f"{name}{_EXT}"


class _RefExtFinder(importlib.abc.MetaPathFinder):
    _names = {"compressai._CXX": "_CXX", "compressai.ans": "ans"}

    def find_spec(self, fullname, path=None, target=None):
        short = self._names.get(fullname)
        if short is None:
            return None
        so = ref_ext_path(short)
        if not so.exists():
            raise ImportError(f"{so} missing: run `make -C oracle ref`")
        loader = importlib.machinery.ExtensionFileLoader(fullname, str(so))
        return importlib.util.spec_from_file_location(fullname, str(so), loader=loader)


def available() -> bool:
    return (REF_ROOT / "coremasic" / "mywork" / "MASIC.py").exists()


def load_ref_ext(name: str):
    """Load one of the compiled reference extensions stand-alone (works on the GPU box too)."""
    so = ref_ext_path(name)
    if not so.exists():
        raise ImportError(f"{so} missing: run `make -C oracle ref` where /root/reference exists")
    modname = f"_masic_ref_{name}"
    if modname in sys.modules:
        return sys.modules[modname]
    loader = importlib.machinery.ExtensionFileLoader(name, str(so))
    spec = importlib.util.spec_from_file_location(name, str(so), loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[modname] = mod
    return mod


def install() -> None:
    """Make `import compressai`, `import kornia`, `import MASIC` resolve to the reference."""
    if not available():
        raise ImportError(f"reference tree not found at {REF_ROOT}")
    if not any(isinstance(f, _RefExtFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefExtFinder())
    for p in (str(ORACLE / "shims"), str(REF_ROOT / "coremasic" / "mywork"), str(REF_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)


def import_masic():
    install()
    import MASIC  # noqa: N811  (coremasic/mywork/MASIC.py, unmodified)
    return MASIC
