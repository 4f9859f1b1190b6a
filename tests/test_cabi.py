"""The C-ABI shared library loads (no GPU needed) and exports every symbol include/eat_b200.h declares."""
import ctypes
import os
import re
import subprocess

from efficientat_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m efficientat_b200.build` first"
    protos = _lib.parse_header()
    assert len(protos) >= 12
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/eat_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (eat_\w+)", out))
    assert exported == set(protos), f"header/library mismatch: {exported ^ set(protos)}"


def test_abi_version_and_error_string():
    l = _lib.lib()
    assert l.abi_version() == 1
    assert isinstance(l.last_error(), bytes)


def test_no_torch_types_in_signatures():
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)     # strip comments
    assert "at::" not in text and "torch" not in text and "Tensor" not in text


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "efficientat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"
