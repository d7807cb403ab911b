"""The C-ABI library loads without a GPU and exports every symbol include/numcl_cuda.h declares."""
import ctypes as C
import os
import re

from numcl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "numcl_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ncl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"libnumcl_cuda.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names


def test_no_cpu_fallback_without_gpu():
    """Compute entry points fail loudly when the library is not bound to a B200."""
    import subprocess, sys
    code = ("import os; os.environ['CUDA_VISIBLE_DEVICES']=''; import ctypes as C; from numcl_b200 import _lib;"
            "L=_lib.load(); rc=L.ncl_init(1, None); print(rc, L.ncl_last_error().decode())")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT).decode()
    assert out.split()[0] == str(_lib.ERR_CUDA) and "no CPU fallback" in out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "numcl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "numcl_oracle" not in text or f == "ncl_device.cuh" or f == "k_fill.cu", f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
