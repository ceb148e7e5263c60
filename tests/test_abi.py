"""CPU: the C-ABI library builds, loads, and exports every symbol include/diffews_b200.h declares; the ctypes table in
diffews_b200/_lib.py mirrors the header (same names, same argument counts).  No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    src = open(os.path.join(ROOT, "include", "diffews_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|long long)\s+(dfw_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(2)] = (m.group(1), n)
    return decls


def test_header_symbols_exported(lib_built):
    decls = _header_decls()
    assert len(decls) >= 20
    lib = ctypes.CDLL(lib_built.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/diffews_b200.h but not exported"


def test_ctypes_table_mirrors_header(lib_built):
    decls = _header_decls()
    sig = lib_built.SIGNATURES
    assert set(sig) == set(decls), set(sig) ^ set(decls)
    for name, (ret, nargs) in decls.items():
        res, args = sig[name]
        assert len(args) == nargs, f"{name}: header has {nargs} args, _lib.py has {len(args)}"
        assert (res is ctypes.c_longlong) == (ret == "long long"), name


def test_version_and_no_gpu_behaviour(lib_built):
    import torch
    assert lib_built.lib.dfw_version() == 2
    assert lib_built.lib.dfw_get_option(6) == 0 and lib_built.lib.dfw_set_option(99, 1) == -1   # option table: explicit, no getenv
    assert lib_built.lib.dfw_launch_count() >= 0
    if not torch.cuda.is_available():
        # no device: every compute entry point must fail loudly (negative status), never fall back
        assert lib_built.lib.dfw_device_ok() < 0
        assert lib_built.lib.dfw_layernorm(None, 0, None, None, None, 0, 1, 8, 1e-5, None) < 0
        assert lib_built.lib.dfw_rthres_workspace_bytes(4) == 256
        assert lib_built.lib.dfw_groupnorm_workspace_bytes(2, 64, 320, 32) > 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no file of the product package may import it."""
    pkg = os.path.join(ROOT, "diffews_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_ops_refuse_cpu_tensors(lib_built):
    import pytest
    import torch
    from diffews_b200 import ops
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(Exception):
        ops.linear(x, x)
    from diffews_b200.unet import MyUNet2DConditionModel  # noqa: F401  (imports without a GPU)


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/diffews_b200.h must compile as C99 (no C++, no torch / CUDA types in signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text('#include "include/diffews_b200.h"\n'
                   "int main(void) { DfwImageDesc d; DfwAdamTensor t; (void)d; (void)t; return dfw_version() < 0; }\n")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", ROOT, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "diffews_b200.h")).read(), flags=re.S)
    for banned in ("torch", "at::", "c10::", "cudaStream_t", "std::", "template"):
        assert banned not in code, banned
