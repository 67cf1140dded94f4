"""CPU: the C-ABI library loads and exports every symbol include/fsb200.h declares; the ctypes table matches the header.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_ops():
    names = _declared()
    for must in ("fsb_gemm_bf16", "fsb_sdpa_fwd", "fsb_sdpa_bwd", "fsb_rmsnorm_fwd", "fsb_rmsnorm_bwd",
                 "fsb_layernorm_fwd", "fsb_layernorm_bwd", "fsb_rope_inplace", "fsb_glu_fwd", "fsb_glu_bwd",
                 "fsb_embedding_fwd", "fsb_embedding_bwd", "fsb_softmax_xent_fwd_bwd", "fsb_adamw_flat",
                 "fsb_last_error", "fsb_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from fsb200 import lib
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in _declared():
        assert hasattr(dll, name), f"{name} declared in include/fsb200.h but not exported by libfsb200.so"


def test_ctypes_table_matches_header():
    from fsb200 import lib
    assert sorted(lib.SIGNATURES) == _declared()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fsb200.h")).read(), flags=re.S)
    for name, (_, argtypes) in lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), f"{name}: header has {n} parameters, ctypes table {len(argtypes)}"


def test_library_reports_version_and_error_string_without_a_gpu():
    from fsb200 import lib
    L = lib.load()
    assert L.fsb_version() >= 1
    assert isinstance(lib.last_error(), str)


def test_ops_refuse_cpu_tensors_loudly():
    import torch
    from fsb200 import ops, lib as L
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(L.GEMM_NT, a, a)


def test_plain_c_client_compiles_against_the_header_and_runs(tmp_path):
    """include/fsb200.h is C (C99, -Wall -Wextra -Werror) and the library is usable without Python / torch: tests/c_abi_client.c
    calls the host entry points (version, error string, index builders, the argument-validation path of a device entry point)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from fsb200 import lib
    libdir = os.path.dirname(lib.LIB_PATH)
    exe = str(tmp_path / "c_abi_client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_client.c"), "-o", exe, "-L", libdir, "-lfsb200", f"-Wl,-rpath,{libdir}"],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=60)
    assert "c_abi_client ok" in out.stdout
