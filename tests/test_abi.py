"""The C-ABI library builds, loads and exports every symbol include/asvd_b200.h declares (no compute calls)."""
import ctypes, os, re
import pytest
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "asvd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asvd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    from asvd4llm_b200 import _lib
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    from asvd4llm_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.asvd_version() == 200


def test_rank_formula_host_entry_point():
    from asvd4llm_b200 import _lib
    from oracle import asvd_oracle as O
    for (m, n) in [(768, 768), (3072, 768), (50272, 768), (4096, 4096), (11008, 4096), (4096, 11008), (32000, 4096),
                   (13824, 5120), (96, 32), (7, 5)]:
        for ratio in [0.1, 0.30000000000000004, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95, 1.9]:
            for align in (1, 8, 128):
                assert _lib.rank_for_ratio(m, n, ratio, align) == O.rank_for_ratio(m, n, ratio, align), (m, n, ratio, align)


def test_workspace_size_is_pure_function_of_shape():
    from asvd4llm_b200 import _lib
    lib = _lib.load()
    a = lib.asvd_svd_workspace_bytes(4096, 4096, 1)
    assert a == lib.asvd_svd_workspace_bytes(4096, 4096, 1)
    assert a >= 4 * 4096 * (4096 + 4096)
    # the working set X and its row-major copy are per matrix; the recovery-GEMM scratch is per call
    assert lib.asvd_svd_workspace_bytes(4096, 4096, 4) > a + 3 * 2 * 4 * 4096 * 4096
    assert lib.asvd_svd_workspace_bytes(0, 5, 1) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import torch.nn as nn
    from asvd4llm_b200 import SVDLinear
    lin = nn.Linear(16, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        SVDLinear.from_linear(lin, 0.9)
    mod = SVDLinear(torch.randn(8, 4), torch.rand(4), torch.randn(16, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        mod(torch.randn(2, 16))
