"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly what
include/mtlora_b200.h declares; the ctypes table covers every declaration."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mtlora_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    path = os.path.join(ROOT, "mtlora_b200", "libmtlora_b200.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(path)


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mtlora_b200.h but not exported"


def test_ctypes_table_matches_header():
    from mtlora_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_functions()


def test_abi_version_and_errors(lib):
    from mtlora_b200 import _native
    assert lib.mtl_abi_version() == _native.MTL_ABI_VERSION
    nat = _native.load()
    cfg = _native.LinearCfg()
    cfg.M, cfg.in_features, cfg.out_features, cfg.n_tasks, cfg.r_shared = 8, 96, 96, 99, 4
    assert nat.mtl_linear_rank_pad(ctypes.byref(cfg)) == -1
    assert b"n_tasks" in nat.mtl_last_error()
    cfg.n_tasks = 2
    cfg.r_task[0], cfg.r_task[1] = 4, 20
    assert nat.mtl_linear_rank_pad(ctypes.byref(cfg)) == 16 + 16 + 32
    assert nat.mtl_linear_rank_offset(ctypes.byref(cfg), 2) == 32


def test_cfg_struct_layout(lib):
    """The ctypes mirror of struct mtl_linear_cfg has the size the library was compiled with (fields were appended this
    round: gelu_aux_is_grad, dy_has_sum), and the field offsets follow the C layout rules."""
    from mtlora_b200 import _native
    assert lib.mtl_linear_cfg_size() == ctypes.sizeof(_native.LinearCfg)
    f = _native.LinearCfg
    assert f.M.offset == 0 and f.in_features.offset == 8 and f.r_task.offset == 32
    assert f.dropout_seed.offset % 8 == 0 and f.rows_per_sample.offset == f.dropout_seed.offset + 8
    assert f.dy_has_sum.offset == f.gelu_aux_is_grad.offset + 4 == f.rows_per_sample.offset + 8
    assert f.u_precomputed.offset == f.dy_has_sum.offset + 4


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    import torch
    from mtlora_b200 import ops
    spec = ops.LinearSpec(96, 96, 8, [])
    x = torch.zeros(1, 16, 96, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.linear_fwd(spec, x, x, None, x, x)
