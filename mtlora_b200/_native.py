"""ctypes binding of libmtlora_b200.so — the C ABI declared in include/mtlora_b200.h.

This is the only place the shared library is opened. There is no CPU or PyTorch fallback: when the library is
missing or a call fails, a RuntimeError carrying mtl_last_error() is raised.
"""
import ctypes
import os

MTL_MAX_TASKS = 7
MTL_ABI_VERSION = 3
MTL_MODE_MATRIX, MTL_MODE_MATRIXV2 = 0, 1
MTL_ACT_NONE, MTL_ACT_GELU, MTL_ACT_GELU_GRAD = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtlora_b200.so")

c_void_p, c_int, c_int32, c_int64, c_float, c_uint64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64)


class LinearCfg(ctypes.Structure):
    """mtl_linear_cfg (include/mtlora_b200.h) == constructor arguments of MTLoRALinear, models/lora.py:161-176."""
    _fields_ = [
        ("M", c_int64),
        ("in_features", c_int32),
        ("out_features", c_int32),
        ("n_tasks", c_int32),
        ("x_tasks_given", c_int32),
        ("shared_mode", c_int32),
        ("r_shared", c_int32),
        ("r_task", c_int32 * MTL_MAX_TASKS),
        ("scale_shared", c_float),
        ("scale_task", c_float * MTL_MAX_TASKS),
        ("dropout_p", c_float),
        ("dropout_seed", c_uint64),
        ("rows_per_sample", c_int32),
        ("gelu_aux_is_grad", c_int32),
        ("dy_has_sum", c_int32),
        ("u_precomputed", c_int32),
    ]


class LinearPlanInfo(ctypes.Structure):
    """mtl_linear_plan_info: the tiling the planner of the fused linear kernel picks (mtl_linear_plan, no GPU needed)."""
    _fields_ = [(n, c_int32) for n in (
        "bn", "n_chunks", "n_splits", "n_stages", "n_slabs", "n_regions", "n_pbuf", "n_dbuf", "d_shared", "tmem_cols",
        "tmem_cols_used", "smem_bytes", "n_work", "n_groups", "s_in", "s_out", "up_pack")]


class OptSeg(ctypes.Structure):
    """mtl_opt_seg: one trainable tensor of the flat optimizer step."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("offset", c_int64), ("numel", c_int64), ("group", c_int32),
                ("pad_", c_int32)]


class OptGroup(ctypes.Structure):
    """mtl_opt_group: hyper-parameters of one torch param_group."""
    _fields_ = [("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("weight_decay", c_float)]


class PackJob(ctypes.Structure):
    """mtl_pack_job: operand staging of one layer inside mtl_linear_pack_many."""
    _fields_ = [("cfg", LinearCfg), ("a_shared", c_void_p), ("b_shared", c_void_p),
                ("a_tasks", c_void_p * MTL_MAX_TASKS), ("b_tasks", c_void_p * MTL_MAX_TASKS),
                ("a_cat", c_void_p), ("b_cat", c_void_p), ("a_cat_t", c_void_p), ("b_cat_t", c_void_p)]


MTL_OPT_CHUNK, MTL_OPT_MAX_GROUPS = 4096, 8
_CFG_P = ctypes.POINTER(LinearCfg)
_PP = ctypes.POINTER(c_void_p)

# name -> (restype, argtypes); must list every function include/mtlora_b200.h declares (tests check this).
SIGNATURES = {
    "mtl_abi_version": (c_int, []),
    "mtl_linear_cfg_size": (c_int, []),
    "mtl_last_error": (ctypes.c_char_p, []),
    "mtl_launch_count": (c_uint64, []),
    "mtl_linear_rank_pad": (c_int, [_CFG_P]),
    "mtl_linear_rank_offset": (c_int, [_CFG_P, c_int]),
    "mtl_linear_pack": (c_int, [_CFG_P, c_void_p, c_void_p, _PP, _PP, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mtl_pack_job_size": (c_int, []),
    "mtl_linear_pack_many": (c_int, [ctypes.POINTER(PackJob), c_int32, c_void_p]),
    "mtl_cast_transpose": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "mtl_linear_rank_project": (c_int, [_CFG_P, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mtl_linear_fwd": (c_int, [_CFG_P, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                               c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "mtl_linear_bwd_input": (c_int, [_CFG_P, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "mtl_linear_plan": (c_int, [_CFG_P, c_int32, c_int32, c_int32, c_int32, ctypes.POINTER(LinearPlanInfo)]),
    "mtl_linear_bwd_params": (c_int, [_CFG_P, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "mtl_xty": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32, c_float,
                        c_void_p]),
    "mtl_window_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                         c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_float,
                                         c_uint64, c_void_p]),
    "mtl_window_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                         c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_float,
                                         c_void_p]),
    "mtl_roll_and_window_partition_forward": (c_int, [c_void_p, c_void_p] + [c_int32] * 7 + [c_void_p]),
    "mtl_roll_and_window_partition_backward": (c_int, [c_void_p, c_void_p] + [c_int32] * 7 + [c_void_p]),
    "mtl_window_merge_and_roll_forward": (c_int, [c_void_p, c_void_p] + [c_int32] * 7 + [c_void_p]),
    "mtl_window_merge_and_roll_backward": (c_int, [c_void_p, c_void_p] + [c_int32] * 7 + [c_void_p]),
    "mtl_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                  c_int64, c_int32, c_float, c_int32, c_int32, c_int32, c_float, c_uint64, c_void_p]),
    "mtl_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "mtl_dropout": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_uint64, c_void_p]),
    "mtl_scale_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p]),
    "mtl_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mtl_patch_embed_fwd": (c_int, [c_void_p] * 10 + [c_int32] * 4 + [c_float, c_void_p]),
    "mtl_scale_rows_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p]),
    "mtl_sum_streams": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p]),
    "mtl_opt_seg_size": (c_int, []),
    "mtl_opt_sqnorm": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "mtl_opt_adamw": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                              ctypes.POINTER(OptGroup), c_int32, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_void_p]),
}

_lib = None
launch_count = 0  # number of C-ABI compute calls issued by this process (bench.py reports kernel launches from it)


def load():
    """Open the library (once) and declare every signature. Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the sm_100a extension is not built. Run `make` at the repo root or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU/PyTorch fallback for this path.")
    import torch  # noqa: F401  (loads the CUDA runtime the library links against)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    got = lib.mtl_abi_version()
    if got != MTL_ABI_VERSION:
        raise RuntimeError(f"libmtlora_b200.so ABI version {got} != expected {MTL_ABI_VERSION}; rebuild")
    if lib.mtl_linear_cfg_size() != ctypes.sizeof(LinearCfg):
        raise RuntimeError(f"struct mtl_linear_cfg is {lib.mtl_linear_cfg_size()} bytes in libmtlora_b200.so but "
                           f"{ctypes.sizeof(LinearCfg)} in the ctypes mirror; rebuild (`make`)")
    if lib.mtl_pack_job_size() != ctypes.sizeof(PackJob):
        raise RuntimeError("struct mtl_pack_job differs between libmtlora_b200.so and the ctypes mirror; rebuild (`make`)")
    _lib = lib
    return lib


profile = None  # set to a list to record (name, meta, start_event, end_event) for every C-ABI call (bench.py --profile-ops)


def call(name, *args, meta=None):
    """Invoke a C-ABI function returning an int status; raise RuntimeError(mtl_last_error()) on failure."""
    global launch_count
    lib = load()
    if profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.mtl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{name} failed (code {rc}): {msg}")
    if profile is not None:
        e1.record()
        profile.append((name, meta, e0, e1))
    launch_count += 1
    return rc


def kernel_launches():
    """Process-wide number of CUDA kernels launched by libmtlora_b200.so (mtl_launch_count)."""
    return int(load().mtl_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_cur_dev = (None, -1)   # (launch_count at which the device was last queried, its index)


def current_device_index():
    """Index of the current CUDA device, queried from torch at most once per C-ABI call (the argument checks of one
    op look at ~10 tensors; torch.cuda.current_device() costs about a microsecond each time)."""
    global _cur_dev
    if _cur_dev[0] != launch_count:
        import torch
        _cur_dev = (launch_count, torch.cuda.current_device())
    return _cur_dev[1]


_raw_stream = None


def stream():
    """cudaStream_t of torch's current stream on the current device (the raw-handle query costs a fraction of a
    microsecond; torch.cuda.current_stream() builds a Python Stream object every time)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(current_device_index())
    return torch.cuda.current_stream().cuda_stream
