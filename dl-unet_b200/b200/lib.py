"""ctypes binding of the C ABI in include/unet_b200.h (libunet_b200.so).

The library is the product path: there is no CPU or PyTorch fallback.  Importing this
module on a box without the built library raises; calling a compute entry point without
a CUDA device raises RuntimeError (the library returns UB_E_DRIVER / a cudaError_t).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UB_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "csrc", "libunet_b200.so")


class Tensor4(C.Structure):
    """ub_tensor4: NHWC bf16 view, strides in elements (include/unet_b200.h)."""

    _fields_ = [
        ("ptr", C.c_void_p),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
        ("sn", C.c_longlong), ("sy", C.c_longlong), ("sx", C.c_longlong),
    ]


_P = C.c_void_p
_T4 = C.POINTER(Tensor4)
_i, _ll, _f = C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes); mirrors include/unet_b200.h one to one
SIGNATURES = {
    "ub_version": (_i, []),
    "ub_error_string": (C.c_char_p, [_i]),
    "ub_device_error": (_i, [C.POINTER(C.c_uint)]),
    "ub_set_tuning": (None, [_i, _i, _i, _i]),
    "ub_set_sm_limit": (None, [_i]),
    "ub_set_prof_buffer": (None, [_P]),
    "ub_set_option": (None, [_i, _i]),
    "ub_pack_weight": (_i, [_P, _P, _i, _i, _i, _P]),
    "ub_unpack_grad": (_i, [_P, _i, _ll, _P, _i, _i, _i, _i, _P]),
    "ub_job_table_bytes": (_i, [_i]),
    "ub_job_table_fill": (_i, [_P, _i, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                               C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ub_pack_all": (_i, [_P, _i, _ll, _P]),
    "ub_unpack_range": (_i, [_P, _i, _i, _i, _i, _P]),
    "ub_conv3x3_fwd": (_i, [_T4, _P, _P, _T4, _i, _P]),
    "ub_conv3x3_fwd_head": (_i, [_T4, _P, _P, _T4, _P, _P, _P, _P, _P]),
    "ub_conv3x3_fwd_pool": (_i, [_T4, _P, _P, _i, _T4, _P, _P, _T4, _i, _P]),
    "ub_conv3x3_dgrad": (_i, [_T4, _P, _T4, _T4, _P]),
    "ub_conv3x3_dgrad_bias": (_i, [_T4, _P, _T4, _T4, _P, _i, _P]),
    "ub_wgrad_slabs": (_i, [_i, _i, _i, _i, _i, _i]),
    "ub_conv3x3_wgrad": (_i, [_T4, _T4, _P, _P]),
    "ub_upconv2x2_fwd": (_i, [_T4, _P, _P, _T4, _P]),
    "ub_upconv2x2_dgrad": (_i, [_T4, _P, _T4, _T4, _P]),
    "ub_upconv2x2_dgrad_bias": (_i, [_T4, _P, _T4, _T4, _P, _P]),
    "ub_upconv2x2_wgrad": (_i, [_T4, _T4, _P, _P]),
    "ub_conv1_fwd": (_i, [_P, _P, _P, _P, _i, _i, _i, _P]),
    "ub_conv1_wgrad_workspace_bytes": (_ll, [_i, _i, _i]),
    "ub_conv1_wgrad": (_i, [_P, _P, _P, _P, _i, _i, _i, _P]),
    "ub_pool2x2_fwd": (_i, [_P, _P, _P, _T4, _i, _i, _i, _i, _i, _P]),
    "ub_pool2x2_bwd": (_i, [_P, _T4, _i, _P, _P, _P, _P, _i, _i, _i, _i, _P]),
    "ub_head_fwd": (_i, [_P, _P, _P, _P, _P, _i, _i, _i, _P]),
    "ub_head_bwd": (_i, [_P, _P, _P, _P, _P, _i, _i, _i, _i, _P]),
    "ub_loss_workspace_bytes": (_ll, []),
    "ub_loss_fwd_bwd": (_i, [_i, _P, _P, _i, _P, _P, _P, _P, _i, _i, _i, _i, _f, _P]),
    "ub_weight_map_workspace_bytes": (_ll, [_i, _i, _i]),
    "ub_weight_map": (_i, [_P, _i, _i, _i, _i, C.c_double, C.c_double, _i, _P, _P, _P, _P]),
    "ub_class_balance": (_i, [_P, _i, _i, _i, _i, _P, _P, _P, _P]),
    "ub_augment_workspace_bytes": (_ll, [_i, _i, _i, C.c_double]),
    "ub_augment_batch": (_i, [_P, _P, _i, _i, _i, _P, _P, _i, _i, _i, C.c_double, C.c_double, _P, _P, _P, _P]),
    "ub_mirror_pad": (_i, [_P, _P, _i, _i, _i, _P]),
    "ub_mask_metrics": (_i, [_P, _i, _P, _i, _ll, _P, _P]),
    "ub_cut_tiles": (_i, [_P, _i, _i, _i, _i, _P, _i, _i, _i, _i, _P, _P]),
    "ub_stitch_tiles": (_i, [_P, _P, _i, _i, _i, _P, _i, _i, _P]),
    "ub_colsum": (_i, [_P, _P, _ll, _i, _ll, _P]),
    "ub_partial_rows": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "ub_reduce_partials": (_i, [_P, _i, _ll, _i, _P, _i, _P]),
    "ub_allreduce_sym": (_i, [C.POINTER(_P), C.POINTER(_P), _P, _P, _P, _i, _i, _ll, _ll, _i, _i, _i, _P]),
    "ub_sgd_momentum": (_i, [_P, _P, _P, _ll, _f, _f, _i, _P]),
}
# include/unet_b200_diag.h: only in libunet_b200_diag.so (make libunet_b200_diag.so), never in the product library
DIAG_SIGNATURES = {
    "ub_last_wgrad_plan": (None, [C.POINTER(_i)]),
    "ub_naive_conv3x3": (_i, [_P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _P]),
    "ub_probe_run": (_i, [_i, C.POINTER(C.c_float), _i]),
}

_lib = None


def load():
    """Load libunet_b200.so (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and os.path.exists("/usr/local/cuda/bin/nvcc"):
            # a source-only checkout on a box with the toolkit: build in-tree once (still no fallback path)
            import subprocess
            subprocess.run(["make", "-C", os.path.dirname(LIB_PATH), "libunet_b200.so"], capture_output=True)
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        # experiment switches, e.g. UB_OPTIONS="3=1,1=0" (see ub_set_option in include/unet_b200.h)
        for kv in filter(None, os.environ.get("UB_OPTIONS", "").split(",")):
            k, v = kv.split("=")
            lib.ub_set_option(int(k), int(v))
    return _lib


_diag = None


def load_diag():
    """The diagnostics build (product library + naive comparator + UMMA probes); bring-up scripts only."""
    global _diag
    if _diag is None:
        path = os.path.join(os.path.dirname(LIB_PATH), "libunet_b200_diag.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is not built: make -C {os.path.dirname(LIB_PATH)} libunet_b200_diag.so")
        lib = C.CDLL(path)
        for name, (res, args) in list(SIGNATURES.items()) + list(DIAG_SIGNATURES.items()):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _diag = lib
    return _diag


def check(code, what=""):
    """Turn a non-zero library return code into RuntimeError (the reference's error convention
    for this path is Python exceptions, e.g. torch.cat's RuntimeError on odd sizes)."""
    if code != 0:
        msg = load().ub_error_string(int(code)).decode()
        raise RuntimeError(f"unet_b200 {what}: error {code}: {msg}")
