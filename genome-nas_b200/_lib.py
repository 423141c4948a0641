"""ctypes binding of include/gnas.h (the C ABI of the sm_100a kernels).

The product path loads `libgnas.so` (built by build.py with nvcc for sm_100a) and
fails loudly when it is missing: there is no CPU or eager-PyTorch fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgnas.so")

MAX_EDGES, N_PRIMS, RNN_EDGES, RNN_PRIMS = 14, 9, 36, 5


class CellDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("C", C.c_int32), ("C_pp", C.c_int32), ("C_p", C.c_int32),
                ("L_pp", C.c_int32), ("L_p", C.c_int32), ("L_out", C.c_int32), ("stride", C.c_int32),
                ("n_edges", C.c_int32), ("has_pre", C.c_int32), ("pre0_factorized", C.c_int32),
                ("n_prims", C.c_int32), ("training", C.c_int32), ("bn_momentum", C.c_float),
                ("sw", (C.c_uint8 * N_PRIMS) * MAX_EDGES),
                ("w_off", ((C.c_int64 * 4) * N_PRIMS) * MAX_EDGES),
                ("pre_w_off", C.c_int64 * 2),
                ("bn_off", ((C.c_int64 * 2) * N_PRIMS) * MAX_EDGES),
                ("pre_bn_off", C.c_int64 * 2)]


class RhnDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("training", C.c_int32),
                ("n_prims", C.c_int32), ("bn_momentum", C.c_float),
                ("sw", (C.c_uint8 * RNN_PRIMS) * RNN_EDGES)]


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_SIGS = {
    "gnas_version": (C.c_int, []),
    "gnas_last_error": (C.c_char_p, []),
    "gnas_kernel_launches": (C.c_int64, []),
    "gnas_profile_enable": (C.c_int, [C.c_int]),
    "gnas_profile_report": (C.c_int64, [C.c_char_p, C.c_int64]),
    "gnas_profile_event_overhead_ms": (C.c_double, [C.c_int, _P]),
    "gnas_side_streams": (C.c_int, [C.c_int]),
    "gnas_cell_workspace": (C.c_int, [C.POINTER(CellDesc), _I64P, _I64P]),
    "gnas_cell_forward": (C.c_int, [C.POINTER(CellDesc), _P, _P, _P, _P, _P, _I64P, _P, _P, _P]),
    "gnas_cell_backward": (C.c_int, [C.POINTER(CellDesc), _P, _P, _P, _P, _P, _I64P, _P, _P, _P, _P, _P, _P, _P,
                                     C.c_int, _P]),
    "gnas_stem_workspace": (C.c_int, [C.c_int, C.c_int, C.c_int, _I64P, _I64P]),
    "gnas_stem_forward": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gnas_stem_backward": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gnas_onehot_expand": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "gnas_rhn_workspace": (C.c_int, [C.POINTER(RhnDesc), _I64P, _I64P]),
    "gnas_rhn_forward": (C.c_int, [C.POINTER(RhnDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gnas_rhn_backward": (C.c_int, [C.POINTER(RhnDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    C.c_int, _P]),
    "gnas_head_workspace": (C.c_int, [C.c_int] * 5 + [_I64P, _I64P]),
    "gnas_head_forward": (C.c_int, [C.c_int] * 6 + [_P] * 9),
    "gnas_head_backward": (C.c_int, [C.c_int] * 6 + [_P] * 12 + [C.c_int, _P]),
    "gnas_loss_workspace": (C.c_int, [_I64P]),
    "gnas_bce_forward": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "gnas_bce_backward": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    "gnas_tar_forward": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P]),
    "gnas_tar_backward": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, _P]),
    "gnas_grad_sumsq": (C.c_int, [_P, C.c_int64, _P, _P]),
    "gnas_sgd_step": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int, _P, C.c_float, _P]),
    "gnas_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int, _P, _P, C.c_float, _P]),
}
EXPORTS = tuple(_SIGS)

_lib = None
_device_type = "cuda"
LAUNCHES = 0  # C-ABI calls made (each enqueues >= 1 of our kernels); bench.py reports the kernel count separately


def _bind(path):
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"genome_nas_b200: {LIB_PATH} is missing. Build the sm_100a extension first "
                "(python genome-nas_b200/build.py or __graft_entry__.build()); there is no fallback path.")
        _lib = _bind(LIB_PATH)
    return _lib


def device_type():
    return _device_type


def _use_library_for_tests(path, device_type):
    """TEST HOOK ONLY (tests/emul): point the binding at the CPU build of the same kernel
    sources running on the execution-model emulator.  Never called by the package."""
    global _lib, _device_type
    _lib = _bind(path) if path else None
    _device_type = device_type


def kernel_launches():
    return int(lib().gnas_kernel_launches())


def profile(enable):
    lib().gnas_profile_enable(1 if enable else 0)


def side_streams(on):
    """Enable / disable the library's side streams (concurrent independent kernel chains); returns the previous setting."""
    return bool(lib().gnas_side_streams(1 if on else 0))


def profile_report():
    """{kernel name: (launches, total device ms)} recorded since profile(True)."""
    buf = C.create_string_buffer(1 << 20)
    n = lib().gnas_profile_report(buf, len(buf))
    out = {}
    if n > 0:
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit("\t", 2)
            out[name] = (int(cnt), float(ms))
    return out


def check(rc, what=""):
    if rc != 0:
        msg = lib().gnas_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"gnas {what} failed with code {rc}: {msg}")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(t):
    if _device_type != "cuda":
        return C.c_void_p(0)
    import torch
    # one process per GPU: kernels, per-device kernel attributes and the library's side streams all belong to the
    # CURRENT device -- a tensor on another device would be launched on the wrong one
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"genome_nas_b200: tensor on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; call torch.cuda.set_device({t.device.index}) first")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require(t, name, dtype=None):
    import torch
    dtype = dtype or torch.float32
    if t.device.type != _device_type:
        raise RuntimeError(f"{name}: expected a {_device_type} tensor, got {t.device}")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: tensor must be contiguous")
    return t
