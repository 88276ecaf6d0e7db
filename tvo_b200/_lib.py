"""ctypes binding of libtvo_b200.so (the C ABI declared in include/tvo_b200.h).

The product has NO CPU fallback: if the shared library is missing or a call fails, a
RuntimeError / ValueError is raised with the library's own message.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtvo_b200.so")

SEL = {"randparents": 0, "batch_fitparents": 1, "fitparents_fixed": 2}
MUT = {"randflip": 0, "sparseflip": 1, "cross": 2, "cross_randflip": 3, "cross_sparseflip": 4}


class TvoError(RuntimeError):
    pass


class EvoConfig(C.Structure):
    _fields_ = [
        ("parent_selection", C.c_int32),
        ("mutation", C.c_int32),
        ("n_parents", C.c_int32),
        ("n_children", C.c_int32),
        ("n_generations", C.c_int32),
        ("flags", C.c_int32),
        ("p_bf", C.c_double),
        ("seed", C.c_uint64),
        ("step", C.c_uint32),
        ("n_offset", C.c_uint32),
        ("fit_norm", C.c_void_p),
    ]


EVO_LPJ_CURRENT = 1


_P, _I32, _I64, _U64, _U32, _F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "tvo_abi_version": [],
    "tvo_last_error": [],
    "tvo_launch_count": [],
    "tvo_reset_launch_count": [],
    "tvo_kernel_timing_enable": [_I32],
    "tvo_kernel_timing_read": [_I32, C.POINTER(C.c_double), C.POINTER(C.c_int64)],
    "tvo_device_sm_count": [C.POINTER(C.c_int)],
    "tvo_pack_states": [_P, _P, _I64, _I32, _P],
    "tvo_unpack_states": [_P, _P, _I64, _I32, _P],
    "tvo_state_matrix": [_P, _I32, _P],
    "tvo_bsc_theta_ws_bytes": [_I32, _I32, _I32],
    "tvo_bsc_stats_len": [_I32, _I32],
    "tvo_bsc_estep_scratch_bytes": [_I64, _I32, _I32],
    "tvo_bsc_mstep_scratch_bytes": [_I64, _I32, _I32, _I32],
    "tvo_bsc_estep_variant": [_I32],
    "tvo_bsc_symmetrize": [_P, _P, _I32, _P],
    "tvo_bsc_epoch_solve_work_doubles": [_I32],
    "tvo_bsc_epoch_solve_variant": [_I32],
    "tvo_bsc_epoch_solve": [_P, _P, _P, _P, _I32, _I32, _P],
    "tvo_random_states": [_P, _P, _I64, _I32, _I32, _F64, _U64, _U32, _U32, _P],
    "tvo_noisyor_theta_ws_bytes": [_I32, _I32, _I32],
    "tvo_noisyor_stats_len": [_I32, _I32],
    "tvo_sssc_theta_ws_bytes": [_I32, _I32, _I32],
    "tvo_sssc_stats_len": [_I32, _I32],
    "tvo_sssc_scratch_bytes": [_I64, _I32, _I32, _I32],
    "tvo_mixture_ws_bytes": [_I32, _I32, _I32],
    "tvo_mixture_stats_len": [_I32, _I32],
    "tvo_mixture_scratch_bytes": [_I64, _I32, _I32, _I32],
}
for _sfx in ("f64", "f32"):
    _PROTOS.update({
        f"tvo_bsc_prepare_{_sfx}": [_P, _P, _I32, _P, _I32, _I32, _P, _P],
        f"tvo_bsc_epoch_finish_{_sfx}": [_P, _P, _P, _P, _P, _I32, _P, _I32, _I32, _P, _P],
        f"tvo_bsc_lpj_{_sfx}": [_P, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P],
        f"tvo_evo_generation_{_sfx}": [_P, _I64, _P, _I64, _P, _I32, _I32, _P, C.POINTER(EvoConfig), _I32, _P, _P,
                                       _I64, _I32, _P],
        f"tvo_fit_norm_{_sfx}": [_P, _I64, _P, _I64, _I32, _P, _P],
        f"tvo_mark_redundant_{_sfx}": [_P, _P, _P, _I64, _P, _I64, _I32, _I32, _I32, _P],
        f"tvo_select_topS_{_sfx}": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P],
        f"tvo_estep_evo_bsc_{_sfx}": [_P, _I64, _P, _P, _P, _P, C.POINTER(EvoConfig), _P, _P, _I64, _I64, _I32, _I32,
                                      _I32, _P, _P],
        f"tvo_lpj2pjc_{_sfx}": [_P, _P, _I64, _I32, _P],
        f"tvo_mixture_prepare_{_sfx}": [_I32, _P, _P, _P, _I32, _I32, _P, _P],
        f"tvo_mixture_lpj_{_sfx}": [_I32, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _I32, _P],
        f"tvo_mixture_mstep_{_sfx}": [_I32, _P, _I64, _P, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _I32, _P],
        f"tvo_mixture_free_energy_{_sfx}": [_I32, _P, _I64, _P, _P, _I64, _P, _I64, _I32, _I32, _P, _P],
        f"tvo_state_marginals_{_sfx}": [_P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _P],
        f"tvo_sample_states_probs_{_sfx}": [_P, _I64, _I32, _P, _I64, _P, _I64, _I32, _I32, _U64, _U32, _U32, _U32, _P],
        f"tvo_mean_posterior_{_sfx}": [_P, _P, _P, _I64, _I32, _I64, _P],
        f"tvo_bsc_free_energy_{_sfx}": [_P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _P, _P],
        f"tvo_noisyor_prepare_{_sfx}": [_P, _P, _I32, _I32, _P, _P],
        f"tvo_noisyor_lpj_{_sfx}": [_P, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P],
        f"tvo_estep_evo_noisyor_{_sfx}": [_P, _I64, _P, _P, _P, _P, C.POINTER(EvoConfig), _P, _I64, _I32, _I32, _I32, _P,
                                          _P],
        f"tvo_noisyor_mstep_{_sfx}": [_P, _I64, _P, _P, _I64, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _P],
        f"tvo_logsumexp_sum_{_sfx}": [_P, _I64, _P, _I64, _I32, _P, _P],
        f"tvo_sssc_prepare_{_sfx}": [_P, _P, _P, _P, _P, _I32, _I32, _P, _P],
        f"tvo_sssc_lpj_{_sfx}": [_P, _I64, _P, _I64, _P, _P, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _P],
        f"tvo_estep_evo_sssc_{_sfx}": [_P, _I64, _P, _P, _P, _P, C.POINTER(EvoConfig), _P, _P, _I64, _P, _I64, _I32, _I32,
                                       _I32, _P, _P],
        f"tvo_sssc_mstep_{_sfx}": [_P, _I64, _P, _P, _I64, _P, _I64, _P, _P, _P, _I64, _P, _I64, _I32, _I32, _I32, _I32,
                                   _P],
        f"tvo_tvae_layer0_{_sfx}": [_P, _I64, _I32, _P, _P, _I32, _P, _P],
        f"tvo_tvae_layer0_grad_{_sfx}": [_P, _I64, _I32, _P, _I32, _P, _P],
        f"tvo_tvae_lpj_{_sfx}": [_I32, _P, _I64, _P, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P],
        f"tvo_bsc_mstep_{_sfx}": [_P, _I64, _P, _P, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _I32, _I32, _I32, _I32, _P],
    })
_RESTYPES = {
    "tvo_last_error": C.c_char_p,
    "tvo_launch_count": _I64,
    "tvo_reset_launch_count": None,
    "tvo_kernel_timing_enable": None,
    "tvo_bsc_theta_ws_bytes": _I64,
    "tvo_bsc_stats_len": _I64,
    "tvo_bsc_estep_scratch_bytes": _I64,
    "tvo_bsc_mstep_scratch_bytes": _I64,
    "tvo_noisyor_theta_ws_bytes": _I64,
    "tvo_noisyor_stats_len": _I64,
    "tvo_sssc_theta_ws_bytes": _I64,
    "tvo_sssc_stats_len": _I64,
    "tvo_sssc_scratch_bytes": _I64,
    "tvo_bsc_epoch_solve_work_doubles": _I64,
    "tvo_mixture_ws_bytes": _I64,
    "tvo_mixture_stats_len": _I64,
    "tvo_mixture_scratch_bytes": _I64,
}

_lib = None


def declared_symbols():
    return sorted(_PROTOS)


def lib():
    """Load the shared library (once).  Raises if it was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TvoError(
                f"{LIB_PATH} not found: the CUDA extension is not built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`); tvo_b200 has no CPU fallback"
            )
        L = C.CDLL(LIB_PATH)
        for name, argtypes in _PROTOS.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        if L.tvo_abi_version() != 2:
            raise TvoError("libtvo_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc == 0:
        return
    msg = lib().tvo_last_error().decode(errors="replace")
    if rc == 1:
        raise ValueError(f"tvo_b200 {what}: {msg}")
    raise TvoError(f"tvo_b200 {what} failed (code {rc}): {msg}")


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


def sfx(dtype):
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise ValueError(f"precision must be torch.float32 or torch.float64, got {dtype}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must live on a CUDA device."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TvoError("tvo_b200 kernels need CUDA tensors (got a CPU tensor); there is no CPU fallback")
    return C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """The current CUDA stream of the current device as a cudaStream_t.  torch.cuda.current_stream() costs ~15 us per
    call in this torch (device-index resolution through os.environ); a dozen calls per EM iteration were 0.18 ms of
    host time, visible at small shards: the raw accessor is used when the build has it."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().tvo_launch_count())


def reset_launch_count():
    lib().tvo_reset_launch_count()


TIMED_ESTEP_EVO_BSC, TIMED_BSC_MSTEP, TIMED_BSC_GEMM_YW, TIMED_BSC_GEMM_ETY = 1, 2, 3, 4


def kernel_timing_enable(on=True):
    lib().tvo_kernel_timing_enable(1 if on else 0)


def kernel_timing_read(which):
    """(total ms, launches) of timed kernel `which` since kernel_timing_enable; synchronises its events."""
    ms, n = C.c_double(0.0), C.c_int64(0)
    check(lib().tvo_kernel_timing_read(which, C.byref(ms), C.byref(n)), "tvo_kernel_timing_read")
    return float(ms.value), int(n.value)
