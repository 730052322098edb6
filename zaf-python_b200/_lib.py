"""ctypes binding of libzafb200.so (the C ABI declared in include/zafb200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzafb200.so")

OK, E_BADARG, E_UNSUPPORTED, E_CUDA, E_NOMEM, E_NCCL = 0, -1, -2, -3, -4, -5
LAYOUT_FRAME_MAJOR, LAYOUT_BIN_MAJOR = 0, 1

_i64, _int, _vp, _sz = C.c_int64, C.c_int, C.c_void_p, C.c_size_t
_pi64 = C.POINTER(C.c_int64)
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes).  Mirrors include/zafb200.h one to one (tests/test_abi.py checks it).
PROTOTYPES = {
    "zafb_last_error": (C.c_char_p, []),
    "zafb_version": (C.c_char_p, []),
    "zafb_device_count": (_int, [C.POINTER(_int)]),
    "zafb_init": (_int, [_int]),
    "zafb_shutdown": (_int, []),
    "zafb_device_info": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_sz), C.c_char_p, _sz]),
    "zafb_malloc": (_int, [_pvp, _sz]),
    "zafb_free": (_int, [_vp]),
    "zafb_host_alloc": (_int, [_pvp, _sz]),
    "zafb_host_free": (_int, [_vp]),
    "zafb_memcpy_h2d": (_int, [_vp, _vp, _sz, _vp]),
    "zafb_memcpy_d2h": (_int, [_vp, _vp, _sz, _vp]),
    "zafb_memcpy_d2d": (_int, [_vp, _vp, _sz, _vp]),
    "zafb_memcpy2d": (_int, [_vp, _sz, _vp, _sz, _sz, _sz, _int, _vp]),
    "zafb_memset": (_int, [_vp, _int, _sz, _vp]),
    "zafb_stream_create": (_int, [_pvp]),
    "zafb_stream_destroy": (_int, [_vp]),
    "zafb_stream_sync": (_int, [_vp]),
    "zafb_device_sync": (_int, []),
    "zafb_event_create": (_int, [_pvp]),
    "zafb_event_destroy": (_int, [_vp]),
    "zafb_event_record": (_int, [_vp, _vp]),
    "zafb_stream_wait_event": (_int, [_vp, _vp]),
    "zafb_event_sync": (_int, [_vp]),
    "zafb_event_elapsed_ms": (_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "zafb_pcm16_to_f32": (_int, [_vp, _i64, _int, _int, _vp, _i64, _vp]),
    "zafb_launch_count": (_i64, []),
    "zafb_host_copy_bytes": (_int, [_vp, _vp]),
    "zafb_stft_geometry": (_int, [_i64, _i64, _i64, _pi64, _pi64, _pi64]),
    "zafb_istft_geometry": (_int, [_i64, _i64, _i64, _pi64, _pi64, _pi64]),
    "zafb_mdct_geometry": (_int, [_i64, _i64, _pi64, _pi64, _pi64]),
    "zafb_imdct_geometry": (_int, [_i64, _i64, _pi64, _pi64]),
    "zafb_cqt_geometry": (_int, [_i64, _i64, _i64, _pi64, _pi64, _pi64]),
    "zafb_stft_plan_create": (_int, [_pvp, _vp, _i64, _i64]),
    "zafb_stft_plan_destroy": (_int, [_vp]),
    "zafb_stft_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "zafb_istft_f32": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _i64, _vp]),
    "zafb_stft_host_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int]),
    "zafb_stft_onesided_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "zafb_istft_onesided_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "zafb_istft_masked_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _i64, _vp, _i64, _vp]),
    "zafb_spec_mirror_f32": (_int, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "zafb_host_mirror_fill": (_int, [_vp, _i64, _i64]),
    "zafb_istft_host_f32": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _i64]),
    "zafb_mdct_plan_create": (_int, [_pvp, _vp, _i64]),
    "zafb_mdct_plan_destroy": (_int, [_vp]),
    "zafb_mdct_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "zafb_imdct_f32": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _i64, _vp]),
    "zafb_mdct_host_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int]),
    "zafb_imdct_host_f32": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _i64]),
    "zafb_dct_plan_create": (_int, [_pvp, _int, _int, _i64]),
    "zafb_dct_plan_destroy": (_int, [_vp]),
    "zafb_dct_f32": (_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "zafb_dct_host_f32": (_int, [_vp, _vp, _i64, _i64, _vp, _i64]),
    "zafb_mel_plan_create": (_int, [_pvp, _vp, _i64, _i64, _vp, _i64, _i64]),
    "zafb_mel_plan_destroy": (_int, [_vp]),
    "zafb_mel_plan_set_route": (_int, [_vp, _int]),
    "zafb_mel_plan_set_precision": (_int, [_vp, _int, _vp]),
    "zafb_melspectrogram_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "zafb_mfcc_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "zafb_melspectrogram_host_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int]),
    "zafb_mfcc_host_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int]),
    "zafb_cqt_plan_create": (_int, [_pvp, _i64, _i64, _vp, _vp, _vp, _i64]),
    "zafb_cqt_plan_destroy": (_int, [_vp]),
    "zafb_cqt_plan_set_route": (_int, [_vp, _int]),
    "zafb_cqt_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _int, _vp]),
    "zafb_cqt_host_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _int]),
    "zafb_spec_abs_f32": (_int, [_vp, _i64, _i64, _i64, _int, _i64, _vp, _vp]),
    "zafb_spec_mask_f32": (_int, [_vp, _i64, _i64, _i64, _int, _vp, _i64, _vp, _vp]),
    "zafb_ratio_min_f32": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "zafb_mul_f32": (_int, [_vp, _vp, _i64, _vp, _vp]),
    "zafb_quantize_f32": (_int, [_vp, _i64, C.c_float, _vp, _vp]),
    "zafb_count_mismatch_u32": (_int, [_vp, _vp, _i64, _pi64, _vp]),
    "zafb_dist_shard_range": (_int, [_i64, _int, _int, _pi64, _pi64]),
    "zafb_dist_nccl_version": (_int, [C.POINTER(_int)]),
    "zafb_dist_unique_id": (_int, [_vp]),
    "zafb_dist_init": (_int, [_pvp, _vp, _int, _int]),
    "zafb_dist_destroy": (_int, [_vp]),
    "zafb_dist_rank": (_int, [_vp, C.POINTER(_int), C.POINTER(_int)]),
    "zafb_dist_broadcast": (_int, [_vp, _vp, _sz, _int, _vp]),
    "zafb_dist_scatter_rows": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "zafb_dist_gather_rows": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "zafb_dist_allgather_rows": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "zafb_dist_max_f64": (_int, [_vp, C.POINTER(C.c_double), _vp]),
    "zafb_dist_peer_export": (_int, [_vp, _vp]),
    "zafb_dist_peer_open": (_int, [_vp, _pvp]),
    "zafb_dist_peer_close": (_int, [_vp]),
}
# not part of the public header: test hooks
_PRIVATE = {
    "zafb_stft_plan_force_kernel": (_int, [_vp, _int]),
    "zafb_dct_plan_force_direct": (_int, [_vp, _int]),
    "zafb_mdct_plan_force_kernel": (_int, [_vp, _int]),
    "zafb_mel_plan_force_kernel": (_int, [_vp, _int]),
    "zafb_cqt_plan_force_kernel": (_int, [_vp, _int]),
}

_lib = None


class ZafbError(RuntimeError):
    pass


def lib():
    """Load libzafb200.so once.  Raises if it has not been built -- by design there is no
    NumPy/CPU substitute behind these functions."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZafbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in {**PROTOTYPES, **_PRIVATE}.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc == OK:
        return
    msg = lib().zafb_last_error().decode("utf-8", "replace")
    if rc == E_BADARG:
        raise ValueError(msg)
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == E_NOMEM:
        raise MemoryError(msg)
    raise ZafbError(f"libzafb200 error {rc}: {msg}")
