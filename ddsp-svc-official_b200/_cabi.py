"""ctypes binding of include/ddsp_b200.h (the C ABI of lib/libddsp_b200.so).

There is no CPU fallback: if the library cannot be loaded the import of any op fails loudly.
"""
import ctypes as C
import os

from . import build as _build

_lib = None

c_f32p = C.c_void_p
i64 = C.c_int64
u64 = C.c_uint64

_PROTOS = {
    'ddsp_b200_version': (C.c_int, []),
    'ddsp_b200_strerror': (C.c_char_p, [C.c_int]),
    'ddsp_b200_last_cuda_error': (C.c_int, []),
    'ddsp_b200_last_launch_count': (C.c_int, []),
    'ddsp_b200_upsample': (C.c_int, [c_f32p, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    'ddsp_b200_fo_to_rot_workspace_bytes': (C.c_size_t, [C.c_int, i64]),
    'ddsp_b200_fo_to_rot': (C.c_int, [c_f32p, C.c_int, i64, C.c_double, c_f32p, C.c_int, c_f32p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    'ddsp_b200_remove_above_fmax': (C.c_int, [c_f32p, i64, i64, c_f32p, i64, i64, C.c_float, C.c_int, C.c_int,
                                              C.c_int, C.c_int, c_f32p, C.c_void_p]),
    'ddsp_b200_phase': (C.c_int, [c_f32p, i64, i64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p, C.c_int,
                                  c_f32p, C.c_void_p, c_f32p, C.c_void_p]),
    'ddsp_b200_combsubfast': (C.c_int, [c_f32p, c_f32p, c_f32p, i64, i64, c_f32p, i64, i64, C.c_void_p, c_f32p,
                                        c_f32p, u64, c_f32p, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p,
                                        C.c_void_p]),
    'ddsp_b200_phase_stream': (C.c_int, [c_f32p, i64, i64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p, C.c_void_p, i64,
                                         c_f32p, C.c_void_p, C.c_void_p]),
    'ddsp_b200_combsubfast_stream': (C.c_int, [c_f32p, c_f32p, c_f32p, i64, i64, c_f32p, i64, i64, C.c_void_p, c_f32p,
                                               u64, C.c_void_p, i64, c_f32p, C.c_int, C.c_int, C.c_int, C.c_double,
                                               c_f32p, C.c_void_p]),
    'ddsp_b200_combsubfast_backward': (C.c_int, [c_f32p, c_f32p, c_f32p, i64, i64, c_f32p, i64, i64, C.c_void_p,
                                                 c_f32p, u64, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_double,
                                                 c_f32p, c_f32p, c_f32p, i64, i64, C.c_void_p]),
    'ddsp_b200_performer_features': (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                               c_f32p, C.c_void_p]),
    'ddsp_b200_performer_project_features': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                       C.c_int, C.c_float, c_f32p, C.c_void_p]),
    'ddsp_b200_performer_attention_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'ddsp_b200_performer_attention': (C.c_int, [c_f32p, c_f32p, c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.c_float, c_f32p, C.c_void_p, C.c_size_t,
                                                C.c_void_p]),
    'ddsp_b200_embed_sum': (C.c_int, [c_f32p, i64, i64, i64, c_f32p, i64, i64, c_f32p, i64, i64, c_f32p, i64, i64,
                                      c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, i64, C.c_int, C.c_int,
                                      C.c_int, c_f32p, C.c_void_p]),
    'ddsp_b200_glu_dwconv_silu': (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p,
                                            C.c_void_p]),
    'ddsp_b200_linear_tf32x3': (C.c_int, [c_f32p, i64, c_f32p, i64, c_f32p, c_f32p, i64, c_f32p, i64, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p]),
    'ddsp_b200_linear_tf32x3_ex': (C.c_int, [c_f32p, i64, c_f32p, c_f32p, i64, c_f32p, c_f32p, i64, c_f32p, i64, c_f32p, c_f32p,
                                             C.c_float, c_f32p, i64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_linear_glu': (C.c_int, [c_f32p, i64, c_f32p, c_f32p, i64, c_f32p, c_f32p, i64, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]),
    'ddsp_b200_dwconv_silu': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    'ddsp_b200_qkv_heads': (C.c_int, [c_f32p, i64, c_f32p, c_f32p, i64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_favor_features': (C.c_int, [c_f32p, c_f32p, C.c_int, C.c_int, C.c_float, c_f32p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p]),
    'ddsp_b200_favor_context': (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_favor_output': (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_tc_microbench': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_apply_frame_mask': (C.c_int, [c_f32p, c_f32p, i64, i64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_apply_volume_mask': (C.c_int, [c_f32p, c_f32p, i64, i64, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ddsp_b200_phase_stream_full': (C.c_int, [c_f32p, i64, i64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p, C.c_void_p, i64,
                                              c_f32p, C.c_void_p, c_f32p, C.c_void_p]),
    'ddsp_b200_combsub_stream': (C.c_int, [c_f32p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, i64, i64, c_f32p, i64, i64,
                                           C.c_void_p, c_f32p, u64, i64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p,
                                           c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'ddsp_b200_sins_stream': (C.c_int, [c_f32p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, i64, i64, c_f32p, i64, i64,
                                        c_f32p, c_f32p, u64, i64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p, c_f32p, c_f32p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    'ddsp_b200_mel_spectrogram': (C.c_int, [c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_float, c_f32p, C.c_int, C.c_void_p]),
    'ddsp_b200_sinc_resample': (C.c_int, [c_f32p, C.c_int, C.c_int, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int, C.c_void_p]),
    'ddsp_b200_interp_frames': (C.c_int, [c_f32p, i64, i64, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_double,
                                          c_f32p, C.c_int, C.c_void_p]),
    'ddsp_b200_sola_splice': (C.c_int, [c_f32p, C.c_int, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p,
                                        C.c_void_p]),
    'ddsp_b200_pad_frames': (C.c_int, [c_f32p, i64, i64, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    'ddsp_b200_groupnorm_leaky': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p]),
    'ddsp_b200_embed_sum_ln': (C.c_int, [c_f32p, i64, i64, c_f32p, i64, i64, c_f32p, i64, i64, c_f32p, i64, i64] + [c_f32p] * 7 +
                               [i64, c_f32p, c_f32p, C.c_float, C.c_int, C.c_int, C.c_int, c_f32p, c_f32p, C.c_void_p]),
    'ddsp_b200_frequency_filter_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'ddsp_b200_frequency_filter': (C.c_int, [c_f32p, c_f32p, i64, i64, C.c_int, C.c_int, C.c_float, C.c_int, c_f32p,
                                             i64, i64, C.c_double, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    'ddsp_b200_combsub_workspace_bytes': (C.c_size_t, [C.c_int] * 5),
    'ddsp_b200_combsub': (C.c_int, [c_f32p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, i64, i64, c_f32p, i64, i64,
                                    C.c_void_p, c_f32p, c_f32p, u64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p,
                                    c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'ddsp_b200_sins_workspace_bytes': (C.c_size_t, [C.c_int] * 5),
    'ddsp_b200_sins': (C.c_int, [c_f32p, C.c_int, c_f32p, C.c_int, c_f32p, C.c_int, i64, i64, c_f32p, i64, i64,
                                 c_f32p, c_f32p, u64, C.c_int, C.c_int, C.c_int, C.c_double, c_f32p, c_f32p, c_f32p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
}


class DDSPB200Error(RuntimeError):
    pass


def lib():
    """The loaded shared library (built on first use if nvcc is available and it is stale)."""
    global _lib
    if _lib is None:
        path = _build.LIB
        override = os.environ.get('DDSP_B200_LIB')           # experiments: load a differently built library
        if override:
            path = override
        elif not os.path.exists(path) or (_build.is_stale() and os.environ.get('DDSP_B200_NO_REBUILD') != '1'):
            try:
                path = _build.build()
            except Exception as e:                       # no nvcc on this box: use the shipped .so if any
                if not os.path.exists(_build.LIB):
                    raise DDSPB200Error(f'libddsp_b200.so is missing and could not be built: {e}') from e
                path = _build.LIB
        handle = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)                    # AttributeError if a symbol is missing
            fn.restype, fn.argtypes = res, args
        if handle.ddsp_b200_version() != 1:
            raise DDSPB200Error('libddsp_b200.so ABI version mismatch')
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_PROTOS)


def check(rc):
    if rc != 0:
        L = lib()
        msg = L.ddsp_b200_strerror(rc).decode()
        if rc == -4:
            msg += f' [cudaError {L.ddsp_b200_last_cuda_error()}]'
        if rc == -5:
            raise ValueError(msg)
        raise DDSPB200Error(msg)
