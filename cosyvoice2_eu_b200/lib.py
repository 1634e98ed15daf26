"""ctypes binding of libcv2eu_b200.so (C ABI declared in include/cv2eu_b200.h).

There is no fallback: if the shared library is missing or no B200 is present, loading / engine creation raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcv2eu_b200.so")

# every symbol include/cv2eu_b200.h declares (checked by tests/test_cabi_symbols.py)
SYMBOLS = [
    "cv2_last_error", "cv2_version", "cv2_engine_create", "cv2_engine_destroy", "cv2_engine_set_tensor", "cv2_engine_finalize",
    "cv2_engine_last_launches", "cv2_engine_set_option", "cv2_engine_read_ranges", "cv2_debug_set_ffn_trace", "cv2_engine_set_seed_ptr", "cv2_engine_set_profiling", "cv2_engine_read_profile", "cv2_estimator_workspace_bytes", "cv2_estimator_forward", "cv2_flow_workspace_bytes",
    "cv2_flow_forward", "cv2_stream_state_bytes", "cv2_stream_state_reset_slot", "cv2_flow_stream_workspace_bytes",
    "cv2_flow_forward_stream", "cv2_encoder_workspace_bytes", "cv2_encoder_forward", "cv2_hift_workspace_bytes", "cv2_hift_forward", "cv2_hift_forward_pcm16", "cv2_crossfade", "cv2_mel_time_stretch", "cv2_prompt_mel_frames", "cv2_prompt_mel_workspace_bytes",
    "cv2_prompt_mel", "cv2_kaldi_fbank_frames", "cv2_kaldi_fbank", "cv2_resample_16k_24k_len", "cv2_resample_16k_24k", "cv2_op_gemm_tap", "cv2_op_flash_attn",
    "cv2_op_rel_attn", "cv2_op_source_stft", "cv2_op_istft", "cv2_op_nsf_source",
]

_lib = None


class Cv2Error(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Cv2Error(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(cosyvoice2_eu_b200 has no fallback path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, u64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong, C.c_size_t
    lib.cv2_last_error.restype = C.c_char_p
    lib.cv2_version.restype = i32
    lib.cv2_engine_create.argtypes = [C.POINTER(vp), i32]
    lib.cv2_engine_destroy.argtypes = [vp]
    lib.cv2_engine_destroy.restype = None
    lib.cv2_engine_set_tensor.argtypes = [vp, C.c_char_p, vp, i32, i32, C.POINTER(C.c_int64)]
    lib.cv2_engine_finalize.argtypes = [vp, i32, i32]
    lib.cv2_engine_last_launches.argtypes = [vp]
    lib.cv2_engine_last_launches.restype = i64
    lib.cv2_engine_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.cv2_engine_read_ranges.argtypes = [vp, C.POINTER(f32), i32]
    lib.cv2_debug_set_ffn_trace.argtypes = [vp]
    lib.cv2_engine_set_seed_ptr.argtypes = [vp, vp]
    lib.cv2_engine_set_profiling.argtypes = [vp, i32]
    lib.cv2_engine_read_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), i32]
    lib.cv2_estimator_workspace_bytes.argtypes = [vp, i32, i32]
    lib.cv2_estimator_workspace_bytes.restype = sz
    lib.cv2_estimator_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz]
    lib.cv2_flow_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.cv2_flow_workspace_bytes.restype = sz
    lib.cv2_flow_forward.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, vp,
                                     C.POINTER(f32), i32, f32, vp, i32, vp, vp, vp, sz]
    lib.cv2_stream_state_bytes.argtypes = [i32, i32, i32]
    lib.cv2_stream_state_bytes.restype = sz
    lib.cv2_stream_state_reset_slot.argtypes = [vp, vp, sz, i32, i32, i32, i32]
    lib.cv2_flow_stream_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.cv2_flow_stream_workspace_bytes.restype = sz
    lib.cv2_flow_forward_stream.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, i64, vp, vp, vp, i32, i32, i32, vp,
                                            C.POINTER(f32), i32, f32, vp, i32, vp, sz, i32, vp, sz]
    lib.cv2_encoder_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.cv2_encoder_workspace_bytes.restype = sz
    lib.cv2_encoder_forward.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, i32, vp, sz]
    lib.cv2_hift_workspace_bytes.argtypes = [vp, i32, i32]
    lib.cv2_hift_workspace_bytes.restype = sz
    lib.cv2_hift_forward.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, u64, vp, vp, vp, i32, vp, sz]
    lib.cv2_hift_forward_pcm16.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, u64, vp, vp, vp, vp, i32, vp, sz]
    lib.cv2_crossfade.argtypes = [vp, vp, vp, vp, i32]
    lib.cv2_mel_time_stretch.argtypes = [vp, vp, i32, vp, i32, i32]
    lib.cv2_prompt_mel_frames.argtypes = [i32]
    lib.cv2_prompt_mel_workspace_bytes.argtypes = [i32, i32]
    lib.cv2_prompt_mel_workspace_bytes.restype = sz
    lib.cv2_prompt_mel.argtypes = [vp, vp, i64, vp, i32, i32, vp, vp, vp, sz]
    lib.cv2_kaldi_fbank_frames.argtypes = [i32]
    lib.cv2_kaldi_fbank.argtypes = [vp, vp, i64, vp, i32, i32, vp, vp, i32]
    lib.cv2_resample_16k_24k_len.argtypes = [i32]
    lib.cv2_resample_16k_24k.argtypes = [vp, vp, i64, vp, i32, i32, vp, i64, vp]
    lib.cv2_op_gemm_tap.argtypes = [vp, vp, i32, i32, i32, i64, vp, i32, i32, vp, i32, i32, C.POINTER(i32), vp, i32, vp, vp, f32,
                                    i32, f32, vp, vp, i32, i32, vp, f32, vp, i32, vp, vp, vp, vp]
    lib.cv2_op_flash_attn.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32]
    lib.cv2_op_rel_attn.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32]
    lib.cv2_op_source_stft.argtypes = [vp, vp, i32, vp, vp, i32, i32]
    lib.cv2_op_istft.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    lib.cv2_op_nsf_source.argtypes = [vp, vp, i32, vp, vp, u64, vp, vp, vp, vp, i32]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("cv2_last_error", "cv2_version", "cv2_engine_destroy", "cv2_engine_last_launches") and \
                not name.endswith("_bytes"):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise Cv2Error(load().cv2_last_error().decode())


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
