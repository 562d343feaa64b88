"""ctypes binding of libemo_b200.so (include/emo_b200.h).  The product path has NO fallback:
if the library is missing or a symbol is absent this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemo_b200.so")

F32, BF16 = 0, 1
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU_NEW, ACT_RELU_MASK_BWD, ACT_GELU_NEW_BWD = 0, 1, 2, 3, 4

vp, i64, i32, f32, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64


class Epilogue(C.Structure):
    _fields_ = [("bias", vp), ("act", i32), ("aux", vp), ("aux_out", vp), ("ld_aux", i64),
                ("aux_scale", f32), ("drop_p", f32), ("seed", u64), ("residual", vp), ("ld_res", i64),
                ("alpha", f32), ("accumulate", i32), ("rowscale", vp), ("colsum", vp),
                ("ln_gamma", vp), ("ln_beta", vp), ("ln_out", vp), ("ld_ln", i64)]


SIGNATURES = {
    "emo_version": ([], i32),
    "emo_last_error": ([], C.c_char_p),
    "emo_embed_fwd": ([vp, vp, i64, i64, vp, vp, vp, vp, i32, i32, i32, f32, f32, u64, i32, vp], i32),
    "emo_embed_rows": ([vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, i32, vp], i32),
    "emo_set_pdl": ([i32], None),
    "emo_stage2_batch": ([vp] * 15 + [i32, i32, i32, i32, i32, vp], i32),
    "emo_stage1_batch": ([vp] * 10 + [i32, i32, i32, vp], i32),
    "emo_performer_decode_step": ([vp, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32,
                                   vp], i32),
    "emo_embed_bwd": ([vp, vp, i64, i64, vp, vp, vp, i32, i32, i32, f32, f32, u64, i64, i32, vp], i32),
    "emo_ln_fwd": ([vp, vp, vp, vp, vp, vp, i64, i32, f32, i32, vp], i32),
    "emo_ln_res_fwd": ([vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, i32, vp], i32),
    "emo_ln_bwd": ([vp, vp, vp, vp, vp, vp, vp, vp, f32, u64, vp, vp, vp, i64, i32, i32, vp], i32),
    "emo_dropout_apply": ([vp, vp, i64, f32, u64, i32, vp], i32),
    "emo_gemm": ([i32, i64, i64, i64, vp, i64, vp, i64, vp, i64, i32, i32, C.POINTER(Epilogue), vp], i32),
    "emo_colsum": ([vp, i64, i64, i64, vp, i32, vp], i32),
    "emo_gemm_ln_res": ([i64, i64, vp, i64, vp, i64, vp, f32, u64, vp, i64, vp, vp, f32, vp, i64, vp, i64, vp, vp, vp], i32),
    "emo_favor_nseg": ([i32, i32, i32, i32], i32),
    "emo_favor_fwd": ([vp, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, i32, i32, i32, i32, vp], i32),
    "emo_favor_bwd": ([vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp], i32),
    "emo_favor_step": ([vp, vp, vp, i64, vp, vp, vp, i64, i32, i32, i32, vp], i32),
    "emo_attn_fwd": ([vp, vp, vp, i64, i64, vp, i64, vp, i32, i32, i32, i32, f32, f32, u64, i32, vp], i32),
    "emo_attn_bwd": ([vp, vp, vp, i64, i64, vp, vp, i64, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, f32, f32,
                      u64, i32, vp], i32),
    "emo_attn_decode_step": ([vp, i64, vp, i64, vp, vp, i64, i32, i32, f32, i32, vp], i32),
    "emo_relattn_decode_step": ([vp, i64, vp, i64, vp, vp, vp, vp, i32, vp, i64, i32, i32, f32, i32, vp], i32),
    "emo_relattn_fwd": ([vp, vp, vp, i64, i64, vp, i64, vp, vp, vp, i64, vp, i32, i32, i32, i32, f32, f32, u64, i32, vp], i32),
    "emo_relattn_bwd": ([vp, vp, vp, i64, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, i64, vp, vp, vp,
                         i32, i32, i32, i32, f32, f32, u64, i32, vp], i32),
    "emo_ce_count": ([vp, i64, i64, i64, i64, i64, vp, vp], i32),
    "emo_ce_fwd_bwd": ([vp, i64, vp, i64, i64, i64, i64, i32, i64, vp, f32, vp, vp, vp, vp, i64, i32, vp], i32),
    "emo_sumsq": ([vp, i64, vp, vp], i32),
    "emo_adam_step": ([vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, i64, vp, f32, f32, i32, vp], i32),
    "emo_cast": ([vp, vp, i64, i32, i32, vp], i32),
    "emo_sample": ([vp, i64, i32, i32, f32, f32, vp, i32, vp, vp, vp, vp], i32),
    "emo_sample_rows": ([vp, i64, i32, i32, vp, f32, vp, i32, vp, vp, vp, vp], i32),
    "emo_logits_sample": ([vp, i64, vp, vp, vp, i64, vp, i32, i32, i32, vp, i64, f32, vp, f32, vp, i32, vp, vp, vp, vp], i32),
}

_lib = None


class EmoError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmoError(
                "libemo_b200.so not found at %s -- build it with `python -m emo_disentanger_b200.build` "
                "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(l, name)      # AttributeError if the ABI is incomplete: fail loudly
            fn.argtypes = args
            fn.restype = res
        _lib = l
    return _lib


LAUNCHES = 0     # C-ABI calls that returned EMO_OK (each launches one kernel of ours); read by bench.py


def check(rc, what=""):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        raise EmoError("%s failed (status %d): %s" % (what, rc, lib().emo_last_error().decode()))
