"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY: ctypes doors onto the two checkers.

  Ref     -> oracle/_ref/libbiogpt_ref.so  (the unmodified reference, see oracle/Makefile)
  Oracle  -> oracle/liboracle.so           (the plain-C restatement, biogpt_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libbiogpt_ref.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_SO)


_ref_lib = None


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        L = C.CDLL(REF_SO)
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_hparams.argtypes = [C.c_void_p, _i32p]
        L.ref_eval.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.ref_time_eval.restype = C.c_int64
        L.ref_time_eval.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_sample.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_double, C.c_double, C.c_uint32]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_quantize.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_type_size.restype = C.c_size_t
        L.ref_type_size.argtypes = [C.c_int]
        L.ref_blck_size.argtypes = [C.c_int]
        L.ref_vec_dot_type.argtypes = [C.c_int]
        L.ref_from_float.argtypes = [C.c_int, _f32p, _u8p, C.c_int]
        L.ref_from_float_reference.argtypes = [C.c_int, _f32p, _u8p, C.c_int]
        L.ref_to_float.argtypes = [C.c_int, _u8p, _f32p, C.c_int]
        L.ref_vec_dot.restype = C.c_float
        L.ref_vec_dot.argtypes = [C.c_int, C.c_int, _u8p, _u8p]
        L.ref_gelu.argtypes = [_f32p, _f32p, C.c_int]
        L.ref_soft_max.argtypes = [_f32p, _f32p, C.c_int, C.c_int]
        L.ref_norm.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_float]
        L.ref_mul_mat.argtypes = [C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        _ref_lib = L
    return _ref_lib


class Ref:
    """The reference's biogpt_model_load / biogpt_eval on the host CPU."""

    def __init__(self, path: str, n_batch: int = 8, n_threads: int = 0):
        self.L = ref_lib()
        self.h = self.L.ref_load(path.encode(), n_batch)
        if not self.h:
            raise RuntimeError(f"reference failed to load {path}")
        hp = np.zeros(7, dtype=np.int32)
        self.L.ref_hparams(self.h, hp)
        self.n_vocab = int(hp[0])
        self.hparams = hp
        self.n_threads = n_threads or (os.cpu_count() or 1)

    def eval(self, tokens, n_past: int) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.n_vocab, dtype=np.float32)
        rc = self.L.ref_eval(self.h, t, len(t), n_past, self.n_threads, out)
        assert rc == 0
        return out

    def time_eval_us(self, tokens, n_past: int, reps: int) -> int:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        return int(self.L.ref_time_eval(self.h, t, len(t), n_past, self.n_threads, reps))

    def sample(self, logits, top_k=1, top_p=1.0, temp=1.0, seed=0) -> int:
        return int(self.L.ref_sample(self.h, np.ascontiguousarray(logits, dtype=np.float32),
                                     top_k, top_p, temp, seed))

    def close(self):
        if self.h:
            self.L.ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_quantize_file(src: str, dst: str, ftype: int) -> None:
    rc = ref_lib().ref_quantize(src.encode(), dst.encode(), ftype)
    if rc != 0:
        raise RuntimeError(f"ref_quantize failed rc={rc}")


# ---------------------------------------------------------------------------------------------
# the plain-C restatement
# ---------------------------------------------------------------------------------------------
_oracle_lib = None
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")


class _Taps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("embed", "layer0_ln", "layer0_q", "layer0_att", "layer0_out", "final_ln")]


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        L = C.CDLL(ORACLE_SO)
        L.bo_load.restype = C.c_void_p
        L.bo_load.argtypes = [C.c_char_p]
        L.bo_free.argtypes = [C.c_void_p]
        L.bo_reset.argtypes = [C.c_void_p]
        L.bo_hparams.argtypes = [C.c_void_p, _i32p]
        L.bo_eval.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, _f32p]
        L.bo_set_taps.argtypes = [C.c_void_p, C.POINTER(_Taps)]
        L.bo_quantize_row_q8_0.argtypes = [_f32p, _u8p, C.c_int]
        L.bo_quantize_row_q8_1.argtypes = [_f32p, _u8p, C.c_int]
        L.bo_fp32_to_fp16_row.argtypes = [_f32p, _u16p, C.c_int]
        L.bo_dequantize_row.argtypes = [C.c_int, _u8p, _f32p, C.c_int]
        L.bo_vec_dot.restype = C.c_float
        L.bo_vec_dot.argtypes = [C.c_int, C.c_int, _u8p, _u8p]
        L.bo_vec_dot_f32.restype = C.c_float
        L.bo_vec_dot_f32.argtypes = [C.c_int, _f32p, _f32p]
        L.bo_mul_mat.argtypes = [C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.bo_norm.argtypes = [_f32p, _f32p, C.c_int, C.c_float]
        L.bo_soft_max.argtypes = [_f32p, _f32p, C.c_int]
        L.bo_gelu.argtypes = [_f32p, _f32p, C.c_int]
        L.bo_tables.argtypes = [_u16p, _u16p]
        _oracle_lib = L
    return _oracle_lib


class Oracle:
    """bo_load / bo_eval: the C restatement of biogpt_eval."""

    TAP_NAMES = ("embed", "layer0_ln", "layer0_q", "layer0_att", "layer0_out", "final_ln")

    def __init__(self, path: str):
        self.L = oracle_lib()
        self.h = self.L.bo_load(path.encode())
        if not self.h:
            raise RuntimeError(f"oracle failed to load {path}")
        hp = np.zeros(7, dtype=np.int32)
        self.L.bo_hparams(self.h, hp)
        self.hparams = hp
        self.n_vocab, self.d_model = int(hp[0]), int(hp[5])

    def eval(self, tokens, n_past: int, taps: bool = False):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.n_vocab, dtype=np.float32)
        bufs = {}
        if taps:
            st = _Taps()
            for n in self.TAP_NAMES:
                bufs[n] = np.zeros((len(t), self.d_model), dtype=np.float32)
                setattr(st, n, bufs[n].ctypes.data)
            self.L.bo_set_taps(self.h, C.byref(st))
        rc = self.L.bo_eval(self.h, t, len(t), n_past, out)
        if taps:
            self.L.bo_set_taps(self.h, None)
        assert rc == 0, "bo_eval failed"
        return (out, bufs) if taps else out

    def reset(self):
        self.L.bo_reset(self.h)

    def close(self):
        if self.h:
            self.L.bo_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
