"""ctypes view of include/bgpt_cuda.h (libbgpt_cuda.so) for tests and bench.py.

This is NOT a second implementation: every function here forwards to the C ABI that the C++
host library (host/biogpt_b200.cpp) also calls.  `Model.load` walks a `.bin` file the way the
reference's biogpt_model_load does (/root/reference/biogpt.cpp:27-453) and hands each tensor's
raw bytes to bgpt_cuda_upload_tensor.  If the shared library is missing, importing succeeds but
any use raises -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import ggml_file as gf

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BGPT_CUDA_LIB") or os.path.join(HERE, "csrc", "libbgpt_cuda.so")   # BGPT_CUDA_LIB: an alternative build, for A/B timing

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")

# every symbol include/bgpt_cuda.h declares (tests/test_capi_symbols.py checks header <-> list)
SYMBOLS = [
    "bgpt_cuda_last_error", "bgpt_cuda_device_count", "bgpt_cuda_version",
    "bgpt_cuda_model_create", "bgpt_cuda_upload_tensor", "bgpt_cuda_set_tables",
    "bgpt_host_build_tables", "bgpt_cuda_model_finalize", "bgpt_cuda_model_free",
    "bgpt_cuda_eval", "bgpt_cuda_eval_topk", "bgpt_cuda_set_chain", "bgpt_cuda_eval_device", "bgpt_cuda_logits_device", "bgpt_cuda_synchronize",
    "bgpt_cuda_decode_greedy", "bgpt_cuda_set_decode_path", "bgpt_cuda_get_decode_path", "bgpt_cuda_decode_kernel_generation", "bgpt_cuda_set_batch_path", "bgpt_cuda_get_batch_path", "bgpt_cuda_debug_read_buffer", "bgpt_cuda_debug_read_prof", "bgpt_cuda_debug_read_trace", "bgpt_cuda_debug_read_rows_trace", "bgpt_cuda_get_eval_path", "bgpt_cuda_set_tc_min_rows", "bgpt_cuda_set_tcx_min_rows", "bgpt_cuda_set_tcw", "bgpt_cuda_set_f16_tc_min_rows", "bgpt_cuda_op_quantize_weights", "bgpt_cuda_set_streams", "bgpt_cuda_eval_streams", "bgpt_cuda_eval_streams_topk", "bgpt_cuda_decode_greedy_streams",
    "bgpt_cuda_hparams", "bgpt_cuda_weight_bytes", "bgpt_cuda_launch_count",
    "bgpt_cuda_last_eval_ms", "bgpt_cuda_set_taps",
    "bgpt_cuda_op_mul_mat", "bgpt_cuda_op_mul_mat_tc", "bgpt_cuda_op_mul_mat_tcx", "bgpt_cuda_op_mul_mat_tcw", "bgpt_cuda_op_topk", "bgpt_cuda_op_quantize_act", "bgpt_cuda_op_norm",
    "bgpt_cuda_op_attention", "bgpt_cuda_op_gelu", "bgpt_cuda_op_dequantize",
]

_lib = None


class BgptError(RuntimeError):
    pass


def lib():
    """Load libbgpt_cuda.so (built by `make -C biogpt.cpp_b200`); fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BgptError(f"{LIB_PATH} is missing: build it with `make -C biogpt.cpp_b200` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.bgpt_cuda_last_error.restype = C.c_char_p
    L.bgpt_cuda_version.restype = C.c_char_p
    L.bgpt_cuda_device_count.restype = C.c_int
    L.bgpt_cuda_model_create.restype = C.c_void_p
    L.bgpt_cuda_model_create.argtypes = [_i32p, C.c_int, C.c_int]
    L.bgpt_cuda_upload_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_size_t]
    L.bgpt_cuda_set_tables.argtypes = [C.c_void_p, _u16p, _u16p]
    L.bgpt_host_build_tables.restype = None
    L.bgpt_host_build_tables.argtypes = [_u16p, _u16p]
    L.bgpt_cuda_model_finalize.argtypes = [C.c_void_p]
    L.bgpt_cuda_model_free.restype = None
    L.bgpt_cuda_model_free.argtypes = [C.c_void_p]
    L.bgpt_cuda_eval.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, _f32p]
    L.bgpt_cuda_eval_topk.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, _f32p, _i32p,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    L.bgpt_cuda_eval_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.bgpt_cuda_logits_device.restype = C.c_void_p
    L.bgpt_cuda_logits_device.argtypes = [C.c_void_p]
    L.bgpt_cuda_synchronize.argtypes = [C.c_void_p]
    L.bgpt_cuda_decode_greedy.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_int, _i32p,
                                          C.POINTER(C.c_float)]
    L.bgpt_cuda_set_chain.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_set_decode_path.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_get_decode_path.argtypes = [C.c_void_p]
    L.bgpt_cuda_decode_kernel_generation.argtypes = [C.c_void_p]
    L.bgpt_cuda_set_batch_path.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_get_batch_path.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_debug_read_buffer.restype = C.c_longlong
    L.bgpt_cuda_debug_read_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    L.bgpt_cuda_debug_read_prof.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.bgpt_cuda_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.bgpt_cuda_op_quantize_weights.argtypes = [C.c_int, _f32p, C.c_longlong, _u8p]
    L.bgpt_cuda_get_eval_path.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_debug_read_rows_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.bgpt_cuda_set_tc_min_rows.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_set_tcx_min_rows.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_set_tcw.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_set_f16_tc_min_rows.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_set_streams.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_cuda_eval_streams.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_void_p]
    L.bgpt_cuda_eval_streams_topk.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bgpt_cuda_decode_greedy_streams.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, _i32p, C.POINTER(C.c_float)]
    L.bgpt_cuda_hparams.restype = None
    L.bgpt_cuda_hparams.argtypes = [C.c_void_p, _i32p]
    L.bgpt_cuda_weight_bytes.restype = C.c_size_t
    L.bgpt_cuda_weight_bytes.argtypes = [C.c_void_p]
    L.bgpt_cuda_launch_count.restype = C.c_uint64
    L.bgpt_cuda_launch_count.argtypes = [C.c_void_p]
    L.bgpt_cuda_last_eval_ms.restype = C.c_float
    L.bgpt_cuda_last_eval_ms.argtypes = [C.c_void_p]
    L.bgpt_cuda_set_taps.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.bgpt_cuda_op_mul_mat.argtypes = [C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    L.bgpt_cuda_op_mul_mat_tc.argtypes = [C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    L.bgpt_cuda_op_mul_mat_tcx.argtypes = [C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    L.bgpt_cuda_op_mul_mat_tcw.argtypes = [C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
    L.bgpt_cuda_op_topk.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _i32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.bgpt_cuda_op_quantize_act.argtypes = [C.c_int, _f32p, _u8p, C.c_int]
    L.bgpt_cuda_op_norm.argtypes = [_f32p, C.c_void_p, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float]
    L.bgpt_cuda_op_attention.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _u16p]
    L.bgpt_cuda_op_gelu.argtypes = [_f32p, _f32p, C.c_int, _u16p]
    L.bgpt_cuda_op_dequantize.argtypes = [C.c_int, _u8p, _f32p, C.c_int, C.c_int]
    _lib = L
    return L


_tools = None


def tools_lib():
    """libbgpt_cuda_tools.so (`make -C biogpt.cpp_b200/csrc tools`): the design micro-benchmarks of
    include/bgpt_cuda_tools.h.  Not part of the product library; only tools/*_bench.py load it."""
    global _tools
    if _tools is None:
        p = os.path.join(HERE, "csrc", "libbgpt_cuda_tools.so")
        if not os.path.exists(p):
            raise BgptError(f"{p} is missing: build it with `make -C biogpt.cpp_b200/csrc tools`")
        T = C.CDLL(p)
        T.bgpt_cuda_last_error.restype = C.c_char_p
        T.bgpt_cuda_debug_quantize_bench.argtypes = [C.c_int, C.c_longlong, C.c_int, C.POINTER(C.c_float)]
        T.bgpt_cuda_debug_icache_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        T.bgpt_cuda_debug_barrier_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        T.bgpt_cuda_debug_gemm_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        _tools = T
    return _tools


def _check(rc: int, what: str):
    if rc != 0:
        raise BgptError(f"{what} failed ({rc}): {lib().bgpt_cuda_last_error().decode()}")


def last_error() -> str:
    return lib().bgpt_cuda_last_error().decode()


def device_count() -> int:
    return int(lib().bgpt_cuda_device_count())


def build_tables():
    """(gelu_f16, exp_f16) built with the host libm by the library's own helper."""
    g = np.zeros(65536, dtype=np.uint16)
    e = np.zeros(65536, dtype=np.uint16)
    lib().bgpt_host_build_tables(g, e)
    return g, e


TAP_NAMES = ("embed", "layer0_q", "layer0_att", "layer0_out", "final_in")


class Model:
    """A `.bin` model resident on one GPU (mirrors biogpt_model_load / biogpt_eval)."""

    def __init__(self, handle, hparams):
        self.h = handle
        self.hparams = hparams
        self.n_vocab = int(hparams[0])
        self.n_positions = int(hparams[3])
        self.d_model = int(hparams[5])

    @classmethod
    def load(cls, path: str, device: int = 0, max_batch: int = 8) -> "Model":
        L = lib()
        mf = gf.read_model(path)
        hp = np.array([mf.hparams.n_vocab, mf.hparams.n_layer, mf.hparams.n_head, mf.hparams.n_positions,
                       mf.hparams.d_ff, mf.hparams.d_model, mf.hparams.ftype], dtype=np.int32)
        h = L.bgpt_cuda_model_create(hp, device, max_batch)
        if not h:
            raise BgptError(f"bgpt_cuda_model_create: {last_error()}")
        m = cls(h, hp)
        try:
            with open(path, "rb") as f:
                for name, e in mf.tensors.items():
                    f.seek(e.offset)
                    raw = f.read(e.nbytes)
                    ne0 = e.ne[0]
                    ne1 = e.ne[1] if len(e.ne) > 1 else 1
                    buf = C.create_string_buffer(raw, len(raw))
                    _check(L.bgpt_cuda_upload_tensor(h, name.encode(), e.ggml_type, ne0, ne1,
                                                     C.cast(buf, C.c_void_p), len(raw)),
                           f"upload_tensor({name})")
            g, ex = build_tables()
            _check(L.bgpt_cuda_set_tables(h, g, ex), "set_tables")
            _check(L.bgpt_cuda_model_finalize(h), "model_finalize")
        except Exception:
            m.close()
            raise
        return m

    def eval(self, tokens: Sequence[int], n_past: int, taps: bool = False):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.n_vocab, dtype=np.float32)
        bufs = {}
        if taps:
            arr = (C.c_void_p * 5)()
            for i, nme in enumerate(TAP_NAMES):
                bufs[nme] = np.zeros((len(t), self.d_model), dtype=np.float32)
                arr[i] = bufs[nme].ctypes.data
            _check(lib().bgpt_cuda_set_taps(self.h, arr), "set_taps")
        _check(lib().bgpt_cuda_eval(self.h, t, len(t), n_past, out), "eval")
        return (out, bufs) if taps else out

    def eval_topk(self, tokens: Sequence[int], n_past: int, k: int, fallback: bool = True):
        """(vals[K], ids[K], exact, full_logits_or_None): the K largest logits of the eval's last row, selected on the device"""
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        st = getattr(self, "_topk_state", None)
        if st is None or st[0] != k:                    # the output buffers are reused from call to call (the returned views are overwritten by the next call)
            st = self._topk_state = (k, np.zeros(k, dtype=np.float32), np.zeros(k, dtype=np.int32), C.c_int(0), C.c_int(0),
                                     np.empty(self.n_vocab, dtype=np.float32))
        _, vals, ids, n_out, exact, full = st
        rc = lib().bgpt_cuda_eval_topk(self.h, t, len(t), n_past, k, vals, ids, C.byref(n_out), C.byref(exact),
                                       full.ctypes.data if fallback else None)
        if rc != 0:
            _check(rc, "eval_topk")
        return vals[:n_out.value], ids[:n_out.value], bool(exact.value), (full if fallback and not exact.value else None)

    def eval_streams(self, tokens: Sequence[int], n_past: int, fetch: bool = True) -> Optional[np.ndarray]:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty((len(t), self.n_vocab), dtype=np.float32) if fetch else None
        _check(lib().bgpt_cuda_eval_streams(self.h, t, len(t), n_past,
                                            out.ctypes.data if fetch else None), "eval_streams")
        return out

    def eval_streams_topk(self, tokens: Sequence[int], n_past: int, k: int):
        """(vals[S, k], ids[S, k], n_out[S], exact[S], full[S, n_vocab]): per stream the k largest logits, selected on the device; full[r] is
        filled only for rows with exact[r] == 0"""
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        S = len(t)
        st = getattr(self, "_stopk_state", None)
        if st is None or st[0] != (S, k):               # the output buffers are reused from call to call (overwritten by the next call)
            st = self._stopk_state = ((S, k), np.zeros((S, k), np.float32), np.zeros((S, k), np.int32), np.zeros(S, np.int32), np.zeros(S, np.int32),
                                      np.zeros((S, self.n_vocab), np.float32))
        _, vals, ids, n_out, exact, full = st
        _check(lib().bgpt_cuda_eval_streams_topk(self.h, t, S, n_past, k, vals.ctypes.data, ids.ctypes.data, n_out.ctypes.data, exact.ctypes.data,
                                                 full.ctypes.data), "eval_streams_topk")
        return vals, ids, n_out, exact, full

    def set_chain(self, on: int):
        """chained launches of eval_topk (the next position's kernel is queued while this call waits): 1 / 0, -1 = BGPT_CHAIN default (on)"""
        _check(lib().bgpt_cuda_set_chain(self.h, on), "set_chain")

    def set_decode_path(self, path: int):
        """1: persistent kernel per token, newest generation (default); 3: generation 4; 2: generation 3; 0: one kernel per fused operator"""
        _check(lib().bgpt_cuda_set_decode_path(self.h, path), "set_decode_path")

    def set_batch_path(self, path: int):
        """1: fused skinny-batch schedule (default); 2: persistent multi-row kernel for 2..8 rows, skinny beyond (opt-in, slower); 0: per-operator kernels"""
        _check(lib().bgpt_cuda_set_batch_path(self.h, path), "set_batch_path")

    def read_buffer(self, which: int, rows: int) -> np.ndarray:
        """debug: raw bytes of an arena buffer after the last eval (0 x, 1 x1, 2 q, 3 act_d records, 4 act_ff records)"""
        buf = np.empty(rows * 16 * 1024 * 4, dtype=np.uint8)
        n = lib().bgpt_cuda_debug_read_buffer(self.h, which, rows, buf.ctypes.data, buf.size)
        if n < 0:
            raise BgptError("read_buffer: " + last_error())
        return buf[:n].copy()

    def batch_path(self, n_rows: int) -> int:
        return int(lib().bgpt_cuda_get_batch_path(self.h, n_rows))

    def set_tcx_min_rows(self, rows: int):
        """bit-exact tcgen05 matmul for quantised evals of `rows`+ rows (default 128; 0 = off)"""
        _check(lib().bgpt_cuda_set_tcx_min_rows(self.h, rows), "set_tcx_min_rows")

    def set_tcw(self, on: int):
        """1 (default): the bit-exact tcgen05 matmul runs as the warp-specialised TMA-fed kernel (k_tcw_exact); 0: k_gemm_tc_xf"""
        _check(lib().bgpt_cuda_set_tcw(self.h, on), "set_tcw")

    def set_f16_tc_min_rows(self, rows: int):
        """F16 weights, opt-in: K-accumulating tcgen05 matmul (tolerance-close) for evals of `rows`+ rows (default 0 = off, all exact)"""
        _check(lib().bgpt_cuda_set_f16_tc_min_rows(self.h, rows), "set_f16_tc_min_rows")

    def set_tc_min_rows(self, rows: int):
        """opt in to the tolerance-close integer tcgen05 matmul for quantised evals of `rows`+ rows (0 = off, the default)"""
        _check(lib().bgpt_cuda_set_tc_min_rows(self.h, rows), "set_tc_min_rows")

    def eval_path(self, n_rows: int) -> int:
        """3 persistent decode kernel, 5 persistent multi-row kernel, 1 fused skinny-batch schedule, 0 per-operator exact SIMT,
        4 per-operator with the bit-exact tcgen05 matmul, 2 per-operator with the tolerance-close tcgen05 matmul (opt-in)"""
        return int(lib().bgpt_cuda_get_eval_path(self.h, n_rows))

    @property
    def decode_path(self) -> int:
        return int(lib().bgpt_cuda_get_decode_path(self.h))

    @property
    def decode_generation(self) -> int:
        """5 / 4 / 3: persistent-kernel generation used for single-token steps; 0: per-operator kernels"""
        return int(lib().bgpt_cuda_decode_kernel_generation(self.h))

    def read_prof(self):
        """[n_layer+1][5][3] clock64 stamps of CTA 0 (needs BGPT_MEGA_PROF=1 before load)"""
        buf = np.zeros(4096, dtype=np.int64)
        n = lib().bgpt_cuda_debug_read_prof(self.h, buf.ctypes.data, buf.size)
        return buf[:n].reshape(-1, 5, 6) if n else None

    def read_trace(self):
        """generation-4 kernel: (stamps [n_cta][n_layer+1][5][12] clock64, calib [n_cta][4]) or None"""
        buf = np.zeros(1 << 20, dtype=np.int64)
        nc, per = C.c_int(0), C.c_int(0)
        n = lib().bgpt_cuda_debug_read_trace(self.h, buf.ctypes.data, buf.size, C.byref(nc), C.byref(per))
        if not n:
            return None
        nc, per = nc.value, per.value
        return buf[:nc * per].reshape(nc, -1, 5, 12), buf[nc * per:nc * per + 4 * nc].reshape(nc, 4)

    def read_rows_trace(self):
        """multi-row kernel: [n_layer + 1][8] clock64 stamps of CTA 0 (needs BGPT_MEGA_PROF=1 before load) or None"""
        buf = np.zeros(4096, dtype=np.int64)
        n = lib().bgpt_cuda_debug_read_rows_trace(self.h, buf.ctypes.data, buf.size)
        return buf[:n].reshape(-1, 8) if n else None

    def set_streams(self, n: int):
        _check(lib().bgpt_cuda_set_streams(self.h, n), "set_streams")

    def decode_greedy_streams(self, first_tokens: Sequence[int], n_past: int, n_steps: int):
        """(ids [n_steps][n_streams], device ms): lock-step greedy decode of the streams, entirely on the device"""
        t = np.ascontiguousarray(first_tokens, dtype=np.int32)
        ids = np.zeros((n_steps, len(t)), dtype=np.int32)
        ms = C.c_float(0)
        _check(lib().bgpt_cuda_decode_greedy_streams(self.h, t, len(t), n_past, n_steps, ids.reshape(-1), C.byref(ms)),
               "decode_greedy_streams")
        return ids, float(ms.value)

    def decode_greedy(self, first_token: int, n_past: int, n_steps: int):
        ids = np.zeros(n_steps, dtype=np.int32)
        ms = C.c_float(0)
        _check(lib().bgpt_cuda_decode_greedy(self.h, first_token, n_past, n_steps, ids, C.byref(ms)),
               "decode_greedy")
        return ids, float(ms.value)

    @property
    def last_eval_ms(self) -> float:
        return float(lib().bgpt_cuda_last_eval_ms(self.h))

    @property
    def launch_count(self) -> int:
        return int(lib().bgpt_cuda_launch_count(self.h))

    @property
    def weight_bytes(self) -> int:
        return int(lib().bgpt_cuda_weight_bytes(self.h))

    def close(self):
        if self.h:
            lib().bgpt_cuda_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- unit-level operators ------------------------------------------------------------------

def op_mul_mat(ggml_type: int, w_bytes: np.ndarray, x: np.ndarray, rows: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, k = x.shape
    y = np.empty((n, rows), dtype=np.float32)
    _check(lib().bgpt_cuda_op_mul_mat(ggml_type, np.ascontiguousarray(w_bytes, dtype=np.uint8), x, y, k, rows, n),
           "op_mul_mat")
    return y


def op_mul_mat_tcx(ggml_type: int, w_bytes: np.ndarray, x: np.ndarray, rows: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, k = x.shape
    y = np.empty((n, rows), dtype=np.float32)
    _check(lib().bgpt_cuda_op_mul_mat_tcx(ggml_type, np.ascontiguousarray(w_bytes, dtype=np.uint8), x, y, k, rows, n),
           "op_mul_mat_tcx")
    return y


def op_topk(logits: np.ndarray, k: int):
    """(vals, ids, exact): the device top-k selection on a host logit row"""
    x = np.ascontiguousarray(logits, dtype=np.float32)
    vals = np.zeros(k, dtype=np.float32); ids = np.zeros(k, dtype=np.int32)
    n_out, exact = C.c_int(0), C.c_int(0)
    _check(lib().bgpt_cuda_op_topk(x, len(x), k, vals, ids, C.byref(n_out), C.byref(exact)), "op_topk")
    return vals[:n_out.value], ids[:n_out.value], bool(exact.value)


def op_mul_mat_tcw(ggml_type: int, w_bytes: np.ndarray, x: np.ndarray, rows: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, k = x.shape
    y = np.empty((n, rows), dtype=np.float32)
    _check(lib().bgpt_cuda_op_mul_mat_tcw(ggml_type, np.ascontiguousarray(w_bytes, dtype=np.uint8), x, y, k, rows, n),
           "op_mul_mat_tcw")
    return y


def op_mul_mat_tc(ggml_type: int, w_bytes: np.ndarray, x: np.ndarray, rows: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, k = x.shape
    y = np.empty((n, rows), dtype=np.float32)
    _check(lib().bgpt_cuda_op_mul_mat_tc(ggml_type, np.ascontiguousarray(w_bytes, dtype=np.uint8), x, y, k, rows, n),
           "op_mul_mat_tc")
    return y


def op_quantize_weights(ggml_type: int, x: np.ndarray) -> np.ndarray:
    """f32 -> Qx blocks (file layout) on the device: the reference's quantize_row_q*_reference bits"""
    from . import ggml_file as gf
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    out = np.empty(gf.row_bytes(ggml_type, x.size), dtype=np.uint8)
    _check(lib().bgpt_cuda_op_quantize_weights(ggml_type, x, x.size, out), "op_quantize_weights")
    return out


def quantize_bench(ggml_type: int, n: int, iters: int = 10) -> float:
    """microseconds per launch of the device quantiser over n weights (device-resident)"""
    us = C.c_float(0)
    if tools_lib().bgpt_cuda_debug_quantize_bench(ggml_type, n, iters, C.byref(us)) != 0:
        raise BgptError("quantize_bench: " + tools_lib().bgpt_cuda_last_error().decode())
    return float(us.value)


def op_quantize_act(weight_type: int, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    k = x.size
    kind_bytes = {gf.GGML_TYPE_F32: 4 * k, gf.GGML_TYPE_F16: 2 * k,
                  gf.GGML_TYPE_Q4_0: k // 32 * 34, gf.GGML_TYPE_Q5_0: k // 32 * 34, gf.GGML_TYPE_Q8_0: k // 32 * 34,
                  gf.GGML_TYPE_Q4_1: k // 32 * 40, gf.GGML_TYPE_Q5_1: k // 32 * 40}[weight_type]
    out = np.zeros(kind_bytes, dtype=np.uint8)
    _check(lib().bgpt_cuda_op_quantize_act(weight_type, x, out, k), "op_quantize_act")
    return out


def op_norm(x: np.ndarray, w=None, b=None, eps: float = 1e-5) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, nc = x.shape
    y = np.empty_like(x)
    wp = np.ascontiguousarray(w, dtype=np.float32) if w is not None else None
    bp = np.ascontiguousarray(b, dtype=np.float32) if b is not None else None
    _check(lib().bgpt_cuda_op_norm(x, wp.ctypes.data if wp is not None else None,
                                   bp.ctypes.data if bp is not None else None, y, rows, nc, eps), "op_norm")
    return y


def op_attention(q, k, v, n_past: int, n_head: int, exp_tab: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    n, d = q.shape
    out = np.empty_like(q)
    _check(lib().bgpt_cuda_op_attention(q, k, v, out, n, n_past, d, n_head,
                                        np.ascontiguousarray(exp_tab, dtype=np.uint16)), "op_attention")
    return out


def op_gelu(x: np.ndarray, gelu_tab: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    y = np.empty_like(x)
    _check(lib().bgpt_cuda_op_gelu(x, y, x.size, np.ascontiguousarray(gelu_tab, dtype=np.uint16)), "op_gelu")
    return y


def op_dequantize(ggml_type: int, w_bytes: np.ndarray, k: int, rows: int) -> np.ndarray:
    y = np.empty((rows, k), dtype=np.float32)
    _check(lib().bgpt_cuda_op_dequantize(ggml_type, np.ascontiguousarray(w_bytes, dtype=np.uint8), y, k, rows),
           "op_dequantize")
    return y
