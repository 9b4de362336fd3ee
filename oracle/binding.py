"""ctypes binding to the CPU oracle (oracle/q3_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (qwen3_rs_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libq3oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "q3_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


class OrcConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "architecture_id", "dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "head_dim", "seq_len",
        "vocab_size", "group_size", "shared_classifier")]


class OrcTrace(C.Structure):
    _fields_ = [
        ("xq_attn_q", C.c_void_p), ("xq_attn_s", C.c_void_p), ("q_post", C.c_void_p), ("k_post", C.c_void_p),
        ("v_row", C.c_void_p), ("att_out", C.c_void_p), ("x_after_attn", C.c_void_p), ("hb_swiglu", C.c_void_p),
        ("hq_q", C.c_void_p), ("hq_s", C.c_void_p), ("x_out", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, f32, u64, sz = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_size_t
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_set_threads": (None, [i32]),
            "orc_get_max_threads": (i32, []),
            "orc_set_perturb": (None, [i32]),
            "orc_quantize": (None, [vp, vp, vp, i32, i32]),
            "orc_dequantize": (None, [vp, vp, vp, sz, i32]),
            "orc_matmul": (None, [vp, vp, vp, vp, vp, i32, i32, i32]),
            "orc_group_dots": (None, [vp, vp, vp, i32, i32, i32]),
            "orc_rmsnorm": (None, [vp, vp, vp, i32]),
            "orc_rope_freqs": (None, [vp, i32, i32]),
            "orc_rope_apply": (None, [vp, vp, i32]),
            "orc_softmax": (None, [vp, i32]),
            "orc_expf_array": (None, [vp, vp, sz]),
            "orc_argmax": (i32, [vp, i32]),
            "orc_sampler_new": (vp, [i32, f32, f32, u64]),
            "orc_sampler_free": (None, [vp]),
            "orc_sampler_random_u32": (C.c_uint32, [vp]),
            "orc_sampler_random_f32": (f32, [vp]),
            "orc_sampler_sample": (i32, [vp, vp]),
            "orc_round_half_to_even": (f32, [f32]),
            "orc_find_optimal_group_size": (i32, [i32, i32]),
            "orc_quantize_q80": (i32, [vp, vp, vp, vp, sz, i32]),
            "orc_model_open": (vp, [C.c_char_p, i32]),
            "orc_model_free": (None, [vp]),
            "orc_model_config": (C.POINTER(OrcConfig), [vp]),
            "orc_model_file_bytes": (sz, [vp]),
            "orc_model_forward": (C.POINTER(C.c_float), [vp, i32, i32]),
            "orc_model_reset": (None, [vp]),
            "orc_model_forward_layers": (C.POINTER(C.c_float), [vp, i32, i32, i32, vp, i32]),
            "orc_model_key_cache": (C.POINTER(C.c_float), [vp]),
            "orc_model_value_cache": (C.POINTER(C.c_float), [vp]),
            "orc_model_set_trace": (None, [vp, i32, C.POINTER(OrcTrace)]),
            "orc_model_set_xdump": (None, [vp, vp]),
            "orc_generate": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def set_perturb(mode: int) -> None:
    """Sensitivity knob: 1 = re-associated float sums (see q3_oracle.c)."""
    lib().orc_set_perturb(mode)


def set_threads(n: int) -> None:
    lib().orc_set_threads(n)


def max_threads() -> int:
    return lib().orc_get_max_threads()


def last_error() -> str:
    return lib().orc_last_error().decode()


# ---- tensor.rs -------------------------------------------------------------------------------
def quantize(x: np.ndarray, gs: int):
    x = np.ascontiguousarray(x, np.float32)
    q = np.empty(x.size, np.int8)
    s = np.empty(x.size // gs, np.float32)
    lib().orc_quantize(_p(q), _p(s), _p(x), x.size, gs)
    return q, s


def dequantize(q: np.ndarray, s: np.ndarray, gs: int):
    x = np.empty(q.size, np.float32)
    lib().orc_dequantize(_p(x), _p(q), _p(s), q.size, gs)
    return x


def matmul(xq, xs, wq, ws, n: int, d: int, gs: int):
    out = np.empty(d, np.float32)
    lib().orc_matmul(_p(out), _p(xq), _p(xs), _p(wq), _p(ws), n, d, gs)
    return out


def group_dots(xq, wq, n: int, d: int, gs: int):
    out = np.empty((d, n // gs), np.int32)
    lib().orc_group_dots(_p(out), _p(xq), _p(wq), n, d, gs)
    return out


# ---- layers.rs -------------------------------------------------------------------------------
def rmsnorm(x, w):
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    out = np.empty_like(x)
    lib().orc_rmsnorm(_p(out), _p(x), _p(w), x.size)
    return out


def rope_freqs(pos: int, head_dim: int):
    cs = np.empty((head_dim // 2, 2), np.float32)
    lib().orc_rope_freqs(_p(cs), pos, head_dim)
    return cs


def rope_apply(v, cs):
    v = np.array(v, np.float32)
    lib().orc_rope_apply(_p(v), _p(np.ascontiguousarray(cs, np.float32)), v.size)
    return v


def softmax(x):
    x = np.array(x, np.float32)
    lib().orc_softmax(_p(x), x.size)
    return x


def expf(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().orc_expf_array(_p(out), _p(x), x.size)
    return out


def argmax(logits) -> int:
    a = np.ascontiguousarray(logits, np.float32)
    return lib().orc_argmax(_p(a), a.size)


# ---- exporter --------------------------------------------------------------------------------
def round_half_to_even(x: float) -> float:
    return lib().orc_round_half_to_even(x)


def find_optimal_group_size(hidden_dim: int, requested: int) -> int:
    return lib().orc_find_optimal_group_size(hidden_dim, requested)


def quantize_q80(w, gs: int):
    w = np.ascontiguousarray(w, np.float32).reshape(-1)
    q = np.empty(w.size, np.int8)
    s = np.empty(max(w.size // gs, 1), np.float32)
    err = C.c_float(0)
    rc = lib().orc_quantize_q80(_p(q), _p(s), C.addressof(err), _p(w), w.size, gs)
    if rc != 0:
        raise ValueError(last_error())
    return q, s[: w.size // gs], err.value


# ---- sampler.rs ------------------------------------------------------------------------------
class Sampler:
    def __init__(self, vocab: int, temperature: float, topp: float, seed: int):
        assert vocab > 0 and temperature >= 0.0 and 0.0 <= topp <= 1.0
        self._h = lib().orc_sampler_new(vocab, temperature, topp, seed)
        self.vocab = vocab

    def random_u32(self) -> int:
        return lib().orc_sampler_random_u32(self._h)

    def random_f32(self) -> float:
        return lib().orc_sampler_random_f32(self._h)

    def sample(self, logits: np.ndarray) -> int:
        a = np.array(logits, np.float32)
        return lib().orc_sampler_sample(self._h, _p(a))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_sampler_free(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass


# ---- model -----------------------------------------------------------------------------------
class Model:
    """Restated Qwen3Transformer (models/qwen3.rs) behind TransformerBuilder semantics."""

    def __init__(self, path: str, ctx_len: Optional[int] = None):
        self._h = lib().orc_model_open(path.encode(), int(ctx_len or 0))
        if not self._h:
            raise RuntimeError(last_error())
        c = lib().orc_model_config(self._h).contents
        self.config = {n: getattr(c, n) for n, _ in OrcConfig._fields_}
        self._keep = []

    def forward(self, token: int, pos: int) -> np.ndarray:
        p = lib().orc_model_forward(self._h, token, pos)
        if not p:
            raise IndexError(last_error())
        return np.ctypeslib.as_array(p, shape=(self.config["vocab_size"],)).copy()

    def reset(self):
        lib().orc_model_reset(self._h)

    def forward_layers(self, x, pos: int, l0: int, l1: int, run_head: bool = False):
        x = np.array(x, np.float32)
        p = lib().orc_model_forward_layers(self._h, pos, l0, l1, _p(x), int(run_head))
        if run_head:
            return x, np.ctypeslib.as_array(p, shape=(self.config["vocab_size"],)).copy()
        return x

    def kv_cache(self):
        c = self.config
        shape = (c["n_layers"], c["seq_len"], c["n_kv_heads"], c["head_dim"])
        k = np.ctypeslib.as_array(lib().orc_model_key_cache(self._h), shape=shape)
        v = np.ctypeslib.as_array(lib().orc_model_value_cache(self._h), shape=shape)
        return k, v

    def dump_residuals(self) -> np.ndarray:
        """Subsequent forward() calls record x entering layer 0 and leaving each layer: [(L+1), dim]."""
        buf = np.zeros((self.config["n_layers"] + 1, self.config["dim"]), np.float32)
        self._keep.append(buf)
        lib().orc_model_set_xdump(self._h, _p(buf))
        return buf

    def trace_layer(self, layer: int) -> dict:
        """Arrange for the next forward() calls to record layer `layer`'s intermediates."""
        c = self.config
        dim, gs, H = c["dim"], c["group_size"], c["hidden_dim"]
        AH, KV = c["n_heads"] * c["head_dim"], c["n_kv_heads"] * c["head_dim"]
        bufs = {
            "xq_attn_q": np.zeros(dim, np.int8), "xq_attn_s": np.zeros(dim // gs, np.float32),
            "q_post": np.zeros(AH, np.float32), "k_post": np.zeros(KV, np.float32),
            "v_row": np.zeros(KV, np.float32), "att_out": np.zeros(AH, np.float32),
            "x_after_attn": np.zeros(dim, np.float32), "hb_swiglu": np.zeros(H, np.float32),
            "hq_q": np.zeros(H, np.int8), "hq_s": np.zeros(H // gs, np.float32),
            "x_out": np.zeros(dim, np.float32),
        }
        t = OrcTrace(**{k: v.ctypes.data for k, v in bufs.items()})
        self._keep.append((bufs, t))
        lib().orc_model_set_trace(self._h, layer, C.byref(t))
        return bufs

    def generate(self, prompt: Sequence[int], max_new: int, temperature: float = 0.0, topp: float = 0.9,
                 seed: int = 0, bos: int = -1, eos: int = -1, with_margins: bool = False):
        s = Sampler(self.config["vocab_size"], temperature, topp, seed)
        pr = np.ascontiguousarray(prompt, np.int32)
        out = np.zeros(max(max_new, 1), np.int32)
        mg = np.zeros(max(max_new, 1), np.float32)
        n = lib().orc_generate(self._h, s._h, _p(pr), pr.size, max_new, bos, eos, _p(out), _p(mg))
        if n < 0:
            raise RuntimeError(last_error())
        return (out[:n].tolist(), mg[:n]) if with_margins else out[:n].tolist()

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass
