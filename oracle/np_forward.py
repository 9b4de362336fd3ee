"""Second, independent restatement of the qwen3-rs forward pass, in numpy float32.

TEST INFRASTRUCTURE ONLY (same rule as q3_oracle.c).  Purpose: the reference ships no tests or
golden vectors for its forward path and cannot be compiled here (no Rust toolchain), so the C
oracle is pinned by agreement with this separately-written restatement instead
(SURVEY.md §8c "How the oracle is trusted").  Written from the reference sources, not from the
C file; deliberately structured differently (vectorised per group / per head, file parsed with
numpy views).  Left-fold f32 sums are reproduced exactly with np.cumsum (sequential).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def lfold_sum(a: np.ndarray) -> np.float32:
    """`.sum::<f32>()`: sequential left fold (np.add.reduce would be pairwise)."""
    a = np.asarray(a, f32).reshape(-1)
    if a.size == 0:
        return f32(0)
    return np.cumsum(a, dtype=f32)[-1]


def round_half_away(x: np.ndarray) -> np.ndarray:
    """f32::round (tensor.rs:116)."""
    return np.copysign(np.floor(np.abs(x) + f32(0.5)), x).astype(f32)


def quantize(x: np.ndarray, gs: int):
    """tensor.rs:91-119.  NB |x|+0.5 is exact for |x| <= 127 (the only range used)."""
    g = np.asarray(x, f32).reshape(-1, gs)
    wmax = np.abs(g).max(axis=1)
    scale = (wmax / f32(127.0)).astype(f32)
    safe = np.where(scale != 0, scale, f32(1))
    qv = np.where(scale[:, None] != 0, g / safe[:, None], f32(0)).astype(f32)
    # round half away from zero without the +0.5 double-rounding hazard
    r = np.trunc(qv)
    frac = np.abs(qv - r)
    r = np.where(frac >= f32(0.5), r + np.sign(qv), r)
    q = np.clip(r, -128, 127).astype(np.int8)
    return q.reshape(-1), scale


def matmul(xq, xs, wq, ws, n: int, d: int, gs: int) -> np.ndarray:
    """tensor.rs:23-62: int32 group dots, (dot*ws)*xs, left fold over groups."""
    ng = n // gs
    w = wq[: d * n].reshape(d, ng, gs).astype(np.int32)
    x = xq[:n].reshape(1, ng, gs).astype(np.int32)
    dots = (w * x).sum(axis=2, dtype=np.int32)  # [d, ng] exact
    terms = (dots.astype(f32) * ws[: d * ng].reshape(d, ng)).astype(f32) * xs[:ng].reshape(1, ng)
    return np.cumsum(terms.astype(f32), axis=1, dtype=f32)[:, -1]


def rmsnorm(x, w) -> np.ndarray:
    """layers.rs:109-119."""
    x = np.asarray(x, f32)
    ss = lfold_sum(x * x)
    f = f32(1.0) / np.sqrt(ss / f32(x.size) + f32(1e-6), dtype=f32)
    return (w * (f * x)).astype(f32)


def rope_freqs(pos: int, head_dim: int):
    """layers.rs:161-171 (numpy's powf/cos/sin may differ from glibc in the last ulp)."""
    half = head_dim // 2
    i = np.arange(half, dtype=f32)
    freq = np.power(f32(1e6), -i / f32(half), dtype=f32)
    ang = (f32(pos) * freq).astype(f32)
    return np.cos(ang, dtype=f32), np.sin(ang, dtype=f32)


def rope_apply(v, cos, sin):
    """layers.rs:173-185, half-split pairs."""
    half = v.size // 2
    x, y = v[:half].copy(), v[half:].copy()
    return np.concatenate([x * cos - y * sin, x * sin + y * cos]).astype(f32)


def softmax(x):
    """layers.rs:495-506."""
    x = np.asarray(x, f32)
    e = np.exp(x - x.max(), dtype=f32)
    return (e * (f32(1.0) / lfold_sum(e))).astype(f32)


class NpModel:
    """models/qwen3.rs load_weights (:199-277) + forward (:62-79, :131-176) + layers.rs."""

    def __init__(self, path: str, ctx_len: int | None = None):
        raw = np.fromfile(path, dtype=np.uint8)
        h = raw[:52].view("<i4")
        assert h[0] == 0x616A6331 and h[1] == 1, "bad magic/version (configuration.rs:116-125)"
        (self.arch, self.dim, self.hidden, self.L, self.n_heads, self.n_kv, self.vocab, self.seq_len,
         self.hd, shared, self.gs) = (int(v) for v in h[2:13])
        self.shared = shared != 0
        if ctx_len:
            self.seq_len = min(ctx_len, self.seq_len)
        self.AH, self.KV = self.n_heads * self.hd, self.n_kv * self.hd
        off = 256

        def take_f32(n):
            nonlocal off
            a = raw[off: off + 4 * n].view("<f4")
            off += 4 * n
            return a

        def take_q(count, size):
            nonlocal off
            out = []
            for _ in range(count):
                q = raw[off: off + size].view(np.int8)
                off += size
                out.append((q, take_f32(size // self.gs)))
            return out

        L, dim = self.L, self.dim
        self.rms_att = take_f32(L * dim).reshape(L, dim)
        self.rms_ffn = take_f32(L * dim).reshape(L, dim)
        self.rms_final = take_f32(dim)
        self.q_ln = take_f32(L * self.hd).reshape(L, self.hd)
        self.k_ln = take_f32(L * self.hd).reshape(L, self.hd)
        self.embed = take_q(1, self.vocab * dim)[0]
        self.wq = take_q(L, dim * self.AH)
        self.wk = take_q(L, dim * self.KV)
        self.wv = take_q(L, dim * self.KV)
        self.wo = take_q(L, self.AH * dim)
        self.w1 = take_q(L, dim * self.hidden)
        self.w2 = take_q(L, self.hidden * dim)
        self.w3 = take_q(L, dim * self.hidden)
        self.wcls = self.embed if self.shared else take_q(1, dim * self.vocab)[0]
        assert off == raw.size, f"trailing bytes: {raw.size - off}"
        self.kc = np.zeros((L, self.seq_len, self.KV), f32)
        self.vc = np.zeros((L, self.seq_len, self.KV), f32)
        self.trace = {}

    def forward(self, token: int, pos: int) -> np.ndarray:
        gs, dim, hd = self.gs, self.dim, self.hd
        eq, es = self.embed
        row = eq[token * dim:(token + 1) * dim].astype(f32).reshape(-1, gs)
        x = (row * es[token * dim // gs:(token + 1) * dim // gs, None]).astype(f32).reshape(-1)  # tensor.rs:72-80
        cos, sin = rope_freqs(pos, hd)
        kv_mul = self.n_heads // self.n_kv
        for l in range(self.L):
            xq, xs = quantize(rmsnorm(x, self.rms_att[l]), gs)
            T = {"xq_attn_q": xq.copy(), "xq_attn_s": xs.copy()}
            self.trace[l] = T
            q = matmul(xq, xs, *self.wq[l], dim, self.AH, gs)
            k = matmul(xq, xs, *self.wk[l], dim, self.KV, gs)
            v = matmul(xq, xs, *self.wv[l], dim, self.KV, gs)
            q = np.concatenate([rope_apply(rmsnorm(q[h * hd:(h + 1) * hd], self.q_ln[l]), cos, sin)
                                for h in range(self.n_heads)])
            k = np.concatenate([rope_apply(rmsnorm(k[h * hd:(h + 1) * hd], self.k_ln[l]), cos, sin)
                                for h in range(self.n_kv)])
            self.kc[l, pos], self.vc[l, pos] = k, v
            T.update(q_post=q.copy(), k_post=k.copy(), v_row=v.copy())
            out = np.zeros(self.AH, f32)
            scale = f32(1.0) / np.sqrt(f32(hd))
            for h in range(self.n_heads):
                kvh = h // kv_mul
                K = self.kc[l, :pos + 1, kvh * hd:(kvh + 1) * hd]
                V = self.vc[l, :pos + 1, kvh * hd:(kvh + 1) * hd]
                prods = (K * q[None, h * hd:(h + 1) * hd]).astype(f32)
                scores = (np.cumsum(prods, axis=1, dtype=f32)[:, -1] * scale).astype(f32)
                a = softmax(scores)
                out[h * hd:(h + 1) * hd] = np.cumsum((a[:, None] * V).astype(f32), axis=0, dtype=f32)[-1]
            xq, xs = quantize(out, gs)
            x = (x + matmul(xq, xs, *self.wo[l], self.AH, dim, gs)).astype(f32)
            T.update(att_out=out.copy(), x_after_attn=x.copy())
            xq, xs = quantize(rmsnorm(x, self.rms_ffn[l]), gs)
            g = matmul(xq, xs, *self.w1[l], dim, self.hidden, gs)
            u = matmul(xq, xs, *self.w3[l], dim, self.hidden, gs)
            hb = ((g * (f32(1.0) / (f32(1.0) + np.exp(-g, dtype=f32)))).astype(f32) * u).astype(f32)
            hq, hs = quantize(hb, gs)
            x = (x + matmul(hq, hs, *self.w2[l], self.hidden, dim, gs)).astype(f32)
            T.update(hb_swiglu=hb.copy(), hq_q=hq.copy(), hq_s=hs.copy(), x_out=x.copy())
        x = rmsnorm(x, self.rms_final)
        xq, xs = quantize(x, gs)
        return matmul(xq, xs, *self.wcls, dim, self.vocab, gs)
