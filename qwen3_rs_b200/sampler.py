"""Host-side sampler with the reference's exact semantics (qwen3-inference/src/sampler.rs).

Greedy argmax returns the LAST index among equal maxima (`Iterator::max_by(total_cmp)`, :57-59);
temperature -> softmax -> xorshift64* coin -> multinomial or top-p (:116-136)."""
from __future__ import annotations

import numpy as np

_M64 = 0xFFFFFFFFFFFFFFFF


def _total_key(a: np.ndarray) -> np.ndarray:
    """f32::total_cmp ordering key."""
    b = np.ascontiguousarray(a, np.float32).view(np.int32)
    return b ^ ((b >> 31).view(np.uint32) >> 1).view(np.int32)


def argmax_last(logits: np.ndarray) -> int:
    a = np.ascontiguousarray(logits, np.float32)
    i = int(np.argmax(a))  # first maximum; the first NaN if there is one
    m = a[i]
    if m == m and m != 0.0 and np.count_nonzero(a == m) == 1:
        return i  # a unique, non-zero, non-NaN maximum: total_cmp and == agree and first == last
    k = _total_key(a)
    return int(k.size - 1 - np.argmax(k[::-1]))


def softmax(x: np.ndarray) -> np.ndarray:
    """layers.rs:495-506 (left-fold sum)."""
    x = np.asarray(x, np.float32)
    e = np.exp(x - x.max(), dtype=np.float32)
    total = np.cumsum(e, dtype=np.float32)[-1]
    return (e * (np.float32(1.0) / total)).astype(np.float32)


class Sampler:
    def __init__(self, vocab_size: int, temperature: float, topp: float, rng_seed: int):
        assert vocab_size > 0, "Vocab size must be positive"
        assert temperature >= 0.0, "Temperature must be non-negative"
        assert 0.0 <= topp <= 1.0, "Top-p must be between 0.0 and 1.0"
        self.vocab_size = vocab_size
        self.temperature = np.float32(temperature)
        self.topp = np.float32(min(max(topp, 0.0), 1.0))
        self.rng_state = rng_seed & _M64

    def random_u32(self) -> int:  # :44-49
        s = self.rng_state
        s ^= s >> 12
        s ^= (s << 25) & _M64
        s ^= s >> 27
        self.rng_state = s
        return ((s * 0x2545F4914F6CDD1D) & _M64) >> 32

    def random_f32(self) -> np.float32:  # :52-54
        return np.float32(self.random_u32() >> 8) / np.float32(16777216.0)

    def sample(self, logits: np.ndarray) -> int:  # :116-136
        if self.temperature == 0.0:
            return argmax_last(logits)
        p = softmax(np.asarray(logits, np.float32) / self.temperature)
        coin = self.random_f32()
        if self.topp <= 0.0 or self.topp >= 1.0:
            return self._sample_mult(p, coin)
        return self._sample_topp(p, coin)

    @staticmethod
    def _sample_mult(p: np.ndarray, coin) -> int:  # :62-71
        cdf = np.cumsum(p, dtype=np.float32)
        hit = np.nonzero(coin < cdf)[0]
        return int(hit[0]) if hit.size else max(p.size - 1, 0)

    def _sample_topp(self, p: np.ndarray, coin) -> int:  # :74-110
        cutoff = (np.float32(1.0) - self.topp) / np.float32(max(p.size - 1, 1))
        idx = np.nonzero(p >= cutoff)[0]
        order = np.argsort(-_total_key(p[idx]).astype(np.int64), kind="stable")
        idx = idx[order]
        probs = p[idx]
        cum = np.cumsum(probs, dtype=np.float32)
        over = np.nonzero(cum > self.topp)[0]
        last = int(over[0]) if over.size else max(idx.size - 1, 0)
        r = coin * cum[last]
        hit = np.nonzero(r < cum[: last + 1])[0]
        return int(idx[hit[0]]) if hit.size else int(idx[last])
