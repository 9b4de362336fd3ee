"""Host-side mirrors of the reference's callers of the hot path (sampler.rs, generation.rs), checked
against the oracle's restatement on CPU -- no GPU needed (a fake transformer replays oracle logits)."""
import numpy as np
import pytest

from oracle import binding as orc
from qwen3_rs_b200 import generation
from qwen3_rs_b200.sampler import Sampler, argmax_last, softmax


def test_argmax_last_matches_oracle_on_ties_and_signed_zeros():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.integers(-3, 4, 200).astype(np.float32)  # many ties
        assert argmax_last(a) == orc.argmax(a)
    for a in ([-0.0, 0.0], [0.0, -0.0], [np.float32("-inf")] * 3, [1.0]):
        a = np.array(a, np.float32)
        assert argmax_last(a) == orc.argmax(a)


def test_softmax_matches_oracle_bitwise():
    x = (np.random.default_rng(1).standard_normal(4096) * 4).astype(np.float32)
    a, b = softmax(x), orc.softmax(x)
    assert np.abs(a - b).max() <= 1e-9 + 2 * np.finfo(np.float32).eps * b.max()  # exp() ulp differences only


@pytest.mark.parametrize("temperature,topp", [(1.0, 0.9), (0.7, 1.0), (1.3, 0.5), (0.0, 0.9), (1.0, 0.0)])
def test_sampler_stream_matches_oracle(temperature, topp):
    """Same xorshift64* stream, same multinomial / top-p decisions (sampler.rs:44-136)."""
    rng = np.random.default_rng(2)
    V = 3000
    a, b = Sampler(V, temperature, topp, 1234), orc.Sampler(V, temperature, topp, 1234)
    agree = 0
    for i in range(60):
        logits = (rng.standard_normal(V) * 3).astype(np.float32)
        agree += a.sample(logits) == b.sample(logits)
    assert agree >= 58  # numpy vs glibc exp() can flip a cdf comparison once in a blue moon
    assert a.random_u32() == b.random_u32()  # RNG streams still in lockstep


class _Replay:
    """Transformer stand-in that runs the oracle underneath (host-logic test only)."""

    class _Cfg:
        def __init__(self, c):
            self.seq_len, self.vocab_size = c["seq_len"], c["vocab_size"]

    def __init__(self, path):
        self.o = orc.Model(path)
        self.calls = []

    def forward(self, token, pos):
        self.calls.append((token, pos))
        return self.o.forward(token, pos)

    def get_config(self):
        return self._Cfg(self.o.config)

    def decode_greedy(self, tok, pos0, n):
        out = []
        for i in range(n):
            tok = orc.argmax(self.o.forward(tok, pos0 + i))
            out.append(tok)
        return out


def test_generate_mirrors_reference_loop(ckpt):
    path = ckpt("tiny-untied", 64, 1)
    prompt = [9, 2, 200]
    want = orc.Model(path).generate(prompt, 12)
    t = _Replay(path)
    got = generation.generate(t, Sampler(t.get_config().vocab_size, 0.0, 0.9, 0), prompt, 12)
    assert got == want
    assert t.calls[0] == (200, 2)  # prompt tokens except the last are never forwarded (generation.rs:26-28)
    assert generation.generate_fast(_Replay(path), prompt, 12) == want
    with pytest.raises(ValueError, match="Please provide a prompt"):
        generation.generate(t, Sampler(8, 0.0, 0.9, 0), [], 4)


def test_generate_stops_on_eos_and_seq_len(ckpt):
    path = ckpt("tiny-untied", 64, 1)
    full = orc.Model(path).generate([9], 10)
    t = _Replay(path)
    k = next(i for i in range(1, len(full)) if full[i] not in full[:i])  # first occurrence of some later token
    got = generation.generate(t, Sampler(t.get_config().vocab_size, 0.0, 0.9, 0), [9], 10, eos_token_id=full[k])
    assert got == full[:k]  # the terminating token is not emitted (generation.rs:33-36)
    assert got == orc.Model(path).generate([9], 10, eos=full[k])
    m = orc.Model(path, 6)
    assert len(m.generate([9], 100)) == 6  # pos < seq_len (generation.rs:25)


class _ReplayPrefill(_Replay):
    """+ q3_prefill's contract: cache and last-token logits as after sequential forwards."""

    def __init__(self, path):
        super().__init__(path)
        self.prefills = []

    def prefill(self, tokens, pos0):
        self.prefills.append((list(tokens), pos0))
        for i, t in enumerate(tokens):
            lg = self.o.forward(t, pos0 + i)
        return lg


@pytest.mark.parametrize("temperature,topp", [(0.0, 0.9), (0.8, 0.9), (1.0, 1.0)])
def test_chat_with_one_prefill_per_turn_matches_the_reference_loop(ckpt, temperature, topp):
    """generation.rs:95-126 forwards and samples every prompt token; the drop-in runs q3_prefill once per turn
    and advances the sampler's RNG by the draws the discarded samples would have consumed: same replies,
    same RNG state afterwards."""
    path = ckpt("tiny", 64, seed=2)
    turns = [[5, 9, 200, 31], [7, 7, 300], [11]]
    V = orc.Model(path).config["vocab_size"]
    a, b = _ReplayPrefill(path), _ReplayPrefill(path)
    sa, sb = Sampler(V, temperature, topp, 42), Sampler(V, temperature, topp, 42)
    ra = generation.chat(a, sa, turns, use_prefill=False, max_new_per_turn=5)
    rb = generation.chat(b, sb, turns, use_prefill=True, max_new_per_turn=5)
    assert ra == rb and [len(r) for r in ra] == [5, 5, 5]
    assert sa.rng_state == sb.rng_state
    assert not a.prefills and [p[1] for p in b.prefills] == [0, 4 + 5, 4 + 5 + 3 + 5]  # one call per turn, at the turn's position
    assert len(a.calls) == len(b.calls) + sum(len(t) for t in turns)


def test_chat_stops_a_turn_on_eos_and_resets_on_a_full_window(ckpt):
    path = ckpt("tiny", 64, seed=2)
    t = _ReplayPrefill(path)
    V, S = t.get_config().vocab_size, t.get_config().seq_len
    free = generation.chat(t, Sampler(V, 0.0, 0.9, 0), [[5, 9]], max_new_per_turn=6)[0]
    eos = free[3]
    stopped = generation.chat(_ReplayPrefill(path), Sampler(V, 0.0, 0.9, 0), [[5, 9]], eos_token_id=eos)[0]
    assert stopped == free[:free.index(eos)]  # the terminating token is not emitted (generation.rs:136-141)
    # no cap and no eos: generation runs to the end of the window, then the position resets and the next turn starts at 0
    t2 = _ReplayPrefill(path)
    r = generation.chat(t2, Sampler(V, 0.0, 0.9, 0), [[5, 9], [7]])
    assert len(r[0]) == S - 2 and t2.prefills[1] == ([7], 0)
