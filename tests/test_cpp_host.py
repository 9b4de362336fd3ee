"""include/qwen3_transformer.hpp: the C++ host mirror of the reference interface (TransformerBuilder / Transformer /
Sampler / generate over the C ABI).  CPU: it compiles against the library, its sampler follows the oracle's restatement of
sampler.rs draw by draw, and construction errors surface with the C ABI's codes.  GPU: generate() through the C++ layer."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import has_cuda
from oracle import binding as orc
from qwen3_rs_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    lib = build.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_check")
    libdir = os.path.dirname(lib)
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "cpp", "host_check.cpp"), "-I", os.path.join(ROOT, "include"),
                           "-L", libdir, "-lqwen3cuda", "-Wl,-rpath," + libdir, "-o", out])
    return out


def _run(exe, *args):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout.split("\n")


@pytest.mark.parametrize("temperature,topp", [(0.0, 0.9), (1.0, 0.9), (0.7, 1.0), (1.3, 0.5), (1.0, 0.0)])
def test_cpp_sampler_follows_the_oracle(host_check, tmp_path, temperature, topp):
    rng = np.random.default_rng(3)
    V, n = 2000, 40
    logits = (rng.standard_normal((n, V)) * 3).astype(np.float32)
    logits[0, :5] = logits[0].max() + 1  # ties: the last maximum wins (sampler.rs:57-59)
    f = tmp_path / "logits.f32"
    logits.tofile(f)
    out = _run(host_check, "sample", V, temperature, topp, 1234, n, f)
    got, rng_state = [int(x) for x in out[:n]], int(out[n])
    ref = orc.Sampler(V, temperature, topp, 1234)
    want = [ref.sample(l.copy()) for l in logits]
    assert got == want  # same libm (glibc expf), same order of operations: identical draws
    if temperature == 0.0:
        assert got[0] == 4 and rng_state == 1234  # argmax consumes no randomness
    else:
        py = __import__("qwen3_rs_b200.sampler", fromlist=["Sampler"]).Sampler(V, temperature, topp, 1234)
        for _ in range(n):
            py.random_u32()
        assert rng_state == py.rng_state  # one draw per sample


def test_cpp_builder_reports_the_abi_error_codes(host_check, tmp_path, ckpt):
    out = _run(host_check, "errors", tmp_path / "missing.bin", ckpt("micro", 32, 7))
    assert out[0].startswith("-2 ")  # Q3_EIO: cannot open the checkpoint (models/mod.rs:56-59)
    if not has_cuda():
        assert out[1].startswith("-4 ")  # Q3_ECUDA: no device, and no CPU fallback
    else:
        assert out[1].startswith("0 ok")


@pytest.mark.gpu
def test_cpp_generate_matches_the_oracle(host_check, ckpt):
    path = ckpt("micro", 32, 7)
    prompt = [3, 17, 5]
    want = orc.Model(path).generate(prompt, 10)
    out = _run(host_check, "generate", path, 10, *prompt)
    assert [int(x) for x in out[0].split()] == want      # generate() through Transformer::forward + host argmax
    assert [int(x) for x in out[1].split()] == want      # the device-resident loop
    assert out[2].strip() == "-1"                         # Q3_EINVAL for an out-of-range token


# ---- tokenizer.rs / render_prompt --------------------------------------------------------------
from oracle import tokenizer_ref as tr  # noqa: E402


def _toy_tokenizer(tmp_path, thinking_templates=True):
    # ids:     0     1     2     3     4      5       6        7              8      9      10       11
    vocab = [b"a", b"b", b"c", b" ", b"ab", b"bc", b"abc", b"<|im_start|>", b"<", b">", b"\xc3\xa9", b"ab"]
    score = [0.0, 0.0, 0.0, 0.0, 5.0, 7.0, 9.0, 0.0, 0.0, 0.0, 0.0, 99.0]   # id 11 duplicates "ab": never found (first position wins)
    base = str(tmp_path / "toy.bin")
    tr.write_tokenizer_file(base, vocab, score, max_token_length=12, bos=7, eos=3)
    open(base + ".template", "w").write("<|im_start|>user\n%s<|im_end|>\n")
    open(base + ".template.with-system", "w").write("S:%s|U:%s|")
    if thinking_templates:
        open(base + ".template.with-thinking", "w").write("T:%s")
    return base, len(vocab)


def test_tokenizer_known_answers(host_check, tmp_path):
    base, V = _toy_tokenizer(tmp_path)
    cases = {
        "abc": [6],                   # round 1: "bc" (7.0) outranks "ab" (5.0) -> [a, bc]; round 2: a + bc = "abc" (9.0) -> [abc]
        "ab": [4],                    # first vocabulary position of "ab" is id 4, not the duplicate 11
        "abab": [4, 4],               # equal scores: the leftmost pair merges first
        "<|im_start|>ab": [7, 4],     # special token found by the "<...>" scan
        "<ab>": [8, 4, 9],            # "<ab>" is not in the vocabulary: falls back to characters
        "a<b": [0, 8, 1],             # no closing ">"
        "<aaaaaaaaaaaa>": [8] + [0] * 12 + [9],  # ">" is beyond max_token_length (12) characters: not even looked up
        "aéz b": [0, 10, 3, 1],  # two-byte character found as one token; unknown "z" skipped
        "": [],
    }
    ref = tr.Tokenizer(base, V, False)
    for text, want in cases.items():
        assert ref.encode(text) == want, text  # the literal restatement agrees with the hand-computed answers
    texts = tmp_path / "texts.txt"
    texts.write_bytes("\n".join(cases).encode("utf-8"))
    out = _run(host_check, "tokenize", base, V, 0, texts)
    for line, (text, want) in zip(out, cases.items()):
        ids, ok = line.split("|")
        assert [int(x) for x in ids.split()] == want, text
        assert ok.strip() == ("0" if "z" in text else "1")  # decode(encode(text)) == text unless a character was dropped
    assert out[len(cases)] == "meta 12 7 3 %d" % V


def test_tokenizer_matches_the_literal_restatement_on_random_vocabularies(host_check, tmp_path):
    rng = np.random.default_rng(11)
    alphabet = [c.encode("utf-8") for c in "abcdefgh é中<>|"]
    vocab = list(alphabet)
    while len(vocab) < 120:  # grow a BPE-like vocabulary by concatenating existing tokens
        a, b = rng.integers(0, len(vocab), 2)
        t = vocab[a] + vocab[b]
        if len(t) <= 10:
            vocab.append(t)
    vocab += [b"<|x|>", b"<y>"]
    score = [0.0] * len(alphabet) + [float(x) for x in rng.integers(1, 40, len(vocab) - len(alphabet))]  # many ties
    base = str(tmp_path / "rnd.bin")
    tr.write_tokenizer_file(base, vocab, score, max_token_length=8, bos=0, eos=1)
    ref = tr.Tokenizer(base, len(vocab), True)
    chars = list("abcdefgh é中<>|") + ["<|x|>", "<y>", "q"]
    lines = ["".join(rng.choice(chars, rng.integers(0, 40))) for _ in range(200)]
    texts = tmp_path / "texts.txt"
    texts.write_bytes("\n".join(lines).encode("utf-8"))
    out = _run(host_check, "tokenize", base, len(vocab), 1, texts)
    for line, text in zip(out, lines):
        assert [int(x) for x in line.split("|")[0].split()] == ref.encode(text), text


def test_tokenizer_short_file_and_templates(host_check, tmp_path):
    base, V = _toy_tokenizer(tmp_path, thinking_templates=False)
    data = open(base + ".tokenizer", "rb").read()
    open(base + ".tokenizer", "wb").write(data[:12 + 3 * 9 + 6])  # cut inside token 3
    ref = tr.Tokenizer(base, V, False)
    assert ref.vocab[:3] == [b"a", b"b", b"c"] and all(v == b"" for v in ref.vocab[3:]) and len(ref.vocab) == V
    texts = tmp_path / "t.txt"
    texts.write_bytes(b"abc cab")
    out = _run(host_check, "tokenize", base, V, 0, texts)
    assert [int(x) for x in out[0].split("|")[0].split()] == ref.encode("abc cab") == [0, 1, 2, 2, 0, 1]
    # render_prompt: every "%s" receives the same text (str::replace), system template only at pos 0
    for pos, sys_p, user in ((0, "be brief", "hi"), (5, "be brief", "hi"), (0, None, "x%sy")):
        got = "\n".join(_run(host_check, "render", base, V, 0, pos, sys_p if sys_p is not None else "-", user))
        assert got == tr.render_prompt(pos, sys_p, user, ref)
    assert tr.render_prompt(0, "be brief", "hi", ref) == "S:be brief\nhi|U:be brief\nhi|"
    # missing template files give empty templates (tokenizer.rs:111-117)
    assert tr.Tokenizer(base, V, True).prompt_template == ""
    assert "\n".join(_run(host_check, "render", base, V, 1, 3, "-", "hello")) == ""


# ---- chat loop with one prefill per user turn, C++ mirror ----------------------------------------
class _Fake:
    """Python twin of host_check's FakeTransformer (same integer-hash logits)."""

    class _Cfg:
        seq_len, vocab_size = 40, 64

    def __init__(self):
        self.forwards = 0

    def get_config(self):
        return self._Cfg

    def forward(self, token, pos):
        self.forwards += 1
        i = np.arange(64, dtype=np.uint64)
        h = (i * 2654435761 + token * 40503 + pos * 69069) & 0xFFFFFFFF
        return (((h >> 8) & 1023).astype(np.float32) / np.float32(64.0)).astype(np.float32)

    def prefill(self, tokens, pos0):
        for k, t in enumerate(tokens):
            lg = self.forward(t, pos0 + k)
        return lg


@pytest.mark.parametrize("temperature,topp", [(0.0, 0.9), (0.8, 0.9), (1.0, 1.0)])
def test_cpp_chat_one_prefill_per_turn(host_check, temperature, topp):
    seq = _run(host_check, "chat", 0, temperature, topp, 42, 5)
    pre = _run(host_check, "chat", 1, temperature, topp, 42, 5)
    assert seq[:3] == pre[:3]                                   # same replies
    assert seq[3].split()[1] == pre[3].split()[1]               # same RNG state afterwards
    assert seq[3].split()[3:] == ["23", "prefills", "0"] and pre[3].split()[3:] == ["15", "prefills", "3"]
    if temperature == 0.0:  # argmax path is exact across languages: the Python mirror must give the same tokens
        from qwen3_rs_b200 import generation
        from qwen3_rs_b200.sampler import Sampler
        want = generation.chat(_Fake(), Sampler(64, 0.0, topp, 42), [[5, 9, 20, 31], [7, 7, 30], [11]], max_new_per_turn=5)
        assert [[int(x) for x in l.split()] for l in pre[:3]] == want
