"""qwen3_rs_b200/tokenizer_export.py against the reference's own unit tests
(qwen3-export/tests/unit/tokenizer_exporter_test.rs, ported; line numbers cited per test) and the template strings of
chat_template_exporter.rs:232-265; then end to end: HF-style tokenizer.json -> .tokenizer/.template files -> the C++
host mirror's Tokenizer reads them and encodes text."""
import json
import math
import os
import struct

import pytest

from qwen3_rs_b200 import tokenizer_export as te


def _read_tokenizer(path):
    data = open(path, "rb").read()
    hdr = struct.unpack_from("<III", data, 0)
    off, toks = 12, []
    while off < len(data):
        score, n = struct.unpack_from("<fI", data, off)
        toks.append((score, data[off + 8:off + 8 + n]))
        off += 8 + n
    return hdr, toks


def _write_json(d, obj, name="tokenizer.json"):
    with open(os.path.join(d, name), "w", encoding="utf-8") as f:
        json.dump(obj, f, ensure_ascii=False)


def test_unicode_to_byte_mapping():  # :10-60
    assert te.token_to_bytes("A") == bytes([65]) and te.token_to_bytes("z") == bytes([122])
    assert te.token_to_bytes("!") == bytes([33]) and te.token_to_bytes("~") == bytes([126])
    assert te.token_to_bytes("¡") == bytes([161]) and te.token_to_bytes("¬") == bytes([172])
    assert te.token_to_bytes("®") == bytes([174]) and te.token_to_bytes("ÿ") == bytes([255])
    assert te.token_to_bytes("hello") == bytes([104, 101, 108, 108, 111]) and te.token_to_bytes("ABC") == b"ABC"
    assert te.token_to_bytes("") == b""
    assert te.token_to_bytes("Ā") == bytes([0])       # chr(256): first unprintable byte
    assert te.token_to_bytes("Ġ") == b" " and te.token_to_bytes("Ċ") == b"\n"   # the GPT-2 space / newline stand-ins
    assert sorted(te.unicode_to_byte_map().values()) == list(range(256))          # a bijection onto the bytes
    assert te.token_to_bytes("中") == "中".encode("utf-8")                          # unmapped character: its UTF-8 bytes


def test_ordered_tokens_vocabulary_and_merges(tmp_path):  # :80-183
    assert te.create_ordered_tokens({"token_c": 3, "token_a": 1, "token_b": 2}) == [(1, "token_a"), (2, "token_b"), (3, "token_c")]
    assert te.create_ordered_tokens({}) == []
    assert te.extract_vocabulary({"model": {"vocab": {"hello": 1, "world": 2, "test": 3}}}) == {"hello": 1, "world": 2, "test": 3}
    with pytest.raises(ValueError, match="Could not find vocabulary"):
        te.extract_vocabulary({"model": {}})
    ranks = te.extract_merge_ranks({"model": {"merges": ["h e", "l l", "o !", "he ll"]}})
    assert ranks == {"h e": 0, "l l": 1, "o !": 2, "he ll": 3}
    assert te.extract_merge_ranks({"model": {"merges": []}}) == {} and te.extract_merge_ranks({"model": {}}) == {}
    assert te.extract_merge_ranks({"model": {"merges": [["h", "e"]]}}) == {}      # pair-style merges are not strings: skipped


def test_scores(tmp_path):  # :193-208
    assert te.DEFAULT_SCORE == -1e6
    assert te.token_score("x", {"x": 0}) == 0.0
    assert abs(te.token_score("x", {"x": 1}) + math.log(2)) < 1e-3 and abs(te.token_score("x", {"x": 10}) + 2.397895) < 1e-3
    assert te.token_score("ab", {"a b": 0}) == te.DEFAULT_SCORE  # the lookup is by token string, the keys are merge strings


def test_errors(tmp_path):  # :220-243, :584-636
    with pytest.raises(FileNotFoundError, match="tokenizer.json not found"):
        te.export_tokenizer(str(tmp_path), str(tmp_path / "out"), 0, 0)
    (tmp_path / "tokenizer.json").write_text("{ not json")
    with pytest.raises(ValueError, match="Failed to parse tokenizer.json"):
        te.export_tokenizer(str(tmp_path), str(tmp_path / "out"), 0, 0)
    _write_json(str(tmp_path), {"model": {"type": "BPE"}})
    with pytest.raises(ValueError, match="Could not find vocabulary"):
        te.export_tokenizer(str(tmp_path), str(tmp_path / "out"), 0, 0)


def test_load_token_data_and_max_token_length(tmp_path):  # :258-351
    _write_json(str(tmp_path), {"model": {"vocab": {"hello": 1, "world": 2, "!": 3}, "merges": ["h e", "l l"]},
                                "added_tokens": [{"id": 100, "content": "<special>"}]})
    vocab, ranks, mx = te.load_token_data(str(tmp_path))
    assert vocab == {"hello": 1, "world": 2, "!": 3, "<special>": 100} and ranks == {"h e": 0, "l l": 1} and mx == 9
    _write_json(str(tmp_path), {"model": {"vocab": {"a": 1, "bb": 2, "this_is_a_long_token!": 3}}})
    assert te.load_token_data(str(tmp_path))[2] == len("this_is_a_long_token!")
    _write_json(str(tmp_path), {"model": {"vocab": {}}})
    assert te.load_token_data(str(tmp_path))[2] == 0


def test_complete_export(tmp_path):  # :378-492
    _write_json(str(tmp_path), {
        "added_tokens": [{"id": 100, "content": "<|endoftext|>", "special": True}, {"id": 101, "content": "<|startoftext|>", "special": True}],
        "model": {"vocab": {"hello": 1, "world": 2, "!": 3, "h e": 4, "l l": 5, "other": 6}, "merges": ["h e", "l l"]}})
    out = te.export_tokenizer(str(tmp_path), str(tmp_path / "output"), 100, 101)
    (mx, bos, eos), toks = _read_tokenizer(out)
    assert (mx, bos, eos) == (15, 100, 101)  # len("<|startoftext|>") == 15
    want = [("hello", te.DEFAULT_SCORE), ("world", te.DEFAULT_SCORE), ("!", te.DEFAULT_SCORE), ("h e", 0.0),
            ("l l", -math.log(2)), ("other", te.DEFAULT_SCORE), ("<|endoftext|>", te.DEFAULT_SCORE), ("<|startoftext|>", te.DEFAULT_SCORE)]
    assert len(toks) == len(want)
    for (score, b), (tok, s) in zip(toks, want):
        assert abs(score - s) < 1e-3 and b == tok.encode()
    # empty vocabulary: header only (:497-530)
    _write_json(str(tmp_path), {"model": {"vocab": {}}})
    hdr, toks = _read_tokenizer(te.export_tokenizer(str(tmp_path), str(tmp_path / "empty"), 0, 0))
    assert hdr == (0, 0, 0) and toks == []
    # a large vocabulary (:697-727)
    _write_json(str(tmp_path), {"model": {"vocab": {f"token_{i}": i for i in range(1000)}}})
    assert os.path.getsize(te.export_tokenizer(str(tmp_path), str(tmp_path / "large"), 0, 1)) > 12 + 1000 * 8


def test_templates(tmp_path):  # chat_template_exporter.rs:71-141, 232-265
    qwen = "{%- if messages[0].role == 'system' %}<|im_start|>system ... <|im_end|> {%- if enable_thinking %}"
    assert te.analyze_template_capabilities(qwen) == ("Qwen3", True, True)
    assert te.analyze_template_capabilities("<|im_start|>user<|im_end|>") == ("Qwen3", False, False)
    assert te.analyze_template_capabilities("<｜User｜>x<｜Assistant｜> think system_prompt") == ("DeepSeek", True, True)
    assert te.analyze_template_capabilities("plain") == ("Unknown", False, False)
    assert te.get_template_configs(True, True) == [(False, False), (False, True), (True, False), (True, True)]
    assert te.get_template_configs(False, True) == [(False, False), (True, False)]
    assert te.render_chat_template("Qwen3", False, True) == "<|im_start|>user\n%s<|im_end|>\n<|im_start|>assistant\n"
    assert te.render_chat_template("Qwen3", False, False) == "<|im_start|>user\n%s<|im_end|>\n<|im_start|>assistant\n<think>\n\n</think>\n\n"
    assert te.render_chat_template("Qwen3", True, True) == "<|im_start|>system\n%s<|im_end|>\n<|im_start|>user\n%s<|im_end|>\n<|im_start|>assistant\n"
    assert te.render_chat_template("DeepSeek", True, False) == "%s<｜User｜>%s<｜Assistant｜><think>\n</think>"
    with pytest.raises(ValueError, match="Unknown template type"):
        te.render_chat_template("Unknown", False, False)
    with pytest.raises(ValueError, match="No chat template found"):
        te.export_templates(str(tmp_path), str(tmp_path / "m.bin"))
    _write_json(str(tmp_path), {"chat_template": qwen}, "tokenizer_config.json")
    written = te.export_templates(str(tmp_path), str(tmp_path / "m.bin"))
    assert [os.path.basename(w) for w in written] == ["m.bin.template", "m.bin.template.with-thinking", "m.bin.template.with-system",
                                                      "m.bin.template.with-system-and-thinking"]
    assert open(written[0]).read().endswith("</think>\n\n")


def test_exported_files_drive_the_cpp_tokenizer(tmp_path):
    """HF-style tokenizer.json (byte-level tokens in GPT-2 spelling) -> exporter -> files -> qwen3::Tokenizer."""
    from test_cpp_host import _run  # the compiled driver
    import shutil
    import subprocess
    from conftest import ROOT
    from qwen3_rs_b200 import build
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "host_check")
    libdir = os.path.dirname(build.build())
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", os.path.join(ROOT, "tests", "cpp", "host_check.cpp"), "-I",
                           os.path.join(ROOT, "include"), "-L", libdir, "-lqwen3cuda", "-Wl,-rpath," + libdir, "-o", exe])
    chars = list("helowrd!") + ["Ġ", "Ċ"]                       # Ġ = space, Ċ = newline in the GPT-2 spelling
    vocab = {c: i for i, c in enumerate(chars)}
    for t in ("he", "ll", "hell", "hello", "Ġw", "or", "ld", "Ġworld"):
        vocab[t] = len(vocab)
    hf = tmp_path / "hf"
    hf.mkdir()
    _write_json(str(hf), {"model": {"vocab": vocab, "merges": ["h e", "l l", "he ll", "hell o"]},
                          "added_tokens": [{"id": len(vocab), "content": "<|im_start|>"}, {"id": len(vocab) + 1, "content": "<|im_end|>"}]})
    _write_json(str(hf), {"chat_template": "<|im_start|>{{ m }}<|im_end|>"}, "tokenizer_config.json")
    base = str(tmp_path / "model.bin")
    te.export_tokenizer(str(hf), base, len(vocab), len(vocab) + 1)
    te.export_templates(str(hf), base)
    V = len(vocab) + 2
    texts = tmp_path / "texts.txt"
    texts.write_bytes("hello world!\n<|im_start|>hello<|im_end|>".encode())
    out = _run(exe, "tokenize", base, V, 0, texts)
    ids = [int(x) for x in out[0].split("|")[0].split()]
    inv = {i: t for t, i in vocab.items()}
    assert b"".join(te.token_to_bytes(inv[i]) for i in ids) == b"hello world!" and out[0].strip().endswith("| 1")
    # all scores are equal (the exporter quirk), so the leftmost mergeable pair wins each round: h+e, he+l? no ("hel" is not a
    # token) ... l+l, he+ll, hell+o; " w", "or", "ld" merge, but no pair of them is in the vocabulary, so " world" is unreachable
    assert ids == [vocab["hello"], vocab["Ġw"], vocab["or"], vocab["ld"], vocab["!"]]
    ids2 = [int(x) for x in out[1].split("|")[0].split()]
    assert ids2 == [len(vocab), vocab["hello"], len(vocab) + 1]  # special tokens found by the "<...>" scan
    assert "\n".join(_run(exe, "render", base, V, 0, 0, "-", "hi")) == "<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\n<think>\n\n</think>\n\n"
