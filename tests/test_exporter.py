"""Exporter restatement vs the reference's own known-answer tests
(qwen3-export/tests/unit/model_exporter_test.rs) -- the one part of the path the reference pins."""
import os
import struct

import numpy as np
import pytest

from oracle import binding as orc
from qwen3_rs_b200 import export, synth

IMPLS = [("numpy", lambda w, gs: export.quantize_q80(np.asarray(w, np.float32), gs)),
         ("c-oracle", lambda w, gs: orc.quantize_q80(np.asarray(w, np.float32), gs))]


def test_round_half_to_even_basic():  # model_exporter_test.rs:27-33
    for x, want in [(1.4, 1.0), (1.6, 2.0), (-1.4, -1.0), (-1.6, -2.0)]:
        assert orc.round_half_to_even(x) == want
        assert np.rint(np.float32(x)) == want


def test_round_half_to_even_halfway_cases():  # :36-45
    for x, want in [(0.5, 0.0), (1.5, 2.0), (2.5, 2.0), (3.5, 4.0), (-0.5, 0.0), (-1.5, -2.0), (-2.5, -2.0)]:
        assert orc.round_half_to_even(x) == want
        assert np.rint(np.float32(x)) == want


@pytest.mark.parametrize("name,q80", IMPLS)
def test_quantize_q80_known_values(name, q80):  # :48-67
    q, s, _ = q80([0.0, 127.0, -127.0, 63.5], 4)
    assert len(s) == 1 and abs(s[0] - 1.0) < 1e-6
    assert q.tolist() == [0, 127, -127, 64]


@pytest.mark.parametrize("name,q80", IMPLS)
def test_quantize_q80_zero_weights(name, q80):  # :70-87
    q, s, err = q80([0.0] * 4, 4)
    assert s[0] == 1.0 and not q.any() and err == 0.0


@pytest.mark.parametrize("name,q80", IMPLS)
def test_quantize_q80_invalid_group_size(name, q80):  # :90-101
    with pytest.raises(ValueError, match="multiple of group_size"):
        q80([1.0, 2.0, 3.0], 4)


def test_find_optimal_group_size():  # :104-134
    table = [((128, 64), 64), ((128, 32), 32), ((128, 16), 16), ((32, 64), 32), ((128, 96), 4), ((60, 40), 20),
             ((60, 30), 30), ((127, 64), 4), ((15, 8), 4), ((128, 2), 4)]
    for (dim, req), want in table:
        assert export.find_optimal_group_size(dim, req) == want
        assert orc.find_optimal_group_size(dim, req) == want


def test_header_constants():  # :137-142
    assert export.MAGIC_NUMBER == 0x616A6331 and export.VERSION == 1
    assert export.HEADER_SIZE == 256 and export.MIN_GROUP_SIZE == 4


@pytest.mark.parametrize("name,q80", IMPLS)
def test_quantization_symmetry(name, q80):  # :145-160
    q, s, _ = q80([100.0, -100.0, 50.0, -50.0], 4)
    assert abs(s[0] - 100.0 / 127.0) < 1e-6
    assert q[0] == -q[1] and q[2] == -q[3]


@pytest.mark.parametrize("name,q80", IMPLS)
def test_quantization_binary_consistency(name, q80):  # :396-457
    w = [1.0, 2.0, 3.0, 4.0, -5.0, 6.0, -7.0, 8.0, 0.1, -0.2, 0.3, -0.4, 100.0, -100.0, 50.0, -25.0]
    q, s, err = q80(w, 4)
    assert len(q) == 16 and len(s) == 4 and err >= 0
    for got, want in zip(s, [4.0 / 127, 8.0 / 127, 0.4 / 127, 100.0 / 127]):
        assert abs(got - want) < 1e-6
    assert q.min() >= -127
    deq = q.astype(np.float32).reshape(4, 4) * s[:, None]
    assert np.abs(deq.reshape(-1) - np.float32(w)).max() <= s.max() * 0.6


@pytest.mark.parametrize("name,q80", IMPLS)
def test_quantization_edge_cases(name, q80):  # :367-393 and :460-515: must not crash, finite scales
    q80([np.nan, 1.0, 2.0, 3.0], 4)
    q80([np.inf, 1.0, 2.0, 3.0], 4)
    _, s, _ = q80([1e-30, 2e-30, 3e-30, 4e-30], 4)
    assert s[0] > 0
    _, s, _ = q80([1e30, -1e30, 1e29, -1e29], 4)
    assert np.isfinite(s[0])
    q, s, err = q80([0.0] * 8, 4)
    assert len(s) == 2 and (s > 0).all() and not q.any() and err == 0
    _, s, _ = q80([1000.0] + [0.0] * 7, 4)
    assert s[0] > s[1]
    _, s, _ = q80([1.0, 2.0, 3.0, 4.0, 1000.0, 2000.0, 3000.0, 4000.0], 4)
    assert s[1] > s[0] * 100


def test_numpy_and_c_quantizers_agree_bitwise():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(64 * 1024) * rng.random(64 * 1024) * 0.1).astype(np.float32)
    w[:64] = 0
    w[100] = np.nan
    # exact .5 ties exercise half-to-even
    w[128:192] = np.arange(64, dtype=np.float32) * 0.5
    w[191] = 127.0
    for gs in (4, 32, 64, 128):
        qa, sa, ea = export.quantize_q80(w, gs)
        qb, sb, eb = orc.quantize_q80(w, gs)
        assert np.array_equal(qa, qb) and np.array_equal(sa, sb)
        assert ea == pytest.approx(eb, rel=1e-6)


def test_checkpoint_layout_and_hf_roundtrip(tmp_path):
    """HF dir (safetensors, bf16) -> export_model gives the same bytes as the streaming path, and the
    header / tensor order match SURVEY appendix A."""
    sh = synth.SHAPES["tiny-untied"]
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    synth.write_hf_dir(sh, str(tmp_path / "hf"), seed=3)
    info = export.export_model(str(tmp_path / "hf"), a, 64)
    synth.export_synthetic(sh, b, 64, seed=3)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert info["shared_classifier"] is False and info["group_size"] == 64
    assert os.path.getsize(a) == synth.checkpoint_bytes(sh, 64)
    hdr = struct.unpack("<13i", open(a, "rb").read(52))
    assert hdr == (0x616A6331, 1, 1, sh.dim, sh.hidden_dim, sh.n_layers, sh.n_heads, sh.n_kv_heads, sh.vocab_size,
                   sh.max_seq_len, sh.head_dim, 0, 64)
    assert open(a, "rb").read(256)[52:] == b"\0" * 204
    # first tensor after the norms is the embedding, quantized with the restated quantizer
    norms = (2 * sh.n_layers * sh.dim + sh.dim + 2 * sh.n_layers * sh.head_dim) * 4
    emb = synth.make_tensor(export.EMBED_TOKENS_KEY, (sh.vocab_size, sh.dim), "embed", 3).numpy()
    q, s, _ = orc.quantize_q80(emb, 64)
    raw = np.fromfile(a, dtype=np.uint8)
    off = 256 + norms
    assert np.array_equal(raw[off:off + q.size].view(np.int8), q)
    assert np.array_equal(raw[off + q.size: off + q.size + 4 * s.size].view("<f4"), s)


def test_tied_checkpoint_detected_as_shared(tmp_path):
    sh = synth.SHAPES["tiny"]
    synth.write_hf_dir(sh, str(tmp_path / "hf"), seed=0)
    info = export.export_model(str(tmp_path / "hf"), str(tmp_path / "a.bin"), 64)
    assert info["shared_classifier"] is True  # no lm_head tensor -> shared (models/qwen3.rs:69)


def test_golden_checkpoints_regenerate_bit_identically(ckpt, golden_meta):
    import hashlib

    for key, meta in golden_meta.items():
        path = ckpt(meta["shape"], meta["group_size"], meta["seed"])
        assert os.path.getsize(path) == meta["bytes"]
        assert hashlib.sha256(open(path, "rb").read()).hexdigest() == meta["sha256"], key


# ---- config loader: the reference's own tests (qwen3-export/tests/unit/config_loader_test.rs) -----------------
_HF_CFG = {"architectures": ["Qwen3ForCausalLM"], "hidden_size": 256, "intermediate_size": 1024, "num_hidden_layers": 4,
           "num_attention_heads": 8, "num_key_value_heads": 8, "vocab_size": 1000, "max_position_embeddings": 512,
           "rms_norm_eps": 1e-6, "head_dim": 32, "bos_token_id": 1, "eos_token_id": 2}


def test_load_hf_config_valid_and_defaults(tmp_path):  # config_loader_test.rs:31-51, :90-118
    import json
    p = tmp_path / "config.json"
    p.write_text(json.dumps(_HF_CFG))
    c = export.ExportConfig.from_hf_json(str(p))
    assert (c.dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads) == (256, 1024, 4, 8, 8)
    assert (c.vocab_size, c.max_seq_len, c.head_dim, c.bos_token_id, c.eos_token_id) == (1000, 512, 32, 1, 2)
    assert abs(c.norm_eps - 1e-6) < 1e-9
    d = {k: v for k, v in _HF_CFG.items() if k not in ("head_dim", "bos_token_id", "eos_token_id")}
    p.write_text(json.dumps(d))
    c = export.ExportConfig.from_hf_json(str(p))
    assert (c.bos_token_id, c.eos_token_id, c.head_dim) == (0, 0, 256 // 8)


def test_load_hf_config_errors(tmp_path):  # :54-87 (message text is serde's in the reference; the failure is what matters)
    import json
    p = tmp_path / "config.json"
    p.write_text("invalid json")
    with pytest.raises(ValueError):
        export.ExportConfig.from_hf_json(str(p))
    p.write_text(json.dumps({"intermediate_size": 1024, "num_hidden_layers": 4}))
    with pytest.raises((ValueError, KeyError)):
        export.ExportConfig.from_hf_json(str(p))
