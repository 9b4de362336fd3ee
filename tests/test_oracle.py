"""The CPU oracle (oracle/q3_oracle.c): hand-computed known answers, invariants, agreement with the
independent numpy restatement (oracle/np_forward.py), and the committed golden vectors.

The reference has no tests for its forward path, so this is how the oracle is pinned."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES
from oracle import binding as orc
from oracle import np_forward as npf


# ---- tensor.rs known answers ---------------------------------------------------------------
def test_quantize_ties_round_half_away_from_zero():
    # scale = 127/127 = 1 -> q = round(x): 0.5 -> 1, -0.5 -> -1, 2.5 -> 3, -2.5 -> -3 (f32::round)
    x = np.array([127.0, 0.5, -0.5, 2.5, -2.5, 1.4999, -126.5, 0.0], np.float32)
    q, s = orc.quantize(x, 8)
    assert s[0] == 1.0
    assert q.tolist() == [127, 1, -1, 3, -3, 1, -127, 0]


def test_quantize_zero_group_has_scale_zero_and_q_zero():
    # unlike the exporter (scale 1.0), runtime quantize stores scale 0.0 (tensor.rs:110-115)
    q, s = orc.quantize(np.zeros(64, np.float32), 32)
    assert s.tolist() == [0.0, 0.0] and not q.any()


def test_quantize_scale_is_max_over_127():
    x = np.linspace(-3, 5, 64, dtype=np.float32)
    q, s = orc.quantize(x, 64)
    assert s[0] == np.float32(5.0) / np.float32(127.0)
    assert q.max() == 127 and q[0] == round(-3 / s[0])


def test_matmul_two_rows_two_groups_by_hand():
    gs, n, d = 4, 8, 2
    xq = np.array([1, 2, 3, 4, -1, -2, -3, -4], np.int8)
    xs = np.array([0.5, 0.25], np.float32)
    wq = np.array([1, 1, 1, 1, 2, 2, 2, 2, -1, 0, 1, 0, 127, 127, 127, 127], np.int8)
    ws = np.array([2.0, 4.0, 1.0, 0.125], np.float32)
    # row0: g0 dot=10 -> 10*2*0.5=10 ; g1 dot=-20 -> -20*4*0.25=-20 ; sum -10
    # row1: g0 dot=2 -> 2*1*0.5=1 ; g1 dot=-1270 -> -1270*0.125*0.25=-39.6875 ; sum -38.6875
    out = orc.matmul(xq, xs, wq, ws, n, d, gs)
    assert out.tolist() == [-10.0, -38.6875]
    assert orc.group_dots(xq, wq, n, d, gs).tolist() == [[10, -20], [2, -1270]]


def test_dequantize_is_single_multiply():
    q = np.array([1, -2, 3, -4], np.int8)
    s = np.array([0.1, 0.3], np.float32)
    assert np.array_equal(orc.dequantize(q, s, 2), q.astype(np.float32) * np.repeat(s, 2))


# ---- layers.rs ------------------------------------------------------------------------------
def test_rmsnorm_known_answer():
    x = np.array([3.0, 4.0], np.float32)  # ss = 25, mean 12.5
    w = np.array([1.0, 2.0], np.float32)
    f = np.float32(1.0) / np.sqrt(np.float32(12.5) + np.float32(1e-6))
    assert np.array_equal(orc.rmsnorm(x, w), w * (f * x))


def test_rope_at_pos0_is_identity_and_rotation_preserves_norm():
    v = np.random.default_rng(1).standard_normal(128).astype(np.float32)
    cs0 = orc.rope_freqs(0, 128)
    assert np.array_equal(cs0[:, 0], np.ones(64, np.float32)) and not cs0[:, 1].any()
    assert np.array_equal(orc.rope_apply(v, cs0), v)
    r = orc.rope_apply(v, orc.rope_freqs(1234, 128))
    assert np.linalg.norm(r) == pytest.approx(np.linalg.norm(v), rel=1e-5)
    # pair (i, i+64) convention: only dims 0 and 64 move when just pair 0 is non-zero
    e = np.zeros(128, np.float32)
    e[0] = 1.0
    cs = orc.rope_freqs(1, 128)
    out = orc.rope_apply(e, cs)
    assert out[0] == cs[0, 0] and out[64] == cs[0, 1] and np.count_nonzero(out) == 2
    # freq_0 = 1 -> angle = pos
    import math
    assert cs[0, 0] == np.float32(math.cos(1.0)) and cs[0, 1] == np.float32(math.sin(1.0))


def test_softmax_sums_to_one_and_is_shift_invariant():
    x = np.random.default_rng(2).standard_normal(1000).astype(np.float32) * 5
    p = orc.softmax(x)
    assert p.sum(dtype=np.float64) == pytest.approx(1.0, abs=1e-5)
    assert np.argmax(p) == np.argmax(x)


def test_argmax_returns_last_of_equal_maxima():  # sampler.rs:57-59 max_by -> last max
    assert orc.argmax(np.array([1.0, 5.0, 5.0, 2.0, 5.0, 0.0], np.float32)) == 4
    assert orc.argmax(np.array([-0.0, 0.0], np.float32)) == 1  # total_cmp: -0 < +0
    assert orc.argmax(np.array([0.0, -0.0], np.float32)) == 0
    assert orc.argmax(np.array([7.0], np.float32)) == 0


def test_sampler_rng_is_xorshift64star():  # sampler.rs:44-54
    s = orc.Sampler(16, 1.0, 0.9, 42)
    st = 42
    for _ in range(5):
        st ^= st >> 12
        st ^= (st << 25) & 0xFFFFFFFFFFFFFFFF
        st ^= st >> 27
        want = ((st * 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF) >> 32
        assert s.random_u32() == want
    f = s.random_f32()
    assert 0.0 <= f < 1.0


def test_sampler_temperature_zero_is_argmax_and_topp_stays_in_nucleus():
    logits = np.array([0.1, 3.0, 2.9, -1.0] + [-9.0] * 12, np.float32)
    assert orc.Sampler(16, 0.0, 0.9, 1).sample(logits) == 1
    s = orc.Sampler(16, 1.0, 0.5, 7)
    assert all(s.sample(logits) in (1, 2) for _ in range(200))
    s = orc.Sampler(16, 1.0, 1.0, 7)  # topp >= 1 -> plain multinomial
    assert set(s.sample(logits) for _ in range(400)) >= {1, 2}


# ---- whole model ----------------------------------------------------------------------------
def test_attention_at_pos0_returns_v0(ckpt):
    """softmax over a single score is exactly 1.0 -> attention output == the V row just written."""
    m = orc.Model(ckpt("tiny", 64, 0))
    tr = m.trace_layer(0)
    m.forward(3, 0)
    kv_mul = m.config["n_heads"] // m.config["n_kv_heads"]
    hd = m.config["head_dim"]
    for h in range(m.config["n_heads"]):
        kvh = h // kv_mul
        assert np.array_equal(tr["att_out"][h * hd:(h + 1) * hd], tr["v_row"][kvh * hd:(kvh + 1) * hd])


def test_generate_skips_prompt_forwards(ckpt):
    """generation.rs:26-28: prompt tokens except the last never reach forward(); their cache rows stay
    zero yet still take part in the softmax, so the result differs from forwarding every prompt token."""
    path = ckpt("tiny-untied", 64, 1)
    m = orc.Model(path)
    m.generate([5, 6, 7], 1)
    k, v = m.kv_cache()
    assert not k[:, :2].any() and not v[:, :2].any() and k[:, 2].any()
    a = orc.Model(path)
    got = a.generate([5, 6, 7], 4)
    b = orc.Model(path)
    want = [orc.argmax(b.forward(7, 2))]  # fresh (zero) cache, only (token 7, pos 2) forwarded
    assert got[0] == want[0]


def test_context_override_and_bounds(ckpt):
    m = orc.Model(ckpt("tiny", 64, 0), 16)
    assert m.config["seq_len"] == 16  # models/mod.rs:65-67 min(ctx, seq_len)
    with pytest.raises(IndexError):
        m.forward(0, 16)
    with pytest.raises(IndexError):
        m.forward(m.config["vocab_size"], 0)
    assert orc.Model(ckpt("tiny", 64, 0), 10 ** 6).config["seq_len"] == 256


def test_bad_checkpoints_rejected(tmp_path, ckpt):
    raw = bytearray(open(ckpt("micro", 32, 7), "rb").read())
    bad = tmp_path / "bad.bin"
    for patch, msg in [((0, b"\0\0\0\0"), "magic"), ((4, b"\2\0\0\0"), "version"), ((12, b"\0\0\0\0"), "dim"),
                       ((8, b"\7\0\0\0"), "architecture_id")]:
        r = bytearray(raw)
        r[patch[0]:patch[0] + 4] = patch[1]
        bad.write_bytes(bytes(r))
        with pytest.raises(RuntimeError, match=msg):
            orc.Model(str(bad))
    bad.write_bytes(bytes(raw[: len(raw) // 2]))
    with pytest.raises(RuntimeError, match="Insufficient data"):
        orc.Model(str(bad))
    with pytest.raises(RuntimeError, match="Failed to open"):
        orc.Model(str(tmp_path / "missing.bin"))


@pytest.mark.parametrize("name,gs,seed", GOLDEN_CASES)
def test_c_oracle_matches_numpy_restatement(ckpt, name, gs, seed):
    """Two independently written restatements: identical int8 activations at layer 0 and logits equal to
    float round-off at the first position (before the last-ulp libm differences between numpy and
    glibc can flip a quantisation step), same argmax along a short teacher-forced run."""
    path = ckpt(name, gs, seed)
    a, b = orc.Model(path), npf.NpModel(path)
    tr = a.trace_layer(0)
    la, lb = a.forward(5, 0), b.forward(5, 0)
    assert np.array_equal(tr["xq_attn_q"], b.trace[0]["xq_attn_q"])
    assert np.array_equal(tr["xq_attn_s"], b.trace[0]["xq_attn_s"])
    assert np.array_equal(tr["hq_q"], b.trace[0]["hq_q"])
    np.testing.assert_allclose(la, lb, rtol=0, atol=2e-5)
    tok = orc.argmax(la)
    for pos in range(1, 4):
        la, lb = a.forward(tok, pos), b.forward(tok, pos)
        assert orc.argmax(la) == int(np.argmax(lb))
        tok = orc.argmax(la)


@pytest.mark.parametrize("name,gs,seed", GOLDEN_CASES)
def test_oracle_reproduces_golden(ckpt, golden, name, gs, seed):
    key = f"{name}_gs{gs}"
    m = orc.Model(ckpt(name, gs, seed))
    prompt = golden[key + "_prompt"].tolist()
    want = golden[key + "_greedy"].tolist()
    assert m.generate(prompt, len(want)) == want
    m.reset()
    seq = prompt + want
    lg = golden[key + "_logits"]
    for p in range(lg.shape[0]):
        np.testing.assert_allclose(m.forward(seq[p], p), lg[p], rtol=0, atol=1e-5)


def test_operator_golden_vectors(golden):
    for gs in (32, 64, 128):
        q, s = orc.quantize(golden["quant_x"], gs)
        assert np.array_equal(q, golden[f"quant_q_gs{gs}"]) and np.array_equal(s, golden[f"quant_s_gs{gs}"])
        q2, s2 = npf.quantize(golden["quant_x"], gs)
        assert np.array_equal(q2, q) and np.array_equal(s2, s)
    out = orc.matmul(golden["mm_xq"], golden["mm_xs"], golden["mm_wq"], golden["mm_ws"], 256, 48, 64)
    assert np.array_equal(out, golden["mm_out"])
    assert np.array_equal(orc.group_dots(golden["mm_xq"], golden["mm_wq"], 256, 48, 64), golden["mm_dots"])
    assert np.array_equal(npf.matmul(golden["mm_xq"], golden["mm_xs"], golden["mm_wq"], golden["mm_ws"], 256, 48, 64), out)


def test_reassociation_sensitivity_is_what_the_tolerances_assume(ckpt):
    """Documents why end-to-end logits cannot be held to 1e-2 on random-init weights: changing only the
    ORDER of float sums (perturb mode) leaves every op within 1e-6 but, once a single int8
    activation lands on the other side of a rounding boundary, whole-model logits move by far more.
    Layer-level (teacher-forced) comparisons do not have that problem -- see tests/test_gpu_parity.py."""
    path = ckpt("small", 128, 2)
    a, b = orc.Model(path), orc.Model(path)
    worst = 0.0
    tok = 11
    try:
        for pos in range(24):
            la = a.forward(tok, pos)
            orc.set_perturb(1)
            lb = b.forward(tok, pos)
            orc.set_perturb(0)
            worst = max(worst, float(np.abs(la - lb).max()))
            tok = orc.argmax(la)
    finally:
        orc.set_perturb(0)
    assert worst < 2.0  # bounded, but NOT tiny; typically 1e-2..5e-1 on these shapes


def test_layer0_of_synthetic_checkpoints_sits_on_quantisation_ties(ckpt):
    """Why the GPU's fast mode flips int8 activations in up to a third of the layer-0 rows of the SYNTHETIC checkpoints
    (tests/test_gpu_parity.py::test_prefill_matches_oracle): the input of layer 0 is an int8 embedding row times its scale times
    a bf16-rounded norm weight, so x / scale lands on exact .5 ties -- and one ulp in the RMSNorm factor (what a differently
    ordered sum of squares costs) decides them.  Shown here with the oracle alone."""
    m = npf.NpModel(ckpt("small", 64, 3))
    gs = m.gs
    eq, es = m.embed
    rng = np.random.default_rng(201)
    near_tie = changed = 0
    toks = rng.integers(0, m.vocab, 200).tolist()
    for t in toks:
        x = (eq[t * m.dim:(t + 1) * m.dim].astype(np.float32).reshape(-1, gs) * es[t * m.dim // gs:(t + 1) * m.dim // gs, None]).reshape(-1)
        y = orc.rmsnorm(x, m.rms_att[0])
        g = y.reshape(-1, gs)
        q = (g / (np.abs(g).max(axis=1, keepdims=True) / np.float32(127.0))).astype(np.float64)
        near_tie += bool((np.abs(np.abs(q - np.floor(q)) - 0.5) < 1e-6).any())
        q1, _ = orc.quantize(y, gs)
        q2, _ = orc.quantize((y * np.float32(1 + 1.2e-7)).astype(np.float32), gs)
        changed += not np.array_equal(q1, q2)
    assert near_tie >= 0.3 * len(toks) and changed >= 0.1 * len(toks), (near_tie, changed)
