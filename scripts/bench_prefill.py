"""BASELINE.json config 3: 2048-token prefill (tcgen05 int8 GEMM) + 256-token decode, one GPU."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import synth, transformer as T

model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-4b"
Tn = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ndec = int(sys.argv[3]) if len(sys.argv) > 3 else 256
gs = 64
shape = synth.SHAPES[model]
m = T.TransformerBuilder.new(bench.bench_checkpoint(model, gs)).with_ctx_length(Tn + ndec + 8).build()
rng = np.random.default_rng(0)
toks = rng.integers(0, shape.vocab_size, Tn).tolist()
ms = min(m.bench_prefill(toks, 0) for _ in range(3))
ah, kv = shape.n_heads * shape.head_dim, shape.n_kv_heads * shape.head_dim
per_tok = shape.n_layers * (2 * shape.dim * ah + 2 * shape.dim * kv + 3 * shape.dim * shape.hidden_dim)
ops = 2.0 * Tn * per_tok
t0 = time.perf_counter(); lg = m.prefill(toks, 0); e2e_ms = (time.perf_counter() - t0) * 1e3
dec_ms = m.bench_decode(int(np.argmax(lg)), Tn, ndec)
out = {"model": model, "group_size": gs, "prefill_tokens": Tn, "prefill_ms": ms, "prefill_tok_s": Tn / ms * 1e3,
       "gemm_int8_TOPS": ops / ms / 1e9, "frac_of_4.5_POPS": ops / ms / 1e9 / 4500.0,
       "prefill_e2e_ms_incl_logits_d2h": e2e_ms, "decode_tokens": ndec, "decode_tok_s": ndec / dec_ms * 1e3,
       "decode_us_per_token": dec_ms / ndec * 1e3,
       "note": "prefill time includes batched norm/quantize, QK-norm+RoPE and f32 causal attention, not only the GEMMs"}
bench.emit(out)
