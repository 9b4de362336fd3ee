"""BASELINE.json config 5: Qwen3-8B decode with a long KV cache (split-K GQA attention) and the group-size sweep.
The cache is pre-filled with synthetic N(0,1) K/V rows (pure-bandwidth number); decode steps are
timed at pos ~ ctx."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import synth, transformer as T

model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-8b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
gss = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64]
steps = 16
shape = synth.SHAPES[model]
peak = bench.measured_peak_gbs()[0]
for gs in gss:
    m = T.TransformerBuilder.new(bench.bench_checkpoint(model, gs)).with_ctx_length(ctx + steps + 8).build()
    c = m.get_config()
    kvd = c.n_kv_heads * c.head_dim
    rng = np.random.default_rng(0)
    blk = 4096
    for l in range(c.n_layers):
        for p0 in range(0, ctx, blk):
            n = min(blk, ctx - p0)
            kv = rng.standard_normal((2, n, kvd)).astype(np.float32)
            m.kv_write(l, p0, kv[0], kv[1])
    for pos0, label in ((8, "short"), (ctx, "long")):
        m.bench_decode(1, pos0, 4)
        ms = m.bench_decode(1, pos0, steps) / steps
        b = shape.bytes_per_token(gs, pos0 + steps // 2)
        bench.emit({"model": model, "group_size": gs, "context": label, "pos": pos0, "us_per_token": ms * 1e3,
                    "tok_s": 1e3 / ms, "bytes_per_token_GB": b / 1e9, "achieved_GBps": b / ms / 1e6,
                    "frac_of_measured_peak": b / ms / 1e6 / peak, "frac_of_8TBps": b / ms / 1e6 / 8000})
    m.close()
