"""Persistent kernel vs multi-kernel graph path, layer by layer on identical inputs: max |difference| relative to the row's scale."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qwen3_rs_b200 import synth, transformer as T
for name, gs, seed in (("micro", 32, 7), ("tiny", 64, 0), ("tiny-untied", 64, 1), ("small", 128, 2), ("small8", 64, 5)):
    path = f"/tmp/diag_{name}_{gs}_{seed}.bin"
    if not os.path.exists(path):
        synth.export_synthetic(synth.SHAPES[name], path, gs, seed=seed)
    m = T.TransformerBuilder.new(path).build()
    c = m.get_config()
    rng = np.random.default_rng(1)
    out = []
    for pos in (0, 3, 40, 100):
        if pos >= c.seq_len: continue
        for l in range(c.n_layers):
            x = rng.standard_normal(c.dim).astype(np.float32)
            m.set_decode_path(0); a = m.forward_layers(x, pos, l, l + 1)
            m.set_decode_path(1); b = m.forward_layers(x, pos, l, l + 1)
            out.append("%d/%d:%.1e" % (pos, l, np.abs(a - b).max() / max(1.0, np.abs(a).max())))
    print(name, gs, "kv_mul", c.n_heads // c.n_kv_heads, " ".join(out))
