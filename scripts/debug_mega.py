"""Compare the persistent decode kernel (path 1) with the multi-kernel graph (path 0) layer by layer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_rs_b200 import synth, transformer as T

def check(shape, gs, seed=0, ctx=None, positions=(0, 1, 2)):
    path = f"/tmp/dbg_{shape.name}_gs{gs}.bin"
    synth.export_synthetic(shape, path, gs, seed=seed)
    m = T.TransformerBuilder.new(path).with_ctx_length(ctx).build()
    c = m.get_config()
    try:
        m.set_decode_path(1)
    except T.Q3Error as e:
        print(shape.name, gs, "skip:", e.message)
        return
    rng = np.random.default_rng(0)
    worst = {}
    for pos in positions:
        for l in range(c.n_layers):
            x = rng.standard_normal(c.dim).astype(np.float32)
            m.set_decode_path(0); a = m.forward_layers(x, pos, l, l + 1)
            m.set_decode_path(1); b = m.forward_layers(x, pos, l, l + 1)
            worst[(pos, l)] = float(np.abs(a - b).max())
        x = rng.standard_normal(c.dim).astype(np.float32)
        m.set_decode_path(0); _, la = m.forward_layers(x, pos, 0, 0, True)
        m.set_decode_path(1); _, lb = m.forward_layers(x, pos, 0, 0, True)
        worst[(pos, 'head')] = float(np.abs(la - lb).max())
    m.reset(); m.set_decode_path(0); t0 = m.decode_greedy(3, 0, 8)
    m.reset(); m.set_decode_path(1); t1 = m.decode_greedy(3, 0, 8)
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    print(f"{shape.name} gs{gs}: max layer diff {max(worst.values()):.3e}; bad {bad}; greedy same {t0 == t1}")
    m.close()

S = synth.Shape
check(S("b", 512, 768, 2, 8, 4, vocab_size=1024, max_seq_len=64), 128)
check(S("c", 1024, 3072, 2, 16, 8, vocab_size=8192, max_seq_len=64), 64)
check(S("d", 1024, 6144, 2, 16, 8, vocab_size=8192, max_seq_len=64), 64)   # down: n_kt = 2
check(S("e", 2560, 9728, 2, 32, 8, vocab_size=8192, max_seq_len=64), 64)   # 4B dims
check(S("f", 4096, 12288, 1, 32, 8, vocab_size=16384, max_seq_len=300), 64, positions=(0, 5, 299))  # 8B dims
