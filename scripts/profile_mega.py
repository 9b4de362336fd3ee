"""Phase-level timeline of the persistent decode kernel from in-kernel clock64 stamps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import transformer as T

model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-8b"
pos = int(sys.argv[2]) if len(sys.argv) > 2 else 64
GHZ = float(sys.argv[3]) if len(sys.argv) > 3 else 1.965
path = bench.bench_checkpoint(model, 64)
m = T.TransformerBuilder.new(path).with_ctx_length(max(256, pos + 8)).build()
for p in range(4):
    m.forward_argmax(1, p)
t = m.debug_profile(1, pos).astype(np.int64)
L = m.get_config().n_layers
names = ["qkv_pro", "qkv_gemv", "qkv_bar", "att", "att_-", "att_bar", "o_pro", "o_gemv", "o_bar", "gu_pro", "gu_gemv", "gu_bar",
         "dn_pro", "dn_gemv", "dn_bar"]
E = len(names)
ev = t[:, 1:1 + E * L].reshape(t.shape[0], L, E)
start = np.concatenate([t[:, :1], ev[:, :-1, -1]], axis=1)[:, :, None]        # [cta, L, 1] end of previous layer
d = np.diff(np.concatenate([start, ev], axis=2), axis=2) / (GHZ * 1e3)         # us, [cta, L, 14]
tot = (t[:, 1 + E * L + 2] - t[:, 0]) / (GHZ * 1e3)
print(f"{model} pos {pos}: kernel {tot.mean():.1f} us (per CTA mean); per layer {d[:, 1:-1].sum(axis=2).mean():.2f} us")
print("phase      mean_us  [min..max over CTAs of the per-CTA mean]   layer-1 only")
for i, n in enumerate(names):
    x = d[:, 2:-1, i].mean(axis=1)
    print(f"{n:9s} {x.mean():7.2f}   [{x.min():6.2f} .. {x.max():6.2f}]   {d[:, 1, i].mean():7.2f}")
h = np.diff(np.concatenate([ev[:, -1, -1:], t[:, 1 + E * L: 1 + E * L + 3]], axis=1), axis=1) / (GHZ * 1e3)
print("head: prologue %.1f gemv %.1f barrier %.1f us" % tuple(h.mean(0)))
np.save("gpurun_out/mega_profile_%s_pos%d.npy" % (model, pos), t)
