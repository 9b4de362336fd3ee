"""Layer-0 K/V rows of q3_prefill against the oracle (default: small gs128 seed 2, T=130; or: name gs seed T): how many rows are off, by how much (Q3_LIB picks the library)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as orc
from qwen3_rs_b200 import synth, transformer as T
name, gs, seed, Tn = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("small", 128, 2, 130)
path = f"/tmp/diag_{name}_{gs}_{seed}.bin"
if not os.path.exists(path):
    synth.export_synthetic(synth.SHAPES[name], path, gs, seed=seed)
m, o = T.TransformerBuilder.new(path).build(), orc.Model(path)
c = o.config
toks = np.random.default_rng(Tn + 1).integers(0, c["vocab_size"], Tn).tolist()
o.reset()
for p, t in enumerate(toks):
    o.forward(t, p)
ko, vo = o.kv_cache()
m.reset()
m.prefill(toks, 0)
kvd = c["n_kv_heads"] * c["head_dim"]
k, v = m.kv_read(0, 0, Tn)
ek = np.abs(k - ko[0, :Tn].reshape(Tn, kvd)).max(axis=1)
ev = np.abs(v - vo[0, :Tn].reshape(Tn, kvd)).max(axis=1)
print(os.environ.get("Q3_LIB", "product library"))
print("K rows > 1e-4: %d / %d; V rows > 1e-4: %d" % ((ek > 1e-4).sum(), Tn, (ev > 1e-4).sum()))
print("sorted K row errors:", np.array2string(np.sort(ek)[::-1][:40], precision=2))
# sequential decode path on the same tokens (same norm/quant arithmetic, GEMV instead of GEMM)
m.reset()
for p, t in enumerate(toks):
    m.forward(t, p)
k2, v2 = m.kv_read(0, 0, Tn)
ek2 = np.abs(k2 - ko[0, :Tn].reshape(Tn, kvd)).max(axis=1)
print("decode path: K rows > 1e-4: %d; rows off in both: %d" % ((ek2 > 1e-4).sum(), ((ek2 > 1e-4) & (ek > 1e-4)).sum()))
