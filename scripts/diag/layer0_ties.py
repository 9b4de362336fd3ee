"""Why layer 0 of the synthetic checkpoints flips so many int8 activations in fast mode: how many embedding rows, after RMSNorm, hold an
element whose x / scale sits on an exact .5 tie, and how many rows change when the normalisation factor moves by one ulp.
  python scripts/diag/layer0_ties.py small 64 3 200"""
import sys, os
import numpy as np
sys.path.insert(0, "/root/repo")
from oracle import np_forward as npf
from oracle import binding as orc
name, gs, seed, Tn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
path = f"/tmp/diag_{name}_{gs}_{seed}.bin"
if not os.path.exists(path):
    from qwen3_rs_b200 import synth
    synth.export_synthetic(synth.SHAPES[name], path, gs, seed=seed)
m = npf.NpModel(path)
attrs = [a for a in dir(m) if not a.startswith("_")]
toks = np.random.default_rng(Tn + 1).integers(0, m.vocab, Tn).tolist()
near = {1e-6: 0, 1e-5: 0, 1e-4: 0}
flip = 0
for t in toks:
    eq, es = m.embed
    x = (eq[t * m.dim:(t + 1) * m.dim].astype(np.float32).reshape(-1, gs) * es[t * m.dim // gs:(t + 1) * m.dim // gs, None]).reshape(-1)
    y = orc.rmsnorm(np.asarray(x, np.float32), m.rms_att[0])
    g = y.reshape(-1, gs).astype(np.float32)
    sc = (np.abs(g).max(axis=1, keepdims=True) / np.float32(127.0)).astype(np.float32)
    q = (g / sc).astype(np.float32).astype(np.float64)
    d = np.abs(np.abs(q - np.floor(q)) - 0.5)
    for k in near:
        if (d < k).any(): near[k] += 1
    # what a 1-ulp different normalisation factor does
    y2 = (y * np.float32(1 + 1.2e-7)).astype(np.float32)
    q1, s1 = orc.quantize(y, gs)
    q2, s2 = orc.quantize(y2, gs)
    if not np.array_equal(q1, q2): flip += 1
print(name, gs, "rows with an element within {1e-6,1e-5,1e-4} of a .5 tie:", near, "of", Tn, "; rows whose int8 change when the norm factor moves by 1 ulp:", flip)
