"""Which reduction causes the fast mode's int8 flips?  

The multi-kernel path can evaluate each float reduction either as a parallel tree (fast mode, same arithmetic as the
persistent kernel) or in the reference's left-fold order, one switch per reduction (q3_set_exact_mask):
    1 RMSNorm sum of squares   2 GEMV group fold   4 QK-norm sum of squares   8 attention (dots, softmax, value mix)   16 SwiGLU expf
For every mask of interest the model is run teacher-forced along the oracle's greedy tokens and compared with the oracle:
   * flips/token : int8 activations that differ from the oracle's at the first quantiser that sees a difference is not observable
                   from outside, so the proxy is the number of (position, layer) residual-stream rows whose max-abs error exceeds
                   1e-3 of the row scale when every layer is fed the ORACLE's input (teacher-forced layerwise, no cascade);
   * end-to-end  : free-running max |dlogit| against the oracle (cascade included), and greedy-token agreement.
Also times each mask (tokens/s through Transformer.forward) -- what ordering a reduction costs.

    python scripts/attribute_flips.py [model=qwen3-0.6b] [tokens=24]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import binding as orc
from qwen3_rs_b200 import transformer as T
from qwen3_rs_b200.sampler import argmax_last

model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-0.6b"
ntok = int(sys.argv[2]) if len(sys.argv) > 2 else 24
path = bench.bench_checkpoint(model, 64)
orc.set_threads(os.cpu_count() or 1)
o = orc.Model(path, 256)
m = T.TransformerBuilder.new(path).with_ctx_length(256).build()
c = o.config
L, kv = c["n_layers"], c["n_kv_heads"] * c["head_dim"]

# oracle trajectory: tokens, logits, residual stream at every layer boundary, KV cache
o.reset()
xd = o.dump_residuals()
seq, ologits, xdumps = [1], [], []
for pos in range(ntok):
    lg = o.forward(seq[pos], pos)
    ologits.append(lg)
    xdumps.append(xd.copy())
    seq.append(orc.argmax(lg))
ko, vo = o.kv_cache()
ko, vo = ko.copy(), vo.copy()

NAMES = {0: "fast (all trees)", 1: "+RMSNorm ordered", 2: "+GEMV fold ordered", 4: "+QK-norm ordered", 8: "+attention ordered", 16: "+SwiGLU glibc expf",
         31 - 1: "all but RMSNorm", 31 - 2: "all but GEMV fold", 31 - 4: "all but QK-norm", 31 - 8: "all but attention", 31 - 16: "all but SwiGLU expf",
         31: "exact (all ordered)"}
print(f"{model} gs64, {ntok} positions, teacher-forced along the oracle's greedy tokens; multi-kernel path (same arithmetic as the persistent kernel)")
print(f"{'mask':>4}  {'reductions in reference order':28s} {'layer rows off >1e-3':>22s} {'layerwise worst':>16s} {'e2e max|dlogit|':>16s} {'median':>8s} {'tokens':>8s} {'tok/s':>8s}")
m.set_decode_path(0)
for mask in (0, 1, 2, 4, 8, 16, 30, 29, 27, 23, 15, 31):
    m.set_exact_mask(mask)
    # (a) layerwise, teacher-forced: every layer gets the oracle's input and cache -> no cascade, a flip shows as one bad row
    off, worst, rows = 0, 0.0, 0
    for pos in range(0, ntok, 3):
        for l in range(L):
            m.kv_write(l, 0, ko[l, :pos + 1].reshape(pos + 1, kv), vo[l, :pos + 1].reshape(pos + 1, kv))
        for l in range(L):
            x = m.forward_layers(xdumps[pos][l], pos, l, l + 1)
            e = float(np.abs(x - xdumps[pos][l + 1]).max()) / max(1.0, float(np.abs(xdumps[pos][l + 1]).max()))
            off += e > 1e-3
            worst = max(worst, e)
            rows += 1
    # (b) free running end to end
    m.reset()
    errs, same = [], 0
    t0 = time.perf_counter()
    outs = [m.forward(seq[p], p) for p in range(ntok)]
    dt = time.perf_counter() - t0
    for p in range(ntok):
        errs.append(float(np.abs(outs[p] - ologits[p]).max()))
        same += argmax_last(outs[p]) == seq[p + 1]
    print(f"{mask:4d}  {NAMES[mask]:28s} {off:10d} / {rows:<9d} {worst:16.2e} {max(errs):16.3e} {np.median(errs):8.1e} {same:4d}/{ntok:<3d} {ntok / dt:8.1f}")
m.set_exact(False)
m.set_decode_path(1)
m.reset()
t0 = time.perf_counter()
outs = [m.forward(seq[p], p) for p in range(ntok)]
dt = time.perf_counter() - t0
errs = [float(np.abs(outs[p] - ologits[p]).max()) for p in range(ntok)]
same = sum(argmax_last(outs[p]) == seq[p + 1] for p in range(ntok))
print(f"   -  {'persistent kernel (fast)':28s} {'':>22s} {'':>16s} {max(errs):16.3e} {np.median(errs):8.1e} {same:4d}/{ntok:<3d} {ntok / dt:8.1f}")
print(f"logit scale: max |logit| {max(float(np.abs(l).max()) for l in ologits):.2f}")
