"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): both decode paths, exact mode, the batched prefill
(tcgen05 GEMM + attention), the device sampler and the exporter quantiser on the small golden shapes.

    compute-sanitizer --tool memcheck  python scripts/sanitize_target.py [shape ...]
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py micro
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qwen3_rs_b200 import synth, transformer as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shapes = sys.argv[1:] or ["micro", "tiny-untied"]
for name in shapes:
    if name == "micro":
        path = os.path.join(ROOT, "tests", "golden", "micro_gs32.bin")
    else:
        path = f"/tmp/q3_sanitize_{name}.bin"
        if not os.path.exists(path):
            synth.export_synthetic(synth.SHAPES[name], path, 64, seed=1)
    m = T.TransformerBuilder.new(path).build()
    c = m.get_config()
    for dp in (1, 0):
        try:
            m.set_decode_path(dp)
        except T.Q3Error as e:
            print(name, "decode path", dp, "unavailable:", e.message)
            continue
        m.reset()
        toks = m.decode_greedy(3, 0, 6)
        lg = m.forward(toks[-1], 6)
        print(name, "path", dp, "greedy", toks, "logit0 %.4f" % lg[0])
    try:
        m.set_decode_path(1)
    except T.Q3Error:
        pass
    m.set_exact(True)
    m.reset()
    print(name, "exact", m.decode_greedy(3, 0, 3))
    m.set_exact(False)
    m.reset()
    lgp = m.prefill(list(range(1, 20)), 0)
    print(name, "prefill logit0 %.4f" % lgp[0])
    m.sampler_set(0.8, 0.9, 7)
    print(name, "sampled", m.decode_sample(3, 19, 4))
    m.close()
rng = np.random.default_rng(0)
w = rng.standard_normal(64 * 64).astype(np.float32)
q, s, _ = T.op_quantize_q80(w, 64)
print("quantize_q80 ok", int(np.abs(q).max()))
xq = rng.integers(-127, 128, (130, 256), dtype=np.int8)
xs = rng.random((130, 4)).astype(np.float32)
wq = rng.integers(-127, 128, 128 * 256, dtype=np.int8)
ws = rng.random(128 * 4).astype(np.float32)
for ex in (True, False):
    out = T.op_gemm_q8(xq, xs, wq, ws, 130, 128, 256, 64, exact=ex)
    print("gemm_q8 exact=%d ok %.3f" % (ex, float(np.abs(out).max())))
print("SANITIZE_TARGET_DONE")
