"""A/B of prefill variants on one box: python scripts/ab_prefill.py [model] [T]  (variants = qwen3_rs_b200/lib/variant_*.so + the product library)"""
import os, sys, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import bench
    from qwen3_rs_b200 import synth, transformer as T
    model, Tn = sys.argv[2], int(sys.argv[3])
    m = T.TransformerBuilder.new(bench.bench_checkpoint(model, 64)).with_ctx_length(Tn + 8).build()
    toks = np.random.default_rng(0).integers(0, synth.SHAPES[model].vocab_size, Tn).tolist()
    ms = min(m.bench_prefill(toks, 0) for _ in range(3))
    sys.stderr.write("RESULT %.2f\n" % ms)
else:
    model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-4b"
    Tn = sys.argv[2] if len(sys.argv) > 2 else "2048"
    libs = [os.path.join(ROOT, "qwen3_rs_b200", "lib", "libqwen3cuda.so")] + sorted(glob.glob(os.path.join(ROOT, "qwen3_rs_b200", "lib", "variant_pf*.so")))
    for rep in range(2):
        for l in libs:
            r = subprocess.run([sys.executable, __file__, "child", model, Tn], env=dict(os.environ, Q3_LIB=l), capture_output=True, text=True, timeout=300)
            v = [x for x in r.stderr.splitlines() if x.startswith("RESULT")]
            print("%-28s %s ms (prefill %s T=%s)" % (os.path.basename(l), v[-1].split()[1] if v else "nan " + r.stderr[-200:], model, Tn), flush=True)
