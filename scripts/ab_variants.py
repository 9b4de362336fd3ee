"""A/B timing of kernel-tuning variants on ONE box: build each variant here (CPU, nvcc), then on the GPU box
`python scripts/ab_variants.py run [model]` times every built variant back to back (interleaved, best of N).

  python scripts/ab_variants.py build name1:DEF1=V,DEF2=V name2: ...     (name 'base' = no defines)
"""
import os, sys, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "qwen3_rs_b200", "lib")

if sys.argv[1] == "build":
    from concurrent.futures import ThreadPoolExecutor
    from qwen3_rs_b200 import build
    for f in glob.glob(os.path.join(VDIR, "variant_*.so")):
        os.remove(f)
    def one(spec):
        name, _, defs = spec.partition(":")
        out = os.path.join(VDIR, "variant_%s.so" % name)
        build.build(defines=[d for d in defs.split(",") if d], out=out)
        return out
    with ThreadPoolExecutor(4) as ex:
        for o in ex.map(one, sys.argv[2:]):
            print("built", o)
elif sys.argv[1] == "child":
    import bench
    from qwen3_rs_b200 import transformer as T
    model, pos0, ntok = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    m = T.TransformerBuilder.new(bench.bench_checkpoint(model, 64)).with_ctx_length(max(256, pos0 + ntok + 8)).build()
    m.bench_decode(1, pos0, 8)
    ms = min(m.bench_decode(1, pos0, ntok) for _ in range(3))
    sys.stderr.write("RESULT %.1f\n" % (ms * 1000 / ntok))
else:
    model = sys.argv[2] if len(sys.argv) > 2 else "qwen3-8b"
    pos0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    ntok = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    libs = sorted(glob.glob(os.path.join(VDIR, "variant_*.so")))
    res = {l: [] for l in libs}
    for rep in range(int(os.environ.get('AB_REPS', '2'))):
        for l in libs:
            r = subprocess.run([sys.executable, __file__, "child", model, str(pos0), str(ntok)], env=dict(os.environ, Q3_LIB=l),
                               capture_output=True, text=True, timeout=300)
            v = [x for x in r.stderr.splitlines() if x.startswith("RESULT")]
            res[l].append(float(v[-1].split()[1]) if v else float("nan"))
            if not v: print(r.stderr[-300:])
    for l in libs:
        print("%-28s %s us/token (pos %d..%d)" % (os.path.basename(l)[8:-3], " ".join("%.1f" % x for x in res[l]), pos0, pos0 + ntok), flush=True)
