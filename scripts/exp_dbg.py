"""Timing experiments on the persistent decode kernel (env Q3_MEGA_DBG bits; outputs are garbage, only the clock matters):
1 = weights always read from the same L2-resident bytes, 2 = no grid barriers, 4 = no prologues."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import bench
    from qwen3_rs_b200 import transformer as T
    model = sys.argv[2]
    path = bench.bench_checkpoint(model, 64)
    m = T.TransformerBuilder.new(path).with_ctx_length(256).build()
    m.bench_decode(1, 0, 8)
    m.reset()
    ms = min(m.bench_decode(1, 0, 32) for _ in range(3))
    sys.stderr.write("RESULT dbg=%s %s: %.1f us/token\n" % (os.environ.get("Q3_MEGA_DBG", "0"), model, ms * 1000 / 32))
else:
    model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-8b"
    for dbg in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "4", "3", "6", "7"]):
        env = dict(os.environ, Q3_MEGA_DBG=dbg)
        r = subprocess.run([sys.executable, __file__, "child", model], env=env, capture_output=True, text=True, timeout=300)
        lines = [l for l in r.stderr.splitlines() if l.startswith("RESULT")]
        print(lines[-1] if lines else "dbg=%s FAILED: %s" % (dbg, r.stderr[-400:]), flush=True)
