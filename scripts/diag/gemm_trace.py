"""In-kernel pipeline trace of the prefill GEMM (MMA warp of CTA 0): needs a library built with -DPF_TRACE=1
(python scripts/ab_variants.py build trace:PF_TRACE=1; Q3_LIB=qwen3_rs_b200/lib/variant_trace.so python scripts/diag/gemm_trace.py [mode]).
The epilogue-side stamps used for profiles/r02_gemm_q8_ceilings.txt were removed again with the drain-loop rewrite."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["Q3_PF_TRACE"] = "1"
from qwen3_rs_b200 import transformer as T
for mode in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]:
    ms = T.bench_gemm_q8(2048, 24576, 4096, 64, mode, 1)
    sys.stderr.write("mode %d: %.1f us\n" % (mode, ms * 1e3))
