import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import synth, transformer as T
model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-4b"
Tn = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
m = T.TransformerBuilder.new(bench.bench_checkpoint(model, 64)).with_ctx_length(Tn + 8).build()
toks = np.random.default_rng(0).integers(0, synth.SHAPES[model].vocab_size, Tn).tolist()
print(m.bench_prefill(toks, 0))
