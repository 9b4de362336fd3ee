import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import synth, transformer as T
path = bench.bench_checkpoint(sys.argv[1] if len(sys.argv) > 1 else "qwen3-8b", 64)
m = T.TransformerBuilder.new(path).with_ctx_length(256).build()
print(m.forward_argmax(1, 0), m.forward_argmax(2, 1))
