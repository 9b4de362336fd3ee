"""Small driver for ncu captures: a few decode steps of the bench model on the chosen path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import transformer as T
model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-8b"
path_id = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
pos0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m = T.TransformerBuilder.new(bench.bench_checkpoint(model, 64)).with_ctx_length(max(256, pos0 + steps + 8)).build()
m.set_decode_path(path_id)
print(m.decode_greedy(1, pos0, steps))
