export PYTHONUNBUFFERED=1
timeout 300 python scripts/diag/mega_vs_graph.py 2>&1 | tail -8
echo "--- old loop"
Q3_LIB=$PWD/qwen3_rs_b200/lib/variant_attnv1.so timeout 300 python scripts/diag/mega_vs_graph.py 2>&1 | tail -8
