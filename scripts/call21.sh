export PYTHONUNBUFFERED=1
cp qwen3_rs_b200/lib/libqwen3cuda.so qwen3_rs_b200/lib/variant_base.so
for d in 0 16 32 64 48 112; do echo "dbg $d"; Q3_MEGA_DBG=$d AB_REPS=1 timeout 600 python scripts/ab_variants.py run qwen3-8b 32768 16 2>&1 | tail -1; done
