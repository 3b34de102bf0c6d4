WLK_ATTN_V5=1 timeout 120 python tools/attn_diag.py 2>&1 | grep -E "max_err|bad rows" | head -6
for v in 0 1 0 1; do WLK_ATTN_V5=$v timeout 100 python tools/bench_attn.py 16 20 2>&1 | tail -1; done
