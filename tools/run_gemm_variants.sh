# decoder token-step GEMM shapes, weights rotated through > L2 worth of copies (HBM-streaming, like the real decoder)
for shape in "96 1280 1280 64" "96 3840 1280 24" "96 5120 1280 16" "96 1280 5120 16"; do
  set -- $shape
  for v in 0 3 4 5 7; do
    echo -n "variant=$v "; ROTATE=$4 WLK_GEMM_VARIANT=$v GELU=0 timeout 60 python tools/bench_gemm.py $1 $2 $3 256 2>&1 | tail -1
  done
done
