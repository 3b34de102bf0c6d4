#!/bin/bash
# the four GEMMs of one encoder layer at 96 streams, old band heuristic (WLK_GEMM_BAND=0) vs the default
cd "$(dirname "$0")/.."
for band in 0 1 2 4; do
  echo "== WLK_GEMM_BAND=$band"
  export WLK_GEMM_BAND=$band
  OUT=bf16 GELU=0 python tools/bench_gemm.py 144000 3840 1280 8 | tail -1
  OUT=f32  GELU=2 python tools/bench_gemm.py 144000 1280 1280 8 | tail -1
  OUT=bf16 GELU=1 python tools/bench_gemm.py 144000 5120 1280 8 | tail -1
  OUT=f32  GELU=2 python tools/bench_gemm.py 144000 1280 5120 8 | tail -1
done
