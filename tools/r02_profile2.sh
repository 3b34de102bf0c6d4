#!/bin/bash
# round-2 follow-up captures: the two-query-tile attention kernel inside the bench tick, the Sortformer attention and
# speaker-cache update kernels inside tools/bench_sortformer.py
cd "$(dirname "$0")/.."
export WLK_NCU=1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_tc2_kernel -c 2 -f \
  -o gpurun_out/r02_attn_tc2_kernel_96streams python bench.py --streams 96 --warmup 3 --no-seam --no-extras --no-cpu-baseline > gpurun_out/ncu_attn2.log 2>&1
tail -1 gpurun_out/ncu_attn2.log
unset WLK_NCU
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sf_attention_mma_kernel --launch-skip 400 -c 2 -f \
  -o gpurun_out/r02_sf_attention_mma_kernel_16streams python tools/bench_sortformer.py 16 12 > gpurun_out/ncu_sf1.log 2>&1
tail -1 gpurun_out/ncu_sf1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sf_update_kernel --launch-skip 9 -c 2 -f \
  -o gpurun_out/r02_sf_update_kernel_16streams python tools/bench_sortformer.py 16 12 > gpurun_out/ncu_sf2.log 2>&1
tail -1 gpurun_out/ncu_sf2.log
ls -la gpurun_out/*.ncu-rep
