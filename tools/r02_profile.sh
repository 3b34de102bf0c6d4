set -x
cd $GRAFT_REPO_ROOT
export WLK_NCU=1
for k in layernorm_kernel attn_tc_kernel gemm_tc2_kernel; do
  c=2; [ $k = gemm_tc2_kernel ] && c=6
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c $c -f -o gpurun_out/r02_${k}_96streams python bench.py --streams 96 --warmup 3 --no-seam --no-extras --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
  tail -2 gpurun_out/ncu_$k.log
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_96streams.csv python bench.py --streams 96 --warmup 3 --no-seam --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
ls -la gpurun_out/
