set -x
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-parity --derep off"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $B --steps 2 --warmup 1 > gpurun_out/r2_launches.log 2>&1
SKB_TRACE=1 timeout 300 $B --steps 2 --warmup 3 > gpurun_out/s2_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'anchor_kernel|chain_kernel|finalize_warp_kernel|cand_pack_kernel|task_setup_kernel|screen_runs_kernel' -s 12 -c 6 -o gpurun_out/r2_final $B --steps 1 --warmup 1 > gpurun_out/s2_ncu.log 2>&1
ls -la gpurun_out
