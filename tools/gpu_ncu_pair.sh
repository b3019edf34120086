# ncu --set full capture of the pair-stage kernels of one mid-run batch (config3) -> gpurun_out/$1.ncu-rep
mkdir -p gpurun_out
B="python bench.py --workload ${WORKLOAD:-config3} --no-cpu --no-parity --derep off"
timeout 900 ncu --set full --clock-control none --import-source on \
  --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,smsp__cycles_active.avg,sm__cycles_elapsed.avg \
  -k regex:"${2:-anchor_kernel|chain_kernel}" -s ${3:-8} -c ${4:-2} -o gpurun_out/$1 $B --steps 1 --warmup 1 > gpurun_out/$1.log 2>&1
tail -3 gpurun_out/$1.log
