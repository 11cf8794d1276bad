# usage (on the GPU box): LIBS="name ..." KERNEL=k_fast_cells bash tools/gpu_ncu_quick.sh
# a handful of ncu counters (no --set full) of one kernel for the in-tree library and variants/libjsfe_<name>.so
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__cycles_elapsed.avg.per_second
cp jetson_slam_b200/libjsfe.so /tmp/base.so
for lib in base $LIBS; do
  [ $lib = base ] || cp variants/libjsfe_$lib.so jetson_slam_b200/libjsfe.so
  echo "== $lib"
  ncu --metrics $M --clock-control none -k regex:^${KERNEL:-k_fast_cells} -s 2 -c 1 python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda --no-gather --no-parity 2>/dev/null | grep -E "^\s+(gpu__|smsp__|l1tex__|sm__)" | sed 's/  */ /g'
done
cp /tmp/base.so jetson_slam_b200/libjsfe.so
