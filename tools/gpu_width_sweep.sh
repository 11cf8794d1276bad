# usage (on the GPU box): VAR=JSFE_FAST_WIDTH VALUES="192 176 ..." LIBS="name ..." bash tools/gpu_width_sweep.sh
# sweep of one k_fast_cells geometry knob (an environment variable read by jsfe_create) for the in-tree library and
# variants/libjsfe_<name>.so; prints the per-kernel times of a short bench run each
VAR=${VAR:-JSFE_FAST_WIDTH}
cp jetson_slam_b200/libjsfe.so /tmp/base.so
for lib in base $LIBS; do
  [ $lib = base ] || cp variants/libjsfe_$lib.so jetson_slam_b200/libjsfe.so
  for w in $VALUES; do
    env $VAR=$w python bench.py --steps 10 --warmup 3 --no-ref-cuda --no-gather > gpurun_out/b_${lib}_$w.json 2> gpurun_out/b_${lib}_$w.err
    echo "$lib $VAR=$w: $(python tools/bench_brief.py gpurun_out/b_${lib}_$w.json | sed -n 2p | cut -c1-60) $(grep -c mismatches.:.0 gpurun_out/b_${lib}_$w.json)"
  done
done
cp /tmp/base.so jetson_slam_b200/libjsfe.so
