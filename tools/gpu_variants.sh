# usage (on the GPU box): VARIANTS="name ..." NCU_KERNELS="k_x ..." NCU_TAG=vN bash tools/gpu_variants.sh
# runs the GPU tests, the default-size bench with the in-tree libjsfe.so and with every variants/libjsfe_<name>.so, then optional
# `ncu --set full` captures of the named kernels into gpurun_out/prof_<tag>_<kernel>.ncu-rep (read here with tools/ncu_summary.py)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
cp jetson_slam_b200/libjsfe.so /tmp/base.so
python bench.py --steps 10 --warmup 3 --no-ref-cuda > gpurun_out/b_base.json 2> gpurun_out/b_base.err; echo base; python tools/bench_brief.py gpurun_out/b_base.json
for v in $VARIANTS; do cp variants/libjsfe_$v.so jetson_slam_b200/libjsfe.so; python bench.py --steps 10 --warmup 3 --no-ref-cuda > gpurun_out/b_$v.json 2> gpurun_out/b_$v.err; echo $v; python tools/bench_brief.py gpurun_out/b_$v.json; done
cp /tmp/base.so jetson_slam_b200/libjsfe.so
for k in $NCU_KERNELS; do ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -f -o gpurun_out/prof_${NCU_TAG}_$k python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda > /dev/null 2>&1; done
ls gpurun_out/*.ncu-rep | tail -3
