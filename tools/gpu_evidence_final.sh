# final evidence pass of a round (run on the GPU box, 1 GPU): full test suite, ncu --set full captures of the five main kernels
# (64-image launches), the launch list of the same command, the default bench line and the reference arm.
T=${NCU_TAG:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for k in k_pyramid k_fast_cells k_blur k_orient_desc k_stereo_match; do ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -f -o gpurun_out/prof_${T}_$k python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda --no-parity > /dev/null 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --pairs 32 --no-ref-cuda --no-parity > gpurun_out/launches_bench.log 2>&1
python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; python tools/bench_brief.py gpurun_out/bench_${T}.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err; tail -c 600 gpurun_out/bench_${T}_ref.json
python tools/latency_stages.py 1 > gpurun_out/latency_stages_${T}.txt 2>&1; cat gpurun_out/latency_stages_${T}.txt
