# final evidence pass: tests, launch list (gpu__time_duration), full captures of every kernel, default bench + reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${NCU_TAG}.csv python bench.py --steps 2 --warmup 1 --pairs 32 --no-ref-cuda > gpurun_out/launches_bench.log 2>&1
for k in k_pyramid k_fast_cells k_blur k_orient_desc k_stereo_match; do ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -f -o gpurun_out/prof_${NCU_TAG}_$k python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda > /dev/null 2>&1; done
ls gpurun_out/prof_${NCU_TAG}_* | wc -l
python bench.py > gpurun_out/bench_${NCU_TAG}.json 2> gpurun_out/bench_${NCU_TAG}.err; python tools/bench_brief.py gpurun_out/bench_${NCU_TAG}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${NCU_TAG}_ref.json 2> gpurun_out/bench_${NCU_TAG}_ref.err; tail -c 600 gpurun_out/bench_${NCU_TAG}_ref.json
