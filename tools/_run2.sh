python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; python tools/bench_brief.py gpurun_out/bench_v10.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 2 --warmup 1 --pairs 32 --no-ref-cuda > gpurun_out/launches_bench.log 2>&1
for k in k_blur k_blur_fix; do ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -f -o gpurun_out/prof_v10_$k python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda > /dev/null 2>&1; done
ls gpurun_out/prof_v10_* | wc -l
