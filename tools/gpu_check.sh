# quick GPU check while tuning (run on the GPU box): GPU tests, a short bench line, optional ncu captures of the named kernels
# usage: NCU_KERNELS="k_x ..." NCU_TAG=r02x bash tools/gpu_check.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time python bench.py --steps ${STEPS:-20} --warmup 5 --no-ref-cuda > gpurun_out/b_${NCU_TAG:-chk}.json 2> gpurun_out/b_${NCU_TAG:-chk}.err ) 2>&1 | grep real
tail -3 gpurun_out/b_${NCU_TAG:-chk}.err
python tools/bench_brief.py gpurun_out/b_${NCU_TAG:-chk}.json
for k in $NCU_KERNELS; do ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -f -o gpurun_out/prof_${NCU_TAG}_$k python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda --no-parity > /dev/null 2>&1; done
ls gpurun_out/*.ncu-rep 2>/dev/null | tail -3
