# round-2 evidence pass (run on the GPU box, 1 GPU): captures for the kernels the main loop's captures do not cover, the reference's own
# kernels on the same input, the launch list of the default bench command, and the cross-scale NMS divergence count.
# gpurun brings back at most 64 MiB: the two many-kernel reports are summarised ON THE BOX (text) and deleted; only the five main
# kernels' .ncu-rep files travel.
T=${NCU_TAG:-r02}
OURS='k_compact|k_stereo_outlier|k_nms_ms_dense|k_nms_ms_buckets|k_repitch|k_frame_grid|k_sbp_match|k_sbp_finish|k_project_points|k_hamming_pairs|k_in_frustum|k_frame_view|k_pack|k_gather_pack|k_cvt_gray|k_mappool'
ncu --set full --clock-control none -k "regex:$OURS" -c 40 -f -o /tmp/prof_minor python tools/exercise_kernels.py --reps 1 > gpurun_out/exercise.log 2>&1
python tools/ncu_multi_summary.py /tmp/prof_minor.ncu-rep > gpurun_out/${T}_ncu_minor_kernels.txt 2>&1
# the reference's own live kernels (oracle/_ref/libjsref.so = its src/cuda compiled unmodified) on the same C2 pair: first launch of each (level 0)
: > gpurun_out/${T}_ncu_reference_kernels.txt
for k in imresize_GPU_pitched FASTComputeScoreGPU_patternSize_16_lookup_mask Tile_unrolling_reduction_kernel_v2 FASTComputeOrientationGPU imgaussian_GPU ORB_compute_descriptorGPU ORB_copy_output_GPU ORBGetDistanceStereoGPU Compute_L1_distance_GPU; do
  ncu --set full --clock-control none -k "regex:^$k" -c 1 -f -o /tmp/prof_ref_$k python tools/exercise_kernels.py --ref --reps 1 > gpurun_out/exercise_ref.log 2>&1
  python tools/ncu_multi_summary.py /tmp/prof_ref_$k.ncu-rep >> gpurun_out/${T}_ncu_reference_kernels.txt 2>&1
done
for k in k_pyramid k_fast_cells k_blur k_orient_desc k_stereo_match; do ncu --set full --clock-control none --import-source on -k regex:^$k -s 2 -c 1 -f -o gpurun_out/prof_${T}_$k python bench.py --pairs 32 --steps 2 --warmup 1 --no-ref-cuda --no-parity > /dev/null 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --pairs 32 --no-ref-cuda --no-parity > gpurun_out/launches_bench.log 2>&1
python tools/nms_ms_divergence.py 24 > gpurun_out/nms_ms_divergence.json 2> gpurun_out/nms_ms_divergence.err
du -sh gpurun_out; ls gpurun_out | head -30; tail -n 2 gpurun_out/exercise.log; tail -n 2 gpurun_out/exercise_ref.log; grep -c kernel: gpurun_out/${T}_ncu_minor_kernels.txt gpurun_out/${T}_ncu_reference_kernels.txt
