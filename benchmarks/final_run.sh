set -x
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_ref.err
for c in config1_1k_256 config2_100k_1080p config3_1m_1080p config4_3m_1080p config5_6m_4k; do python benchmarks/stage_probe.py $c 2>/dev/null | tail -1; done > gpurun_out/r02_all_configs_1gpu.jsonl
python benchmarks/stage_probe.py config3_1m_1080p gsplat 2>/dev/null | tail -1 >> gpurun_out/r02_all_configs_1gpu.jsonl
python benchmarks/stage_probe.py config3_1m_1080p gsplat packed 2>/dev/null | tail -1 >> gpurun_out/r02_all_configs_1gpu.jsonl
BSPLAT_LIB=$PWD/mojosplat_b200/csrc/ab/libbsplat_phases.so python benchmarks/sort_phases.py > gpurun_out/r02_sort_phases.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_frame python benchmarks/one_frame.py > gpurun_out/ncu_frame.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_bwd_pair|raster_pair_kernel" -c 2 -f -o gpurun_out/r02_train python scratch/bwd_time.py > gpurun_out/ncu_train.log 2>&1
tail -c 600 gpurun_out/r02_bench_final.json
