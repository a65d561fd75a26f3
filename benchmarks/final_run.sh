#!/bin/bash
# One gpurun call that refreshes the judged artefacts (copied from gpurun_out/ into profiles/ afterwards):
#   bench line + reference arm, per-config stage times, sort phase clocks, ncu launch list of the bench command,
#   ncu --set full of three frames (benchmarks/make_profiles.sh turns it into the summaries) and of the training kernels
set -x
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_bwd_pair -c 1 -f -o gpurun_out/r02_bwd_pair python benchmarks/bwd_time.py > gpurun_out/ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_pair_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02_fwd_train python benchmarks/bwd_time.py > gpurun_out/ncu_train2.log 2>&1
if [ "$1" = "full" ]; then
for c in config1_1k_256 config2_100k_1080p config3_1m_1080p config4_3m_1080p config5_6m_4k; do python benchmarks/stage_probe.py $c 2>/dev/null | tail -1; done > gpurun_out/r02_all_configs_1gpu.jsonl
python benchmarks/stage_probe.py config3_1m_1080p gsplat 2>/dev/null | tail -1 >> gpurun_out/r02_all_configs_1gpu.jsonl
python benchmarks/stage_probe.py config3_1m_1080p gsplat packed 2>/dev/null | tail -1 >> gpurun_out/r02_all_configs_1gpu.jsonl
BSPLAT_LIB=$PWD/mojosplat_b200/csrc/ab/libbsplat_phases.so python benchmarks/sort_phases.py > gpurun_out/r02_sort_phases.txt 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_frame python benchmarks/one_frame.py > gpurun_out/ncu_frame.log 2>&1
fi
tail -c 300 gpurun_out/r02_bench_final.json
