#!/bin/bash
# usage: sweep.sh N
N=$1
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
$TR benchmarks/bench_configs.py --config 4 2>gpurun_out/cfg4_n$N.err | grep "^{" > gpurun_out/cfg4_n$N.json
for ex in nccl p2p; do
  if [ "$N" = "1" ] && [ "$ex" = "p2p" ]; then continue; fi
  $TR benchmarks/bench_configs.py --config 5 --exchange $ex 2>gpurun_out/cfg5_n${N}_$ex.err | grep "^{" > gpurun_out/cfg5_n${N}_$ex.json
done
$TR bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_final_n$N.err | grep "^{" > gpurun_out/bench_final_n$N.json
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/cfg*_n$N*.json"))+["gpurun_out/bench_final_n$N.json"]:
    try:
        d=json.load(open(f)); print(f, d.get("value"), d.get("unit"), d.get("exchange",""), d.get("bit_identical_to_single_gpu",""), d.get("ms_per_step",""))
    except Exception as e: print(f, "ERR", e)
PY
