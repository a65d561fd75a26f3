#!/bin/bash
# profiles/ summaries from the captures a gpurun call left in gpurun_out/ (run here, after the call):
#   benchmarks/make_profiles.sh <frame.ncu-rep> <tag>
# writes profiles/<tag>_frame_summary.txt (one line per kernel of the captured frames), <tag>_raster_hot_sass.txt,
# <tag>_raster_details.txt, <tag>_projection_details.txt and refreshes profiles/ncu_traffic.json.
set -e
rep=$1; tag=$2
cd "$(dirname "$0")/.."
{
  echo "# ncu --set full --clock-control none --import-source on python benchmarks/one_frame.py   (config 3: 1 M Gaussians @1080p)"
  echo "# 3 fused frames (12 kernels each) + the stand-alone projection stage; per-launch times are cold-cache and serialised"
  python benchmarks/ncu_summary.py $rep
} > profiles/${tag}_frame_summary.txt
python benchmarks/ncu_src.py $rep raster_pair 48 > profiles/${tag}_raster_hot_sass.txt 2>&1 || true
ncu -i $rep --page details --kernel-name regex:raster_pair_kernel 2>/dev/null | awk '/raster_pair_kernel/{n++} n<=1' > profiles/${tag}_raster_details.txt
ncu -i $rep --page details --kernel-name regex:project_kernel 2>/dev/null | awk '/project_kernel/{n++} n<=1' > profiles/${tag}_projection_details.txt
python benchmarks/ncu_traffic.py $rep profiles/ncu_traffic.json profiles/${tag}_frame_summary.txt
