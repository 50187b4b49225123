#!/bin/bash
# ncu evidence for the slab kernels on ONE GPU (all slabs in one process: peer stores land in the
# same device): launch list of one step of 4 slabs of a 4096^2 grid, and --set full of the
# transpose / halo / barrier kernels.  Run through gpurun from the repo root.
R=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/${R}_qg3_4096_slab4local_launches.csv python tools/prof_slab_local.py 4096 4 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:seg_copy -c 14 -o gpurun_out/${R}_slab_full -f \
  python tools/prof_slab_local.py 4096 4 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${R}_slab_full.ncu-rep gpurun_out/${R}_qg3_4096_slab4local_ncu_full.csv
# the segmented y-sweeps (probe / apply passes) of the same run
SOMAX_B200_SLAB_NSEG=8 timeout 600 ncu --set full --clock-control none -k regex:thomas_sweep -c 16 \
  -o gpurun_out/${R}_slab_sweeps -f python tools/prof_slab_local.py 4096 4 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${R}_slab_sweeps.ncu-rep gpurun_out/${R}_qg3_4096_slab4local_sweeps_ncu_full.csv
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -5
