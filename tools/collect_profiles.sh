#!/bin/bash
# Round profile collection on a B200 box (run through gpurun from the repo root):
#   bench JSON lines, the ncu launch list of the bench command, ncu --set full of one step.
# Outputs land in gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into profiles/*.csv.
R=${1:-r02}
mkdir -p gpurun_out
for w in qg3_8192 qg3_1024 qg3_128 qg1_64 qg3_8192_f64 qg3_256_ens1024 swm2_4096 swm2_64; do
  steps=10; [ $w = qg3_128 ] && steps=100; [ $w = qg1_64 ] && steps=200; [ $w = swm2_64 ] && steps=200
  [ $w = qg3_1024 ] && steps=40; [ $w = qg3_8192_f64 ] && steps=4
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${R}_qg3_8192_launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -c 40 -o gpurun_out/${R}_qg3_8192_full -f \
  python tools/prof_run.py qg3_8192 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:qg_rhs -c 6 -o gpurun_out/${R}_qg3_8192_full_rhs -f \
  python tools/prof_run.py qg3_8192 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:swm_rhs -c 6 -o gpurun_out/${R}_swm2_4096_full -f \
  python tools/prof_run.py swm2_4096 1 > /dev/null 2>&1
# summarise on the box; the .ncu-rep files are too large to bring back (64 MiB cap)
python tools/ncu_summary.py gpurun_out/${R}_qg3_8192_full.ncu-rep gpurun_out/${R}_qg3_8192_ncu_full.csv
python tools/ncu_summary.py gpurun_out/${R}_qg3_8192_full_rhs.ncu-rep gpurun_out/${R}_qg3_8192_ncu_full_rhs.csv
python tools/ncu_summary.py gpurun_out/${R}_swm2_4096_full.ncu-rep gpurun_out/${R}_swm2_4096_ncu_full.csv
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -12
