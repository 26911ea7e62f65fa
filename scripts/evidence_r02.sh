#!/bin/bash
# Regenerates the round-2 evidence under gpurun_out/ on ONE B200 (copy what is cited into profiles/):
#   gpurun --timeout 2400 -- 'bash scripts/evidence_r02.sh'            everything
#   gpurun --timeout 700  -- 'QUICK=1 bash scripts/evidence_r02.sh'    only what depends on the scorer kernels
# 1. full GPU suite, 2. driver-style bench + reference arm, 3. ncu launch lists of every workload, 4. --set full captures of the
# kernels DESIGN.md quotes.  Nothing printed under ncu is a bench value.  Every step runs under its own `timeout`.
set -u
O=gpurun_out
Q=${QUICK:-0}
mkdir -p $O
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r02_gpu_tests_final.log
timeout 300 python bench.py --steps 20 --warmup 3 2>$O/bench_default.err | tail -1 > $O/r02_bench_default_1gpu.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 2>$O/bench_ref.err | tail -1 > $O/r02_bench_reference_arm.json
NCU="timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$NCU --log-file $O/r02_launches_4dof_1pct.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > /dev/null 2>&1
$NCU --log-file $O/r02_launches_4dof_47pct.csv python bench.py --flag-pct 53 --windows 524288 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > /dev/null 2>&1
$NCU --log-file $O/r02_launches_openlab.csv python bench.py --workload openlab_hybrid --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > /dev/null 2>&1
if [ "$Q" != "1" ]; then
  $NCU --log-file $O/r02_launches_train.csv python scripts/prof_train.py > /dev/null 2>&1
  $NCU --log-file $O/r02_launches_cnn_train_4dof.csv python scripts/prof_cnn_train.py 4dof > /dev/null 2>&1
  $NCU --log-file $O/r02_launches_cnn_train_openlab.csv python scripts/prof_cnn_train.py openlab > /dev/null 2>&1
  $NCU --log-file $O/r02_launches_cnnol_8192.csv python scripts/prof_cnnol_ncu.py > /dev/null 2>&1
  $NCU --log-file $O/r02_percentile_launches.csv python scripts/prof_percentile.py > /dev/null 2>&1
  timeout 120 python scripts/membound_bench.py > $O/r02_membound.jsonl 2>/dev/null
fi
python scripts/launch_summary.py $O/r02_launches_*.csv > $O/r02_launch_summary.txt 2>&1
FULL="timeout 200 ncu --set full --import-source on --clock-control none"
[ "$Q" != "1" ] && $FULL -k regex:ol_conv_gemm_kernel -s 3 -c 3 -o $O/r02_cnnol_fused python scripts/prof_cnnol_ncu.py > /dev/null 2>&1
$FULL -k regex:vae_score_tc_dual_kernel -s 1 -c 1 -o $O/r02_vae_tc_dual python scripts/prof_tc_ol.py > /dev/null 2>&1
$FULL -k regex:vae_score_tc_kernel -s 1 -c 1 -o $O/r02_vae_tc python scripts/prof_tc.py > /dev/null 2>&1
for r in r02_cnnol_fused r02_vae_tc_dual r02_vae_tc; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ls -la $O | tail -30
