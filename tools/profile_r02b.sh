#!/bin/bash
# Round-2 profile pass #2 (after the CTA-pair reverse sweep): everything lands in gpurun_out/.
set -x
O=gpurun_out
python tools/diag_bwd_speed.py > $O/r2b_bwd_speed.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_bwd_sweep_pair_kernel -s 2 -c 1 -o $O/ncu_bwd_sweep_pair_r2 -f python tools/diag_bwd_speed.py > /dev/null 2>&1
ncu -i $O/ncu_bwd_sweep_pair_r2.ncu-rep --page raw --csv > $O/ncu_bwd_sweep_pair_r2_raw.csv 2>/dev/null
python tools/ncu_src_summary.py $O/ncu_bwd_sweep_pair_r2.ncu-rep 14 > $O/ncu_bwd_sweep_pair_r2_src.txt 2>&1
MVSDF_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file $O/launches_train32k_r2.csv python bench.py --workload train32k --steps 1 --warmup 1 --no-cpu-baseline > $O/r2b_ncu_train.log 2>&1
timeout 300 python bench.py --workload train32k --steps 20 --warmup 3 2>&1 | tail -1 > $O/bench_train32k_r2b.json
compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_backward.py::test_pair_sweep_matches_single_cta_sweep" -x -q 2>&1 | tail -6 > $O/r2b_sanitizer_memcheck.log
ls -la $O | tail -12
