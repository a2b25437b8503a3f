#!/bin/bash
# Round-2 profile pass #3 (final code: fused SDF head in the exact kernel): everything lands in gpurun_out/.
set -x
O=gpurun_out
python tests/diag/diag_speed.py > $O/r2c_fwd_speed.log 2>&1
SCREEN=1 python tests/diag/diag_speed.py >> $O/r2c_fwd_speed.log 2>&1
MVSDF_FUSE_HEAD=0 python tests/diag/diag_speed.py >> $O/r2c_fwd_speed.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_pair2_kernel -s 3 -c 1 -o $O/ncu_pair2_r2c -f python tests/diag/diag_speed.py > /dev/null 2>&1
ncu -i $O/ncu_pair2_r2c.ncu-rep --page raw --csv > $O/ncu_pair2_r2c_raw.csv 2>/dev/null
python tools/ncu_src_summary.py $O/ncu_pair2_r2c.ncu-rep 14 > $O/ncu_pair2_r2c_src.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $O/launches_dram_cfg2_r2c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2c_ncu_bench.log 2>&1
timeout 400 python bench.py 2>&1 | tail -1 > $O/bench_cfg2_r2c.json
timeout 300 python bench.py --workload train32k --steps 20 --warmup 3 2>&1 | tail -1 > $O/bench_train32k_r2c.json
ls -la $O | tail -12
