#!/bin/bash
# Round-2 profile pass on the GPU box (run through gpurun); everything lands in gpurun_out/.
set -x
O=gpurun_out
python tools/diag_bwd_speed.py > $O/r2_bwd_speed.log 2>&1
python tests/diag/diag_speed.py > $O/r2_fwd_speed.log 2>&1
SCREEN=1 python tests/diag/diag_speed.py >> $O/r2_fwd_speed.log 2>&1
# launch lists (durations + DRAM bytes): the headline bench and one training step
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $O/launches_dram_cfg2_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2_ncu_bench.log 2>&1
MVSDF_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $O/launches_train8k_r2.csv python bench.py --workload train8k --steps 1 --warmup 1 --no-cpu-baseline > $O/r2_ncu_train.log 2>&1
# full captures of the backward kernels and of the exact forward kernel
ncu --set full --clock-control none --import-source on -k regex:mlp_bwd_sweep_kernel -s 2 -c 1 -o $O/ncu_bwd_sweep_r2 -f python tools/diag_bwd_speed.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_bwd_dw_kernel -s 2 -c 1 -o $O/ncu_bwd_dw_r2 -f python tools/diag_bwd_speed.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_pair2_kernel -s 3 -c 1 -o $O/ncu_pair2_r2 -f python tests/diag/diag_speed.py > /dev/null 2>&1
# sanitizers on the new kernels (small cases)
compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_backward.py::test_sdf_backward_vs_explicit_chain[256-300-True-True]" "tests/test_gpu_backward.py::test_render_backward_vs_explicit_chain[256-200]" tests/test_gpu_backward.py::test_fused_adam_matches_torch_adam_with_clipping "tests/test_featext.py::test_native_featext_matches_oracle_at_other_shapes[3-40-72-2]" -x -q 2>&1 | tail -8 > $O/r2_sanitizer_memcheck.log
compute-sanitizer --tool synccheck python -m pytest "tests/test_gpu_backward.py::test_sdf_backward_vs_explicit_chain[256-300-True-True]" "tests/test_gpu_mlp.py" -x -q 2>&1 | tail -8 > $O/r2_sanitizer_synccheck.log
compute-sanitizer --tool synccheck python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -6 >> $O/r2_sanitizer_synccheck.log
ls -la $O | tail -15
