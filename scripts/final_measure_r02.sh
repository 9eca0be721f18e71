# the single-GPU measurement batch of round 2 (gpurun, one B200): everything profiles/r02_* is made from
set -x
cd $GRAFT_REPO_ROOT
if [ -z "$SKIP_TESTS" ]; then (time timeout 1100 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_tests.log 2>&1; grep -E "passed|failed" gpurun_out/r02_tests.log; fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_render|k_encode|k_tile_certs" -s 6 -c 3 -f -o gpurun_out/r02_k1_k2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__sass_average_branch_targets_threads_uniform.pct,l1tex__t_bytes.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:k_render -s 4 -c 1 --csv --log-file gpurun_out/r02_k1_flop_counters.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
KIND=stress W=3840 H=2160 timeout 900 ncu --metrics sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_render -s 2 -c 1 --csv --log-file gpurun_out/r02_stress_flop_counters.csv python scripts/one_k1.py > gpurun_out/ncu4.log 2>&1
timeout 600 python scripts/bounds_check.py gpurun_out/r02_bounds_check.txt > /dev/null 2>&1; tail -3 gpurun_out/r02_bounds_check.txt
bash scripts/r02_single.sh
ls -la gpurun_out/r02_*
