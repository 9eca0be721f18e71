set -x
cd $GRAFT_REPO_ROOT
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01f_tests.log 2>&1; tail -2 gpurun_out/r01f_tests.log; fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r01f_smoke.log 2>&1; tail -1 gpurun_out/r01f_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01f_bench_reference_arm.json 2> gpurun_out/r01f_ref.err
timeout 600 python bench.py > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_render|k_encode" -s 4 -c 2 -f -o gpurun_out/r01f_k1_k2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum,sm__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__sass_average_branch_targets_threads_uniform.pct,l1tex__t_bytes.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_render -s 4 -c 1 --csv --log-file gpurun_out/r01f_k1_flop_counters.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 900 python scripts/run_configs.py gpurun_out/r01f_configs.json > gpurun_out/r01f_configs.log 2>&1; tail -3 gpurun_out/r01f_configs.log
ls -la gpurun_out/r01f_*
