set -x
cd $GRAFT_REPO_ROOT
python scripts/ab_k1.py libtrt_b200.so > gpurun_out/r02n_ab.log 2>&1; cat gpurun_out/r02n_ab.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02n_tests.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r02n_tests.log | head
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_render -s 2 -c 1 -f -o gpurun_out/r02n_k1 python scripts/one_k1.py > gpurun_out/r02n_ncu.log 2>&1; tail -1 gpurun_out/r02n_ncu.log
timeout 300 python scripts/bounds_check.py gpurun_out/r02_bounds_check.txt > /dev/null 2>&1; tail -4 gpurun_out/r02_bounds_check.txt
