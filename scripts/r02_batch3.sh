set -x
cd $GRAFT_REPO_ROOT
python scripts/ab_k1.py libtrt_b200_nofuse.so libtrt_b200.so > gpurun_out/r02i_ab.log 2>&1; cat gpurun_out/r02i_ab.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02i_tests.log 2>&1; grep -E "passed|failed" gpurun_out/r02i_tests.log
timeout 300 python scripts/orbit_trace.py > gpurun_out/r02i_orbit_trace.txt 2>&1; cat gpurun_out/r02i_orbit_trace.txt
