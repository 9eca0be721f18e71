set -x
cd $GRAFT_REPO_ROOT
python scripts/ab_k1.py libtrt_b200.so libtrt_b200_uv2.so libtrt_b200_u2.so libtrt_b200_uinl.so libtrt_b200_pw.so libtrt_b200_pwuv2.so libtrt_b200_c5.so libtrt_b200_w6c3.so libtrt_b200_clni.so > gpurun_out/r02g_ab.log 2>&1; cat gpurun_out/r02g_ab.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02g_tests.log 2>&1; tail -6 gpurun_out/r02g_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; cut -c1-400 gpurun_out/r02g_bench.json; tail -3 gpurun_out/r02g_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err; cut -c1-200 gpurun_out/r02g_ref.json
bash scripts/run_sanitizer.sh gpurun_out/r02_sanitizer.txt
