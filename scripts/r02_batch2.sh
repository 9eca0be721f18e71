set -x
cd $GRAFT_REPO_ROOT
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02h_tests.log 2>&1; tail -4 gpurun_out/r02h_tests.log
compute-sanitizer --version > gpurun_out/r02h_san_debug.txt 2>&1; timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py >> gpurun_out/r02h_san_debug.txt 2>&1; echo "exit $?" >> gpurun_out/r02h_san_debug.txt; head -30 gpurun_out/r02h_san_debug.txt
bash scripts/r02_single.sh
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; cut -c1-300 gpurun_out/r02h_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_render|k_encode|k_tile_certs" -s 6 -c 3 -f -o gpurun_out/r02h_k1_k2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_ncu2.log 2>&1; tail -2 gpurun_out/r02h_ncu2.log
KIND=stress W=1920 H=1080 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_render -s 2 -c 1 -f -o gpurun_out/r02h_stress python scripts/one_k1.py > gpurun_out/r02h_ncu3.log 2>&1; tail -2 gpurun_out/r02h_ncu3.log
