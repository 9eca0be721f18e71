set -x
cd $GRAFT_REPO_ROOT
python scripts/ab_k1.py libtrt_b200_topdown.so libtrt_b200.so libtrt_b200_f3.so > gpurun_out/r02l_ab.log 2>&1; cat gpurun_out/r02l_ab.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02l_tests.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r02l_tests.log | head
timeout 300 python scripts/band_overhead.py > gpurun_out/r02l_band_overhead.txt 2>&1; cat gpurun_out/r02l_band_overhead.txt
timeout 300 python bench.py --config orbit --warmup 3 --no-cpu-baseline > gpurun_out/r02l_orbit.json 2> gpurun_out/r02l_orbit.err; python -c "
import json; d=json.load(open('gpurun_out/r02l_orbit.json')); print('orbit', d['value'], d['e2e']['value'], d['e2e']['ms_per_frame'])"
timeout 300 python bench.py --config stress --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02l_stress.json 2> gpurun_out/r02l_stress.err; python -c "
import json; d=json.load(open('gpurun_out/r02l_stress.json')); print('stress', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
