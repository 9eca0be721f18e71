set -x
cd $GRAFT_REPO_ROOT
for t in default 1; do
  if [ $t = 1 ]; then export OMP_NUM_THREADS=1; fi
  timeout 300 python bench.py --config orbit --steps 120 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_orbit_$t.json 2> gpurun_out/r02j_orbit_$t.err
  python -c "
import json; d=json.load(open('gpurun_out/r02j_orbit_$t.json')); print('$t', d['value'], d['e2e']['value'], d['e2e']['ms_per_frame'])"
done
