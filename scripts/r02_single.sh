# single-GPU configs of round 2 (BASELINE configs 1, 3, 4 with clocks and the CPU beside them) + the C host orbit
set -x
cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --config demo4k --steps 20 --warmup 3 > gpurun_out/r02_config1_n1.json 2> gpurun_out/r02_config1_n1.err; cut -c1-200 gpurun_out/r02_config1_n1.json
timeout 600 python bench.py --config stress --steps 5 --warmup 2 > gpurun_out/r02_config3_n1.json 2> gpurun_out/r02_config3_n1.err; cut -c1-200 gpurun_out/r02_config3_n1.json; tail -2 gpurun_out/r02_config3_n1.err
timeout 600 python bench.py --config orbit --warmup 3 > gpurun_out/r02_config4_n1.json 2> gpurun_out/r02_config4_n1.err; cut -c1-200 gpurun_out/r02_config4_n1.json; tail -2 gpurun_out/r02_config4_n1.err
# the C host: trt_render_orbit -> fwrite -> /dev/null
python -c "
from terminalraytracer_b200 import scene as S; import os
sky=S.synthetic_cubemap('milky_way',1024); os.makedirs('skybox/milky_way',exist_ok=True)
[S.write_ppm('skybox/milky_way/'+n, sky.face(i)) for i,n in enumerate(S.FACE_FILES)]"
gcc -O2 -Iinclude host/trt_demo.c -Lterminalraytracer_b200 -ltrt_b200 -Wl,-rpath,$PWD/terminalraytracer_b200 -lm -o /tmp/trt_demo
for i in 1 2; do /tmp/trt_demo --orbit 360 milky_way 1920 1080 2>> gpurun_out/r02_c_host_orbit.txt > /dev/null; done; cat gpurun_out/r02_c_host_orbit.txt
rm -rf skybox/milky_way
