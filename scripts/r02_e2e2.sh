set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29600
run() { port=$((port+1)); timeout 200 $TR --master-port $port bench.py --gpus 2 --steps 6 --warmup 5 --e2e-steps 20 > gpurun_out/r02k_$1.json 2> gpurun_out/r02k_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r02k_$1.json')); print('$1', 'dev %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], d['host_stream_identical_to_single_gpu'])"; }
run default
TRT_L2_PERSIST=1 run l2persist
TRT_HOST_PIECES=0.1,0.3,0.3,0.2,0.1 run p5
TRT_HOST_PIECES=0.15,0.35,0.3,0.2 run p4
TRT_HOST_PIECES=0.05,0.1,0.15,0.2,0.2,0.15,0.1,0.05 run p8
TRT_HOST_PIECES=0.4,0.3,0.2,0.1 run r01
