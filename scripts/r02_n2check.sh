set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 300 $TR --master-port 29901 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02p_n2.json 2> gpurun_out/r02p_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r02p_n2.json')); print('n2', 'dev %.3f' % d['ms_per_step'], 'K1 max %.3f mean %.3f' % (d['bands']['k1_ms_max_rank'], d['bands']['k1_ms_mean_rank']), 'e2e %.3f' % d['e2e']['ms_per_step'], d['stream_identical_to_single_gpu'], d['host_stream_identical_to_single_gpu'])"
timeout 300 $TR --master-port 29902 bench.py --gpus 2 --steps 20 --warmup 5 --fused 1 > gpurun_out/r02p_n2f.json 2> gpurun_out/r02p_n2f.err; python -c "
import json; d=json.load(open('gpurun_out/r02p_n2f.json')); print('n2 fused', 'dev %.3f' % d['ms_per_step'], 'K1 max %.3f mean %.3f' % (d['bands']['k1_ms_max_rank'], d['bands']['k1_ms_mean_rank']), 'e2e %.3f' % d['e2e']['ms_per_step'], d['stream_identical_to_single_gpu'], d['host_stream_identical_to_single_gpu'])"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather or orbit or pipeline" 2>&1 | tail -2
