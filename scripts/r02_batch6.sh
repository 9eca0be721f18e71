set -x
cd $GRAFT_REPO_ROOT
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02m_tests.log 2>&1; grep -E "passed|failed|Error" gpurun_out/r02m_tests.log | head
python scripts/ab_k1.py libtrt_b200.so > gpurun_out/r02m_ab.log 2>&1; cat gpurun_out/r02m_ab.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29700
for f in 0 1 2; do port=$((port+1)); timeout 200 $TR --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 10 --fused $f > gpurun_out/r02m_n2_f$f.json 2> gpurun_out/r02m_n2_f$f.err; python -c "
import json; d=json.load(open('gpurun_out/r02m_n2_f$f.json')); print('fused=$f', 'dev %.3f' % d['ms_per_step'], 'K1 %.3f' % d['roofline']['kernel_ms'], 'e2e %.3f' % d['e2e']['ms_per_step'], d['stream_identical_to_single_gpu'], d['host_stream_identical_to_single_gpu'])"; done
