# the multi-GPU measurements of round 2, final code: bash scripts/r02_multi2.sh   (inside gpurun --gpus 8)
set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29800
run() { n=$1; name=$2; shift 2; port=$((port+1)); timeout 300 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python -c "
import json; d=json.load(open('gpurun_out/$name.json')); e=d['e2e']; print('$name', d['unit'], '%.1f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'K1 %.3f' % d['roofline']['kernel_ms'], 'e2e %.3f ms' % e.get('ms_per_step', e.get('ms_per_frame', 0)), d.get('stream_identical_to_single_gpu', d.get('ordered_stream_identical_to_single_gpu')), d.get('host_stream_identical_to_single_gpu'))" || tail -5 gpurun_out/$name.err; }
run 8 r02f_scale_n8_fused --steps 20 --warmup 5 --fused 1
run 8 r02f_scale_n8_direct --steps 20 --warmup 5 --fused 2
run 4 r02f_scale_n4_pieces --steps 20 --warmup 5 --fused 0
run 4 r02f_scale_n4_direct --steps 20 --warmup 5 --fused 2
run 2 r02f_scale_n2_pieces --steps 20 --warmup 5 --fused 0
run 2 r02f_scale_n2_direct --steps 20 --warmup 5 --fused 2
run 8 r02f_config3_n8 --config stress --steps 10 --warmup 4 --fused 2
run 8 r02f_config4_n8 --config orbit --warmup 3
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_scale_n1.json 2> gpurun_out/r02f_scale_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r02f_scale_n1.json')); print('n1', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
