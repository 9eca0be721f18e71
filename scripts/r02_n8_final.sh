# the last 8-GPU check of round 2 (gpurun --gpus 8): the default gather with the committed band-feedback policy, and the direct gather in pieces
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29900
run() { n=$1; name=$2; shift 2; port=$((port+1)); timeout 100 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python -c "
import json; d=json.load(open('gpurun_out/$name.json')); e=d['e2e']; print('$name', d['unit'], '%.1f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'K1 %.3f' % d['roofline']['kernel_ms'], 'e2e %.3f ms' % e.get('ms_per_step', e.get('ms_per_frame', 0)), d.get('stream_identical_to_single_gpu'), d.get('host_stream_identical_to_single_gpu'), d.get('bands', {}).get('k1_ms_mean_rank'))" || tail -5 gpurun_out/$name.err; echo "elapsed $SECONDS"; }
run 8 r02z_scale_n8_fused --steps 20 --warmup 5 --no-cpu-baseline
if [ $SECONDS -lt 105 ]; then run 8 r02z_scale_n8_direct_pieces --steps 20 --warmup 5 --no-cpu-baseline --fused 2 --pieces 0.5,0.3,0.2; fi
if [ $SECONDS -lt 120 ]; then run 8 r02z_config3_n8 --config stress --steps 8 --warmup 4 --no-cpu-baseline --fused 2; fi
