# multi-GPU measurements of round 2: bash scripts/r02_multi.sh <N list> (inside gpurun --gpus N)
set -x
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29510
for n in $1; do
  port=$((port+1)); timeout 300 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err; cut -c1-260 gpurun_out/r02_scale_n$n.json; tail -2 gpurun_out/r02_scale_n$n.err
done
nmax=$(echo $1 | awk '{print $1}')
port=$((port+1)); timeout 300 $TR --nproc-per-node $nmax --master-port $port bench.py --gpus $nmax --steps 20 --warmup 5 --fused 0 > gpurun_out/r02_scale_n${nmax}_pieces.json 2> gpurun_out/r02_scale_n${nmax}_pieces.err; cut -c1-260 gpurun_out/r02_scale_n${nmax}_pieces.json
port=$((port+1)); timeout 300 $TR --nproc-per-node $nmax --master-port $port bench.py --gpus $nmax --steps 20 --warmup 5 --fused 1 > gpurun_out/r02_scale_n${nmax}_fused.json 2> gpurun_out/r02_scale_n${nmax}_fused.err; cut -c1-260 gpurun_out/r02_scale_n${nmax}_fused.json
port=$((port+1)); timeout 400 $TR --nproc-per-node $nmax --master-port $port bench.py --gpus $nmax --config stress --steps 10 --warmup 3 > gpurun_out/r02_config3_n$nmax.json 2> gpurun_out/r02_config3_n$nmax.err; cut -c1-260 gpurun_out/r02_config3_n$nmax.json; tail -2 gpurun_out/r02_config3_n$nmax.err
port=$((port+1)); timeout 400 $TR --nproc-per-node $nmax --master-port $port bench.py --gpus $nmax --config orbit --warmup 3 > gpurun_out/r02_config4_n$nmax.json 2> gpurun_out/r02_config4_n$nmax.err; cut -c1-260 gpurun_out/r02_config4_n$nmax.json; tail -2 gpurun_out/r02_config4_n$nmax.err
for n in $1; do
  port=$((port+1)); timeout 120 $TR --nproc-per-node $n --master-port $port scripts/pcie_bw.py >> gpurun_out/r02_pcie_bw.txt 2>/dev/null
done
cat gpurun_out/r02_pcie_bw.txt
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
