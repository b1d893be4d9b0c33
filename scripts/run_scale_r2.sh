# round-2 scaling evidence: default bench line at 8 GPUs, quick loss-only lines at 8 (no rebalance, for A/B), 4, 2 and 1 GPUs
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/r2b_n8.json 2> gpurun_out/r2b_n8.err; echo "n8 default rc=$?"
Q="--no-e2e --no-train --no-sparse --no-cpu-baseline --no-measure-copy --configs c4 --steps 200 --warmup 5"
$TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 $Q --no-rebalance > gpurun_out/r2b_n8_norebalance.json 2> gpurun_out/r2b_n8_norebalance.err; echo "n8 norebalance rc=$?"
for G in 4 2; do $TR --nproc-per-node $G --master-port 29703 bench.py --gpus $G $Q > gpurun_out/r2b_n$G.json 2> gpurun_out/r2b_n$G.err; echo "n$G rc=$?"; done
python bench.py $Q > gpurun_out/r2b_n1q.json 2> gpurun_out/r2b_n1q.err; echo "n1 rc=$?"
