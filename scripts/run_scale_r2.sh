TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_multi_test.log 2>&1; echo "multi test rc=$?"; tail -3 gpurun_out/r2_multi_test.log | cut -c1-200
$TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/r2_n8.json 2> gpurun_out/r2_n8.err; echo "n8 default rc=$?"
Q="--no-e2e --no-train --no-sparse --no-cpu-baseline --no-measure-copy --configs c4 --steps 200 --warmup 5"
for T in p2p_oneshot nccl; do $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 $Q --transport $T > gpurun_out/r2_n8_$T.json 2> gpurun_out/r2_n8_$T.err; echo "n8 $T rc=$?"; done
for G in 4 2; do $TR --nproc-per-node $G --master-port 29703 bench.py --gpus $G $Q > gpurun_out/r2_n$G.json 2> gpurun_out/r2_n$G.err; echo "n$G rc=$?"; done
python bench.py $Q > gpurun_out/r2_n1q.json 2> gpurun_out/r2_n1q.err; echo "n1 rc=$?"
