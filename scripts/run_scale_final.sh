# final-code scaling evidence with the driver's own step counts (--steps 20 --warmup 5), one 8-GPU box
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/r2f_multi_tests.log 2>&1; echo "tests rc=$?"
$TR --nproc-per-node 8 --master-port 29721 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2f_n8.json 2> gpurun_out/r2f_n8.err; echo "n8 rc=$?"
$TR --nproc-per-node 4 --master-port 29722 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2f_n4.json 2> gpurun_out/r2f_n4.err; echo "n4 rc=$?"
Q="--no-e2e --no-train --no-sparse --no-cpu-baseline --no-measure-copy --configs none"
$TR --nproc-per-node 8 --master-port 29723 bench.py --gpus 8 --steps 200 --warmup 5 $Q > gpurun_out/r2f_n8_k200.json 2> gpurun_out/r2f_n8_k200.err; echo "n8 k200 rc=$?"
python bench.py --steps 20 --warmup 5 $Q > gpurun_out/r2f_n1q.json 2> gpurun_out/r2f_n1q.err; echo "n1 rc=$?"
