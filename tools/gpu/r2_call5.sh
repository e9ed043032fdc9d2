set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_cancel.py -x -q > gpurun_out/r2_gputests_multidev_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_multidev_n2.log
tail -15 gpurun_out/r2_gputests_multidev_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2_bench_n2_a.err
head -c 6000 gpurun_out/r2_bench_n2_a.json
