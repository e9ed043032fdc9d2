set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
nproc; free -g | head -2; cat /sys/fs/cgroup/cpu.max
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_1.log
tail -15 gpurun_out/r2_gputests_1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench_n1_a.err
cat gpurun_out/r2_bench_n1_a.json | head -c 6000
timeout 600 python tools/quick_codec.py 4 compress,decompress > gpurun_out/r2_quick_codec_a.log 2>&1; tail -20 gpurun_out/r2_quick_codec_a.log
