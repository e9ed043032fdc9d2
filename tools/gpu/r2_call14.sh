set -x
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l; nproc; cat /sys/fs/cgroup/cpu.max; free -g | head -2

timeout 600 python -m pytest tests/test_gpu_multidev.py -x -q > gpurun_out/r2_gputests_multidev_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_multidev_n$N.log
tail -4 gpurun_out/r2_gputests_multidev_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
grep -v "^\[W\|^W0" gpurun_out/r2_bench_n$N.err | tail -c 1500
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','logical_gibs','e2e','e2e_stream_api','fanout','idempotent_at_full_size','failed'):
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2>/dev/null; head -c 700 gpurun_out/r2_bench_ref_n$N.json
