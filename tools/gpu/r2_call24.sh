set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | wc -l; nproc; cat /sys/fs/cgroup/cpu.max
export MTZ_WATCHDOG_MS=10000
if [ "$2" != "notests" ]; then
timeout 400 python -m pytest tests/test_gpu_multidev.py -x -q -o timeout=120 > gpurun_out/r2_gputests_multidev_final_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_multidev_final_n$N.log
grep -v "^  File\|^    \|^~~~" gpurun_out/r2_gputests_multidev_final_n$N.log | tail -30
fi
unset MTZ_WATCHDOG_MS
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_final_n$N.json 2> gpurun_out/r2_bench_final_n$N.err; echo "bench rc=$?"
grep -v "^\[W\|^W0\|UserWarning\|d_in\[" gpurun_out/r2_bench_final_n$N.err | tail -c 1000
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_final_n$N.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','logical_gibs','e2e','e2e_stream_api','fanout','idempotent_at_full_size','failed'):
    print(k, json.dumps(d.get(k))[:700])
print(json.dumps(d['workload_detail'])[:200])
PY
