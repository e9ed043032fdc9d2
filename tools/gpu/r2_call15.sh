set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; cat /sys/fs/cgroup/cpu.max
timeout 600 python -m pytest tests/test_gpu_codec.py tests/test_gpu_stream_api.py -x -q > gpurun_out/r2_gputests_prio.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_prio.log
tail -4 gpurun_out/r2_gputests_prio.log
for cfg in "A=1" "MTZ_LZ4_PERSISTENT=1" "MTZ_STREAM_PRIORITIES=0" "MTZ_LZ4_PERSISTENT=1 MTZ_STREAM_PRIORITIES=0"; do
  echo "== quick_codec [$cfg]"
  env $cfg timeout 300 python tools/quick_codec.py 2>&1 | tail -9
done > gpurun_out/r2_prio_ab.log 2>&1
cat gpurun_out/r2_prio_ab.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_prio_n1.json 2> gpurun_out/r2_bench_prio_n1.err; echo "bench rc=$?"
grep -v "^\[W\|^W0" gpurun_out/r2_bench_prio_n1.err | tail -c 1500
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_prio_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','logical_gibs','e2e','e2e_stream_api','roofline','failed'):
    print(k, json.dumps(d.get(k))[:900])
PY
