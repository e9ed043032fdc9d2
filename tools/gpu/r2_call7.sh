set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_2.log
tail -6 gpurun_out/r2_gputests_2.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; echo "bench rc=$?"
tail -c 1000 gpurun_out/r2_bench_n1_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_d.json').read().strip().splitlines()[-1])
for k in ('value','e2e','e2e_stream_api','failed'):
    print(k, json.dumps(d.get(k))[:900])
print(json.dumps(d['workloads']['verify'].get('ring_acquire_commit')))
PY
# small records: the combination that took a box down in round 1, bounded
timeout 150 python bench.py --recsize 8192 --gib 1 --steps 2 --warmup 1 --verify-gib 0 --no-cpu > gpurun_out/r2_small_8k.json 2> gpurun_out/r2_small_8k.err; echo "small8k rc=$?"
tail -c 600 gpurun_out/r2_small_8k.err; head -c 1500 gpurun_out/r2_small_8k.json
timeout 150 python bench.py --recsize 4096 --gib 1 --steps 2 --warmup 1 --verify-gib 0 --no-cpu > gpurun_out/r2_small_4k.json 2> gpurun_out/r2_small_4k.err; echo "small4k rc=$?"
tail -c 600 gpurun_out/r2_small_4k.err; head -c 1500 gpurun_out/r2_small_4k.json
nvidia-smi --query-gpu=name,memory.used --format=csv
