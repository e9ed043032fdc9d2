set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_3.log
tail -6 gpurun_out/r2_gputests_3.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err; echo "bench rc=$?"
tail -c 1000 gpurun_out/r2_bench_n1_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_e.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_stream_api','roofline','failed'):
    print(k, json.dumps(d.get(k))[:1000])
print(json.dumps(d['workloads']['verify'].get('ring_acquire_commit')))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_decompress.csv python tools/prof_codec.py decompress 32768 > gpurun_out/ncu_c.log 2>&1
grep -c stamp gpurun_out/r2_launches_decompress.csv
timeout 300 python tools/bench_plumbing.py 1 cpump,off > gpurun_out/r2_plumbing_config0.json 2>gpurun_out/plumb.err; cat gpurun_out/r2_plumbing_config0.json
