set -x
mkdir -p gpurun_out
timeout 600 python tools/ring_probe.py 8 > gpurun_out/r2_ring_probe_b.log 2>&1; cat gpurun_out/r2_ring_probe_b.log
timeout 300 python tools/chain_probe.py 4 > gpurun_out/r2_chain_probe.log 2>&1; cat gpurun_out/r2_chain_probe.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2_bench_n1_f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_f.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_stream_api','failed'):
    print(k, json.dumps(d.get(k))[:1000])
print(json.dumps(d['workloads']['verify'].get('ring_acquire_commit')))
PY
