set -x
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 --verify-gib 16 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n1_c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_c.json').read().strip().splitlines()[-1])
for k in ('value','e2e','e2e_stream_api','failed'):
    print(k, json.dumps(d.get(k))[:900])
print(json.dumps(d['workloads']['verify'].get('ring_acquire_commit')))
PY
# launch list of one small RECOMPRESS pass (4096 records) and a full capture of one K3 launch
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_recompress.csv python tools/prof_codec.py recompress 20000 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_lz4_encode -s 1 -c 1 -o gpurun_out/r2_k3 python tools/prof_codec.py compress 4096 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
