set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lz4_certificate.py tests/test_gpu_codec.py -x -q -m gpu > gpurun_out/r2_gputests_cert2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_cert2.log
tail -4 gpurun_out/r2_gputests_cert2.log
timeout 300 python tools/quick_codec.py 2>&1 | tail -8 > gpurun_out/r2_cert_b.log; cat gpurun_out/r2_cert_b.log
ncu --set full --clock-control none --import-source on -k regex:k3c_lz4_certify -s 1 -c 1 -o gpurun_out/r2_k3c python tools/prof_codec.py recompress 8192 > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/r2_k3c.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_recompress_cert.csv python tools/prof_codec.py recompress 20000 > gpurun_out/ncu_g.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cert_n1.json 2> gpurun_out/r2_bench_cert_n1.err; echo "bench rc=$?"
grep -v "^\[W\|^W0" gpurun_out/r2_bench_cert_n1.err | tail -c 1200
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_cert_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','logical_gibs','e2e','e2e_stream_api','roofline','failed'):
    print(k, json.dumps(d.get(k))[:700])
print(json.dumps(d['workloads'].get('recompress_reencode_all'))[:600])
print(json.dumps(d['workload_detail'])[:300])
PY
