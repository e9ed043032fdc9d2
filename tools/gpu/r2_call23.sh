set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; cat /sys/fs/cgroup/cpu.max
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_final.log
tail -4 gpurun_out/r2_gputests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; echo "bench rc=$?"
grep -v "^\[W\|^W0" gpurun_out/r2_bench_final_n1.err | tail -c 800
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_final_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','logical_gibs','steps','e2e','e2e_stream_api','roofline','cpu_baseline','clocks','failed'):
    print(k, json.dumps(d.get(k))[:600])
print(json.dumps(d['workloads'].get('recompress_reencode_all'))[:400])
print(json.dumps(d['workloads'].get('verify'))[:1500])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2>/dev/null; head -c 600 gpurun_out/r2_bench_final_ref.json
