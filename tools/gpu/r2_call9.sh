set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_4.log
tail -5 gpurun_out/r2_gputests_4.log
timeout 600 python tools/ring_probe.py 8 > gpurun_out/r2_ring_probe.log 2>&1; cat gpurun_out/r2_ring_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_decompress_b.csv python tools/prof_codec.py decompress 32768 > gpurun_out/ncu_d.log 2>&1
grep stamp_chain gpurun_out/r2_launches_decompress_b.csv | awk -F'","' '{print $NF}' | head -8
QUICK_RESIDENT_ONLY=1 timeout 300 python tools/quick_codec.py 4 decompress 2>&1 | grep resident
