set -x
mkdir -p gpurun_out
timeout 300 python tools/chain_probe.py 4 > gpurun_out/r2_chain_probe_b.log 2>&1; cat gpurun_out/r2_chain_probe_b.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_5.log
tail -4 gpurun_out/r2_gputests_5.log
