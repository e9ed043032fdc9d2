set -x
mkdir -p gpurun_out
timeout 300 python tools/chain_probe.py > gpurun_out/r2_chain_probe_v7.log 2>&1; cat gpurun_out/r2_chain_probe_v7.log | tail -6
timeout 300 python -m pytest tests/test_gpu_codec.py -x -q -m gpu 2>&1 | tail -2
