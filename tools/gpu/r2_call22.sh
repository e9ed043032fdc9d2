set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lz4_certificate.py -x -q -m gpu > gpurun_out/r2_gputests_cert5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_cert5.log
tail -3 gpurun_out/r2_gputests_cert5.log
QUICK_RESIDENT_ONLY=1 timeout 300 python tools/quick_codec.py 2>&1 | tail -4 > gpurun_out/r2_cert_e.log; cat gpurun_out/r2_cert_e.log
