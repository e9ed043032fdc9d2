set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lz4_certificate.py tests/test_gpu_codec.py -x -q -m gpu > gpurun_out/r2_gputests_cert.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_cert.log
tail -5 gpurun_out/r2_gputests_cert.log
for cfg in "A=1" "MTZ_CERTIFY=0"; do
  echo "== quick_codec [$cfg]"
  env $cfg timeout 300 python tools/quick_codec.py 2>&1 | tail -9
done > gpurun_out/r2_cert_ab.log 2>&1
cat gpurun_out/r2_cert_ab.log
