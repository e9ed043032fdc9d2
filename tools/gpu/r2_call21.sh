set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lz4_certificate.py -x -q -m gpu > gpurun_out/r2_gputests_cert4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_cert4.log
tail -3 gpurun_out/r2_gputests_cert4.log
QUICK_RESIDENT_ONLY=1 timeout 300 python tools/quick_codec.py 2>&1 | tail -4 > gpurun_out/r2_cert_d.log; cat gpurun_out/r2_cert_d.log
ncu --set full --clock-control none --import-source on -k regex:k3c_lz4_certify -s 1 -c 1 -o gpurun_out/r2_k3c_c python tools/prof_codec.py recompress 8192 > gpurun_out/ncu_i.log 2>&1
ls -la gpurun_out/r2_k3c_c.ncu-rep
