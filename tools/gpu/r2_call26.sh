set -x
mkdir -p gpurun_out
for i in 1 2; do
for so in "" "manatee_b200/libmanatee_gpu_old.so"; do
  echo "== MTZ_SO=[$so]"
  MTZ_SO=$so QUICK_RESIDENT_ONLY=1 timeout 300 python tools/quick_codec.py 2>&1 | grep "recompress resident" | tail -1
done
done > gpurun_out/r2_cert_f.log 2>&1
cat gpurun_out/r2_cert_f.log
