set -x
mkdir -p gpurun_out
for v in v4_128 v4_256 v7_128 v7_256; do
  echo "== $v"
  MTZ_SO=manatee_b200/libmanatee_gpu_$v.so timeout 200 python tools/chain_probe.py 2 2>&1 | grep decompress
done > gpurun_out/r2_chain_probe_ab.log 2>&1
cat gpurun_out/r2_chain_probe_ab.log
