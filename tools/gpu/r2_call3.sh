set -x
mkdir -p gpurun_out
export QUICK_RESIDENT_ONLY=1
for v in "" _v3 _nospec _noprobe _w26 _v3w26; do
  echo "=== variant libmanatee_gpu$v.so"
  MTZ_SO=$PWD/manatee_b200/libmanatee_gpu$v.so timeout 300 python tools/quick_codec.py 4 compress 2>&1 | grep -E "resident|oracle|Error|error" 
done > gpurun_out/r2_k3_variants_a.log 2>&1
cat gpurun_out/r2_k3_variants_a.log
