set -x
mkdir -p gpurun_out
timeout 600 python tools/k3_bound.py 16384 > gpurun_out/r2_k3_bound.log 2>&1; cat gpurun_out/r2_k3_bound.log
QUICK_RESIDENT_ONLY=1 MTZ_SO=$PWD/manatee_b200/libmanatee_gpu_prof.so timeout 300 python tools/quick_codec.py 2 compress 2>&1 | grep -E "K3PROF|resident" | head -12 > gpurun_out/r2_k3_prof.log; cat gpurun_out/r2_k3_prof.log
