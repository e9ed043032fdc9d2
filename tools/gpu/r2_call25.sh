set -x
mkdir -p gpurun_out
timeout 400 python tools/foreign_probe.py 2 2>&1 | grep -v "^\[W\|Warning" | tail -6 > gpurun_out/r2_foreign_probe.log; cat gpurun_out/r2_foreign_probe.log
