set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; cat /sys/fs/cgroup/cpu.max
for cfg in "A=1" "MTZ_LZ4_PERSISTENT=1 MTZ_STREAM_PRIORITIES=0" "MTZ_LZ4_PERSISTENT=1" "MTZ_STREAM_PRIORITIES=0"; do
  echo "== fanout_probe [$cfg]"
  env $cfg timeout 300 python tools/fanout_probe.py 16 2 2 2>&1 | grep -v "^\[W\|^W0" | tail -8
done > gpurun_out/r2_fanout_probe.log 2>&1
cat gpurun_out/r2_fanout_probe.log
MTZ_TRACE=gpurun_out/r2_fanout_trace.txt timeout 300 python tools/fanout_probe.py 4 2 2 2>&1 | tail -4
head -c 20000 gpurun_out/r2_fanout_trace.txt > gpurun_out/r2_fanout_trace_head.txt; rm -f gpurun_out/r2_fanout_trace.txt
