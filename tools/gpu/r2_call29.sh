mkdir -p gpurun_out
timeout 120 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_final2.log
tail -3 gpurun_out/r2_gputests_final2.log
