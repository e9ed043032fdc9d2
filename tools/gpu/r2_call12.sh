set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_stamp_chain -s 1 -c 1 -o gpurun_out/r2_chain python tools/prof_codec.py decompress 20000 > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/r2_chain.ncu-rep
