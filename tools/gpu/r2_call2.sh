set -x
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$?"
tail -c 2000 gpurun_out/r2_bench_n1_b.err
head -c 7000 gpurun_out/r2_bench_n1_b.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_b.json 2> gpurun_out/r2_bench_ref_b.err; echo "ref rc=$?"
head -c 1500 gpurun_out/r2_bench_ref_b.json
