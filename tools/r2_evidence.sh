set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
UGSM_USE_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_bench_one_pair.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 1 -c 1 -f -o gpurun_out/r2/k1_level0 python tools/k1_one.py 4928 3264 3 > gpurun_out/r2/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_ -s 1 -c 1 -f -o gpurun_out/r2/k2_level0_s10 python tools/k2_one.py 4928 3264 10 3 > gpurun_out/r2/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_ -s 1 -c 1 -f -o gpurun_out/r2/k2_level2_s5 python tools/k2_one.py 2463 1631 5 3 > gpurun_out/r2/ncu_k2b.log 2>&1; echo "ncu k2b rc=$?"
python tools/bench_kernels.py > gpurun_out/r2/kernel_times.txt 2>&1
python tools/measure_configs.py > gpurun_out/r2/configs.jsonl 2> gpurun_out/r2/configs.err
python tools/small_levels.py > gpurun_out/r2/small_levels.txt 2>&1
