# Round-2 evidence of the current build, one B200 (run through gpurun; outputs under gpurun_out/r2/).
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2/bench_reference_arm.json 2> gpurun_out/r2/bench_ref.err; echo "ref rc=$?"
UGSM_USE_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_bench_one_pair.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu_launch.log 2>&1; echo "ncu list rc=$?"
# level-0 K1 / K2 launches of a real pair: k1_stream runs 64 times per pair (levels 0-6), k2_stream 42 times (levels 0-5);
# the last launch of the second pair of tools/one_pair.py is skipped to with -s 126 / -s 82
UGSM_USE_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_stream -s 126 -c 1 -f -o gpurun_out/r2/k1_level0 python tools/one_pair.py > gpurun_out/r2/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
UGSM_USE_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_stream -s 82 -c 1 -f -o gpurun_out/r2/k2_level0_s10 python tools/one_pair.py > gpurun_out/r2/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_stream -s 1 -c 1 -f -o gpurun_out/r2/k2_level2_s5 python tools/k2_one.py 2463 1631 5 3 > gpurun_out/r2/ncu_k2b.log 2>&1; echo "ncu k2b rc=$?"
python tools/bench_kernels.py > gpurun_out/r2/kernel_times.txt 2>&1
UGSM_USE_GRAPH=0 python tools/bench_kernels.py | tail -3 >> gpurun_out/r2/kernel_times.txt 2>&1
python tools/measure_configs.py > gpurun_out/r2/configs.jsonl 2> gpurun_out/r2/configs.err
python tools/small_levels.py > gpurun_out/r2/small_levels.txt 2>&1
python tools/cloud_bench.py > gpurun_out/r2/cloud.jsonl 2>/dev/null
