# Round-2 evidence on the 8 GPUs of one box (gpurun --gpus 8): bench.py at N = 8 (pair-parallel + the row-band block) and
# BASELINE configs[4] through ugsm_batch (tools/config4_batch.py).  Outputs under gpurun_out/r2/.
mkdir -p gpurun_out/r2
nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err; echo "bench n8 rc=$?"
python tools/config4_batch.py 8 1050 > gpurun_out/r2/config4_batch_n8.jsonl 2> gpurun_out/r2/config4_batch.err; echo "batch rc=$?"
cat gpurun_out/r2/config4_batch_n8.jsonl
