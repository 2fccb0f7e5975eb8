set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 4 8; do
  $TR --nproc-per-node $n --master-port 2960$n tools/bench_shared_gpc.py 2> gpurun_out/c5_${n}gpu.err | grep -v "^NCCL" > gpurun_out/c5_${n}gpu.json
done
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2961$n bench.py --gpus $n --steps 3 --warmup 3 --no-c2 2> gpurun_out/tgt_${n}gpu.err | grep "^{" > gpurun_out/tgt_${n}gpu.json
done
for n in 4 8; do
  $TR --nproc-per-node $n --master-port 2962$n tools/bench_lung.py --gpus $n 2> gpurun_out/lung_${n}gpu.err | grep "^{" > gpurun_out/lung_${n}gpu.jsonl
done
ls -la gpurun_out/*gpu.json*
