#!/bin/bash
# Round-end measurement runs (development aid): bench.py at N ranks, output into gpurun_out/.
N=${1:-1}
TAG=${2:-final}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_${TAG}_bench_1gpu.json 2> gpurun_out/r2_${TAG}_bench_1gpu.err
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_${TAG}_bench_reference.json 2> gpurun_out/r2_${TAG}_bench_reference.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_${TAG}_bench_${N}gpu.json 2> gpurun_out/r2_${TAG}_bench_${N}gpu.err
fi
tail -c 600 gpurun_out/r2_${TAG}_bench_${N}gpu.json
