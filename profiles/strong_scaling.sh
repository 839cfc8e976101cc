#!/bin/bash
# Strong-scaling lines for BASELINE configs[3] and [4] (bench.py --strong): usage  profiles/strong_scaling.sh N
# Writes gpurun_out/strong_c{4,5}_n{N}.json.  One dataset of the config's full size, 16 batches dealt to N ranks,
# counter allreduce inside every step.
N=${1:-1}
mkdir -p gpurun_out
run() {
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@"; else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$N" "$@"; fi
}
run --strong --config 4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/strong_c4_n$N.json 2> gpurun_out/strong_c4_n$N.err
run --strong --config 5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/strong_c5_n$N.json 2> gpurun_out/strong_c5_n$N.err
tail -n 2 gpurun_out/strong_c4_n$N.err gpurun_out/strong_c5_n$N.err
