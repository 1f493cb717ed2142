#!/bin/bash
# Experiment: the N = 8 slab shape (512 x 512 x 64 points per rank) reproduced on 2 ranks, overlap on / off.
mkdir -p gpurun_out/r2q
run() { tag=$1; shift; env "$@" MG_BENCH_SHAPE=512,512,128 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 --no-parity > gpurun_out/r2q/$tag.json 2> gpurun_out/r2q/$tag.err; python - gpurun_out/r2q/$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "ms/step %.3f"%d['ms_per_step'], "gap %.2f"%d['step_minus_kernel_sum_ms'], {k:(round(v['avg_ms'],3),v['launches']) for k,v in d['kernels'].items()})
except Exception as e: print(sys.argv[2], "failed", e)
PY
}
run ov1 MG_OVERLAP=1
run ov1_c3 MG_OVERLAP=1 MG_CHUNKS=3
