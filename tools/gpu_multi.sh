#!/bin/bash
# N-GPU box batch: multi-rank parity (tools/multi_gpu_check.py through tests/test_gpu_multi.py) and the weak-scaling
# bench at N ranks next to N = 1 on the same box.   usage: tools/gpu_multi.sh <tag> <N> [pytest] [bench]
tag=$1; N=$2; shift 2
what=${@:-pytest bench}
out=gpurun_out/$tag
mkdir -p $out
for w in $what; do
case $w in
pytest)
  python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > $out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $out/pytest_multi.log
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_check.py > $out/multi_gpu_check_n$N.log 2>&1; echo "multi_gpu_check N=$N rc=$?"; grep multi_gpu_check: $out/multi_gpu_check_n$N.log ;;
bench)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err; echo "bench N=$N rc=$?"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
  python - $out/bench_n$N.json $out/bench_n1.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("  N=%d ms/step %.3f value %.3e e2e %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value",0)), d.get("halo_overlap"))
    except Exception as e: print("no json", f, e)
PY
  ;;
esac
done
