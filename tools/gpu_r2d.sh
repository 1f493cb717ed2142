mkdir -p gpurun_out/r2d
python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider > gpurun_out/r2d/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2d/pytest_multi.log
for ov in 1 0; do
MG_OVERLAP=$ov python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d/bench_n2_ov$ov.json 2> gpurun_out/r2d/bench_n2_ov$ov.err; echo "bench n2 overlap=$ov rc=$?"
python - gpurun_out/r2d/bench_n2_ov$ov.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {n:round(v["ms"]/10,3) for n,v in d["kernels"].items()})
except Exception as e: print("no json", e)
PY
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d/bench_n1.json 2> gpurun_out/r2d/bench_n1.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d/bench_n1.json").read().strip().splitlines()[-1]); print("N=1 ms/step %.3f value %.3e"%(d["ms_per_step"],d["value"]))
PY
