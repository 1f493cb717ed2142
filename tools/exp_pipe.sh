out=gpurun_out/r3f; mkdir -p $out
MG_SWEEPA_PIPE=1 timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused_variants.py tests/test_adjoint_relation.py -m gpu -q -p no:cacheprovider > $out/pytest_pipe.log 2>&1; echo "pytest(pipe) rc=$?"; tail -2 $out/pytest_pipe.log
for rep in 1 2; do
for v in 0 1; do
  MG_SWEEPA_PIPE=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/b_${v}_${rep}.json 2> $out/b_${v}_${rep}.err
  python - $out/b_${v}_${rep}.json $v <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("pipe", sys.argv[2], "ms/step %.3f" % d["ms_per_step"], {n:round(v["avg_ms"],4) for n,v in d["kernels"].items()}, d["clocks"]["sm_mhz"])
PY
done; done
