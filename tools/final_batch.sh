out=gpurun_out/${1:-r4c}; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench_1.json 2> $out/bench_1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c1 --steps 3 --warmup 1 > $out/bench_c1.json 2> $out/bench_c1.err; echo "bench c1 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench ref rc=$?"
python - $out/bench_1.json $out/bench_c1.json $out/bench_ref.json <<'PY'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k={n:round(v["avg_ms"],3) for n,v in (d.get("kernels") or {}).items()}
        print(f, "ms/step %.3f value %.4e e2e %.4e" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value",0)), k, (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print("no json", f, e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1; echo "launches rc=$?"
