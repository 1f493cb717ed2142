#!/bin/bash
# One GPU-box batch: parity tests, bench variants (kernel generations / tile heights), ncu captures.
# usage: tools/gpu_batch.sh <tag> [pytest|bench|ncu ...]   (default: all three)
tag=${1:-batch}; shift
what=${@:-pytest bench ncu}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
for w in $what; do
case $w in
pytest)
  timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $out/pytest.log 2>&1
  echo "pytest rc=$?"; tail -5 $out/pytest.log ;;
bench)
  i=0
  while IFS= read -r cfg; do
    [ -z "$cfg" ] && continue
    i=$((i+1))
    env $cfg timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --no-cpu-baseline > $out/bench_$i.json 2> $out/bench_$i.err
    echo "== $i: $cfg rc=$?"
    python - "$out/bench_$i.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k={n:round(v["avg_ms"],3) for n,v in d["kernels"].items()}
    print("  ms/step %.3f value %.3e e2e %.3e kernels %s fwd_frac %.3f adj_frac %.3f" % (d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value",0), k, d["roofline_path"]["forward"]["frac_of_hbm_peak"], d["roofline_path"].get("adjoint",{}).get("frac_of_hbm_peak",0)))
except Exception as e:
    print("  (no json)", e)
PY
  done <<< "${BENCH_CFGS:-MG_FWD=2}" ;;
ncu)
  for k in ${NCU_KERNELS:-k_sweepBD k_adjoint1v2}; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-1} -c 1 -f -o $out/prof_$k \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu_$k.log 2>&1
    echo "ncu $k rc=$?"
    # the reports are too large to travel in numbers (64 MiB cap on gpurun_out): keep the raw page as CSV
    ncu -i $out/prof_$k.ncu-rep --page raw --csv > $out/raw_$k.csv 2>/dev/null
    ncu -i $out/prof_$k.ncu-rep --page source --csv > $out/source_$k.csv 2>/dev/null
    [ -z "$KEEP_REP" ] && rm -f $out/prof_$k.ncu-rep
  done ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_bench.log 2>&1
  echo "launches rc=$?" ;;
esac
done
