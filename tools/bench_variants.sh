#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v['avg_ms'],3) for k,v in d['kernels'].items()}, 'fwd_frac', round(d['roofline_path']['forward']['frac_of_hbm_peak'],4))"; }
run MG_PREFETCH=1 MG_MINB_B=2
run MG_PREFETCH=1 MG_MINB_B=1
run MG_PREFETCH=0 MG_MINB_B=1
run MG_PREFETCH=1 MG_MINB_B=1 MG_CHUNKS=4
