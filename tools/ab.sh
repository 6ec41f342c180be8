#!/bin/bash
# Developer tool: A/B bench lines of kernel variants selected by environment variables (one line per variant).
# usage: tools/ab.sh OUT "VAR=VAL ..." "VAR=VAL ..." ...      (extra bench flags in $BENCH_FLAGS)
out=$1; shift
: > "$out"
for v in "$@"; do
  echo "== $v" >> "$out"
  env $v python bench.py --steps ${STEPS:-200} --warmup 5 --no-cpu --e2e-steps 2 $BENCH_FLAGS 2>>"$out.err" | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    s=d.get('secondary_n20_rank0') or {}
    print(json.dumps({'ms':d['ms_per_step'],'value':d['value'],'frac':d['roofline']['frac'],'n20_ms':s.get('ms_per_step'),'n20_frac':s.get('roofline_frac'),'clk':d['clocks']['sm_mhz']}))
" >> "$out"
done
cat "$out"
