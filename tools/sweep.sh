#!/bin/bash
# Developer tool: bench every variant library at N=50 and N=20 (short runs), one JSON line each.
# usage: tools/sweep.sh out.log tag1 tag2 ...
out=$1; shift
PKG=$(ls -d /root/repo/*_b200)
: > $out
for tag in "$@"; do
  lib=$PKG/variants/libuavsim_$tag.so
  [ "$tag" = product ] && lib=$PKG/libuavsim.so
  for n in 50 20; do
    r=$(UAVSIM_LIB=$lib python bench.py --steps 300 --warmup 5 --no-cpu --e2e-steps 2 --sensors $n 2>&1 | tail -1)
    echo "$tag N=$n $(echo "$r" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f M  %.4f ms  frac %.4f  clk %s' % (d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" 2>&1 | tail -1)" >> $out
  done
done
