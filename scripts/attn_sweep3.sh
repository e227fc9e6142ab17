#!/bin/bash
# fwd/bwd attention kernel time vs batch (fixed overhead vs streaming rate); tuning aid, run under gpurun
run() { python bench.py --workload cfg3 --hw 128 --batch $1 --steps 6 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('batch $1 fwd', round(r['fwd_gbs']), 'us', round($1/64*172e6/r['fwd_gbs']/1e3,1), 'bwd', round(r['bwd_gbs']), 'us', round($1/64*201.3e6/r['bwd_gbs']/1e3,1), 'step ms', round(d['ms_per_step'],4))"; }
for b in 32 64 128 256; do run $b; done
