#!/bin/bash
run() { AGB_BENCH_FLUSH=$3 AGB_ATTN_DEBUG=$2 AGB_ATTN_BWD_CPS=4 python bench.py --workload cfg3 --hw 128 --batch $1 --steps 6 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('batch $1 dbg $2 flush $3 fwd', round(r['fwd_gbs']), 'us', round($1/64*172e6/r['fwd_gbs']/1e3,1), 'bwd', round(r['bwd_gbs']))"; }
for f in write write_read; do run 64 16 $f; run 64 0 $f; run 128 0 $f; done
