#!/bin/bash
# forward attention kernel: CTAs x ring depth x ablation (tuning aid, run under gpurun)
run() { AGB_ATTN_FWD_CTAS=$1 AGB_ATTN_FWD_STAGES=$2 AGB_ATTN_DEBUG=$3 AGB_ATTN_BWD_CPS=4 python bench.py --workload cfg3 --hw 128 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('ctas $1 stages $2 dbg $3 fwd', round(r['fwd_gbs']), 'us', round(172e6/r['fwd_gbs']/1e3,1))"; }
for dbg in 16 0; do
  run 444 3 $dbg; run 444 4 $dbg; run 296 4 $dbg; run 296 6 $dbg; run 296 8 $dbg; run 148 8 $dbg
done
