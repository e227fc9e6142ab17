#!/bin/bash
# cfg3 attention kernel timings for a few CTAs-per-sample settings (tuning aid, run under gpurun)
for f in "" 4 5 6 7 8 12; do for b in ""; do
  AGB_ATTN_FWD_CPS=$f AGB_ATTN_BWD_CPS=$f python bench.py --workload cfg3 --hw 128 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cps=$f', 'ms', round(d['ms_per_step'],4), 'fwd', round(r['fwd_gbs']), 'bwd', round(r['bwd_gbs']), 'frac', round(r['frac'],3))"
done; done
