#!/bin/bash
# Round-1 closing measurements on one B200 (run through gpurun); everything lands in gpurun_out/s5_*.
# Benches run WITHOUT a profiler; the ncu passes repeat the same commands afterwards.
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/s5_pytest.log
timeout 200 python bench.py > $O/s5_bench_cfg2.log 2>&1
timeout 200 python bench.py --workload cfg3 > $O/s5_bench_cfg3.log 2>&1
timeout 200 python bench.py --workload cfg1 > $O/s5_bench_cfg1.log 2>&1
timeout 200 python bench.py --math fp32 > $O/s5_bench_cfg2_fp32.log 2>&1
timeout 300 python bench.py --workload cfg4 --gpus 1 --steps 3 --warmup 3 > $O/s5_scale_n1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s5_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 > $O/s5_ncu_a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s5_launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 > $O/s5_ncu_b.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"damsm_fwd2|damsm_bwd3|tc_gemm" --launch-skip 12 -c 4 \
    -o $O/s5_ncu_cfg2_damsm -f python bench.py --steps 1 --warmup 3 --no-graph > $O/s5_ncu_c.log 2>&1
tail -n 3 $O/s5_pytest.log
