#!/bin/bash
# One round's measurement set on one B200 (run under gpurun; writes gpurun_out/):  bash tools/profile_round.sh
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
# launch list of the bench step (per-launch times are cold-cache and serialised: only the SHARES are comparable)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-networks > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fi_bwd_rows -s 3 -c 1 -o $O/bench_fi_bwd -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-networks > $O/ncu_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fi_fwd_patch -s 3 -c 1 -o $O/bench_fi_fwd -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-networks > $O/ncu_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fp_pipeline -s 1 -c 1 -o $O/fp_pipeline -f \
    python tools/run_one.py fp_fwd 16 > $O/ncu_fp.log 2>&1
C=64 timeout 300 ncu --set full --clock-control none -k regex:fi_fwd_cols_chunked -s 2 -c 1 -o $O/fi_fwd_c64 -f \
    python tools/run_one.py fi_fwd 1 > $O/ncu_c64.log 2>&1
C=64 timeout 300 ncu --set full --clock-control none -k regex:fi_bwd_chunked -s 1 -c 1 -o $O/fi_bwd_c64 -f \
    python tools/run_one.py fi_bwd 1 > $O/ncu_bwd_c64.log 2>&1
timeout 200 python tools/sweep_fi.py --iters 20 --out $O/sweep_fi.json > $O/sweep_fi.log 2>&1
timeout 200 python tools/sweep_c64.py > $O/sweep_c64.log 2>&1
cat $O/bench_n1.json
