#!/bin/bash
# One GPU call that refreshes the round's evidence: bench lines (ours + reference arm), the ncu launch list of the bench
# command, and one `ncu --set full` capture per dominant kernel.  Outputs land in gpurun_out/ (scratch); the summaries are
# condensed into profiles/ by tools/make_profile_summary.py afterwards.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:alo_fast -c 1 -f -o gpurun_out/prof_fast_r1 \
    python tools/profile_run.py --config c2 --n 262144 --reps 1 > gpurun_out/ncu_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:alo_clen -c 1 -f -o gpurun_out/prof_clen_r1 \
    python tools/profile_run.py --config c3 --n 32768 --reps 1 > gpurun_out/ncu_clen.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_scan -c 1 -f -o gpurun_out/prof_cn_r1 \
    python tools/profile_cn.py 592 2 > gpurun_out/ncu_cn.log 2>&1
python tools/throughput_table.py > gpurun_out/throughput_r1.txt 2>&1
python tools/time_cn.py 8192 100000 > gpurun_out/time_cn.log 2>&1
python tools/time_c3.py > gpurun_out/time_c3.log 2>&1
python tools/time_c1_latency.py > gpurun_out/c1_latency.txt 2>&1
python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1
python tools/parity_full.py > gpurun_out/parity_full.txt 2>&1
tail -2 gpurun_out/bench_r1.json | cut -c1-400
