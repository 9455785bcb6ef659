#!/bin/bash
# One GPU call that refreshes the round's evidence: bench lines (ours + reference arm), the ncu launch list of the bench
# command, and one `ncu --set full` capture per dominant kernel.  Outputs land in gpurun_out/ (scratch); afterwards, in the dev
# container, tools/round_profile_summaries.sh condenses them into profiles/ (text summaries + profiles/kernel_profile.json,
# which bench.py reads instead of carrying ncu numbers as literals).
R=${R:-r2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_b_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:alo_fast -c 1 -f -o gpurun_out/prof_fast_$R \
    python tools/profile_run.py --config c2 --n 262144 --reps 1 > gpurun_out/ncu_fast_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qd_seed_hybr -c 1 -f -o gpurun_out/prof_seed_$R \
    python tools/profile_run.py --config c2 --n 262144 --reps 1 > gpurun_out/ncu_seed_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:alo_clen -c 1 -f -o gpurun_out/prof_clen_$R \
    python tools/profile_run.py --config c3 --n 32768 --reps 1 > gpurun_out/ncu_clen_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cn_scan -c 1 -f -o gpurun_out/prof_cn_$R \
    python tools/profile_cn.py 592 2 > gpurun_out/ncu_cn_$R.log 2>&1
python tools/throughput_table.py > gpurun_out/throughput_$R.txt 2>&1
python tools/seed_report.py > gpurun_out/seed_report_$R.txt 2>&1
python tools/parity_report.py > gpurun_out/parity_report_$R.txt 2>&1
python tools/parity_full.py > gpurun_out/parity_full_$R.txt 2>&1
python tools/time_c1_latency.py > gpurun_out/c1_latency_$R.txt 2>&1
tail -c 600 gpurun_out/bench_$R.json
