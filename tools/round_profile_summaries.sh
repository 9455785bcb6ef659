#!/bin/bash
# Dev-container half of tools/round_profile.sh: condense gpurun_out/*.ncu-rep into profiles/ (committed).
R=${R:-r2}
cd "$(dirname "$0")/.."
it() { grep -E "^ok " "$1" | tail -1 | awk '{print $3}'; }
python tools/make_profile_summary.py launches gpurun_out/launches_$R.csv profiles/${R}_bench_launches.csv
python tools/make_profile_summary.py full gpurun_out/prof_fast_$R.ncu-rep profiles/${R}_alo_fast_A12_ncu_full.csv 262144
python tools/make_profile_summary.py full gpurun_out/prof_seed_$R.ncu-rep profiles/${R}_qd_seed_hybr_ncu_full.csv 262144
python tools/make_profile_summary.py full gpurun_out/prof_clen_$R.ncu-rep profiles/${R}_alo_clen_B24_ncu_full.csv 32768
python tools/make_profile_summary.py full gpurun_out/prof_cn_$R.ncu-rep profiles/${R}_cn_scan_2000_ncu_full.csv 592
python tools/make_profile_summary.py json gpurun_out/prof_fast_$R.ncu-rep profiles/kernel_profile.json alo_fast_A12_kernel 262144 $(it gpurun_out/ncu_fast_$R.log)
python tools/make_profile_summary.py json gpurun_out/prof_seed_$R.ncu-rep profiles/kernel_profile.json qd_seed_hybr_kernel 262144 $(it gpurun_out/ncu_seed_$R.log)
python tools/make_profile_summary.py json gpurun_out/prof_clen_$R.ncu-rep profiles/kernel_profile.json alo_clen_kernel 32768 $(it gpurun_out/ncu_clen_$R.log)
python tools/make_profile_summary.py json gpurun_out/prof_cn_$R.ncu-rep profiles/kernel_profile.json cn_scan_kernel 592 2000
for f in bench_$R.json bench_ref_$R.json throughput_$R.txt seed_report_$R.txt parity_report_$R.txt parity_full_$R.txt c1_latency_$R.txt; do
    [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/${R}_${f/_$R/}
done
# SASS of the hot kernels (instruction-diet work is reviewable against these)
for k in alo_fast_A12_kernelILi16ELb0ELi0ELb0ELb0 qd_seed_hybr_kernel alo_clen_kernelILi1ELi2ELb1 cn_scan_kernelILi8ELb1; do
    for o in fastamericanoptionpricing_b200/csrc/_build/*.o; do
        fn=$(cuobjdump -elf $o 2>/dev/null | grep -o "_ZN4faop[A-Za-z0-9_]*$k[A-Za-z0-9_]*" | head -1)
        [ -n "$fn" ] && cuobjdump -sass -fun "$fn" $o | grep -E "^\s+/\*[0-9a-f]{4}\*/|Function" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' > profiles/${R}_sass_$(echo $k | cut -c1-24).txt
    done
done
ls -la profiles | tail -30
