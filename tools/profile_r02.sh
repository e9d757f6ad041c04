#!/bin/bash
# Round-2 evidence capture (run under gpurun, one GPU): launch list of the bench command + one `ncu --set full`
# capture of every stage kernel that ships.  Outputs under gpurun_out/; summaries are made with tools/ncu_report.py /
# tools/ncu_launch_summary.py / tools/ncu_by_line.py in the build container and committed under profiles/.
set -x
cd "$(dirname "$0")/.."
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${PFX:-r02}_launches_cfg4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${PFX:-r02}_launches_bench.log 2>&1
for k in mt_fft_kernel csm_tc_ta_kernel power_from_csm_kernel coherence_epilogue_vec_kernel granger_herm_kernel; do
    $NCU --set full --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${PFX:-r02}_$k \
        python bench.py --workload cfg4w8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${PFX:-r02}_ncu_$k.log 2>&1
done
$NCU --set full --import-source on -k regex:csm_kernel -s 1 -c 1 -f -o gpurun_out/${PFX:-r02}_pli_csm_kernel \
    python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${PFX:-r02}_ncu_pli.log 2>&1
ls -la gpurun_out/*.ncu-rep
