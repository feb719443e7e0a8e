#!/bin/bash
# Round measurement pass (run under gpurun): bench lines, ncu launch lists and one full capture per dominant kernel.
# Outputs land in gpurun_out/ (scratch); tools/ncu_summary.py turns them into profiles/*.md here.
R=${1:-r2}
O=gpurun_out
mkdir -p $O
B="--no-cpu-baseline --also none --sustain 0"
# the default driver line (cfg2 + workloads sub-records + cpu baseline), then one full line per workload
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${R}_bench_default.json 2> $O/${R}_bench_default.err
for w in cfg5 cfg4 cfg1m cfg1l cfg1b cfg3; do
  timeout 600 python bench.py --workload $w $B > $O/${R}_bench_$w.json 2> $O/${R}_bench_$w.err
  tail -c 300 $O/${R}_bench_$w.json; echo
done
for w in cfg2 cfg5 cfg4 cfg1m cfg1b cfg1l cfg3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_${w}_launches.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-e2e $B > /dev/null 2>&1
done
cap() {  # workload, kernel regex, skip, count, tag
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o $O/${R}_$5 \
    python bench.py --workload $1 --steps 2 --warmup 3 --no-e2e $B > /dev/null 2>&1
}
cap cfg2 stft_tdoa_warp_kernel 3 1 cfg2_stft_tdoa
cap cfg5 "srp_tc_small_kernel|ds_select_kernel|stft_hw_kernel" 9 3 cfg5_srp
cap cfg4 "srp_tc_kernel|srp_prepare_kernel" 6 2 cfg4_srp_tc
cap cfg3 "ds_fan_tc_kernel" 3 1 cfg3_ds_fan
cap cfg1m "mask_fused_kernel" 3 1 cfg1m_fused
cap cfg1b "mb_fused_kernel|mb_gate_kernel" 6 2 cfg1b_kernels
cap cfg1l "gcc_tau_tc_kernel|curve_scan|stft_hw_kernel" 9 3 cfg1l_kernels
ls -la $O | grep "${R}_" | tail -40
