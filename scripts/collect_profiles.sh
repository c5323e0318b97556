#!/bin/bash
# copies what scripts/profile_round.sh brought back in gpurun_out/ into profiles/ (tracked), with the command lines on top
G=gpurun_out; TAG=${1:-r2}; R=${2:-r02}
hdr() { echo "$1"; echo "($2; under ncu a kernel runs cold-cache and serialised; summary by scripts/ncu_summary.py)"; echo; }
( hdr "ncu --set full --clock-control none --import-source on -k regex:k_step_hmm -s 20 -c 1  python bench.py --steps 2 --warmup 3 --T 40" "round 2, final build; N = 2^24 particles, K = 16"; cat $G/${TAG}_step_hmm_summary.txt ) > profiles/${R}_step_hmm_ncu_full_summary.txt
( hdr "ncu --set full ... -k regex:k_step_lg -s 20 -c 1  python bench.py --workload lg --steps 2 --warmup 3 --T 40" "round 2; linear-Gaussian fp32, N = 2^24"; cat $G/${TAG}_step_lg_summary.txt ) > profiles/${R}_step_lg_ncu_full_summary.txt
( hdr "ncu --set full ... -k regex:k_mh_bnet -s 1 -c 1  python scripts/mh_case.py" "round 2, final build; BASELINE cfg5: 65536 chains x 10^4 steps, Bayes net C,S,R | W, k_mh_bnet_tab"; cat $G/${TAG}_mh_summary.txt ) > profiles/${R}_mh_ncu_full_summary.txt
( hdr "ncu --set full ... -k regex:k_tilescan -s 20 -c 1  python bench.py --steps 2 --warmup 3 --T 40" "round 2; 8192 tiles, one 8-CTA cluster"; cat $G/${TAG}_tilescan_summary.txt ) > profiles/${R}_tilescan_ncu_full_summary.txt
( hdr "ncu --set full ... -k regex:k_sweep_hmm -s 1 -c 1  python scripts/res_trace.py" "round 2, final build; the resident sweep kernel: ONE cooperative launch = 40 steps of 2^20 particles, 512 CTAs"; cat $G/${TAG}_sweep_resident_summary.txt ) > profiles/${R}_sweep_resident_ncu_full_summary.txt
( hdr "ncu --set full ... -k regex:k_count_back_all -s 1 -c 1  python scripts/cfg3_split.py" "round 2; the backward pass of one PMCMC iteration at cfg3's size (2^20 particles, T = 1000): counting phase + walking phase"; cat $G/${TAG}_count_back_summary.txt ) > profiles/${R}_count_back_ncu_full_summary.txt
cp $G/${TAG}_launches.csv profiles/${R}_launch_list.csv
( echo "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  python bench.py --steps 2 --warmup 3 --T 40"; echo "(round 2, final build; first 400 launches = 5 sweeps of 40 steps at N = 2^24; per-launch times are cold-cache and serialised: compare SHARES)"; echo; python scripts/ncu_summary.py launches $G/${TAG}_launches.csv ) > profiles/${R}_launch_list_summary.txt
tail -1 $G/${TAG}_bench.log > profiles/${R}_bench_line.json
tail -1 $G/${TAG}_bench_lg32.log > profiles/${R}_bench_line_lg_fp32.json; tail -1 $G/${TAG}_bench_lg64.log > profiles/${R}_bench_line_lg_fp64.json
cp $G/${TAG}_configs.log profiles/${R}_configs_full_size.jsonl
