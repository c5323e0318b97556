#!/bin/bash
# One gpurun call: the round's bench lines, launch list, ncu --set full captures of every hot kernel (reports stay in
# /tmp on the GPU box -- gpurun_out/ carries back at most 64 MiB -- and come back as text summaries + gzipped source pages),
# the five BASELINE configs at full size.     usage (on the box): bash scripts/profile_round.sh [tag]
set -x
cd ${GRAFT_REPO_ROOT:-.}
TAG=${1:-r2}
P=/tmp/prof; mkdir -p $P gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # cap <name> <kernel regex> <skip> <command...>
  local name=$1 re=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$re -s $skip -c 1 -o $P/$name "$@" > gpurun_out/ncu_${name}.log 2>&1
  python scripts/ncu_summary.py full $P/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  ncu -i $P/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
}
timeout 300 python bench.py > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload lg > gpurun_out/${TAG}_bench_lg32.log 2>&1
timeout 300 python bench.py --workload lg --precision fp64 > gpurun_out/${TAG}_bench_lg64.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --T 40 > gpurun_out/ncu_launch_${TAG}.log 2>&1
cap ${TAG}_step_hmm k_step_hmm 20 python bench.py --steps 2 --warmup 3 --T 40
cap ${TAG}_step_lg k_step_lg 20 python bench.py --workload lg --steps 2 --warmup 3 --T 40
cap ${TAG}_mh k_mh_bnet 1 python scripts/mh_case.py
cap ${TAG}_tilescan k_tilescan 20 python bench.py --steps 2 --warmup 3 --T 40
cap ${TAG}_sweep_resident k_sweep_hmm 1 python scripts/res_trace.py
cap ${TAG}_count_back k_count_back_all 1 python scripts/cfg3_split.py
timeout 600 python scripts/run_configs.py > gpurun_out/${TAG}_configs.log 2>&1
timeout 60 python scripts/mh_case.py > gpurun_out/${TAG}_mh_case.log 2>&1
timeout 60 python scripts/cfg3_split.py > gpurun_out/${TAG}_cfg3_split.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -40
