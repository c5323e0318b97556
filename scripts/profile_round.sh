set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 python bench.py > gpurun_out/s4_bench.log 2> gpurun_out/s4_bench.err
timeout 300 python bench.py --workload lg > gpurun_out/s4_bench_lg32.log 2>&1
timeout 300 python bench.py --workload lg --precision fp64 > gpurun_out/s4_bench_lg64.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --T 40 > gpurun_out/ncu_launch_r2.log 2>&1
timeout 300 $NCU -k regex:k_step_lg -s 20 -c 1 -o gpurun_out/prof_step_lg_r2 python bench.py --workload lg --steps 2 --warmup 3 --T 40 > gpurun_out/ncu_lg_r2.log 2>&1
timeout 300 $NCU -k regex:k_mh_bnet -s 1 -c 1 -o gpurun_out/prof_mh_r2 python scripts/mh_case.py > gpurun_out/ncu_mh_r2.log 2>&1
timeout 300 $NCU -k regex:k_tilescan -s 20 -c 1 -o gpurun_out/prof_tilescan_r2 python bench.py --steps 2 --warmup 3 --T 40 > gpurun_out/ncu_ts_r2.log 2>&1
timeout 300 $NCU -k regex:k_sweep_hmm -s 1 -c 1 -o gpurun_out/prof_sweep_res_r2b python scripts/res_trace.py > gpurun_out/ncu_res_r2b.log 2>&1
timeout 300 $NCU -k regex:k_count_back_all -s 1 -c 1 -o gpurun_out/prof_count_back_r2 python scripts/cfg3_split.py > gpurun_out/ncu_cb_r2.log 2>&1
timeout 600 python scripts/run_configs.py > gpurun_out/configs_r2.log 2>&1
timeout 60 python scripts/mh_case.py > gpurun_out/mh_case_r2.log 2>&1
timeout 60 python scripts/cfg3_split.py > gpurun_out/cfg3_split_r2.log 2>&1
ls -la gpurun_out/*r2* | tail -20
