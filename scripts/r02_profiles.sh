#!/bin/bash
# Round-2 ncu evidence, one tool invocation per GPU call:  gpurun --timeout 900 -- 'bash scripts/r02_profiles.sh list|search|run|features'
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
case "${1:-list}" in
list)
	python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_list_plain.log 2> gpurun_out/r02_list_plain.err || { tail -n 5 gpurun_out/r02_list_plain.err; exit 1; }
	ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_list_ncu.log 2>&1
	wc -l gpurun_out/r02_launches.csv ;;
search)
	python scripts/prof_search.py 50000 10 2 > gpurun_out/r02_search_plain.log 2>&1 || { tail -n 5 gpurun_out/r02_search_plain.log; exit 1; }
	ncu --set full --clock-control none --import-source on -k regex:'k_sweep_ss|k_flip_prefix|k_partition2|k_rank_class|k_keys|k_pass_table|k_level_decide|k_level_jobs|k_scaf_sides|k_finalize_terminal|k_rs_scatter|k_scan_onepass' -c 24 -o gpurun_out/r02_search python scripts/prof_search.py 50000 10 1 > gpurun_out/r02_search_ncu.log 2>&1
	ls -la gpurun_out/r02_search.ncu-rep ;;
run)
	python scripts/prof_search.py 50000 10 2 > gpurun_out/r02_run_plain.log 2>&1 || { tail -n 5 gpurun_out/r02_run_plain.log; exit 1; }
	ncu --set full --clock-control none --import-source on -k regex:'k_sweep_ss|k_flip_prefix|k_partition2|k_pass_table|k_level_decide|k_level_jobs|k_scaf_sides|k_finalize_terminal|k_count_low|k_reduce_best' -c 20 -o gpurun_out/r02_run python scripts/prof_search.py 50000 10 1 > gpurun_out/r02_run_ncu.log 2>&1
	ls -la gpurun_out/r02_run.ncu-rep ;;
features)
	python scripts/prof_coverage.py > gpurun_out/r02_features_plain.log 2>&1 || { tail -n 5 gpurun_out/r02_features_plain.log; exit 1; }
	ncu --set full --clock-control none --import-source on -k regex:'k_cov_sum|k_cov_quotient|k_cov_tile_filter|k_cov_pairs|k_cov_accumulate|k_kmer|k_pack$|k_seg_fill' -c 14 -o gpurun_out/r02_features python scripts/prof_coverage.py > gpurun_out/r02_features_ncu.log 2>&1
	ls -la gpurun_out/r02_features.ncu-rep ;;
esac
