mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02_tests.log 2>&1
(timeout 300 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_score|k_distmap|k_select|k_prep_lines|k_vp_support" -s 5 -c 5 -o gpurun_out/r02_detect python tools/distmap_phases.py > gpurun_out/r02_detect_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_chol_solve" -c 2 -o gpurun_out/r02_chol python tools/ba_time.py 2 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_ed_draw_warp|k_ed_fit_warp|k_ed_emit" -c 3 -o gpurun_out/r02_edlines python tests/diag/edlines_time.py > /dev/null 2>&1
tail -3 gpurun_out/r02_tests.log
head -c 400 gpurun_out/r02_bench.json
