timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_bench_reference_final.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 170 --csv --log-file gpurun_out/r2_launches_final.csv python tools/profile_target.py 64 > gpurun_out/pt1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_inter_mb|csc_bgra|k_cavlc_mb|k_slice|k_pack' -s 160 -c 15 -f -o gpurun_out/r2_full_final python tools/profile_target.py 40 > gpurun_out/pt2.log 2>&1
ncu -i gpurun_out/r2_full_final.ncu-rep --page raw --csv > gpurun_out/r2_full_final_raw.csv 2>/dev/null
rm -f gpurun_out/r2_full_final.ncu-rep
head -c 1500 gpurun_out/r2_bench_final.json
