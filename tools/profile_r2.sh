set -x
python tools/merge_only.py 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:merge_pf_kernel -c 1 -f -o gpurun_out/r2c_merge_pf python tools/merge_only.py 2 > gpurun_out/r2c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lk_iteration_kernel -c 1 -f -o gpurun_out/r2c_lk python tools/merge_only.py 1 > gpurun_out/r2c_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:merge_band_kernel -c 1 -f -o gpurun_out/r2c_merge_band python tools/merge_only.py 1 > gpurun_out/r2c_ncu3.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench_short.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_ncu4.log 2>&1
ls -la gpurun_out | tail -12
