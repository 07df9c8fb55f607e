# final single-GPU evidence of round 2 (run under gpurun): bench line, reference arm, launch list of the timed steps, other configs
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2g_bench_reference_arm.json 2> gpurun_out/r2g_ref.err
python tools/other_configs.py --all > gpurun_out/r2g_other_configs.txt 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
tail -c 600 gpurun_out/r2g_bench.json; grep -c merge_pf gpurun_out/r2g_launches.csv
