set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-parity > gpurun_out/r02_bench_under_ncu.log 2>&1
# 2. full capture of the dominant kernel (same per-sample behaviour at 65536 samples)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_ws4 -c 1 -s 1 -o gpurun_out/r02_ws4_4096x65536 -f python tools/profile_run.py 4096 65536 0 2 > gpurun_out/ncu_r02c.log 2>&1
# 3. the two-CTA regime
timeout 600 ncu --set full --clock-control none --import-source on -k regex:demod_ws4 -c 1 -s 1 -o gpurun_out/r02_ws4_2cta_9472x65536 -f python tools/profile_run.py 9472 65536 0 2 > gpurun_out/ncu_r02d.log 2>&1
# 4. DRAM traffic of one bench-size launch
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:demod_ws4 -c 1 -s 1 --csv --log-file gpurun_out/r02_ws4_traffic_4096x4000000.csv python tools/profile_run.py 4096 4000000 0 2 > gpurun_out/ncu_r02e.log 2>&1
# 5. sanitizers
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?" >> gpurun_out/r02_sanitizer_$tool.txt; tail -3 gpurun_out/r02_sanitizer_$tool.txt; done
