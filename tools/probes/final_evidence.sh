set -x
python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2g_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_1080p-10s.json 2> gpurun_out/r2g_bench_1080p-10s.err; cut -c1-160 gpurun_out/r2g_bench_1080p-10s.json
for wl in 540p-8s 2160p-20s lsvq-mix; do python bench.py --steps 10 --warmup 3 --workload $wl > gpurun_out/r2g_bench_$wl.json 2> gpurun_out/r2g_bench_$wl.err; cut -c1-160 gpurun_out/r2g_bench_$wl.json; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k4_pyr_fused -s 3 -c 1 -o gpurun_out/r2g_pyr_fused -f python tools/flow_ab.py --impls 0 --reps 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k5_flow_rgb_patchsum_band -s 3 -c 1 -o gpurun_out/r2g_rgb_band -f python tools/rgb_ab.py --child /tmp/x > /dev/null 2>&1
compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/r2g_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r2g_sanitizer_memcheck.log
