# Round-end measurement on one B200 (run from the repo root through gpurun): bench lines of every workload, the CPU arm, ncu launch
# lists (our kernels only) and one full capture; outputs under gpurun_out/r02k_*.
cd ${GRAFT_REPO_ROOT:-.}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02k_pytest_gpu.log 2>&1; tail -1 gpurun_out/r02k_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1; tail -1 gpurun_out/r02k_smoke.log
for w in metric cfg2 cfg3 cfg4 biased; do
  extra=""; [ $w != metric ] && extra="--no-e2e --no-cpu"
  timeout 600 python bench.py --workload $w $extra > gpurun_out/r02k_bench_$w.log 2> gpurun_out/r02k_bench_$w.err || echo "bench $w failed"
done
timeout 600 python bench.py --impl reference > gpurun_out/r02k_bench_ref.log 2> gpurun_out/r02k_bench_ref.err
K='regex:l1_kernel|quantize_warp|decode_mean|decode_lut|rz_|fwht_|fill_uniforms|bump_seed'
for w in metric cfg2 cfg3 cfg4 biased; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r02k_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu --eager > gpurun_out/r02k_ncu_$w.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k "$K" -s 7 -c 7 -o gpurun_out/r02k_biased python tools/run_once.py 128x16777216 1 biased > gpurun_out/r02k_ncu_full_biased.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k "$K" -s 6 -c 3 -o gpurun_out/r02k_metric python tools/run_once.py 128x16777216 1 > gpurun_out/r02k_ncu_full_metric.log 2>&1
tail -2 gpurun_out/r02k_ncu_full_metric.log
