mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "encode or pipeline or bench_object or config5 or chained" > gpurun_out/s2f_pytest.log 2>&1; tail -15 gpurun_out/s2f_pytest.log
timeout 120 ./tools/tc_rate > gpurun_out/s2f_tc_rate.log 2>&1; tail -12 gpurun_out/s2f_tc_rate.log
for c in blockpush buttonpress; do timeout 600 python bench.py --config $c --skip-extras > gpurun_out/s2f_bench_$c.json 2> gpurun_out/s2f_bench_$c.err; head -c 330 gpurun_out/s2f_bench_$c.json; echo; done
