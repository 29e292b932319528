mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2r_pytest.log 2>&1; tail -3 gpurun_out/s2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s2r_smoke.log 2>&1; tail -2 gpurun_out/s2r_smoke.log
timeout 900 python bench.py > gpurun_out/s2r_bench.json 2> gpurun_out/s2r_bench.err; head -c 300 gpurun_out/s2r_bench.json; echo
for c in blockpush buttonpress sketchy sketchy_hires; do timeout 900 python bench.py --config $c > gpurun_out/s2r_bench_$c.json 2> gpurun_out/s2r_bench_$c.err; head -c 300 gpurun_out/s2r_bench_$c.json; echo; done
