set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s2a_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/s2a_bench.json 2> gpurun_out/s2a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2a_ref.json 2> gpurun_out/s2a_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s2a_launches.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/s2a_ncu.log 2>&1
tail -3 gpurun_out/s2a_pytest.log; tail -2 gpurun_out/s2a_smoke.log; head -c 600 gpurun_out/s2a_bench.json
