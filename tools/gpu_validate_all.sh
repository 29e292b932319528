# development script: the whole GPU validation in one gpurun call -- pytest -m gpu, smoke(), bench.py for every workload (usage: gpurun -- bash tools/gpu_validate_all.sh)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/val_pytest.log 2>&1; tail -3 gpurun_out/val_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/val_smoke.log 2>&1; tail -2 gpurun_out/val_smoke.log
timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; head -c 300 gpurun_out/val_bench.json; echo
for c in blockpush buttonpress sketchy sketchy_hires; do timeout 900 python bench.py --config $c > gpurun_out/val_bench_$c.json 2> gpurun_out/val_bench_$c.err; head -c 300 gpurun_out/val_bench_$c.json; echo; done
