mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stack or transformer_block or pipeline or bench_object" > gpurun_out/s2c_pytest.log 2>&1; tail -5 gpurun_out/s2c_pytest.log
timeout 300 python tools/trace_tblock.py > gpurun_out/s2c_trace.log 2>&1; grep -A32 "layer 3" gpurun_out/s2c_trace.log
timeout 600 python bench.py --skip-extras > gpurun_out/s2c_bench.json 2> gpurun_out/s2c_bench.err; head -c 400 gpurun_out/s2c_bench.json
