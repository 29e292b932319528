mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_savi_training.py -m gpu -x -q > gpurun_out/s2y_pytest.log 2>&1; tail -3 gpurun_out/s2y_pytest.log | cut -c1-250
timeout 600 python tools/bench_savi_train.py > gpurun_out/s2y_savi_train.json 2> gpurun_out/s2y_savi_train.err; cat gpurun_out/s2y_savi_train.json | cut -c1-1000
