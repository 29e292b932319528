mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q > gpurun_out/s2g_pytest.log 2>&1; tail -5 gpurun_out/s2g_pytest.log
timeout 300 python tools/profile_train_aten.py 8 > gpurun_out/s2g_aten.log 2>&1; tail -5 gpurun_out/s2g_aten.log
timeout 600 python bench.py --mode train --skip-extras > gpurun_out/s2g_train.json 2> gpurun_out/s2g_train.err; head -c 300 gpurun_out/s2g_train.json
