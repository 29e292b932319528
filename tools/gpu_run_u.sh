mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "enc_mlp or encode or pipeline" > gpurun_out/s2u_pytest.log 2>&1; tail -3 gpurun_out/s2u_pytest.log
timeout 300 python tools/bench_enc_mlp.py > gpurun_out/s2u_encmlp.log 2>&1; tail -6 gpurun_out/s2u_encmlp.log
