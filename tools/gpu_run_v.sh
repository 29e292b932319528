mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2v_pytest.log 2>&1; tail -2 gpurun_out/s2v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s2v_smoke.log 2>&1; tail -1 gpurun_out/s2v_smoke.log
timeout 900 python bench.py > gpurun_out/s2v_bench.json 2> gpurun_out/s2v_bench.err; head -c 300 gpurun_out/s2v_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2v_ref.json 2> gpurun_out/s2v_ref.err; head -c 200 gpurun_out/s2v_ref.json; echo
timeout 600 python bench.py --mode train > gpurun_out/s2v_train.json 2> gpurun_out/s2v_train.err; head -c 250 gpurun_out/s2v_train.json; echo
timeout 600 python bench.py --mode train --train-arith bf16 > gpurun_out/s2v_train_bf16.json 2> gpurun_out/s2v_train_bf16.err; head -c 250 gpurun_out/s2v_train_bf16.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/s2v_launches.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/s2v_ncu.log 2>&1
