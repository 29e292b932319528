mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mha or config5 or pipeline" > gpurun_out/s2l_pytest.log 2>&1; tail -5 gpurun_out/s2l_pytest.log
timeout 600 python bench.py --config sketchy_hires --skip-extras > gpurun_out/s2l_bench_hires.json 2> gpurun_out/s2l_bench_hires.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2l_bench_hires.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["clocks"], {k:round(v*d["ms_per_step"],2) for k,v in d["kernel_time_shares"].items()})
PY
