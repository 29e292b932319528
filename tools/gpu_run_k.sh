mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s2k_pytest.log 2>&1; tail -5 gpurun_out/s2k_pytest.log
timeout 600 python bench.py --skip-extras > gpurun_out/s2k_bench.json 2> gpurun_out/s2k_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2k_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["clocks"], {k:round(v*d["ms_per_step"],2) for k,v in d["kernel_time_shares"].items()})
PY
timeout 600 python bench.py --config buttonpress --skip-extras > gpurun_out/s2k_bench_bp.json 2> gpurun_out/s2k_bench_bp.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2k_bench_bp.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["clocks"], {k:round(v*d["ms_per_step"],2) for k,v in d["kernel_time_shares"].items()})
PY
