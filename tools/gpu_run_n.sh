mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s2n_pytest.log 2>&1; tail -5 gpurun_out/s2n_pytest.log
for c in gridshapes5 sketchy_hires; do
timeout 600 python bench.py --config $c --skip-extras > gpurun_out/s2n_bench_$c.json 2> gpurun_out/s2n_bench_$c.err; python - <<PY
import json
d=json.loads(open("gpurun_out/s2n_bench_$c.json").read().strip().splitlines()[-1])
print("$c", d["ms_per_step"], d["clocks"]["sm_mhz"], {k:round(v*d["ms_per_step"],2) for k,v in d["kernel_time_shares"].items()})
PY
done
