mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s2q_pytest.log 2>&1; tail -4 gpurun_out/s2q_pytest.log
timeout 900 python bench.py > gpurun_out/s2q_bench.json 2> gpurun_out/s2q_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2q_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["clocks"]["sm_mhz"], {k:round(v*d["ms_per_step"],2) for k,v in d["kernel_time_shares"].items()})
c=d["slot_attention_config2"]
print({k:(round(v["forward_ms"],4), round(v.get("reference_torch_cuda_ms",0),3), round(v.get("vs_reference_torch_cuda",0),2)) for k,v in c.items() if k.startswith("S")})
print(c["encode_only"])
PY
