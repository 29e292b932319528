mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/s2d_bench.json 2> gpurun_out/s2d_bench.err; tail -c 1500 gpurun_out/s2d_bench.err
timeout 600 python bench.py --mode train > gpurun_out/s2d_train.json 2> gpurun_out/s2d_train.err; tail -c 600 gpurun_out/s2d_train.err
timeout 600 python bench.py --mode train --train-arith bf16 > gpurun_out/s2d_train_bf16.json 2> gpurun_out/s2d_train_bf16.err
timeout 600 python bench.py --mode train --batch 64 > gpurun_out/s2d_train_b64.json 2> gpurun_out/s2d_train_b64.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2d_bench.json").read().strip().splitlines()[-1])
print(json.dumps(d["slot_attention_config2"],indent=1)[:3500])
for f in ("s2d_train","s2d_train_bf16","s2d_train_b64"):
    t=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, t["ms_per_step"], t["value"], t.get("torch_cuda_baseline"), t.get("vs_torch_cuda"))
PY
