# development script: 8-GPU bench lines (usage: gpurun --gpus 8 -- bash tools/gpu_bench_8gpu.sh)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --skip-extras > gpurun_out/val_n8_gridshapes5.json 2> gpurun_out/val_n8_gridshapes5.err; head -c 250 gpurun_out/val_n8_gridshapes5.json; echo
timeout 400 $TR bench.py --gpus 8 --config sketchy_hires --skip-extras > gpurun_out/val_n8_sketchy_hires.json 2> gpurun_out/val_n8_sketchy_hires.err; head -c 250 gpurun_out/val_n8_sketchy_hires.json; echo
timeout 400 $TR bench.py --gpus 8 --mode train --train-arith bf16 --skip-extras > gpurun_out/val_n8_train_bf16.json 2> gpurun_out/val_n8_train_bf16.err; head -c 250 gpurun_out/val_n8_train_bf16.json; echo
timeout 400 $TR bench.py --gpus 8 --mode train --skip-extras > gpurun_out/val_n8_train_fp32.json 2> gpurun_out/val_n8_train_fp32.err; head -c 250 gpurun_out/val_n8_train_fp32.json; echo
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 400 $TR4 bench.py --gpus 4 --mode train --train-arith bf16 --skip-extras > gpurun_out/val_n4_train_bf16.json 2> gpurun_out/val_n4_train_bf16.err; head -c 250 gpurun_out/val_n4_train_bf16.json; echo
