mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --skip-extras > gpurun_out/s2s_n8_gridshapes5.json 2> gpurun_out/s2s_n8_gridshapes5.err; head -c 250 gpurun_out/s2s_n8_gridshapes5.json; echo
timeout 400 $TR bench.py --gpus 8 --config sketchy_hires --skip-extras > gpurun_out/s2s_n8_sketchy_hires.json 2> gpurun_out/s2s_n8_sketchy_hires.err; head -c 250 gpurun_out/s2s_n8_sketchy_hires.json; echo
