mkdir -p gpurun_out
python tools/trace_tblock.py > gpurun_out/s2b_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tblock_stack -s 2 -c 1 -o gpurun_out/s2b_tblock -f python tools/trace_tblock.py > gpurun_out/s2b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv5x5_wgrad_tc -s 1 -c 1 -o gpurun_out/s2b_wgrad -f python tools/bench_savi_train.py > gpurun_out/s2b_ncu2.log 2>&1
tail -60 gpurun_out/s2b_trace.log
