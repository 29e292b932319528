mkdir -p gpurun_out
for B in 8 64; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2e_train_b${B}.csv python bench.py --mode train --batch $B --steps 1 --warmup 3 --skip-extras > gpurun_out/s2e_ncu_b${B}.log 2>&1
done
ls -la gpurun_out/s2e*
