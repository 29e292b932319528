mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2p_pytest.log 2>&1; tail -3 gpurun_out/s2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s2p_smoke.log 2>&1; tail -2 gpurun_out/s2p_smoke.log
timeout 900 python bench.py > gpurun_out/s2p_bench.json 2> gpurun_out/s2p_bench.err; head -c 300 gpurun_out/s2p_bench.json; echo
for c in blockpush buttonpress sketchy sketchy_hires; do timeout 900 python bench.py --config $c > gpurun_out/s2p_bench_$c.json 2> gpurun_out/s2p_bench_$c.err; head -c 300 gpurun_out/s2p_bench_$c.json; echo; done
ncu --set full --clock-control none --import-source on -k regex:tblock_stack -s 2 -c 1 -o gpurun_out/s2p_tblock -f python tools/trace_tblock.py > gpurun_out/s2p_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_composite_tc -s 1 -c 1 -o gpurun_out/s2p_composite -f python tools/ncu_target.py composite > gpurun_out/s2p_ncu2.log 2>&1
