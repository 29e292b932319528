mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s2t_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/s2t_memcheck.log
