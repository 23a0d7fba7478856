mkdir -p gpurun_out/dbg
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/dbg/sanitizer.log 2>&1
tail -60 gpurun_out/dbg/sanitizer.log
