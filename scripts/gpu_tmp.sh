O=gpurun_out/r2final
mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > $O/memcheck.log 2>&1; grep -c "Invalid\|out of bounds\|misaligned" $O/memcheck.log; tail -3 $O/memcheck.log
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > $O/racecheck.log 2>&1; tail -2 $O/racecheck.log
