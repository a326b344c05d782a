mkdir -p gpurun_out
for i in 1 2; do
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r43_tests$i.log 2>&1; echo "run $i rc=$?"
grep -c "frame #" gpurun_out/r43_tests$i.log; tail -n 3 gpurun_out/r43_tests$i.log | cut -c1-200
done
