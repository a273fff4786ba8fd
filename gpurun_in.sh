cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_r1.json
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1.json
cut -c1-400 gpurun_out/bench_ref_r1.json; echo; cut -c1-300 gpurun_out/bench_r1.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'accumulate|fft_|kgen|get_HI|radial' -s 30 -c 10 -f -o gpurun_out/prof_r1_all python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu2.log 2>&1
ls -la gpurun_out | head -30
