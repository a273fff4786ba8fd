cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
GH_FFT_BATCH_MB=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fft_' -s 12 -c 3 -f -o gpurun_out/prof_fft1024 python bench.py --grid 1024 --nside 512 --shells 150 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
