cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
cat gpurun_out/accumulate_audit_256.json | tr '\n' ' '
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench6.log
python -c "
import json;d=json.loads(open('gpurun_out/bench6.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value']);print(d['roofline']['stage_ms_alone'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'accumulate' -s 3 -c 1 -f -o gpurun_out/prof_r1e_acc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu3.log 2>&1
