#!/bin/bash
# round-2 closing GPU pass: ablation sweep of the one-pass kernel, ncu capture of the stand-alone belief kernel in the padded layout,
# the whole GPU suite, smoke() and the default bench
mkdir -p gpurun_out
for d in 0 1 2 4 6 8 14 15; do
  echo "OLFNAV_F2_DBG=$d" >> gpurun_out/r2q_fused2_ablation.log
  OLFNAV_F2_DBG=$d timeout 120 python scripts/fused2_profile.py 100000 75 5 2>&1 | tail -3 >> gpurun_out/r2q_fused2_ablation.log
done
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2q_gpu_tests.log 2>&1; tail -5 gpurun_out/r2q_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 600 gpurun_out/r2q_bench.json
python bench.py --config C4 --steps 3 --warmup 1 --no-backup --no-cpu-baseline > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:belief_step_tma -s 2 -c 1 -o gpurun_out/r2q_belief python bench.py --config C4 --steps 3 --warmup 1 --no-backup --no-cpu-baseline > gpurun_out/r2q_ncu.log 2>&1
ls -la gpurun_out/r2q*
