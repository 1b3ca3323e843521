#!/bin/bash
# Round 2, first GPU pass: parity tests, smoke, bench (pipelined / serialised batches), ncu launch list + one full capture.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -40 gpurun_out/a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
cat gpurun_out/a_bench.json | head -c 3000
ELLC_OVERLAP=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/a_bench_nooverlap.json 2> gpurun_out/a_bench_nooverlap.err; echo "bench nooverlap rc=$?"
cat gpurun_out/a_bench_nooverlap.json | head -c 600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/a_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track -s 2 -c 1 -f -o gpurun_out/prof_r2_a python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/a_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
