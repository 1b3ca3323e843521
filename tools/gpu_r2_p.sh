#!/bin/bash
# Round 2, final pass on one GPU: parity tests, smoke, bench lines of the final build (forward, loop-closure mode, reference arm), ncu launch list + full capture.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/p_tests.log
tail -6 gpurun_out/p_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"
head -c 400 gpurun_out/p_bench.json; echo
timeout 600 python bench.py > gpurun_out/p_bench_default.json 2> gpurun_out/p_bench_default.err; echo "bench default rc=$?"
head -c 300 gpurun_out/p_bench_default.json; echo
timeout 600 python bench.py --lc-mode const_weight --steps 20 --warmup 3 > gpurun_out/p_bench_lc.json 2> gpurun_out/p_bench_lc.err; echo "lc rc=$?"
head -c 300 gpurun_out/p_bench_lc.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/p_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gn_track -s 2 -c 1 -f -o gpurun_out/prof_r2_p python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/p_ncu_full.log 2>&1; echo "ncu full rc=$?"
