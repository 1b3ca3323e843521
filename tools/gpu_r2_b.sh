#!/bin/bash
# Round 2, second GPU pass (1 GPU): parity tests incl. the C++ shim surface, bench on every BASELINE config, loop-closure mode.
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b_tests.log
tail -30 gpurun_out/b_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
head -c 1500 gpurun_out/b_bench.json; echo
timeout 600 python bench.py --config 720p_single --steps 50 --warmup 5 > gpurun_out/b_bench_720p.json 2> gpurun_out/b_bench_720p.err; echo "720p rc=$?"
head -c 700 gpurun_out/b_bench_720p.json; echo
timeout 600 python bench.py --config 1080p_stress --steps 10 --warmup 3 > gpurun_out/b_bench_1080p.json 2> gpurun_out/b_bench_1080p.err; echo "1080p rc=$?"
head -c 700 gpurun_out/b_bench_1080p.json; echo
timeout 600 python bench.py --lc-mode const_weight --steps 10 --warmup 3 > gpurun_out/b_bench_lc.json 2> gpurun_out/b_bench_lc.err; echo "lc rc=$?"
head -c 700 gpurun_out/b_bench_lc.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; echo "ref rc=$?"
head -c 1500 gpurun_out/b_bench_ref.json; echo
tail -5 gpurun_out/b_bench*.err
