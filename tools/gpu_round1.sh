#!/bin/bash
# One gpurun call: parity suite, reference differential, sweep, bench, ncu evidence.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv 2>/dev/null &
SMI=$!
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== ref golden"; timeout 600 python -m tests.golden.make_ref_golden > gpurun_out/ref_golden.log 2>&1; tail -12 gpurun_out/ref_golden.log
echo "== sweep"; timeout 1200 python tools/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.jsonl; tail -5 gpurun_out/sweep.err
echo "== bench"; timeout 900 python bench.py --verbose > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -8 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 200 --warmup 5 --verbose > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full on search kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 20 -c 3 -f -o gpurun_out/prof_search \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "== ncu full on insert kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:insert_flat_kernel -s 40 -c 2 -f -o gpurun_out/prof_insert \
    python bench.py --steps 40 --warmup 3 --graph 0 --no-cpu > gpurun_out/ncu_full_ins.log 2>&1
kill $SMI
ls -la gpurun_out
