python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r02bf_bench_n1.jsonl 2> gpurun_out/r02bf_bench_n1.err; echo bench rc=$?
python tools/determinism_probe.py 2>&1 | tail -3
