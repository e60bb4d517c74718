set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r02as_bench_n1.jsonl 2> gpurun_out/r02as_bench_n1.err; echo bench rc=$?
PIO_BANK=65536 PIO_STEPS=10 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 700 --csv --log-file gpurun_out/r02ar_decode_warm.csv python tools/stage_probe.py text 64 518 1 > gpurun_out/r02ar.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/r02at_fc1 -f python tools/gemm_one.py 87936 3072 768 gelu > gpurun_out/r02at.log 2>&1; echo rc=$?
