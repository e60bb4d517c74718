python tools/beam_probe.py 64 20 12 64 5 2>&1 | tee gpurun_out/r02az_beam_probe.txt
python tools/beam_probe.py 512 20 12 64 5 2>&1 | tee -a gpurun_out/r02az_beam_probe.txt
python bench.py --batch 8 --boxes 4 --size 224 --steps 20 --warmup 5 > gpurun_out/r02ba_bench_n1_config0.jsonl 2>gpurun_out/r02ba.err; echo rc=$?
python bench.py --workload regionset-viecap --steps 5 --warmup 3 > gpurun_out/r02ba_bench_n1_regionset_viecap.jsonl 2>>gpurun_out/r02ba.err; echo rc=$?
python bench.py --workload traces --steps 5 --warmup 3 > gpurun_out/r02ba_bench_n1_traces.jsonl 2>>gpurun_out/r02ba.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/r02ay_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-eager --no-parity-sample > gpurun_out/r02ay.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:vit_attention -s 2 -c 1 -o gpurun_out/r02bb_attention -f python tools/attn_probe.py 64 1374 2 > gpurun_out/r02bb.log 2>&1; echo rc=$?
