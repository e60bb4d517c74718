for v in 0 1; do echo RING_STAGING=$v; PIO_GEMM_RING_STAGING=$v PIO_BANK=65536 python tools/stage_probe.py text 64 518 3 2>&1 | tail -2; PIO_GEMM_RING_STAGING=$v python tools/stage_probe.py vit 64 518 3 | tail -2;  done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
