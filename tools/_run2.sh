for db in 0 1; do echo "PIO_ATTN_DB=$db"; PIO_ATTN_DB=$db python tools/attn_probe.py 64 1374 10; done
for db in 0 2; do echo "PIO_ATTN_DB=$db"; PIO_ATTN_DB=$db python tools/attn_probe.py 256 261 10; done
PIO_ATTN_DB=1 timeout 600 python -m pytest tests -m gpu -x -q -k "attention or vit" 2>&1 | tail -5
PIO_ATTN_DB=2 timeout 600 python -m pytest tests -m gpu -x -q -k "attention or vit" 2>&1 | tail -5
