export PIO_VIT_PDL=1
for i in 1 2 3; do python -m pytest tests -m gpu -q -k "determin or reproducible" 2>&1 | tail -2; done
python -m pytest tests -m gpu -q 2>&1 | tail -3
TAG=vit_pdl_on python tools/determinism_probe.py 2>&1 | tail -3
python tools/stage_probe.py vit 64 518 3 | tail -2
PIO_VIT_PDL=0 python tools/stage_probe.py vit 64 518 3 | tail -2
