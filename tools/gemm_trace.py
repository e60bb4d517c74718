"""Where does the time of one small-M dense layer go?  %globaltimer stamps of every CTA of gemm_tc_kernel (1-CTA tcgen05 GEMM).

    python tools/gemm_trace.py build                       # here: compiles gemm_sm100.cu with -DPIO_GEMM_TRACE into tools/_trace/libpio_trace.so
    python tools/gemm_trace.py run [R] [N] [K]             # on the GPU box: a real decode of R rows, stamps of the LAST launch with that (N, K)

Slots: 0 entry | 1 set-up done | 2 producer past griddepcontrol.wait | 3 last load issued | 4 MMA: first stage landed |
5 last commit issued | 6 epilogue: first accumulator ready | 7 later accumulator ready | 8 stores issued | 9 stores drained | 10 exit
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "_trace")
LIB = os.path.join(OUT, "libpio_trace.so")


def build():
    sys.path.insert(0, os.path.join(ROOT, "patch-ioner_b200"))
    import build as B

    B.build()
    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(OUT, "gemm_sm100.o")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + ["-DPIO_GEMM_TRACE", "-c", os.path.join(B.CSRC, "gemm_sm100.cu"), "-o", obj])
    objs = [obj if s == "gemm_sm100.cu" else os.path.join(B.OBJ, s.replace(".cu", ".o")) for s in B.SOURCES]
    subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"])
    print(LIB)


def run(R, N, K):
    import torch

    sys.path.insert(0, ROOT)
    from patchioner_b200 import _lib as L

    L.LIB_PATH = LIB
    from patchioner_b200 import ops, synth

    lib = L.lib()
    dev = torch.device("cuda:0")
    dec = ops.Decoder(synth.make_decoder_weights(1234), dev, "bf16")
    pre = torch.randn(R, 768, device=dev)
    dec.decode(pre, 4)
    torch.cuda.synchronize()
    lib.pio_debug_gemm_trace_clear(N, K)
    dec.decode(pre, 6)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (160 * 16))()
    assert lib.pio_debug_gemm_trace_read(buf, 160 * 16) == 0
    rows = [[buf[c * 16 + s] for s in range(11)] for c in range(160)]
    rows = [r for r in rows if r[0]]
    t0 = min(r[0] for r in rows)
    names = ["entry", "setup", "pdl_wait", "loads_issued", "first_stage", "last_commit", "acc0_ready", "accN_ready", "stores_issued",
             "drained", "exit"]
    print(f"decode R={R}: last gemm_tc_kernel launch with N={N} K={K}: {len(rows)} CTAs; us after the first CTA's entry")
    print("cta " + " ".join(f"{n:>13s}" for n in names))
    for i in list(range(0, len(rows), max(1, len(rows) // 12))):
        print(f"{i:3d} " + " ".join(f"{(v - t0) / 1e3:13.2f}" if v else f"{'-':>13s}" for v in rows[i]))
    import statistics as st

    print("med " + " ".join(f"{st.median([(r[s] - t0) / 1e3 for r in rows if r[s]]):13.2f}" if any(r[s] for r in rows) else f"{'-':>13s}"
                            for s in range(11)))
    print("max " + " ".join(f"{max([(r[s] - t0) / 1e3 for r in rows if r[s]]):13.2f}" if any(r[s] for r in rows) else f"{'-':>13s}"
                            for s in range(11)))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        a = [int(v) for v in sys.argv[2:]]
        run(a[0] if a else 4096, a[1] if len(a) > 1 else 768, a[2] if len(a) > 2 else 768)
