"""Where does a fused decode step spend its time?  Runs pio_decode_greedy (persistent kernel) at R rows with
PIO_FUSED_TIMELINE set, then prints, for one steady-state step, the critical path per phase from the kernel's own
globaltimer stamps (stamp 0: CTA enters the phase, 1: its input is ready, 2: work done, 3: arrival posted).

    python tools/fused_probe.py [R] [step]
"""
import os
import struct
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TYPES = ["LN1r", "QKV", "ATTN", "PROJ", "FC", "FC2", "LNFr", "LMHEAD", "PICK"]


def load(path):
    raw = open(path, "rb").read()
    steps, pps, G, first, L, R, k, nph = struct.unpack("8i", raw[:32])
    table = struct.unpack(f"{nph}i", raw[32:32 + 4 * nph])
    flat = np.frombuffer(raw[32 + 4 * nph:], dtype=np.uint64).astype(np.int64)
    t = flat[:steps * pps * G * k].reshape(steps * pps, G, k)
    tail = flat[steps * pps * G * k:]
    if tail.size >= 4 and tail[3] > tail[1]:
        print(f"SM clock during the kernel: {(tail[2] - tail[0]) / (tail[3] - tail[1]) * 1e3:.0f} MHz ({(tail[3] - tail[1]) / 1e3:.0f} us)")
    names = [(f"L{v % 16}." if v // 16 < 6 else "") + TYPES[v // 16] for v in table[:pps]]
    return dict(steps=steps, pps=pps, G=G, first=first, L=L, R=R, names=names), t


def analyse(meta, t, step):
    """stamps per (phase, CTA): 0 enters the phase, 1 previous phase seen complete, 4 activation tiles in shared memory (GEMM
    phases), 5 first accumulator ready (GEMM phases), 2 work done, 3 arrival posted"""
    pps = meta["pps"]
    prev_done = None
    tot = 0
    print(f"{'phase':10s} {'ctas':>4s} {'total':>7s} {'detect':>7s} {'acts':>6s} {'mma':>6s} {'work':>6s} {'arrive':>6s}   (ns; slowest CTA of each phase; total = last arrival - last arrival of the previous phase)")
    for p in range(pps):
        gp = step * pps + p - meta["first"]
        a = t[gp]
        part = a[:, 3] > 0
        if not part.any():
            continue
        done = a[part, 3].max()
        c = int(np.argmax(np.where(part, a[:, 3], 0)))  # the CTA that arrived last
        gemm = a[c, 5] > 0
        if prev_done is not None:
            detect = a[c, 1] - prev_done
            acts = (a[c, 4] - a[c, 1]) if gemm else 0
            mma = (a[c, 5] - a[c, 4]) if gemm else 0
            work = a[c, 2] - (a[c, 5] if gemm else a[c, 1])
            arr = a[c, 3] - a[c, 2]
            dur = done - prev_done
            tot += dur
            if gemm and a.shape[1] >= 16 and a[c, 8] > 0:
                rel = [int(a[c, i] - a[c, 6]) for i in range(8, 15)]
                print(f"{'':10s} MMA warp after waking: accumulator free +{rel[0]}; weights of group 0/1/2 present +{rel[1]}/+{rel[2]}/+{rel[3]}; group issued +{rel[4]}/+{rel[5]}/+{rel[6]}")
            extra = f"   [MMA warp: wakes +{a[c, 6] - a[c, 4]}, all issued +{a[c, 7] - a[c, 6]}, accumulator seen +{a[c, 5] - a[c, 7]}]" if gemm and a[c, 6] > 0 else ""
            print(f"{meta['names'][p]:10s} {int(part.sum()):4d} {dur:7d} {detect:7d} {acts:6d} {mma:6d} {work:6d} {arr:6d}{extra}")
        prev_done = done
    print(f"step total (without its first phase) {tot / 1e3:.1f} us")


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    import __graft_entry__ as ge

    ge.build()
    from patchioner_b200 import ops, synth

    dev = torch.device("cuda:0")
    dec = ops.Decoder(synth.make_decoder_weights(1234), dev, "bf16")
    g = torch.Generator().manual_seed(5)
    prefix = torch.randn(R, 768, generator=g)
    prefix = (prefix / prefix.norm(dim=-1, keepdim=True)).to(dev)
    for _ in range(3):
        dec.decode(prefix, 30)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dec.decode(prefix, 30)
    e1.record()
    torch.cuda.synchronize()
    print(f"R={R}: {e0.elapsed_time(e1) / 10:.3f} ms per 30-step decode = {e0.elapsed_time(e1) / 300 * 1e3:.1f} us/step")
    path = "/tmp/fused_timeline.bin"
    os.environ["PIO_FUSED_TIMELINE"] = path
    dec.decode(prefix, 30)
    torch.cuda.synchronize()
    del os.environ["PIO_FUSED_TIMELINE"]
    meta, t = load(path)
    print(meta)
    analyse(meta, t, step)


if __name__ == "__main__":
    main()
