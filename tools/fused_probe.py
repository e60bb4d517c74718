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

NAMES = ["LN1", "QKV", "ATTN", "PROJ", "LN2", "FC", "FC2"]


def load(path):
    raw = open(path, "rb").read()
    steps, pps, G, first, L, R, k, _ = struct.unpack("8i", raw[:32])
    t = np.frombuffer(raw[32:], dtype=np.uint64).reshape(steps * pps, G, k).astype(np.int64)
    return dict(steps=steps, pps=pps, G=G, first=first, L=L, R=R), t


def analyse(meta, t, step):
    pps, L = meta["pps"], meta["L"]
    prev_done = None
    rows = []
    for p in range(pps):
        gp = step * pps + p - meta["first"]
        a = t[gp]
        part = a[:, 3] > 0
        if not part.any():
            continue
        name = (f"L{p // 7}." + NAMES[p % 7]) if p < 7 * L else ["LNF", "LMHEAD", "PICK"][p - 7 * L]
        done = a[part, 3].max()
        ready = a[part, 1].max()
        enter = a[part, 0].max()
        work = (a[part, 2] - a[part, 1]).max()
        arr = (a[part, 3] - a[part, 2]).max()
        dur = (done - prev_done) if prev_done is not None else 0
        extra = ""
        if a.shape[1] >= 8 and prev_done is not None and (a[part, 4] > 0).any():
            # GEMM phases: act producer saw the phase (4), after the proxy fence (5), first unit's act loads issued (6); MMA warp: first weight tile present (7)
            c = int(np.argmax(np.where(part, a[:, 1], 0)))  # the CTA whose accumulator was ready last
            extra = "  [slowest CTA: detect %d  +proxy fence %d  +TMA issued %d  | W tile there at %d | acc ready %d]" % (
                a[c, 4] - prev_done, a[c, 5] - a[c, 4], a[c, 6] - a[c, 5], a[c, 7] - prev_done, a[c, 1] - prev_done)
        rows.append((name, int(part.sum()), dur, (ready - prev_done) if prev_done is not None else 0, work, arr, enter - (prev_done or enter), extra))
        prev_done = done
    print(f"{'phase':10s} {'ctas':>4s} {'total':>8s} {'->ready':>8s} {'work':>8s} {'arrive':>8s} {'late-enter':>10s}   (ns; total = last arrival of this phase - last arrival of the previous)")
    tot = 0
    for name, n, dur, rdy, work, arr, late, extra in rows:
        print(f"{name:10s} {n:4d} {dur:8d} {rdy:8d} {work:8d} {arr:8d} {late:10d}{extra}")
        tot += dur
    print(f"step total {tot / 1e3:.1f} us")


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
