"""Micro-benchmark of mgf_conv_tc on representative layer shapes (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morphganformer_b200 import tc

SHAPES = [  # name, B, H, W, Cin, Cout, per_sample
    ("vgg1_2", 8, 1024, 1024, 64, 64, False), ("vgg2_2", 8, 512, 512, 128, 128, False), ("vgg3_2", 8, 256, 256, 256, 256, False),
    ("vgg4_2", 8, 128, 128, 512, 512, False), ("vgg5_1", 8, 64, 64, 512, 512, False),
    ("gen64", 8, 64, 64, 512, 512, True), ("gen128", 8, 128, 128, 256, 256, True), ("gen256", 8, 256, 256, 128, 128, True),
    ("gen512", 8, 512, 512, 64, 64, True), ("gen1024", 8, 1024, 1024, 32, 32, True),
]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name, b, h, w, ci, co, ps in SHAPES:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        x = torch.randn(b, h, w, ci, device="cuda").to(torch.bfloat16)
        wt = (torch.randn(b if ps else 1, 9, co, ci, device="cuda") * 0.05).to(torch.bfloat16)
        out = torch.empty(b, h, w, co, dtype=torch.bfloat16, device="cuda")
        for bn in ([0, co] if co <= 128 else [0]):      # bn=0: automatic (halo variant where eligible); explicit bn: generic kernel
            ts = []
            for it in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                tc.conv_tc([x], wt, tc.TAPS_3X3, (b, h, w), 1, co, out, bn=bn)
                e1.record(); torch.cuda.synchronize()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            ms = min(ts)
            fl = 2.0 * b * h * w * ci * co * 9
            by = 2.0 * b * h * w * (ci + co)
            rows.append(dict(name=name, bn=bn, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1), gbs=round(by / ms / 1e6, 1)))
            print(rows[-1], flush=True)
    return rows


if __name__ == "__main__":
    main()
