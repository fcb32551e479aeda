"""Dev probe: cc_label against the C oracle on blobby masks, many repetitions; prints what differs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cc as cc_oracle
from video_llava_seg_b200.utils.misc import get_connected_components

def blobby(n, h, w, seed, thr=0.0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, 1, max(h // 8, 1), max(w // 8, 1), generator=g)
    z = torch.nn.functional.interpolate(z, size=(h, w), mode="bilinear", align_corners=False)
    return (z + 0.15 * torch.randn(n, 1, h, w, generator=g)) > thr

dev = "cuda:0"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for rep in range(3):
    m = blobby(32, 256, 256, 7 * rep + 256)
    rl, rc = cc_oracle.cc_label(m)
    md, rld, rcd = m.to(dev), rl.to(dev), rc.to(dev)
    junk = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    for it in range(iters):
        if it % 3 == 0:
            junk.zero_()
        labels, counts = get_connected_components(md)
        bl, bc = (labels != rld), (counts != rcd)
        if bool(bl.any()) or bool(bc.any()):
            bad += 1
            l, c, bl, bc = labels.cpu(), counts.cpu(), bl.cpu(), bc.cpu()
            imgs = sorted(set(bl.flatten(1).any(1).nonzero().flatten().tolist()) | set(bc.flatten(1).any(1).nonzero().flatten().tolist()))
            print(f"rep {rep} it {it}: label mismatches {int(bl.sum())}, count mismatches {int(bc.sum())}, images {imgs}")
            i = imgs[0]
            pairs = sorted(set(zip(l[i, 0][bl[i, 0]].tolist(), rl[i, 0][bl[i, 0]].tolist())))[:6]
            cpairs = sorted(set(zip(c[i, 0][bc[i, 0]].tolist(), rc[i, 0][bc[i, 0]].tolist())))[:6]
            print("   (got, want) labels", pairs, " counts", cpairs)
            if bad > 6:
                sys.exit(1)
print("mismatching launches:", bad)
