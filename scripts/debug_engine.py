import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, util
from oracle import ganformer
res, cb, cm, B = 64, 2048, 64, 2
G = util.build_G(res, 0, cb, cm); sd = util.state_dict_cpu(G)
ws = util.case_tensor((B, 17, G.num_ws, 32), 11); mask = torch.ones(B, 16)
tr = {}
ref = ganformer.synthesis(sd, ws, sd["pos"], mask, res, trace=tr)
Gc = G.cuda(); Gc.synthesis.engine = "tc"
img, _ = Gc.synthesis(ws.cuda(), pos=Gc.pos, mask=mask.cuda(), noise_mode="const")
st = Gc.synthesis._tc._states[B]
for k in sorted(tr, key=lambda s: (len(s), s)):
    if k in st:
        a = st[k].float().cpu().permute(0, 3, 1, 2); b = tr[k]
        print("%-8s max|err| %.4f rms err %.5f  rms ref %.4f  rel rms %.4f" % (k, (a-b).abs().max(), (a-b).square().mean().sqrt(), b.square().mean().sqrt(), (a-b).square().mean().sqrt()/b.square().mean().sqrt()))
e = img.cpu() - ref
print("img max|err| %.4f rms %.5f rms ref %.4f rel rms %.4f" % (e.abs().max(), e.square().mean().sqrt(), ref.square().mean().sqrt(), e.square().mean().sqrt()/ref.square().mean().sqrt()))
