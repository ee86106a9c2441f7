import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, util
from oracle import ganformer
res, cb, cm, B = 64, 2048, 64, 2
G = util.build_G(res, 0, cb, cm); sd = util.state_dict_cpu(G)
z = util.case_tensor((B, 17, 32), 5) * 0.05 + util.case_tensor((1, 17, 32), 6)
ref, ws = ganformer.generator(sd, z, res)
tr = {}
ref2 = ganformer.synthesis(sd, ws, sd["pos"], torch.ones(B, 16), res, trace=tr)
Gc = G.cuda(); Gc.synthesis.engine = "tc"
wsg = Gc.mapping(z.cuda(), None, pos=Gc.pos, mask=torch.ones(B, 16, device="cuda"))
print("ws diff", (wsg.cpu() - ws).abs().max().item())
img, _ = Gc.synthesis(wsg, pos=Gc.pos, mask=torch.ones(B, 16, device="cuda"), noise_mode="const")
img = img.cpu()
print("mean img^2 tc %.5f ref %.5f ; mean tc %.5f ref %.5f" % (img.square().mean(), ref.square().mean(), img.mean(), ref.mean()))
st = Gc.synthesis._tc._states[B]
for k in sorted(tr, key=lambda s: (len(s), s)):
    if k in st:
        a = st[k].float().cpu().permute(0, 3, 1, 2); b = tr[k]
        print("%-8s meansq tc %.5f ref %.5f ratio %.5f | mean tc %.5f ref %.5f" % (k, a.square().mean(), b.square().mean(), a.square().mean()/b.square().mean(), a.mean(), b.mean()))
import math
from oracle import ops as O
pre = "synthesis.b64"
f = O.setup_filter([1, 3, 3, 1])
L = G.num_ws
xo = tr["xout64"]
yl_ref, _ = ganformer.synthesis_layer(sd, pre + ".conv_last", xo, ws[:, :, L - 2], sd["pos"], torch.ones(B, 16), 64, attention=False, bias=False, noise=False, f=f)
yl = st["yl"].float().cpu().permute(0, 3, 1, 2)
print("yl meansq tc %.5f ref %.5f ratio %.5f relrms %.5f" % (yl.square().mean(), yl_ref.square().mean(), yl.square().mean()/yl_ref.square().mean(), (yl-yl_ref).square().mean().sqrt()/yl_ref.square().mean().sqrt()))
# torgb from the reference yl through the oracle vs from tc yl
wr = sd[pre + ".torgb.weight"]
styles = ganformer._fc(sd, pre + ".torgb.affine", ws[:, -1, L - 1]) * (1.0 / math.sqrt(wr[0].numel()))
img_from_tc_yl = O.bias_act(ganformer.modulated_conv2d(yl, wr, styles, demodulate=False), sd[pre + ".torgb.biasAct.bias"])
print("img(oracle torgb on tc yl) vs tc img: max diff %.5f ; vs ref img rel rms %.5f" % ((img_from_tc_yl - img).abs().max(), (img_from_tc_yl-ref).square().mean().sqrt()/ref.square().mean().sqrt()))
print("s_rgb diff", (st["s_rgb"].cpu() - styles).abs().max().item())
