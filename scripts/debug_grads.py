import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, util
from oracle import ganformer
res, cb, cm, B = 64, 2048, 64, 2
G = util.build_G(res, 0, cb, cm); sd = util.state_dict_cpu(G)
ws = util.case_tensor((B, 17, G.num_ws, 32), 11); mask = torch.ones(B, 16)
wsr = ws.clone().requires_grad_(True)
ref = ganformer.synthesis(sd, wsr, sd["pos"], mask, res)
tgt = torch.tanh(util.case_tensor(ref.shape, 12))
gref, = torch.autograd.grad((ref - tgt).square().mean(), [wsr])
Gc = G.cuda(); Gc.synthesis.engine = "tc"
wsg = ws.cuda().requires_grad_(True)
img, _ = Gc.synthesis(wsg, pos=Gc.pos, mask=mask.cuda(), noise_mode="const")
g, = torch.autograd.grad((img - tgt.cuda()).square().mean(), [wsg]); g = g.cpu()
# same loss gradient but pushed through the fp32 ops engine from the tc image (isolates backward from forward error)
for l in range(G.num_ws):
    for nm, sl in (("comp", slice(0, 16)), ("glob", slice(16, 17))):
        a, b = g[:, sl, l], gref[:, sl, l]
        cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
        print("slot %2d %s err %.3e scale %.3e cos %.5f ratio %.4f" % (l, nm, (a-b).abs().max(), b.abs().max(), cos, (a.norm()/b.norm()).item()))
