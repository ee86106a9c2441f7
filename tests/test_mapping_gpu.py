"""Fused mapping-network kernels (mgf_mapping_fwd / mgf_mapping_bwd) vs the oracle (pinned to the reference MappingNetwork) for ws,
and vs PyTorch autograd through the mirror module for d(loss)/dz.  fp32 kernels: tolerance 1e-5 relative."""
import pytest
import torch

import util
from oracle import ganformer

pytestmark = pytest.mark.gpu


def _G(seed):
    G = util.build_G(16, seed, 512, 32)
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():                       # make every bias live (FC biases are zero at init)
        for n, p in G.mapping.named_parameters():
            if n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * (30.0 if ".l" in n or "out_layer" in n else 0.3))   # lrmul 0.01 layers scale their bias by 0.01
    return G


@pytest.mark.parametrize("B", [1, 5])
def test_mapping_forward_matches_oracle_and_backward_matches_autograd(B):
    from morphganformer_b200.mapping_engine import MappingEngine, supported
    G = _G(3)
    assert supported(G)
    z = util.case_tensor((B, 17, 32), 40)
    with torch.no_grad():
        _, ws_ref = ganformer.generator(util.state_dict_cpu(G), z, 16)
    G = G.cuda()
    mask = torch.ones(B, 16, device="cuda")
    E = MappingEngine(G)
    ws = E.forward(z.cuda(), mask)
    assert tuple(ws.shape) == (B, 17, G.num_ws, 32)
    assert (ws.cpu() - ws_ref).abs().max().item() < 2e-5 * max(1.0, ws_ref.abs().max().item())
    # backward, with a non-trivial mask
    mask[:, 3] = 0; mask[0, 7] = 0
    zc = z.cuda().requires_grad_(True)
    with torch.enable_grad():
        w_t = G.mapping(zc, None, pos=G.pos, mask=mask)
    dws = util.case_tensor((B, 17, G.num_ws, 32), 41).cuda()
    gz_ref, = torch.autograd.grad(w_t, [zc], grad_outputs=[dws])
    ws2 = E.forward(z.cuda(), mask)
    assert (ws2 - w_t.detach()).abs().max().item() < 2e-5 * max(1.0, w_t.abs().max().item())
    gz = E.backward(dws)
    assert (gz - gz_ref).abs().max().item() < 5e-5 * gz_ref.abs().max().item()


def test_mapping_engine_rejects_other_configurations():
    from morphganformer_b200.mapping_engine import MappingEngine, supported
    from morphganformer_b200.training import networks as N
    from morphganformer_b200 import _lib
    kw = N.ganformer_default_kwargs(16, 512, 32)
    kw["mapping_kwargs"] = dict(kw["mapping_kwargs"], ltnt2ltnt=False)
    G = N.Generator(**kw).cuda()
    assert not supported(G)
    with pytest.raises(_lib.MgfError):
        MappingEngine(G)
