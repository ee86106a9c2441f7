"""B200-native (sm_100a) kernels for the GANformer synthesis + latent-projection hot path, behind the reference's
torch_utils/ops surface and the G.synthesis(ws, ...) signature.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
