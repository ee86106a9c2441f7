"""Build-container only: writes tests/golden/ref_snapshot_16.pkl -- a snapshot pickle produced by the REAL reference's
`persistence` machinery (torch_utils/persistence.py:110-119 `__reduce__`) for a tiny GANformer generator (16x16, 32 channels),
plus the image that reference generator produces for a fixed z, so the loader test can pin weights AND behaviour.

The pickled `module_src` strings (the reference's source text, which its own loader re-executes) are blanked before writing:
this repo's loader never executes them, and reference sources must not be copied into the repo.

    python tests/golden/make_golden_ckpt.py                 # ref_snapshot_16.pkl (resnet architecture)
    python tests/golden/make_golden_ckpt.py 64 skip         # ref_snapshot_64_skip.pkl: 64x64, architecture='skip' (what TensorFlow snapshots
                                                            # convert to, reference loader.py:129) -- the fixture the GPU loader test runs
"""
import os, pickle, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np, torch
from oracle import refimport
import util

ref = refimport.load()
from torch_utils import persistence          # the reference's (sys.path set by refimport.load)
RES = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ARCH = sys.argv[2] if len(sys.argv) > 2 else "resnet"
STEM = "ref_snapshot_%d%s" % (RES, "" if ARCH == "resnet" else "_" + ARCH)
Gr = util.randomize(refimport.build_generator(RES, seed=11, channel_base=512 if RES == 16 else 2048, channel_max=32, architecture=ARCH), 12)
assert persistence.is_persistent(Gr)

# blank the embedded source text: patch the reduce meta on the fly
orig = persistence._module_to_src
blob = pickle.dumps(dict(G=Gr, Gs=Gr, training_set_kwargs=dict(note="synthetic")))
class Blank(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch_utils.persistence" and name == "_reconstruct_persistent_obj":
            return lambda meta: _Node(meta)
        return super().find_class(module, name)
class _Node:
    def __init__(self, meta): self.meta = dict(meta, module_src="")
    def __reduce__(self): return (_reconstruct_persistent_obj, (self.meta,))
def _reconstruct_persistent_obj(meta):
    raise RuntimeError("fixture is read with morphganformer_b200.loader only")
_reconstruct_persistent_obj.__module__ = "torch_utils.persistence"
persistence._reconstruct_persistent_obj = _reconstruct_persistent_obj      # so pickle finds the global by reference
import io
tree = Blank(io.BytesIO(blob)).load()
out = pickle.dumps(tree)
assert b"def modulated_conv2d" not in out and b"class Generator" not in out
z = util.case_tensor((2, 17, 32), 13)
with torch.no_grad():
    img = Gr(z, noise_mode="const")[0]
open(os.path.join(HERE, STEM + ".pkl"), "wb").write(out)
np.savez_compressed(os.path.join(HERE, STEM + "_io.npz"), z=z.numpy(), img=img.numpy(),
                    checksum=np.float64(util.sd_checksum({k: v.detach() for k, v in Gr.state_dict().items()})))
print("wrote", len(out), "bytes; img", tuple(img.shape), "torch", torch.__version__)
