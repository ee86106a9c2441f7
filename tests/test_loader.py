"""Checkpoint loading: reference-format snapshot pickles and TensorFlow variable dictionaries."""
import io
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

import util
from morphganformer_b200 import loader
from oracle import ganformer, refimport


def test_reference_snapshot_fixture_loads_without_executing_source():
    """tests/golden/ref_snapshot_16.pkl was pickled by the real reference's persistence machinery (make_golden_ckpt.py)."""
    io_ = np.load(os.path.join(util.GOLDEN, "ref_snapshot_16_io.npz"))
    data = loader.load_network(os.path.join(util.GOLDEN, "ref_snapshot_16.pkl"))
    G = data["Gs"]
    from morphganformer_b200.training import networks as N
    assert isinstance(G, N.Generator) and data["training_set_kwargs"] == dict(note="synthetic")
    sd = util.state_dict_cpu(G)
    assert abs(util.sd_checksum(sd) - float(io_["checksum"])) < 1e-6 * abs(float(io_["checksum"]))
    # behaviour: the oracle (pinned to the reference) on the loaded weights reproduces the reference generator's image
    with torch.no_grad():
        img, _ = ganformer.generator(sd, torch.from_numpy(io_["z"]), 16)
    assert (img - torch.from_numpy(io_["img"])).abs().max() < 2e-5


def test_unpickler_refuses_foreign_globals():
    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    with pytest.raises(pickle.UnpicklingError):
        loader.load_network_pkl(pickle.dumps(dict(G=Evil())))


@pytest.mark.parametrize("target", ["numpy.testing._private.utils:runstring", "torch.utils.collect_env:run", "torch.storage:_load_from_bytes_nested",
                                    "torch.hub:load", "numpy:load", "torch:load", "collections:namedtuple"])
def test_unpickler_refuses_code_executing_members_of_allowed_packages(target, tmp_path):
    """A prefix test on the top-level package is not enough: torch / numpy hold functions that run code.  Each of these pickles used
    to execute its payload through load_network_pkl; all must now be refused without side effects."""
    import importlib
    marker = tmp_path / "pwned"
    modname, fname = target.split(":")
    if fname == "_load_from_bytes_nested":
        # the storage hook is reachable, but its nested archive goes through torch's weights-only unpickler: a nested full pickle with a
        # foreign global must fail instead of executing
        class Inner:
            def __reduce__(self):
                return (os.system, ("touch %s" % marker,))
        buf = io.BytesIO(); torch.save(Inner(), buf)
        payload = (torch.storage._load_from_bytes, (buf.getvalue(),))
    else:
        fn = getattr(importlib.import_module(modname), fname)
        args = {"runstring": ("open(%r, 'w').close()" % str(marker), {}), "run": ("touch %s" % marker,), "load": (str(marker),),
                "namedtuple": ("X", "a b")}[fname]
        payload = (fn, args)

    class Evil:
        def __reduce__(self):
            return payload
    with pytest.raises(Exception) as ei:
        loader.load_network_pkl(pickle.dumps(dict(G=Evil())))
    assert isinstance(ei.value, (pickle.UnpicklingError, RuntimeError)), ei.value
    assert not marker.exists()


def _fake_persistent_pickle(G):
    """Pickles `G` the way torch_utils/persistence.py:110-119 does, without the reference tree."""
    mod = types.ModuleType("torch_utils.persistence")

    def _reconstruct_persistent_obj(meta):
        raise AssertionError("must not be called by the loader")
    _reconstruct_persistent_obj.__module__ = "torch_utils.persistence"
    _reconstruct_persistent_obj.__qualname__ = "_reconstruct_persistent_obj"
    mod._reconstruct_persistent_obj = _reconstruct_persistent_obj
    saved = {k: sys.modules.get(k) for k in ("torch_utils", "torch_utils.persistence")}
    sys.modules["torch_utils"] = types.ModuleType("torch_utils"); sys.modules["torch_utils.persistence"] = mod

    class Node:
        def __init__(self, m, kwargs=None):
            st = {k: v for k, v in m.__dict__.items() if not k.startswith("_") or k in ("_parameters", "_buffers", "_non_persistent_buffers_set")}
            st["_modules"] = {k: Node(c) for k, c in m._modules.items()}
            st["_init_args"], st["_init_kwargs"] = (), kwargs or {}
            self.meta = dict(type="class", version=6, module_src="", class_name=type(m).__name__, state=st)

        def __reduce__(self):
            return (_reconstruct_persistent_obj, (self.meta,))
    try:
        from morphganformer_b200.training import networks as N
        return pickle.dumps(dict(G=Node(G, N.ganformer_default_kwargs(8, 512, 32)), D=None))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_persistent_protocol_roundtrip():
    G = util.build_G(8, 3, 512, 32)
    data = loader.load_network_pkl(_fake_persistent_pickle(G))
    a, b = G.state_dict(), data["G"].state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def _tf_stub_from(G, res, channel_base, channel_max):
    """Inverts the name map: a TensorFlow-layout variable dictionary holding G's weights."""
    inv_T = lambda v: v.transpose()
    variables = {}
    for name, t in loader.named_params_and_buffers(G):
        src = loader.tf_source_of(name)
        if src is None:
            continue
        tf_name, tr = src
        v = t.detach().numpy()
        if tr is loader._T:
            v = v.transpose()
        elif tr is loader._CONV:
            v = v.transpose(2, 3, 1, 0)
        elif tr is loader._CONV_FLIP:
            v = v.transpose(2, 3, 1, 0)[::-1, ::-1]
        elif tr is loader._PLUS1:
            v = v - 1
        elif tf_name.endswith("Const/const"):
            v = v[None]
        elif tf_name.startswith("synthesis/noise"):
            v = v[None, None]
        variables[tf_name] = np.array(v)
    static = dict(latent_size=32, dlatent_size=32, components_num=16, transformer=True, resolution=res, fmap_base=channel_base // 2,
                  fmap_max=channel_max, architecture="resnet", mapping_resnet=True, mapping_ltnt2ltnt=True, use_pos=True,
                  norm="layer", integration="mul", kmeans=True, mapping_layersnum=8, dlatent_avg_beta=None)
    return static, variables


def _mk_stub(cls, static, variables):
    s = cls()
    s.version, s.static_kwargs, s.components, s.variables = 5, static, {}, list(variables.items())
    return s


def test_tf_conversion_roundtrip():
    G = util.build_G(16, 7, 512, 32)
    static, variables = _tf_stub_from(G, 16, 512, 32)
    G2 = loader.convert_tf_generator(_mk_stub(loader.TFNetworkStub, static, variables))
    a, b = G.state_dict(), G2.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert loader._conv_scope(4, 1).endswith("4x4/Conv") and loader._conv_scope(8, 0).endswith("Conv0_up")


@pytest.mark.skipif(not refimport.available(), reason="reference tree not present")
def test_tf_conversion_matches_reference_converter():
    """Same TF stub through the reference's convert_tf_generator (loader.py:87-246) and through ours."""
    refimport.load()
    import importlib
    ref_loader = importlib.import_module("loader")            # the reference's top-level loader.py
    G = util.build_G(256, 7, 512, 32)     # the reference's map only names conv_last at 256 / 512 / 1024 (loader.py:236-245)
    static, variables = _tf_stub_from(G, 256, 512, 32)
    Gr = ref_loader.convert_tf_generator(_mk_stub(ref_loader._TFNetworkStub, static, variables))
    Gm = loader.convert_tf_generator(_mk_stub(loader.TFNetworkStub, static, variables))
    a, b = Gr.state_dict(), Gm.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_copy_params_and_buffers():
    A, B = util.build_G(8, 1, 512, 32), util.build_G(8, 2, 512, 32)
    loader.copy_params_and_buffers(A, B, require_all=True)
    assert all(torch.equal(v, B.state_dict()[k]) for k, v in A.state_dict().items())
    with pytest.raises(KeyError):
        loader.copy_params_and_buffers({"pos": A.pos}, B, require_all=True)


def test_skip_architecture_snapshot_fixture_loads_on_cpu():
    """tests/golden/ref_snapshot_64_skip.pkl (make_golden_ckpt.py 64 skip): the architecture TensorFlow snapshots convert to."""
    io_ = np.load(os.path.join(util.GOLDEN, "ref_snapshot_64_skip_io.npz"))
    G = loader.load_network(os.path.join(util.GOLDEN, "ref_snapshot_64_skip.pkl"))["Gs"]
    assert G.synthesis.architecture == "skip" and G.img_resolution == 64
    sd = util.state_dict_cpu(G)
    assert abs(util.sd_checksum(sd) - float(io_["checksum"])) < 1e-6 * abs(float(io_["checksum"]))


@pytest.mark.gpu
@pytest.mark.parametrize("stem", ["ref_snapshot_16", "ref_snapshot_64_skip"])
def test_loaded_reference_snapshot_runs_on_both_cuda_engines(stem):
    """SURVEY 8f-1 end to end: a snapshot pickled by the REAL reference -> loader -> .cuda() -> G(z) on the exact-fp32 ops engine (1e-4)
    and on the tcgen05 engine (1e-2, absolute) against the image the reference generator itself produced for the same z."""
    io_ = np.load(os.path.join(util.GOLDEN, stem + "_io.npz"))
    G = loader.load_network(os.path.join(util.GOLDEN, stem + ".pkl"))["Gs"].cuda()
    z, want = torch.from_numpy(io_["z"]).cuda(), torch.from_numpy(io_["img"])
    for engine, tol in (("ops", 1e-4), ("tc", 1e-2)):
        G.synthesis.engine = engine
        with torch.no_grad():
            img = G(z, noise_mode="const")[0]
        err = (img.cpu() - want).abs().max().item()
        print(stem, engine, "max-abs vs the reference's image", err)
        assert err < tol, (engine, err)
