"""Load reference-format generator checkpoints into the B200 generator mirror.

What the reference offers (loader.py:40-56 `load_network_pkl`, :182-246 the TF->torch name map, torch_utils/misc.py:137-144
`copy_params_and_buffers`, torch_utils/persistence.py:171-194 the pickled-object protocol) and what is here:

* `load_network_pkl(f)` / `load_network(path)` -- reads a snapshot pickle `{G, D, Gs}`.  PyTorch snapshots hold
  `persistence`-decorated modules: each object is pickled as `_reconstruct_persistent_obj(meta)` with `meta.state` = the module's
  `__dict__` and `meta.module_src` = the source text of its defining module.  The reference re-executes that source; this loader
  does NOT execute anything from the file: a restricted unpickler (explicit (module, name) allowlist of tensor / ndarray /
  container constructors, `_allowed_globals`) turns every persistent object into an inert `PersistentStub`,
  the parameter/buffer tree is flattened to a `state_dict`, and the generator is rebuilt from `init_kwargs` with this package's
  `training.networks.Generator` (same parameter names, so `load_state_dict(strict=True)` is the parity check).
  TensorFlow snapshots (`dnnlib.tflib.network.Network` triples) go through `convert_tf_generator`.
  The discriminator is outside the hot path: `D` is returned as its flat `state_dict`, not as a module.
* `convert_tf_generator(tf_G)` -- the TF variable-name map of the reference restated as a rule table (`_TF_RULES`).
* `copy_params_and_buffers(src, dst, require_all)` -- same semantics as the reference helper; `src` may also be a state_dict.

Only local files are opened (the reference's `gdrive:` aliases need a network; they raise here).
"""
import collections
import io
import pickle
import re

import numpy as np
import torch


class EasyDict(dict):
    """dict with attribute access (stand-in for dnnlib.EasyDict, dnnlib/util.py:32-43)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


class TFNetworkStub(EasyDict):
    """Unpickled `dnnlib.tflib.network.Network`: fields version, static_kwargs, components, variables [(name, ndarray)]."""


class PersistentStub:
    """Inert image of a `persistence`-decorated object: class name, constructor arguments and the raw state dict."""

    def __init__(self, meta):
        meta = dict(meta)
        self.class_name = meta.get("class_name")
        self.version = meta.get("version")
        self.state = dict(meta.get("state") or {})

    @property
    def init_args(self):
        return tuple(self.state.get("_init_args", ()))

    @property
    def init_kwargs(self):
        return EasyDict(self.state.get("_init_kwargs", {}))

    def flat_state(self, prefix=""):
        """Flattens the torch.nn.Module state held by this stub (and its children) into `state_dict` naming."""
        out = collections.OrderedDict()
        for kind in ("_parameters", "_buffers"):
            for name, t in (self.state.get(kind) or {}).items():
                if t is not None and name not in (self.state.get("_non_persistent_buffers_set") or ()):
                    out[prefix + name] = t.detach() if isinstance(t, torch.Tensor) else torch.as_tensor(t)
        for name, child in (self.state.get("_modules") or {}).items():
            if child is None:
                continue
            if isinstance(child, PersistentStub):
                out.update(child.flat_state(prefix + name + "."))
            elif isinstance(child, torch.nn.Module):          # plain (non-persistent) torch containers, e.g. ModuleList
                for k, v in child.state_dict().items():
                    out[prefix + name + "." + k] = v
            else:
                raise pickle.UnpicklingError("unexpected child %r of %s" % (type(child), self.class_name))
        return out


def _reconstruct(meta):
    return PersistentStub(meta)


def _load_storage_from_bytes(b):
    """Stand-in for torch.storage._load_from_bytes (how a pickled tensor's storage arrives): the nested archive is read with
    torch's weights-only unpickler, so it cannot name arbitrary globals either."""
    return torch.load(io.BytesIO(b), weights_only=True)


def _allowed_globals():
    """Explicit (module, name) -> object allowlist: tensor / ndarray / container constructors only.  Everything else in a checkpoint is
    refused -- a prefix test on the top-level package is not enough (numpy.testing._private.utils.runstring, torch.utils.collect_env.run
    and friends execute code)."""
    import copyreg
    import _codecs
    table = {
        ("collections", "OrderedDict"): collections.OrderedDict,
        ("_codecs", "encode"): _codecs.encode,
        ("copyreg", "_reconstructor"): copyreg._reconstructor,
        ("torch._utils", "_rebuild_tensor_v2"): torch._utils._rebuild_tensor_v2,
        ("torch._utils", "_rebuild_tensor"): torch._utils._rebuild_tensor,
        ("torch._utils", "_rebuild_parameter"): torch._utils._rebuild_parameter,
        ("torch.storage", "_load_from_bytes"): _load_storage_from_bytes,
        ("torch", "Size"): torch.Size,
        ("torch", "device"): torch.device,
        ("torch.nn.parameter", "Parameter"): torch.nn.Parameter,
        ("torch.nn.modules.container", "ModuleList"): torch.nn.ModuleList,
        ("torch.nn.modules.container", "ModuleDict"): torch.nn.ModuleDict,
        ("torch.nn.modules.container", "Sequential"): torch.nn.Sequential,
        ("torch.nn.modules.container", "ParameterList"): torch.nn.ParameterList,
        ("torch.nn.modules.dropout", "Dropout"): torch.nn.Dropout,
        ("torch.nn.modules.linear", "Identity"): torch.nn.Identity,
        ("numpy", "ndarray"): np.ndarray,
        ("numpy", "dtype"): np.dtype,
    }
    for name in ("float16", "float32", "float64", "bfloat16", "int8", "uint8", "int16", "int32", "int64", "bool"):
        table[("torch", name)] = getattr(torch, name)                 # torch.dtype objects pickle as getattr(torch, name)
    for name in ("FloatStorage", "DoubleStorage", "HalfStorage", "BFloat16Storage", "LongStorage", "IntStorage", "ShortStorage",
                 "CharStorage", "ByteStorage", "BoolStorage"):
        if hasattr(torch, name):
            table[("torch", name)] = getattr(torch, name)
    import numpy.core.multiarray as _ma_old      # numpy 1.x pickles name numpy.core.*, numpy 2.x numpy._core.*: same functions
    for mod in ("numpy.core.multiarray", "numpy._core.multiarray"):
        table[(mod, "_reconstruct")] = _ma_old._reconstruct
        table[(mod, "scalar")] = _ma_old.scalar
    return table


_ALLOWED = None


class _SafeUnpickler(pickle.Unpickler):
    """Resolves only the allowlisted tensor/ndarray/container constructors; maps the reference's own classes to inert stubs."""

    def find_class(self, module, name):
        global _ALLOWED
        if module == "torch_utils.persistence" and name == "_reconstruct_persistent_obj":
            return _reconstruct
        if module == "dnnlib.tflib.network" and name == "Network":
            return TFNetworkStub
        if module in ("dnnlib.util", "dnnlib") and name == "EasyDict":
            return EasyDict
        if module == "builtins" and name in ("set", "frozenset", "dict", "list", "tuple", "slice", "complex", "bytearray", "object"):
            return super().find_class(module, name)
        if _ALLOWED is None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                _ALLOWED = _allowed_globals()
        obj = _ALLOWED.get((module, name))
        if obj is None:
            raise pickle.UnpicklingError("refusing to import %s.%s from a checkpoint" % (module, name))
        return obj


def named_params_and_buffers(module):
    """reference torch_utils/misc.py:133-135."""
    assert isinstance(module, torch.nn.Module)
    return list(module.named_parameters()) + list(module.named_buffers())


@torch.no_grad()
def copy_params_and_buffers(src, dst_module, require_all=False):
    """Copies every tensor of `src` (module or state_dict) whose name also exists in `dst_module` (reference
    torch_utils/misc.py:137-144); `require_all` asserts that nothing of the destination is left untouched."""
    assert isinstance(dst_module, torch.nn.Module)
    src_tensors = dict(named_params_and_buffers(src)) if isinstance(src, torch.nn.Module) else dict(src)
    for name, tensor in named_params_and_buffers(dst_module):
        if name in src_tensors:
            tensor.copy_(torch.as_tensor(src_tensors[name]).detach().to(tensor.dtype)).requires_grad_(tensor.requires_grad)
        elif require_all:
            raise KeyError("source has no tensor named %r" % name)


def _networks():
    from .training import networks
    return networks


def generator_from_stub(stub):
    """PersistentStub of a reference `training.networks.Generator` -> this package's Generator with the same weights."""
    if stub.class_name != "Generator":
        raise ValueError("expected a pickled Generator, got %r" % (stub.class_name,))
    G = _networks().Generator(*stub.init_args, **_plain(stub.init_kwargs)).eval().requires_grad_(False)
    sd = stub.flat_state()
    missing, unexpected = G.load_state_dict(sd, strict=False)
    # constant buffers (FIR taps, positional grids) are rebuilt identically by the constructor; anything else must match
    hard = [k for k in missing if not re.fullmatch(r".*\.(resample_kernel|grid_pos)", k)]
    if hard or unexpected:
        raise RuntimeError("checkpoint does not fit the generator: missing %s unexpected %s" % (hard[:5], list(unexpected)[:5]))
    return G


def _plain(obj):
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_plain(v) for v in obj)
    return obj


def load_network_pkl(f):
    """f: binary file object (or bytes).  Returns dict(G=Generator, Gs=Generator, D=state_dict or None, ...extra entries)."""
    if isinstance(f, (bytes, bytearray)):
        f = io.BytesIO(f)
    data = _SafeUnpickler(f).load()
    if isinstance(data, tuple) and len(data) == 3 and all(isinstance(n, TFNetworkStub) for n in data):
        tf_G, tf_D, tf_Gs = data
        return dict(G=convert_tf_generator(tf_G), D=collect_tf_params(tf_D), Gs=convert_tf_generator(tf_Gs))
    if not isinstance(data, dict) or not any(k in data for k in ("G", "Gs", "G_ema")):
        raise ValueError("not a generator snapshot (expected a dict with G / Gs or a TensorFlow (G, D, Gs) triple)")
    out = {}
    for key, val in data.items():
        if isinstance(val, PersistentStub):
            out[key] = generator_from_stub(val) if val.class_name == "Generator" else val.flat_state()
        elif isinstance(val, torch.nn.Module):
            out[key] = val
        else:
            out[key] = val
    return out


def load_network(filename):
    """Local-file counterpart of reference loader.py:26-30."""
    if str(filename).startswith(("gdrive:", "http://", "https://")):
        raise ValueError("remote checkpoints are not fetched here (no network): download %r first" % (filename,))
    with open(filename, "rb") as f:
        return load_network_pkl(f)


# ----------------------------------------------------------------------------------------------------------- TensorFlow snapshots
def collect_tf_params(tf_net):
    """Flattens a TF network stub into {"scope/.../var": ndarray} (reference loader.py:60-68)."""
    out = {}
    stack = [("", tf_net)]
    while stack:
        prefix, net = stack.pop()
        for name, value in net.variables:
            out[prefix + name] = value
        for name, comp in net.components.items():
            stack.append((prefix + name + "/", comp))
    return out


def _tf_generator_kwargs(static):
    """TF static_kwargs -> Generator constructor kwargs (reference loader.py:98-153)."""
    g = lambda name, default=None: static[name] if static.get(name) is not None else default
    transformer = bool(g("transformer", False))
    att = dict(num_heads=g("num_heads", 1), attention_dropout=g("attention_dropout", 0.12), ltnt_gate=g("ltnt_gate", False),
               use_pos=g("use_pos", False))
    w_avg_beta = static.get("dlatent_avg_beta", 0.995)
    mapping = dict(num_layers=g("mapping_layersnum", 8), layer_dim=g("mapping_dim"), act=g("mapping_nonlinearity", "lrelu"),
                   lrmul=g("mapping_lrmul", 0.01), w_avg_beta=1 if w_avg_beta is None else w_avg_beta,
                   resnet=g("mapping_resnet", False), ltnt2ltnt=g("mapping_ltnt2ltnt", False), transformer=transformer,
                   normalize_global=False, **att)
    synthesis = dict(channel_base=2 * g("fmap_base", 16 << 10), channel_max=g("fmap_max", 512), architecture=g("architecture", "skip"),
                     resample_kernel=g("resample_kernel", [1, 3, 3, 1]), local_noise=g("local_noise", True),
                     act=g("nonlinearity", "lrelu"), latent_stem=g("latent_stem", False), style=g("style", True),
                     transformer=transformer, start_res=g("start_res", 0), end_res=g("end_res", 8), img_gate=g("img_gate", False),
                     integration=g("integration", "add"), norm=g("norm"), kmeans=g("kmeans", False),
                     kmeans_iters=g("kmeans_iters", 1), iterative=g("iterative", False), pos_dim=g("pos_dim"),
                     pos_type=g("pos_type", "sinus"), pos_init=g("pos_init", "uniform"),
                     pos_directions_num=g("pos_directions_num", 2), **att)
    return dict(z_dim=g("latent_size", 512), c_dim=g("label_size", 0), w_dim=g("dlatent_size", 512),
                k=g("components_num", 1) + int(transformer), img_resolution=g("resolution", 1024), img_channels=g("num_channels", 3),
                mapping_kwargs=mapping, synthesis_kwargs=synthesis)


_T = lambda v: np.asarray(v).transpose()                                   # dense [in, out] -> [out, in]
_CONV = lambda v: np.asarray(v).transpose(3, 2, 0, 1)                      # HWIO -> OIHW
_CONV_FLIP = lambda v: np.asarray(v)[::-1, ::-1].transpose(3, 2, 0, 1)     # transposed-conv kernels are stored spatially flipped
_ID = lambda v: np.asarray(v)
_PLUS1 = lambda v: np.asarray(v) + 1                                       # TF stores the style bias around 0, torch around 1
_SINGULAR = {"queries": "query", "keys": "key", "values": "value"}


def _att_scope_rules(torch_prefix, tf_scope):
    """Rules for one attention layer: torch `<prefix>.<leaf>` <- TF `<scope>/<var>` (reference loader.py:195-204, :221-230)."""
    p = torch_prefix
    return [
        (p + r"\.to_(queries|keys|values)\.weight", lambda m: (tf_scope(m) + "/weight_" + _SINGULAR[m[-1]], _T)),
        (p + r"\.to_(queries|keys|values)\.bias", lambda m: (tf_scope(m) + "/bias_" + _SINGULAR[m[-1]], _ID)),
        (p + r"\.(from|to)_pos_map\.weight", lambda m: (tf_scope(m) + "/weight_%s_pos" % m[-1], _T)),
        (p + r"\.(from|to)_pos_map\.bias", lambda m: (tf_scope(m) + "/bias_%s_pos" % m[-1], _ID)),
        (p + r"\.modulation\.weight", lambda m: (tf_scope(m) + "/weight_out", _T)),
        (p + r"\.modulation\.bias", lambda m: (tf_scope(m) + "/bias_out", _ID)),
        (p + r"\.centroids", lambda m: (tf_scope(m) + "/toasgn_init", _ID)),
        (p + r"\.queries2centroids\.weight", lambda m: (tf_scope(m) + "/weight_key2", _T)),
        (p + r"\.queries2centroids\.bias", lambda m: (tf_scope(m) + "/bias_key2", _ID)),
        (p + r"\.att_weight", lambda m: (tf_scope(m) + "/iter_0/st_weights", _ID)),
    ]


def _conv_scope(r, i):
    """TF scope of torch `synthesis.b{r}.conv{i}`: 4x4 has the single `Conv`; above, conv0 is `Conv0_up`, conv1 is `Conv1`."""
    r, i = int(r), int(i)
    return "synthesis/%dx%d/Conv%s" % (r, r, "" if r == 4 else ("0_up" if i == 0 else str(i)))


def _mapping_scope(sub):
    return "mapping/global/" if "global" in sub else "mapping/"


_TF_RULES = (
    [(r"pos", lambda m: ("ltnt_emb/emb", _ID)),
     (r"mapping\.w_avg", lambda m: ("dlatent_avg", _ID)),
     (r"mapping\.embed\.weight", lambda m: ("mapping/LabelConcat/weight", _T)),
     (r"mapping\.([a-z_]+)\.l(\d+)\.fc(\d+)\.(weight|bias)",
      lambda m: ("%sDense%s_%s/%s" % (_mapping_scope(m[0]), m[1], m[2], m[3]), _T if m[3] == "weight" else _ID)),
     (r"mapping\.([a-z_]+)\.out_layer\.(weight|bias)",
      lambda m: ("%sDense3/%s" % (_mapping_scope(m[0]), m[1]), _T if m[1] == "weight" else _ID))]
    + _att_scope_rules(r"mapping\.mlp\.sa(\d+)", lambda m: "mapping/AttLayer_%s" % m[0])
    + [(r"synthesis\.b4\.const", lambda m: ("synthesis/4x4/Const/const", lambda v: np.asarray(v)[0])),
       (r"synthesis\.b(\d+)\.conv0\.weight", lambda m: (_conv_scope(m[0], 0) + "/weight", _CONV_FLIP)),
       (r"synthesis\.b(\d+)\.conv1\.weight", lambda m: (_conv_scope(m[0], 1) + "/weight", _CONV)),
       (r"synthesis\.b(\d+)\.conv(\d+)\.biasAct\.bias", lambda m: (_conv_scope(m[0], m[1]) + "/bias", _ID)),
       (r"synthesis\.b(\d+)\.conv(\d+)\.noise_const",
        lambda m: ("synthesis/noise%d" % (int(np.log2(int(m[0]))) * 2 - 5 + int(m[1])), lambda v: np.asarray(v)[0, 0])),
       (r"synthesis\.b(\d+)\.conv(\d+)\.noise_strength", lambda m: (_conv_scope(m[0], m[1]) + "/noise_strength", _ID)),
       (r"synthesis\.b(\d+)\.conv(\d+)\.affine\.weight", lambda m: (_conv_scope(m[0], m[1]) + "/mod_weight", _T)),
       (r"synthesis\.b(\d+)\.conv(\d+)\.affine\.bias", lambda m: (_conv_scope(m[0], m[1]) + "/mod_bias", _PLUS1))]
    + _att_scope_rules(r"synthesis\.b(\d+)\.conv(\d+)\.transformer", lambda m: _conv_scope(m[0], m[1]) + "/AttLayer_l2n")
    + [(r"synthesis\.b(\d+)\.torgb\.weight", lambda m: ("synthesis/%sx%s/ToRGB/weight" % (m[0], m[0]), _CONV)),
       (r"synthesis\.b(\d+)\.torgb\.biasAct\.bias", lambda m: ("synthesis/%sx%s/ToRGB/bias" % (m[0], m[0]), _ID)),
       (r"synthesis\.b(\d+)\.torgb\.affine\.weight", lambda m: ("synthesis/%sx%s/ToRGB/mod_weight" % (m[0], m[0]), _T)),
       (r"synthesis\.b(\d+)\.torgb\.affine\.bias", lambda m: ("synthesis/%sx%s/ToRGB/mod_bias" % (m[0], m[0]), _PLUS1)),
       (r"synthesis\.b(\d+)\.skip\.weight", lambda m: ("synthesis/%sx%s/Skip/weight" % (m[0], m[0]), _CONV_FLIP)),
       (r"synthesis\.b(\d+)\.conv_last\.weight", lambda m: ("synthesis/%sx%s/ToRGB/extraLayer/weight" % (m[0], m[0]), _CONV)),
       (r"synthesis\.b(\d+)\.conv_last\.affine\.weight", lambda m: ("synthesis/%sx%s/ToRGB/extraLayer/mod_weight" % (m[0], m[0]), _T)),
       (r"synthesis\.b(\d+)\.conv_last\.affine\.bias", lambda m: ("synthesis/%sx%s/ToRGB/extraLayer/mod_bias" % (m[0], m[0]), _PLUS1))]
)
_TF_RULES = [(re.compile(p), fn) for p, fn in _TF_RULES]
_TF_KEEP = re.compile(r".*\.(resample_kernel|grid_pos)")     # constants the constructor already built


def tf_source_of(torch_name):
    """torch parameter/buffer name -> (TF variable name, transform) or None when the tensor keeps its constructed value."""
    if _TF_KEEP.fullmatch(torch_name):
        return None
    for pat, fn in _TF_RULES:
        m = pat.fullmatch(torch_name)
        if m:
            return fn(m.groups())
    raise KeyError("no TensorFlow source known for %r" % torch_name)


@torch.no_grad()
def convert_tf_generator(tf_G):
    """TensorFlow GANformer generator stub -> Generator (reference loader.py:87-246)."""
    if tf_G.version < 4:
        raise ValueError("TensorFlow pickle version too low")
    kwargs = _tf_generator_kwargs(tf_G.static_kwargs)
    params = collect_tf_params(tf_G)
    # per-lod ToRGB layers mean the 'orig' architecture (reference loader.py:169-174)
    for name in list(params):
        m = re.fullmatch(r"ToRGB_lod(\d+)/(.*)", name)
        if m:
            r = kwargs["img_resolution"] >> int(m.group(1))
            params["synthesis/%dx%d/ToRGB/%s" % (r, r, m.group(2))] = params[name]
            kwargs["synthesis_kwargs"]["architecture"] = "orig"
    G = _networks().Generator(**kwargs).eval().requires_grad_(False)
    for name, tensor in named_params_and_buffers(G):
        if name == "mapping.embed.bias":
            tensor.zero_()
            continue
        src = tf_source_of(name)
        if src is None:
            continue
        tf_name, transform = src
        if tf_name not in params:
            raise KeyError("TensorFlow snapshot has no variable %r (needed for %r)" % (tf_name, name))
        value = torch.from_numpy(np.array(transform(params[tf_name]))).to(tensor.dtype)
        if tuple(value.shape) != tuple(tensor.shape):
            raise ValueError("%s: TensorFlow variable %s has shape %s, expected %s" % (name, tf_name, tuple(value.shape), tuple(tensor.shape)))
        tensor.copy_(value)
    return G
