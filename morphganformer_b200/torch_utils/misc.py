"""The two helpers of the reference's torch_utils/misc.py that the hot path calls (assert_shape :70-83,
profiled_function :86-91) plus suppress_tracer_warnings."""
import contextlib
import torch


def assert_shape(tensor, ref_shape):
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f"Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}")
    for idx, (size, ref_size) in enumerate(zip(tensor.shape, ref_shape)):
        if ref_size is not None and int(size) != int(ref_size):
            raise AssertionError(f"Wrong size for dimension {idx}: got {size}, expected {ref_size}")


def profiled_function(fn):
    def decorator(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    decorator.__name__ = fn.__name__
    return decorator


@contextlib.contextmanager
def suppress_tracer_warnings():
    yield
