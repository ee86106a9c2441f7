"""Drop-in mirror of the reference's `torch_utils.ops` package: same module and function names and keyword
signatures, bodies replaced by calls into libmgf_sm100a.so through the C ABI in include/mgf.h."""
