"""Host side of the B200-native GAST-Net lifting path.

`engine`  : binds nn.Module parameter trees to the C-ABI library and launches the kernels.
`_lib`    : ctypes loader for csrc/libgast_b200.so (fails loudly when missing).
`synth`   : deterministic, torch-RNG-independent weights/inputs for tests and bench.
`dist`    : clip sharding + gradient all-reduce over torch.distributed.
"""
__all__ = ['engine', 'synth']
