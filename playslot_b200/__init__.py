"""
playslot_b200 -- B200-native (sm_100a) kernels behind PlaySlot's per-frame slot pipeline.

`playslot_b200.models` mirrors the reference's `src/models` package (same class names, constructor
kwargs, forward signatures and state_dict layout); every forward runs hand-written CUDA kernels from
`libplayslot_b200.so` through the C ABI declared in include/playslot_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
