"""rdcfes_b200 -- B200-native hot path of rdcFEs (FE assembly of the RDC operators + Krylov solve).

The package holds only what the path needs: csrc/ (CUDA kernels + the C ABI of include/rdc.h), lib.py
(ctypes loader; fails loudly when the CUDA library is missing), system.py (host-side mirror of the
reference's TransientLinearImplicitSystem usage), params.py / synth.py (input keys, synthetic cases).
"""
from . import params, synth  # noqa: F401

__all__ = ["params", "synth"]
