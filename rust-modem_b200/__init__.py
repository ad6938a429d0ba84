"""rust-modem_b200: B200 (sm_100a) implementation of rust-modem's batched
modulate -> (AWGN) -> demodulate sample path.

The product is the C-ABI library (include/modem_gpu.h, csrc/); this package is only the
ctypes binding used by tests/ and bench.py.  The directory name contains a hyphen, so it
is loaded with `__graft_entry__.load_package()` (importlib) under the module name
`rust_modem_b200`.  There is no CPU fallback: without the built library or without an
sm_100 GPU every compute call raises.
"""
from . import capi
from .capi import (ModemError, Modem, ModemCfg, lib, build_library, library_path, host_constellation,
                   lowpass_taps, rrc_taps, hilbert_taps, host_phasor, Phasor, sample_freq, samples_per_symbol, FLAG_FUSED_MAC, FLAG_NO_TMEM,
                   STATEFUL_NAMES)

from .sharding import shard_range, shard_channels

__all__ = ["shard_range", "shard_channels", "ModemError", "Modem", "ModemCfg", "lib", "build_library", "library_path", "host_constellation",
           "lowpass_taps", "rrc_taps", "hilbert_taps", "host_phasor", "Phasor", "STATEFUL_NAMES", "sample_freq", "samples_per_symbol", "FLAG_FUSED_MAC", "FLAG_NO_TMEM"]
