"""auxgf_b200 -- B200-native DF-AGF2(None,0) self-energy moment build (drop-in for the hot path of
obackhouse/auxgf: ``OptRAGF2.build_part`` / ``OptUAGF2.build_part`` and ``auxgf/lib/libagf2``)."""
__version__ = "0.1.0"
