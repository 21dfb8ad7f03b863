"""plantcaduceus_b200 -- B200-native (sm_100a) engine for the PlantCaduceus masked-LM forward pass and
zero-shot variant scoring.  Host-side mirror of the reference's HF interface over libpcad.so."""
from .configuration import CaduceusConfig, PRESETS, preset, build_complement_map, DEFAULT_VOCAB
from .tokenizer import CharDNATokenizer
from .weights import random_init_state_dict, count_parameters

__all__ = ["CaduceusConfig", "PRESETS", "preset", "build_complement_map", "DEFAULT_VOCAB", "CharDNATokenizer",
           "random_init_state_dict", "count_parameters", "CaduceusForMaskedLM", "MaskedLMOutput"]


def __getattr__(name):
    # modeling imports (and loads) libpcad.so; keep `import plantcaduceus_b200` usable for host-only tools.
    if name in ("CaduceusForMaskedLM", "MaskedLMOutput"):
        from . import modeling
        return getattr(modeling, name)
    raise AttributeError(name)
