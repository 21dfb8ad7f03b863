"""Caduceus model configuration, mirroring the keys of the HF-hub ``config.json``
the reference loads through ``AutoModelForMaskedLM.from_pretrained(..., trust_remote_code=True)``
(reference: src/zero_shot_score.py:91; key list: SURVEY.md section 5 "Config / flags").

Only the combination the PlantCaduceus checkpoints use is supported by the engine
(rcps=True, bidirectional=True, strategy "add", tied in/out projections, RMSNorm, fused add+norm);
anything else is rejected at construction instead of silently diverging.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

# Tokenizer default (songlab/tokenizer-dna-mlm, 7 raw tokens padded to 8 rows:
# reference pretrain/scripts/train_plant_BERT.py:27-28, notebooks/examples.ipynb:66).
DEFAULT_VOCAB = {"[PAD]": 0, "[MASK]": 1, "[UNK]": 2, "a": 3, "c": 4, "g": 5, "t": 6}


def build_complement_map(vocab: Dict[str, int], vocab_size: int) -> Dict[int, int]:
    """Complement map over token ids: a<->t, c<->g, everything else fixed.

    Follows reference pretrain/llmlib/architectures/models/mamba/caduceus.py:100-105, extended
    with identity entries up to the padded vocab size (same file, :124-125).
    """
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a"}
    out = {}
    for tok, idx in sorted(vocab.items(), key=lambda kv: kv[1]):
        out[idx] = vocab[comp[tok]] if tok in comp else idx
    for idx in range(vocab_size):
        out.setdefault(idx, idx)
    return out


@dataclass
class CaduceusConfig:
    d_model: int = 384
    n_layer: int = 20
    vocab_size: int = 8
    ssm_cfg: Dict = field(default_factory=lambda: dict(
        d_state=16, d_conv=4, expand=2, dt_rank="auto", conv_bias=True, bias=False))
    rms_norm: bool = True
    residual_in_fp32: bool = False
    fused_add_norm: bool = True
    norm_epsilon: float = 1e-5
    pad_vocab_size_multiple: int = 8
    bidirectional: bool = True
    bidirectional_strategy: str = "add"
    bidirectional_weight_tie: bool = True
    rcps: bool = True
    complement_map: Optional[Dict[int, int]] = None
    model_type: str = "caduceus"

    def __post_init__(self):
        if self.vocab_size % self.pad_vocab_size_multiple != 0:
            self.vocab_size += self.pad_vocab_size_multiple - self.vocab_size % self.pad_vocab_size_multiple
        if self.complement_map is None:
            self.complement_map = build_complement_map(DEFAULT_VOCAB, self.vocab_size)
        else:
            self.complement_map = {int(k): int(v) for k, v in self.complement_map.items()}
            for idx in range(self.vocab_size):
                self.complement_map.setdefault(idx, idx)

    # derived sizes -----------------------------------------------------------------------
    @property
    def d_state(self) -> int:
        return int(self.ssm_cfg.get("d_state", 16))

    @property
    def d_conv(self) -> int:
        return int(self.ssm_cfg.get("d_conv", 4))

    @property
    def expand(self) -> int:
        return int(self.ssm_cfg.get("expand", 2))

    @property
    def d_inner(self) -> int:
        return self.expand * self.d_model

    @property
    def dt_rank(self) -> int:
        r = self.ssm_cfg.get("dt_rank", "auto")
        return math.ceil(self.d_model / 16) if r == "auto" else int(r)

    # Mamba-2 / SSD mixer (PlantCAD2, reference docs/PlantCAD2-overview.md:17-21).  mamba_ssm's block factory reads the
    # layer type from ssm_cfg["layer"] ("Mamba1" default, "Mamba2"); the Mamba2 hyper-parameters keep its names.
    @property
    def mixer(self) -> str:
        return str(self.ssm_cfg.get("layer", "Mamba1"))

    @property
    def is_mamba2(self) -> bool:
        return self.mixer == "Mamba2"

    @property
    def headdim(self) -> int:
        return int(self.ssm_cfg.get("headdim", 64))

    @property
    def ngroups(self) -> int:
        return int(self.ssm_cfg.get("ngroups", 1))

    @property
    def nheads(self) -> int:
        return self.d_inner // self.headdim

    @property
    def conv_dim(self) -> int:
        """Channels that go through the depthwise conv: x | B | C (Mamba2.conv1d)."""
        return self.d_inner + 2 * self.ngroups * self.d_state

    @property
    def d_in_proj(self) -> int:
        """Mamba2.in_proj output width: z | x | B | C | dt."""
        return 2 * self.d_inner + 2 * self.ngroups * self.d_state + self.nheads

    def validate_supported(self) -> None:
        """Raise ValueError for configurations the engine does not implement."""
        bad = []
        if not self.rcps:
            bad.append("rcps=False")
        if not self.bidirectional:
            bad.append("bidirectional=False")
        if self.bidirectional_strategy != "add":
            bad.append(f"bidirectional_strategy={self.bidirectional_strategy!r}")
        if not self.bidirectional_weight_tie:
            bad.append("bidirectional_weight_tie=False")
        if not self.rms_norm:
            bad.append("rms_norm=False")
        if self.mixer not in ("Mamba1", "Mamba2"):
            bad.append(f"ssm_cfg.layer={self.mixer!r}")
        if self.is_mamba2:
            # the SSD kernel is specialised for the PlantCAD2 shape: 64-wide heads, 64 states, one B/C group
            if self.d_state != 64:
                bad.append(f"d_state={self.d_state} (Mamba2 engine is specialised for 64)")
            if self.headdim != 64:
                bad.append(f"headdim={self.headdim} (Mamba2 engine is specialised for 64)")
            if self.ngroups != 1:
                bad.append(f"ngroups={self.ngroups} (Mamba2 engine is specialised for 1)")
            if self.d_inner % (2 * self.headdim) != 0:
                bad.append(f"nheads={self.nheads} must be even")
            for key, want in (("rmsnorm", True), ("norm_before_gate", False), ("D_has_hdim", False)):
                if self.ssm_cfg.get(key, want) != want:
                    bad.append(f"ssm_cfg.{key}={self.ssm_cfg.get(key)!r}")
        elif self.d_state != 16:
            bad.append(f"d_state={self.d_state} (engine is specialised for 16)")
        if self.d_conv != 4:
            bad.append(f"d_conv={self.d_conv} (engine is specialised for 4)")
        if self.ssm_cfg.get("bias", False):
            bad.append("ssm_cfg.bias=True")
        if not self.ssm_cfg.get("conv_bias", True):
            bad.append("ssm_cfg.conv_bias=False")
        # the engine's own limits (pcad_create, csrc/pcad.cu): reject here instead of failing later at .to("cuda")
        if self.d_model % 128 != 0 or not (0 < self.d_model <= 2048):
            bad.append(f"d_model={self.d_model} (engine needs a multiple of 128, <= 2048)")
        if self.expand != 2:
            bad.append(f"expand={self.expand} (engine is specialised for 2)")
        if not self.is_mamba2 and (self.dt_rank <= 0 or self.dt_rank % 8 != 0):
            bad.append(f"dt_rank={self.dt_rank} (engine needs a multiple of 8)")
        if not self.fused_add_norm:
            bad.append("fused_add_norm=False (different module tree, key names and rounding order)")
        if self.vocab_size != 8:
            bad.append(f"vocab_size={self.vocab_size} (engine LM head is specialised for 8 rows)")
        if bad:
            raise ValueError("unsupported Caduceus configuration: " + ", ".join(bad))

    def to_dict(self) -> Dict:
        return dict(d_model=self.d_model, n_layer=self.n_layer, vocab_size=self.vocab_size,
                    ssm_cfg=dict(self.ssm_cfg), rms_norm=self.rms_norm,
                    residual_in_fp32=self.residual_in_fp32, fused_add_norm=self.fused_add_norm,
                    norm_epsilon=self.norm_epsilon, pad_vocab_size_multiple=self.pad_vocab_size_multiple,
                    bidirectional=self.bidirectional, bidirectional_strategy=self.bidirectional_strategy,
                    bidirectional_weight_tie=self.bidirectional_weight_tie, rcps=self.rcps,
                    complement_map={str(k): v for k, v in self.complement_map.items()},
                    model_type=self.model_type)

    @classmethod
    def from_dict(cls, d: Dict) -> "CaduceusConfig":
        keys = {f for f in cls.__dataclass_fields__}
        return cls(**{k: v for k, v in d.items() if k in keys})


# The four published sizes (reference README.md:60-63; shapes SURVEY.md Appendix A).
PRESETS = {
    "PlantCaduceus_l20": dict(d_model=384, n_layer=20),
    "PlantCaduceus_l24": dict(d_model=512, n_layer=24),
    "PlantCaduceus_l28": dict(d_model=768, n_layer=28),
    "PlantCaduceus_l32": dict(d_model=1024, n_layer=32),
}
# PlantCAD2 (reference docs/PlantCAD2-overview.md:17-21; 8192-bp context).  The Mamba-2 hyper-parameters are not stated in
# the reference; d_state 64 / headdim 64 / ngroups 1 / expand 2 with tied in/out projections are the values that reproduce
# the published parameter counts 88 M / 311 M / 694 M (img/PlantCAD2-difference.jpg) to three digits (tests/test_oracle.py).
_MAMBA2 = dict(layer="Mamba2", d_state=64, d_conv=4, expand=2, headdim=64, ngroups=1, conv_bias=True, bias=False, chunk_size=256)
PRESETS.update({
    "PlantCAD2-Small-l24-d0768": dict(d_model=768, n_layer=24, ssm_cfg=dict(_MAMBA2)),
    "PlantCAD2-Medium-l48-d1024": dict(d_model=1024, n_layer=48, ssm_cfg=dict(_MAMBA2)),
    "PlantCAD2-Large-l48-d1536": dict(d_model=1536, n_layer=48, ssm_cfg=dict(_MAMBA2)),
})
_ALIASES = {"cad2-small": "PlantCAD2-Small-l24-d0768", "cad2-medium": "PlantCAD2-Medium-l48-d1024",
            "cad2-large": "PlantCAD2-Large-l48-d1536"}


def preset(name: str, **overrides) -> CaduceusConfig:
    key = name.split("/")[-1]
    key = _ALIASES.get(key, key)
    if key in PRESETS:
        kw = dict(PRESETS[key])
    elif "PlantCaduceus_" + key in PRESETS:
        kw = dict(PRESETS["PlantCaduceus_" + key])
    else:
        raise KeyError(f"unknown preset {name!r}; known: {sorted(PRESETS)}")
    if "ssm_cfg" in kw:
        kw["ssm_cfg"] = dict(kw["ssm_cfg"])
    kw.update(overrides)
    return CaduceusConfig(**kw)
