"""GenieConfig — field-for-field mirror of the reference dataclass (genie/config.py:7-55) so the
reference's JSON files (e.g. genie/configs/magvit_n32_h8_d256.json) and `config.json` of HF checkpoints
load unchanged.  `factored_vocab_size` is derived exactly like the reference (config.py:54-55)."""
import json
from dataclasses import dataclass, asdict


def nth_root(x, n):
    root = round(x ** (1 / n))
    assert root ** n == x, (x, n, root)
    return root


@dataclass
class GenieConfig:
    num_layers: int
    num_heads: int
    d_model: int
    T: int = 16
    S: int = 256
    image_vocab_size: int = 262144
    use_mup: bool = False

    num_factored_vocabs: int = 1
    factored_vocab_size: int = None

    max_corrupt_rate: float = 0.2
    non_mlm_ratio: float = 0.5
    num_prompt_frames: int = 8

    qkv_bias: bool = False
    proj_bias: bool = True
    attn_drop: float = 0.0
    qk_norm: bool = True

    mlp_ratio: float = 4.0
    mlp_drop: float = 0.0
    mlp_bias: bool = True

    def save_pretrained(self, json_path):
        with open(json_path, "w") as f:
            json.dump(vars(self), f)

    @classmethod
    def from_pretrained(cls, json_path):
        with open(json_path, "r") as f:
            config = json.load(f)
        return cls(**config)

    def shallow_copy(self):
        return GenieConfig(**vars(self))

    def to_dict(self):
        return asdict(self)

    def __post_init__(self):
        self.factored_vocab_size = nth_root(self.image_vocab_size, self.num_factored_vocabs)
        # attn_drop / mlp_drop are accepted and ignored: dropout is the identity in eval mode, and this path is
        # inference-only (the reference loads such checkpoints for generate.py / evaluate.py too).  Training
        # (train.py) is out of scope: SURVEY.md section 2 row 9.
