"""Pin oracle/magvit_oracle.py to the fixture produced by the reference's own Encoder / LFQ / Decoder."""
import hashlib

import numpy as np
import torch

from helpers import load_golden, rel_fro
from oracle import magvit_oracle as MO


def _sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def test_magvit_oracle_matches_reference_fixture():
    z = load_golden("magvit")
    cfg = MO.VQOracleConfig()
    sd = MO.init_vq_state_dict(cfg, seed=int(z["seed"]))
    assert _sha(sd) == str(z["sd_sha"])
    g = torch.Generator().manual_seed(int(z["img_seed"]))
    img = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    torch.set_num_threads(8)
    with torch.no_grad():
        lat = MO.encoder_forward(sd, cfg, img)
        assert rel_fro(lat, torch.from_numpy(z["z"])) < 1e-5
        ids = MO.lfq_indices(lat)
        assert torch.equal(ids, torch.from_numpy(z["ids"]).long())
        rec = MO.decoder_forward(sd, cfg, MO.codebook_entry(ids, cfg.z_channels))
        assert rel_fro(rec[:, :, ::8, ::8], torch.from_numpy(z["rec_sub"])) < 1e-5
        tok = torch.randint(0, 2 ** 18, (1, 16, 16), generator=g)
        assert torch.equal(tok, torch.from_numpy(z["tok_le"]).long())
        img_le = MO.decode_tokens(sd, cfg, tok, little_endian=True)
        assert rel_fro(img_le[:, :, ::8, ::8], torch.from_numpy(z["img_le_sub"])) < 1e-5


def test_lfq_bit_order_roundtrip():
    ids = torch.randint(0, 2 ** 18, (3, 4, 4), generator=torch.Generator().manual_seed(0))
    q = MO.codebook_entry(ids, 18)
    assert set(q.unique().tolist()) <= {-1.0, 1.0}
    assert torch.equal(MO.lfq_indices(q), ids)                       # big-endian index of the +-1 codes
    # little-endian (dataset) convention = channel flip (visualize.py:115)
    le = ((q.flip(1) > 0).long() * (2 ** torch.arange(18)).view(1, 18, 1, 1)).sum(1)
    assert torch.equal(le, ids)


def test_depth_to_space_dcr():
    x = torch.arange(2 * 8 * 3 * 3, dtype=torch.float32).reshape(2, 8, 3, 3)
    y = MO.depth_to_space(x, 2)
    assert y.shape == (2, 2, 6, 6)
    # channel index = (b1*2 + b2)*C' + c'  ->  pixel (2h+b1, 2w+b2)
    assert y[1, 1, 2 * 1 + 1, 2 * 2 + 0] == x[1, (1 * 2 + 0) * 2 + 1, 1, 2]
