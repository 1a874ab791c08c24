"""Live cross-check of the CPU oracle against the REFERENCE on randomly drawn small configurations.

Runs only where /root/reference exists (the build container); on the GPU box it is skipped - the
committed fixtures (tests/golden/*.json, test_oracle_golden.py) carry the same pin there.  The
configurations sweep what the fixtures cannot enumerate: widths, head counts, depths, sequence
lengths, batch sizes, pad fractions, patch dropout and every combination of the loss flags."""
import importlib.util
import random
import sys
from pathlib import Path

import pytest
import torch

from oracle import clip_oracle as O

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present")


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", Path(__file__).parent / "golden" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _draw(rng: random.Random):
    heads_t, heads_i = rng.choice([1, 2, 4]), rng.choice([1, 2, 4])
    patch = rng.choice([8, 16])
    side = patch * rng.choice([2, 3, 4])
    cfg = dict(dim_text=64 * rng.choice([1, 2, 4]), dim_image=64 * rng.choice([1, 2, 4]),
               dim_latent=64 * rng.choice([1, 2, 4]), num_text_tokens=rng.choice([32, 97, 300]),
               text_enc_depth=rng.choice([1, 2, 3]), text_seq_len=rng.choice([5, 16, 33]), text_heads=heads_t,
               visual_enc_depth=rng.choice([1, 2]), visual_heads=heads_i, visual_image_size=side,
               visual_patch_size=patch)
    flags = dict(decoupled_contrastive_learning=rng.random() < 0.4, extra_latent_projection=rng.random() < 0.4,
                 use_all_token_embeds=rng.random() < 0.3)
    # (FILIP with a causal text tower does not run in the reference itself: 15 text tokens after the CLS
    #  strip against a 16-wide mask, x_clip.py:705 / :806)
    text_kind = rng.choice(["plain", "plain", "rotary"] + ([] if flags["use_all_token_embeds"] else ["causal"]))
    if text_kind == "rotary":
        flags["text_rotary_pos_emb"] = True
    elif text_kind == "causal":
        flags.update(text_causal_mask=True, text_eos_id=cfg["num_text_tokens"] - 1)
    cfg.update({k: v for k, v in flags.items() if v})
    batch = rng.choice([2, 3, 5, 8])
    pad = rng.choice([0.0, 0.2, 0.5])
    n_patches = (side // patch) ** 2
    drop = rng.choice([0.0, 0.0, 0.5]) if n_patches >= 4 else 0.0
    return cfg, batch, pad, drop


@pytest.mark.parametrize("seed", range(16))
def test_oracle_equals_live_reference_on_random_configuration(seed):
    mg = _mg()
    x_clip = mg.import_reference()
    rng = random.Random(1000 + seed)
    cfg_kwargs, batch, pad, drop = _draw(rng)
    torch.set_num_threads(4)
    out = mg.run_case(x_clip, f"random_{seed}", cfg_kwargs, batch, pad, drop)   # reference, + its own oracle cross-check
    cfg = O.ClipConfig(**cfg_kwargs)
    state = O.protocol_state_dict(cfg, mg.WEIGHT_SEED)
    text, image = O.protocol_inputs(cfg, batch, mg.INPUT_SEED, pad)
    keep = None if out["keep"] is None else torch.tensor(out["keep"])
    p = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    loss = O.clip_forward(p, text, image, cfg, keep=keep)
    loss.backward()
    assert abs(loss.item() - out["loss"]) <= 5e-6 * max(1.0, abs(out["loss"])), (cfg_kwargs, loss.item(), out["loss"])
    assert abs(p["temperature"].grad.item() - out["dtemperature"]) <= 5e-6
    for k, n in out["grad_norms"].items():
        g = p[k].grad
        assert g is not None, (cfg_kwargs, k)
        assert abs(g.double().norm().item() - n) <= 5e-4 * max(n, 1e-5) + 1e-8, (cfg_kwargs, k)
