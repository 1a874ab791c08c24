"""FILIP fine-grained loss (x_clip/x_clip.py:799-811) - scheduled after the CLS path."""
from .clip import Unsupported


def filip_loss(clip, zt, zi, zt_x, zi_x, text_mask):
    raise Unsupported("x_clip_b200: use_all_token_embeds (FILIP) loss kernel is not built yet")
