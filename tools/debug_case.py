"""Stage-by-stage comparison of x_clip_b200 (GPU) with the CPU oracle on a golden case."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import clip_oracle as O  # noqa: E402
import x_clip_b200  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for case in sys.argv[1:]:
        gold = json.loads((ROOT / "tests" / "golden" / f"{case}.json").read_text())
        cfg = O.ClipConfig(**gold["cfg"])
        state = O.protocol_state_dict(cfg, gold["weight_seed"])
        text, image = O.protocol_inputs(cfg, gold["batch"], gold["input_seed"], gold["pad_fraction"])
        clip = x_clip_b200.CLIP(**gold["cfg"], visual_patch_dropout=0.).to(dev)
        clip.load_state_dict(state)
        clip.train()
        p = {k: v.clone() for k, v in state.items()}
        with torch.no_grad():
            o_loss, parts = O.clip_forward(p, text, image, cfg, return_parts=True)
            et, ei = clip(text.to(dev), image.to(dev), return_encodings=True)
            lat = clip(text.to(dev), image.to(dev), return_latents=True)
        loss = clip(text.to(dev), image.to(dev), return_loss=True)

        def rel(a, b):
            return ((a.float().cpu() - b).norm() / b.norm()).item()
        print(f"== {case}: B={gold['batch']} loss {loss.item():.6f} oracle {o_loss.item():.6f} gold {gold['loss']:.6f}")
        print("   enc_text rel", rel(et, parts["enc_text"]), " enc_image rel", rel(ei, parts["enc_image"]))
        print("   zt rel", rel(lat[0], parts["text_latents"]), " zi rel", rel(lat[1], parts["image_latents"]))
        if len(lat) == 4:
            print("   zt_x rel", rel(lat[2], parts["text_latents_extra"]), " zi_x rel", rel(lat[3], parts["image_latents_extra"]))


if __name__ == "__main__":
    main()
