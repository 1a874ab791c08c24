"""2-GPU NCCL run of x_clip_b200.CLIP against the per-rank golden values recorded from two gloo
ranks of the (patched) reference: same global loss on every rank, per-rank gradients =
d loss / d(local latents) through the local encoders, full d temperature on every rank
(x_clip/distributed.py:41-56, x_clip.py:759-769).  Skipped on boxes with < 2 GPUs."""
import json
import os
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _worker(rank, world, case, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import clip_oracle as O
        import x_clip_b200
        gold = json.loads((GOLD / f"{case}.json").read_text())
        cfg = O.ClipConfig(**gold["cfg"])
        state = O.protocol_state_dict(cfg, gold["weight_seed"])
        text, image = O.protocol_inputs(cfg, world * gold["per_rank"], gold["input_seed"], gold["pad_fraction"])
        t, im = text.chunk(world)[rank].to(dev), image.chunk(world)[rank].to(dev)
        clip = x_clip_b200.CLIP(**gold["cfg"], visual_patch_dropout=0.).to(dev)   # after init_process_group
        assert clip.requires_all_gather
        clip.load_state_dict(state)
        clip.train()
        loss = clip(t, im, return_loss=True)
        loss.backward()
        torch.cuda.synchronize()
        norms = {k: p.grad.double().norm().item() for k, p in clip.named_parameters() if p.grad is not None}
        gn = sum(v * v for v in norms.values()) ** 0.5
        q.put((rank, loss.item(), gn, clip.temperature.grad.item(), norms))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["sharded_plain", "sharded_dcl_extra"])
def test_two_gpu_matches_reference_ranks(case):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    gold = json.loads((GOLD / f"{case}.json").read_text())
    world = gold["world"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, case, 29811, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _q, time as _t
    t0 = _t.time()
    while len(res) < world and _t.time() - t0 < 300:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert len(res) == world, "a rank died"
    res.sort()
    for (rank, loss, gn, dtemp, norms), ref in zip(res, gold["ranks"]):
        assert abs(loss - ref["loss"]) <= 1e-3 * abs(ref["loss"]), (rank, loss, ref["loss"])
        assert abs(gn - ref["grad_norm"]) <= 1.5e-2 * ref["grad_norm"], (rank, gn, ref["grad_norm"])
        assert abs(dtemp - ref["dtemperature"]) <= 1e-2 * abs(ref["dtemperature"]) + 1e-3
        for k, v in ref["grad_norms"].items():
            if v > 1e-2 * ref["grad_norm"]:
                assert abs(norms[k] - v) <= 5e-2 * v, (rank, k, norms[k], v)
    assert abs(res[0][1] - res[1][1]) < 1e-6          # identical global loss on both ranks
    assert abs(res[0][3] - res[1][3]) < 1e-6          # identical (all-reduced) d temperature


def _filip_worker(rank, world, case, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import clip_oracle as O
        import x_clip_b200
        gold = json.loads((GOLD / f"{case}.json").read_text())
        cfg = O.ClipConfig(**gold["cfg"])
        state = O.protocol_state_dict(cfg, gold["weight_seed"])
        text, image = O.protocol_inputs(cfg, gold["batch"], gold["input_seed"], gold["pad_fraction"])
        t, im = text.chunk(world)[rank].to(dev), image.chunk(world)[rank].to(dev)
        clip = x_clip_b200.CLIP(**gold["cfg"], visual_patch_dropout=0.).to(dev)
        clip.load_state_dict(state)
        clip.train()
        loss = clip(t, im, return_loss=True)
        loss.backward()
        bad = []
        if True:
            # sum of the per-rank gradients == gradient of the full-batch loss (temperature: W x)
            grads = {}
            for k, p in clip.named_parameters():
                if p.grad is None:
                    continue
                g = p.grad.clone()
                dist.all_reduce(g)
                grads[k] = (g / world if k == "temperature" else g).float().cpu()
            if rank == 0:
                pr = {k: v.clone().requires_grad_(True) for k, v in state.items()}
                ref = O.clip_forward(pr, text, image, cfg)
                ref.backward()
                total = sum((v.grad.double() ** 2).sum() for v in pr.values() if v.grad is not None).sqrt().item()
                for k, g in grads.items():
                    og = pr[k].grad
                    if k == "temperature" or og is None or og.norm().item() < 1e-3 * total:
                        continue
                    cos = torch.nn.functional.cosine_similarity(g.flatten().double(), og.flatten().double(), dim=0).item()
                    nrel = abs(g.norm().item() - og.norm().item()) / og.norm().item()
                    if cos < 0.99 or nrel > 5e-2:
                        bad.append((k, round(cos, 4), round(nrel, 4)))
                q.put(("ref", ref.item(), pr["temperature"].grad.item(), bad))
        q.put((rank, loss.item(), clip.temperature.grad.item()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["tiny_filip", "tiny_filip_dcl_extra"])
def test_two_gpu_filip_matches_full_batch(case):
    """FILIP under data parallelism (which the reference cannot run): 2 ranks x 2 pairs must give
    the single-process loss of the 4-pair golden case and gradients that sum to its gradients."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    gold = json.loads((GOLD / f"{case}.json").read_text())
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_filip_worker, args=(r, world, case, 29821, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue as _q, time as _t
    t0 = _t.time()
    while len(res) < world + 1 and _t.time() - t0 < 300:
        try:
            res.append(q.get(timeout=2))
        except _q.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert len(res) == world + 1, "a rank died"
    ref = [r for r in res if r[0] == "ref"][0]
    ranks = sorted(r for r in res if r[0] != "ref")
    assert abs(ref[1] - gold["loss"]) < 1e-5
    for rank, loss, dtemp in ranks:
        assert abs(loss - gold["loss"]) <= 1e-3 * abs(gold["loss"]), (rank, loss, gold["loss"])
        assert abs(dtemp - gold["dtemperature"]) <= 1e-2 * abs(gold["dtemperature"]) + 1e-3
    assert abs(ranks[0][1] - ranks[1][1]) < 1e-6
    assert not ref[3], ref[3]
