#!/usr/bin/env python
"""BASELINE.json configs[4]: end-to-end OneProt train step on B200 - random-init ESM-2 650M towers + projection heads +
fused ClipLoss - and the loss path's share of it.

    python tests/perf_e2e_step.py [--batch 256] [--seq-len 128] [--steps 5]                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/perf_e2e_step.py

What is timed (CUDA events, max over ranks), mirroring the reference's training_step for ONE modality pair
(oneprot_module.py:92-108) with the shipped configuration (configs/model/oneprot.yaml, components/sequence.yaml:
ESM-2 t33 650M frozen, no LoRA, mean pooling, MLP head 1280 -> 1152 -> 1024, Normalize; the modality tower gets
LearnableLogitScaling; ClipLoss(local_loss=True, gather_with_grad=True); AdamW on the trainable head parameters;
gradient-norm clipping at 1.0):

    tower fwd (x 2)   transformers' EsmModel built OFFLINE from a hand-written EsmConfig (no checkpoint, no network),
                      bf16 autocast, under no_grad because the towers are frozen - out of scope code (SURVEY.md
                      section 2), plain library PyTorch here
    heads fwd         oneprot_b200.BaseEncoder (pooling, LayerNorm, Linear on the tcgen05 GEMM, GELU, L2-normalise/scale)
    loss fwd + bwd    oneprot_b200.ClipLoss          (<- the hot path of this repository)
    heads bwd, clip_grad_norm_, AdamW step

and the same step with the loss (and only the loss) swapped for the unmodified reference ClipLoss from oracle/_ref run
eagerly - the reference's own arrangement.  The encoders are not this repository's product: the number that matters
here is the share of the step the loss path takes and how the swap changes the step.  Lives under tests/ because it
executes oracle/ (the reference classes of oracle/_ref as the comparison arm) - nothing under oneprot_b200/ does."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def build_tower(layers, hidden, heads, inter, dev):
    from transformers import EsmConfig, EsmModel
    cfg = EsmConfig(vocab_size=33, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=inter,
                    position_embedding_type="rotary", token_dropout=True, pad_token_id=1, mask_token_id=32,
                    emb_layer_norm_before=False, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = EsmModel(cfg, add_pooling_layer=False).to(dev).eval()
    for p in m.parameters():
        p.requires_grad = False                       # components/sequence.yaml:12  frozen: true
    return m, cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU")
    ap.add_argument("--seq-len", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=33)          # ESM-2 t33 650M: 33 x 1280, 20 heads, FFN 5120
    ap.add_argument("--hidden", type=int, default=1280)
    ap.add_argument("--heads", type=int, default=20)
    ap.add_argument("--inter", type=int, default=5120)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_step.json"))
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from oneprot_b200 import ClipLoss
    from oneprot_b200.heads import BaseEncoder
    from oracle.make_ref import import_reference

    torch.manual_seed(1234)                                       # same random-init weights on every rank (what DDP broadcasts)
    tower_seq, cfg = build_tower(args.layers, args.hidden, args.heads, args.inter, dev)
    tower_mod, _ = build_tower(args.layers, args.hidden, args.heads, args.inter, dev)   # struct-token tower: same architecture
    head_seq = BaseEncoder(args.hidden, 1024, proj_type="mlp", pooling_type="mean").to(dev).to(torch.bfloat16)
    head_mod = BaseEncoder(args.hidden, 1024, proj_type="mlp", pooling_type="mean", use_logit_scale=True,
                           learnable_logit_scale=True).to(dev).to(torch.bfloat16)
    params = [p for m in (head_seq, head_mod) for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4)
    g = torch.Generator().manual_seed(77 + rank)
    B, L = args.batch, args.seq_len
    tok_seq = torch.randint(4, 24, (B, L), generator=g).to(dev)
    tok_mod = torch.randint(4, 24, (B, L), generator=g).to(dev)
    lens = torch.randint(L // 2, L + 1, (B,), generator=g).to(dev)
    mask = (torch.arange(L, device=dev)[None, :] < lens[:, None])
    tok_seq = torch.where(mask, tok_seq, torch.full_like(tok_seq, cfg.pad_token_id))
    tok_mod = torch.where(mask, tok_mod, torch.full_like(tok_mod, cfg.pad_token_id))
    maskf = mask.float()

    ours = ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    ref = import_reference()
    theirs = ref[0].ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world) if ref else None

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step(loss_mod, marks=None):
        m = marks.append if marks is not None else (lambda e: None)
        m(ev())
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            h_seq = tower_seq(input_ids=tok_seq, attention_mask=mask.long()).last_hidden_state
            h_mod = tower_mod(input_ids=tok_mod, attention_mask=mask.long()).last_hidden_state
        m(ev())
        f_seq = head_seq.norm(head_seq.proj(head_seq.pooling(h_seq.to(torch.bfloat16), maskf)))
        f_mod = head_mod.norm(head_mod.proj(head_mod.pooling(h_mod.to(torch.bfloat16), maskf)))
        m(ev())
        opt.zero_grad(set_to_none=True)
        f_seq_l, f_mod_l = f_seq.detach().requires_grad_(True), f_mod.detach().requires_grad_(True)
        loss = loss_mod(f_seq_l, f_mod_l)                         # oneprot_module.py:103: loss_fn(sequence_features, modality_features)
        loss.backward()
        m(ev())                                                   # [2, 3] = the loss path alone: fwd + bwd to the features
        torch.autograd.backward([f_seq, f_mod], [f_seq_l.grad, f_mod_l.grad])
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        m(ev())
        return loss

    def measure(loss_mod):
        for _ in range(args.warmup):
            step(loss_mod)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        acc = [0.0] * 4
        loss = None
        for _ in range(args.steps):
            marks = []
            loss = step(loss_mod, marks)
            torch.cuda.synchronize()
            for i in range(4):
                acc[i] += marks[i].elapsed_time(marks[i + 1]) / args.steps
        t = torch.tensor(acc, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        towers, heads_f, loss_ms, rest = [float(x) for x in t]
        total = towers + heads_f + loss_ms + rest
        return {"ms_per_step": total, "towers_fwd_ms": towers, "heads_fwd_ms": heads_f, "loss_fwd_bwd_ms": loss_ms,
                "heads_bwd_clip_adamw_ms": rest, "loss_share_of_step": loss_ms / total,
                "trainable_part_ms": heads_f + loss_ms + rest, "loss_share_of_trainable_part": loss_ms / (heads_f + loss_ms + rest),
                "loss": float(loss.detach().float())}

    res = {"config": {"workload": "e2e OneProt train step, one modality pair: 2 x random-init ESM-2 650M towers (frozen, bf16 autocast, "
                                  "transformers EsmModel) + MLP heads + ClipLoss(local_loss=True, gather_with_grad=True) + AdamW on the heads",
                      "n_gpus": world, "batch_per_gpu": B, "seq_len": L, "global_batch": B * world, "layers": args.layers, "hidden": args.hidden,
                      "tower_params_M": sum(p.numel() for p in tower_seq.parameters()) / 1e6,
                      "trainable_params_M": sum(p.numel() for p in params) / 1e6},
           "ours": measure(ours)}
    if theirs is not None:
        res["reference_loss_eager"] = measure(lambda a, b: theirs(a, b, 1.0))
        res["step_speedup_from_swapping_the_loss"] = res["reference_loss_eager"]["ms_per_step"] / res["ours"]["ms_per_step"]
        res["loss_path_speedup"] = res["reference_loss_eager"]["loss_fwd_bwd_ms"] / res["ours"]["loss_fwd_bwd_ms"]
    res["samples_per_s"] = B * world / (res["ours"]["ms_per_step"] * 1e-3)
    if rank == 0:
        print(json.dumps(res), flush=True)
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
