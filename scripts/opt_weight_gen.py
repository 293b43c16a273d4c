#!/usr/bin/env python
"""Dummy-weight generator for models no public checkpoint exists for (OPT-66B / OPT-175B).

Same job as the reference's examples/cpu/inference/python/llm/utils/opt-weight-gen.py (every parameter
~ U[0,1) in bf16, :61-62; ``--model opt-66b|opt-175b --save_dir DIR``), but it never builds the model
in memory: layers are generated one at a time from per-layer seeds and appended to this build's
native slab files (isca-2025-lia_b200/checkpoint.py), one file per tensor-parallel rank, ready to be
read straight into HBM or the pinned host arena by ``OPTForCausalLM.from_pretrained(DIR)``.

  python scripts/opt_weight_gen.py --model opt-175b --save_dir /data/opt-175b-tp8 --tp 8
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", type=str, default="opt-66b")
    ap.add_argument("--save_dir", type=str, required=True)
    ap.add_argument("--tp", type=int, default=1, help="tensor-parallel world size to shard for")
    ap.add_argument("--kind", choices=["dummy", "normal"], default="dummy",
                    help="dummy: U[0,1) as the reference; normal: N(0, 0.02) random-init (lia/modeling_opt.py:895-904)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--num-layers", type=int, default=0, help="override depth (debug)")
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args(argv)
    import lia_b200  # noqa: F401
    from lia_b200 import checkpoint
    from lia_b200.modeling_opt import get_config
    from lia_b200.weights import random_embeddings, random_layer
    cfg = get_config(args.model)
    if args.num_layers:
        cfg.num_hidden_layers = args.num_layers
    emb = random_embeddings(cfg.vocab_size, cfg.hidden_size, cfg.max_position_embeddings, args.seed * 100003 + 17,
                            args.device, args.kind, cfg.init_std, 0.0, cfg.pad_token_id)
    meta = checkpoint.write_slabs(
        args.save_dir, cfg,
        lambda i: random_layer(cfg.hidden_size, cfg.ffn_dim, args.seed * 100003 + 1000 + i, args.device, args.kind, cfg.init_std),
        emb, tp_world=args.tp)
    gb = meta["layout"]["numel"] * 2 * cfg.num_hidden_layers * args.tp / 1e9
    print(f"Model saved to {args.save_dir} ({cfg.name}, {cfg.num_hidden_layers} layers, tp {args.tp}, {gb:.2f} GB of layer slabs)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
