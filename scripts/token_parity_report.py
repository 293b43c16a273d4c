"""Teacher-forced greedy-token parity report (run on a GPU): every one of the new x B argmax decisions of the product is
compared with the oracle's (= the reference's eager op sequence on cuBLAS/ATen, same GPU, same weights), not only the
decisions up to a sequence's first divergence as tests/test_gpu_model.py does.

The oracle generates `new` tokens greedily; the product is then driven through its reference-shaped forward face with
the ORACLE's token at every step (so both see the same context) and its argmax is compared position by position.  For
every disagreement the oracle's own top-2 margin is printed in bf16 ulps of the winning logit: the claim to verify is
that every disagreement is a near-tie (margin of a few ulps), i.e. a decision that fp32 summation order alone flips.

  python scripts/token_parity_report.py [model=opt-1.3b] [layers=3] [B=8] [S=256] [new=32]
Prints one JSON line; nothing is asserted (thresholds for a test are to be calibrated from this output)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lia_b200  # noqa: E402
from oracle import opt_ref  # noqa: E402


def oracle_model(m, device):
    dec = m.model.decoder
    hq = dec.layout.hq
    layers = []
    for v in dec.resident_views:
        w = {k: v[k] for k in ("ln1_w", "ln1_b", "o_w", "o_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")}
        w["q_w"], w["k_w"], w["v_w"] = v["qkv_w"][:hq], v["qkv_w"][hq:2 * hq], v["qkv_w"][2 * hq:]
        w["q_b"], w["k_b"], w["v_b"] = v["qkv_b"][:hq], v["qkv_b"][hq:2 * hq], v["qkv_b"][2 * hq:]
        layers.append({k: t.to(device) for k, t in w.items()})
    return {"H": m.config.num_attention_heads, "layers": layers, "embed_tokens": dec.embed_tokens, "embed_positions": dec.embed_positions,
            "final_ln_w": dec.final_ln_w, "final_ln_b": dec.final_ln_b}


def ulp(x):
    return 2.0 ** (torch.floor(torch.log2(x.abs().clamp_min(1e-30))) - 7)


def main(argv):
    name = argv[0] if argv else "opt-1.3b"
    L, B, S, new = (int(a) for a in (argv[1:5] + ["3", "8", "256", "32"][len(argv[1:5]):]))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cfg = lia_b200.modeling_opt.get_config(name)
    if L > 0:
        cfg.num_hidden_layers = L
    m = lia_b200.OPTForCausalLM(cfg, dev).init_weights(seed=3, bias_std=0.02, ln_std=0.05)
    om = oracle_model(m, dev)
    ids = torch.randint(3, cfg.vocab_size, (B, S), generator=torch.Generator().manual_seed(1234)).to(dev)
    ref_logits = []
    with torch.no_grad():
        ref = opt_ref.greedy_generate(om, ids, new, collect_logits=ref_logits)
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    logits, past = m(input_ids=ids, attention_mask=mask, max_new_tokens=new, prefill_policy=0, decoding_policy=0)
    agree, flips = 0, []
    for t in range(new):
        lg = logits[:, -1].float()
        lg[:, cfg.eos_token_id] = float("-inf")
        ours = lg.argmax(-1)
        want = ref[:, S + t]
        rl = ref_logits[t].float().clone()
        rl[:, cfg.eos_token_id] = float("-inf")
        for b in range(B):
            if ours[b] == want[b]:
                agree += 1
            else:
                margin = (rl[b, want[b]] - rl[b, ours[b]]).item()
                flips.append({"b": b, "t": t, "margin": margin, "ulps": margin / ulp(rl[b, want[b]]).item()})
        if t + 1 < new:
            mask = torch.cat([mask, mask.new_ones(B, 1)], dim=-1)
            logits, past = m(input_ids=want[:, None].contiguous(), attention_mask=mask, past_key_values=past, max_new_tokens=new)
    worst = max((f["ulps"] for f in flips), default=0.0)
    print(json.dumps({"model": name, "layers": cfg.num_hidden_layers, "B": B, "S": S, "new": new, "decisions": B * new, "identical": agree,
                      "flips": len(flips), "worst_flip_margin_ulps": worst, "flip_list": flips[:40]}))


if __name__ == "__main__":
    main(sys.argv[1:])
