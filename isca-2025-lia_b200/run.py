"""Command line of the reference's benchmark driver, on the B200 build.

Same flags as examples/cpu/inference/python/llm/run.py:169-215 and
single_instance/run_generation.py:59-118, same printed summary lines
(run_generation.py:330-354) -- plus tokens/s and roofline fractions, which the reference
never prints.  When -m is a checkpoint directory that carries a tokenizer, --prompt (or the
"opt" pool of a prompt.json next to run.py / in the working directory) is tokenised and the output
decoded exactly as run_generation.py:260-285, 319 does; otherwise (random-init models, no
tokenizer offline) prompts are synthetic token ids: ``randint(3, vocab)`` of length --input-tokens,
identical for every batch row as in run_generation.py:285.

Policy mapping: 0/2/3/4 run everything on the GPU (non-resident layers streamed from pinned
host memory); 1 (the reference's default: full-CPU IPEX/AMX) is refused with a pointer to
``bench.py --impl reference``.
"""
import argparse
import json
import os
import sys
import time


def build_parser():
    p = argparse.ArgumentParser("Generation script (bf16 path, B200 build)")
    p.add_argument("-m", "--model-id", "--model-name-or-path", dest="model_id", type=str, default="facebook/opt-30b")
    p.add_argument("--dtype", type=str, choices=["float32", "bfloat16"], default="bfloat16")
    p.add_argument("--input-tokens", default="32", type=str)
    p.add_argument("--max-new-tokens", default=32, type=int)
    p.add_argument("--prompt", default=None, type=str)
    p.add_argument("--greedy", action="store_true")
    p.add_argument("--ipex", action="store_true")
    p.add_argument("--deployment-mode", action="store_true")
    p.add_argument("--profile", action="store_true")
    p.add_argument("--benchmark", action="store_true")
    p.add_argument("--num-iter", default=100, type=int)
    p.add_argument("--num-warmup", default=10, type=int)
    p.add_argument("--batch-size", default=1, type=int)
    p.add_argument("--token-latency", action="store_true")
    p.add_argument("--prefill-policy", default=1, type=int)
    p.add_argument("--decoding-policy", default=1, type=int)
    p.add_argument("--no-overlap", action="store_true")
    p.add_argument("--pin-weight", action="store_true")
    p.add_argument("--gpu-percentage", default=0, type=int)
    p.add_argument("--num-minibatch", default=1, type=int)
    p.add_argument("--enable-cxl", action="store_true")
    # additions of this build
    p.add_argument("--tp", default=0, type=int, help="tensor-parallel world size (default: WORLD_SIZE)")
    p.add_argument("--dummy-weights", action="store_true", help="U[0,1) weights as utils/opt-weight-gen.py")
    p.add_argument("--num-layers", default=0, type=int, help="override depth (debug)")
    p.add_argument("--seed", default=0, type=int)
    return p


def load_tokenizer(model_id):
    """The tokenizer the reference loads next to the model (run_generation.py:169-171), when the checkpoint directory
    carries one (tokenizer.json or vocab.json + merges.txt); None otherwise -- then prompts are synthetic token ids."""
    if not os.path.isdir(model_id):
        return None
    names = set(os.listdir(model_id))
    if not ({"tokenizer.json"} & names or {"vocab.json", "merges.txt"} <= names):
        return None
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(model_id, local_files_only=True)


def pick_prompt(args, search_dirs):
    """run_generation.py:260-276: --prompt, else the entry of prompt.json's "opt" pool for --input-tokens."""
    if args.prompt is not None:
        return args.prompt
    for d in search_dirs:
        pj = os.path.join(d, "prompt.json")
        if os.path.exists(pj):
            pool = json.load(open(pj)).get("opt", {})
            if int(args.input_tokens) > 8192 and "8192" in pool:
                return pool["8192"] * int(int(args.input_tokens) / 8192)
            if str(args.input_tokens) in pool:
                return pool[str(args.input_tokens)]
    return None


def build_inputs(args, vocab_size, tokenizer=None, search_dirs=()):
    """-> (input_ids [B, S] int64 on the host, prompt text or None).  With a tokenizer and a prompt the ids are the
    tokenizer's (every batch row the same text, run_generation.py:285); otherwise synthetic ids of --input-tokens."""
    import torch
    text = pick_prompt(args, search_dirs) if tokenizer is not None else None
    if tokenizer is not None and text is None and args.prompt is None:
        print("[WARN] tokenizer found but no --prompt / prompt.json entry for --input-tokens %s: using synthetic token ids"
              % args.input_tokens, file=sys.stderr)
    if text is not None:
        row = tokenizer(text, return_tensors="pt").input_ids                       # [1, S]
        print("---- Prompt size:", row.size(dim=1))                                # run_generation.py:279
    else:
        g = torch.Generator().manual_seed(1234)
        row = torch.randint(3, vocab_size, (1, int(args.input_tokens)), generator=g)
    return row.expand(args.batch_size, row.shape[1]).contiguous().to(torch.int64), text


def main(argv=None):
    args = build_parser().parse_args(argv)
    print(args)
    if args.dtype != "bfloat16":
        print("error: only --dtype bfloat16 is implemented on the B200 path", file=sys.stderr)
        return 2
    if args.prefill_policy == 1 or args.decoding_policy == 1:
        print("error: policy 1 is the reference's full-CPU (IPEX/AMX) path, which this build does not contain; "
              "use --prefill-policy 0 --decoding-policy 0 (all-GPU, streamed weights) or time the CPU baseline with "
              "`python bench.py --impl reference`.", file=sys.stderr)
        return 2
    import torch
    import lia_b200
    from lia_b200 import tp
    from lia_b200.modeling_opt import get_config
    rank, world = tp.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    S = int(args.input_tokens)
    dev = torch.device("cuda", local)
    if os.path.isdir(args.model_id):
        # a checkpoint directory (HF safetensors / pytorch_model.bin, or native slabs): run_generation.py:159-167
        from lia_b200 import checkpoint
        cfg = checkpoint.open_checkpoint(args.model_id).config
        if args.num_layers:
            cfg.num_hidden_layers = args.num_layers
        cfg.token_latency = bool(args.token_latency)
        cfg.text_max_length = S + args.max_new_tokens                 # run_generation.py:149
        model = lia_b200.OPTForCausalLM.from_pretrained(args.model_id, dev, gpu_percentage=args.gpu_percentage,
                                                        tp_rank=rank, tp_world=world, config=cfg)
    else:
        cfg = get_config(args.model_id)
        if args.num_layers:
            cfg.num_hidden_layers = args.num_layers
        cfg.token_latency = bool(args.token_latency)
        cfg.text_max_length = S + args.max_new_tokens                 # run_generation.py:149
        model = lia_b200.OPTForCausalLM(cfg, dev, tp_rank=rank, tp_world=world)
        model.init_weights(seed=args.seed, kind="dummy" if args.dummy_weights else "normal",
                           gpu_percentage=args.gpu_percentage)
    tokenizer = load_tokenizer(args.model_id)
    here = os.path.dirname(os.path.abspath(__file__))
    input_ids, text = build_inputs(args, cfg.vocab_size, tokenizer, (os.getcwd(), here, os.path.dirname(here)))
    input_ids = input_ids.pin_memory()                                            # run_generation.py:285
    S = input_ids.shape[1]
    generate_kwargs = dict(do_sample=False, temperature=0.9, num_beams=1 if args.greedy else 1,
                           max_new_tokens=args.max_new_tokens, min_new_tokens=args.max_new_tokens,
                           prefill_policy=args.prefill_policy, decoding_policy=args.decoding_policy,
                           no_overlap=args.no_overlap, pin_weight=args.pin_weight, gpu_percentage=args.gpu_percentage,
                           num_minibatch=args.num_minibatch, enable_cxl=args.enable_cxl)      # run_generation.py:179-182
    total_time, total_list = 0.0, []
    num_iter, num_warmup = args.num_iter, args.num_warmup
    if args.profile:
        # run_generation.py:220-221, 291-307: five profiled generate() calls (wait 1, warm-up 3, active 1), then the
        # operator table -- here with the CUDA activity as well, sorted by device time, since the work is on the GPU
        from torch.profiler import ProfilerActivity, profile, schedule

        def trace_handler(prof):
            if rank == 0:
                print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30))
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], schedule=schedule(wait=1, warmup=3, active=1),
                     on_trace_ready=trace_handler) as prof:
            for _ in range(5):
                model.generate(input_ids, **generate_kwargs)
                prof.step()
    for i in range(num_iter):
        tic = time.time()
        if text is not None:                    # the reference times tokenizer -> generate -> batch_decode (run_generation.py:308-319)
            input_ids = tokenizer([text] * args.batch_size, return_tensors="pt").input_ids.pin_memory()
        output = model.generate(input_ids, **generate_kwargs)
        gen_ids = output[0] if args.token_latency else output
        gen_text = tokenizer.batch_decode(gen_ids, skip_special_tokens=True) if text is not None else None   # run_generation.py:319
        toc = time.time()
        total_new_tokens = [int(o.shape[0]) - S for o in gen_ids]
        if rank == 0:
            if gen_text is not None:
                print(gen_text[:2], total_new_tokens[:4], flush=True)             # run_generation.py:329
            else:
                print(total_new_tokens[:4], flush=True)
            print("Iteration: %d, Time: %.6f sec" % (i, toc - tic), flush=True)
        if i >= num_warmup:
            total_time += toc - tic
            if args.token_latency:
                total_list.append(output[1])
    if rank != 0:
        return 0
    print("\n", "-" * 10, "Summary:", "-" * 10)
    latency = total_time / max(1, num_iter - num_warmup)
    print("Inference latency: %.3f sec." % latency)
    if args.token_latency and total_list:
        import numpy as np
        from itertools import chain
        first_latency = np.mean([x[0] for x in total_list])
        average_2n = sorted(chain(*[x[1:] for x in total_list]))
        print("First token average latency: %.3f sec." % first_latency)
        if average_2n:
            print("Average 2... latency: %.3f sec." % np.mean(average_2n))
            print("P90 2... latency: %.3f sec." % average_2n[int(len(average_2n) * 0.9)])
            print("P99 2... latency: %.3f sec." % average_2n[int(len(average_2n) * 0.99)])
    print("Throughput: %.1f tokens/sec (batch %d x %d new tokens)" % (args.batch_size * args.max_new_tokens / latency,
                                                                     args.batch_size, args.max_new_tokens))
    st = model.model.decoder.streamer
    if st is not None:
        s = st.stats()
        print("Streamed weights: %.2f GB at %.1f GB/s (pinned host -> HBM)" % (s["bytes"] / 1e9, s["gbps"]))
        if model.model.decoder.host_pool:
            print("NOTE: LIA_HOST_LAYER_POOL=%d -- streamed layers alias %d distinct pinned slabs (host RAM < model); "
                  "bytes per step are those of the full model, outputs are not" % (model.model.decoder.host_pool,
                                                                                  model.model.decoder.host_pool))
    for gs in model._states.values():
        if gs.spill is not None:
            k = gs.spill.stats()
            print("KV cache: %d of %d layers in HBM, %d spilled to %.2f GB of pinned host memory (%.2f GB in, %.2f GB out)"
                  % (gs.kv_resident, cfg.num_hidden_layers, k["layers"], k["host_bytes"] / 1e9, k["h2d_bytes"] / 1e9,
                     k["d2h_bytes"] / 1e9))
    return 0


if __name__ == "__main__":
    sys.exit(main())
