#!/usr/bin/env python
"""asvd.py — upstream's CLI (asvd.py:14-203: same flags, same defaults, same call order) on the B200 path.

Additions are namespaced and optional: `--calib_dataset synthetic` (random token ids; the box has no datasets),
`--synthetic_model {opt-125m,llama-2-7b,llama-2-13b}` (random-init architecture instead of a checkpoint) and
multi-GPU through torchrun (layers sharded per asvd4llm_b200.sharding).  Out of scope here, as in SURVEY.md §2:
quantization and the lm-eval harness (`--eval_*` are accepted and reported as skipped).
"""
import argparse
import os

import numpy as np
import torch

from asvd4llm_b200.act_aware_utils import calib_fisher_info, calib_input_distribution
from asvd4llm_b200.binary_search import binary_search_truncation_rank, search_allocation, LinearIndex
from asvd4llm_b200.evaluate_utils import evaluate_perplexity
from asvd4llm_b200.sensitivity import calib_sensitivity_ppl, calib_sensitivity_stable_rank
from asvd4llm_b200 import sharding

SYNTHETIC = {
    "opt-125m": dict(kind="opt", vocab_size=50272, hidden_size=768, ffn_dim=3072, num_hidden_layers=12, num_attention_heads=12,
                     max_position_embeddings=2048, word_embed_proj_dim=768),
    "llama-2-7b": dict(kind="llama", vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                       num_attention_heads=32, max_position_embeddings=4096),
    "llama-2-13b": dict(kind="llama", vocab_size=32000, hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                        num_attention_heads=40, max_position_embeddings=4096),
}


def build_model(args, device):
    if args.synthetic_model:
        cfg = dict(SYNTHETIC[args.synthetic_model])
        kind = cfg.pop("kind")
        torch.manual_seed(args.seed)
        if kind == "opt":
            from transformers import OPTConfig, OPTForCausalLM
            model = OPTForCausalLM(OPTConfig(**cfg))
        else:
            from transformers import LlamaConfig, LlamaForCausalLM
            model = LlamaForCausalLM(LlamaConfig(**cfg))
        model.config._name_or_path = "synthetic/" + args.synthetic_model
        return model.half().to(device).eval(), None
    from transformers import AutoModelForCausalLM, AutoTokenizer
    tok = AutoTokenizer.from_pretrained(args.model_id, trust_remote_code=True)
    model = AutoModelForCausalLM.from_pretrained(args.model_id, torch_dtype=torch.float16, trust_remote_code=True)
    return model.to(device).eval(), tok


def get_calib_data(args, tokenizer, vocab_size, seqlen=2048):
    """list of {"input_ids": [1, seqlen], "attention_mask": [1, seqlen]} — upstream datautils.py:146-160."""
    if args.calib_dataset == "synthetic" or tokenizer is None:
        g = torch.Generator().manual_seed(args.seed)
        return [dict(input_ids=torch.randint(0, vocab_size, (1, seqlen), generator=g),
                     attention_mask=torch.ones(1, seqlen, dtype=torch.long)) for _ in range(args.n_calib_samples)]
    raise SystemExit(f"--calib_dataset {args.calib_dataset} needs the upstream datautils.py samplers (HF datasets + network); "
                     "use --calib_dataset synthetic or provide the cache files upstream publishes")


def main(args):
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed_all(args.seed)
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        torch.distributed.init_process_group("nccl")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    rank = int(os.environ.get("RANK", 0))
    os.makedirs("cache", exist_ok=True)

    import time
    phase = {}

    def tick(name, t0):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        phase[name] = round(time.time() - t0, 2)

    model, tokenizer = build_model(args, device)
    if not args.raw_model:
        calib_loader = get_calib_data(args, tokenizer, model.config.vocab_size)
        t0 = time.time()
        if "fisher" in args.scaling_method:
            calib_fisher_info(model, calib_loader, args.use_cache)
        if "abs" in args.scaling_method:
            calib_input_distribution(model, calib_loader, args.scaling_method, args.use_cache)
        tick("calibration_s", t0)
        t0 = time.time()
        if args.sensitivity_metric == "stable_rank":
            # one sigma-only factorisation per linear (no model forwards): every rank computes the whole table
            sensitivity = calib_sensitivity_stable_rank(model, calib_loader, args, args.use_cache)
            if world == 1:
                binary_search_truncation_rank(model, sensitivity, calib_loader, args)
            else:
                index = LinearIndex(model)
                chosen, default = search_allocation(model, sensitivity, calib_loader, args, index=index)
                sharding.decompose_sharded(model, chosen, default, args, index=index)
        elif world == 1:
            sensitivity = calib_sensitivity_ppl(model, calib_loader, args, args.use_cache)
            binary_search_truncation_rank(model, sensitivity, calib_loader, args)
        else:
            # (layer, ratio) units in contiguous ranges: equal unit counts are equal work (the sweep is almost entirely
            # model forwards) and a layer's six ratios stay on one rank, which computes its SVD once -- only the
            # world - 1 layers that straddle a boundary are factorised twice
            from asvd4llm_b200.sensitivity import enumerate_linears, RATIOS, KV_RATIOS
            n_units = len(enumerate_linears(model)) * len(KV_RATIOS if args.compress_kv_cache else RATIOS)
            lo, hi = rank * n_units // world, (rank + 1) * n_units // world
            shard = calib_sensitivity_ppl(model, calib_loader, args, args.use_cache, unit_filter=lambda u: lo <= u < hi)
            sensitivity = sharding.gather_sensitivity(model, shard)
            index = LinearIndex(model)
            chosen, default = search_allocation(model, sensitivity, calib_loader, args, index=index)
            stats = sharding.decompose_sharded(model, chosen, default, args, index=index)
            if rank == 0:
                print(f"sharded final pass: decompose {stats['decompose_s']:.2f} s, wait for the slowest rank {stats['imbalance_wait_s']:.2f} s, "
                      f"factor exchange {stats['exchange_s']:.2f} s "
                      f"({stats['bytes'] / 1e9:.2f} GB received in {stats['collectives']} collectives)")
        tick("sensitivity_search_decompose_s", t0)
        if rank == 0:
            print(f"phase times on {world} GPU(s): {phase}")
        if args.weight_quant != "none":
            print("weight quantization is out of scope of the B200 path; skipped")
    if rank == 0:
        ids = torch.cat([b["input_ids"] for b in get_calib_data(args, tokenizer, model.config.vocab_size)], 0)
        result = {"calib_ppl": evaluate_perplexity(model, ids, args.n_calib_samples),
                  "eval": "skipped: lm-eval harness and datasets are out of scope / unavailable offline"}
        print(result)
        os.makedirs("output", exist_ok=True)
        with open("output/result.txt", "a+") as f:
            f.write(f"{args}\n{result}\n")


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--model_id", type=str, default="facebook/opt-1.3b", help="Pretrained model ID")
    parser.add_argument("--ppl_target", type=float, default=-1, help="target ppl")
    parser.add_argument("--param_ratio_target", type=float, default=-1, help="target param ratio")
    parser.add_argument("--act_aware", action="store_true", help="use act aware svd (ASVD)")
    parser.add_argument("--alpha", type=float, default=0.5, help="hyper-parameter alpha for ASVD")
    parser.add_argument("--n_calib_samples", type=int, default=32, help="number of samples used for calibration")
    parser.add_argument("--calib_dataset", type=str, default="wikitext2",
                        choices=["wikitext2", "c4", "ptb", "alpaca", "selfgen", "synthetic"], help="calibration dataset")
    parser.add_argument("--scaling_method", type=str, default="abs_mean",
                        choices=["abs_mean", "abs_max", "fisher", "fisher_abs_mean"], help="scaling method")
    parser.add_argument("--sensitivity_metric", type=str, default="ppl", choices=["ppl", "stable_rank"], help="search metric")
    parser.add_argument("--use_cache", action="store_true", help="use cached calibration results")
    parser.add_argument("--weight_quant", type=str, default="none",
                        choices=["none", "rtn_int8", "rtn_int6", "awq_int8", "awq_int4"], help="weight quantization method")
    parser.add_argument("--eval_mmlu", action="store_true", help="evaluate mmlu")
    parser.add_argument("--eval_ppl", default="wikitext2,ptb", type=str)
    parser.add_argument("--eval_tasks", type=str, default="")
    parser.add_argument("--sigma_fuse", type=str, default="UV", help="sigma fuse method", choices=["U", "V", "UV"])
    parser.add_argument("--seed", type=int, default=233, help="random seed, which can significantly affect the calibration results")
    parser.add_argument("--compress_kv_cache", action="store_true", help="compress kv cache by asvd for k_proj and v_proj")
    parser.add_argument("--kv_cache_ratio_target", type=float, default=-1, help="kv cache ratio")
    parser.add_argument("--rank_align", type=int, default=1, help="align rank in SVD")
    parser.add_argument("--raw_model", action="store_true", help="use the raw model without ASVD")
    parser.add_argument("--use_bos", action="store_true", help="use bos token in calibration")
    parser.add_argument("--eval_batch_size", type=int, default=8,
                        help="(extension) calibration samples per forward in the sensitivity sweep's perplexity evaluations (upstream: 1)")
    parser.add_argument("--synthetic_model", type=str, default="", choices=[""] + list(SYNTHETIC), help="(extension) random-init architecture")
    main(parser.parse_args())
