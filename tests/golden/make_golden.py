"""Generates tests/golden/*.pt by running the UNMODIFIED upstream code (mounted read-only at /root/reference)
in the build container.  Run once:  python tests/golden/make_golden.py
The fixtures are committed; /root/reference does not exist on the GPU box and nothing else reads it.

Upstream import notes: evaluate_utils.py imports lm_eval at module top, which is not
installed; three stub modules in sys.modules are enough.  Upstream functions write ./cache/*.pt, so we run
them inside a temporary working directory that already holds cache/.
"""
import os, sys, types, tempfile, argparse
import torch, torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("lm_eval", "lm_eval.base", "lm_eval.evaluator", "lm_eval.tasks"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["lm_eval.base"].BaseLM = object
    sys.modules["lm_eval"].evaluator = sys.modules["lm_eval.evaluator"]
    sys.modules["lm_eval"].tasks = sys.modules["lm_eval.tasks"]
    sys.path.insert(0, REF)
    import modules.svd_linear as svd_linear, act_aware_utils, evaluate_utils, sensitivity, binary_search
    return svd_linear, act_aware_utils, evaluate_utils, sensitivity, binary_search


def tiny_opt(seed=233):
    from transformers import OPTConfig, OPTForCausalLM
    cfg = OPTConfig(vocab_size=128, hidden_size=64, ffn_dim=128, num_hidden_layers=2, num_attention_heads=4,
                    max_position_embeddings=64, word_embed_proj_dim=64)
    torch.manual_seed(seed)
    model = OPTForCausalLM(cfg).float().eval()
    model.config._name_or_path = "synthetic/tiny-opt"
    return cfg, model


def main():
    svd_linear, act_aware_utils, evaluate_utils, sensitivity, binary_search = import_reference()
    SVDLinear = svd_linear.SVDLinear

    # ---------------------------------------------------------------- from_linear cases
    cases = []
    shapes = [(96, 64, True), (64, 96, False), (128, 128, True), (80, 200, True), (200, 80, False)]
    idx = 0
    for (m, n, has_bias) in shapes:
        for act_aware, alpha, fuse, ratio, align, wdtype, sdtype in [
            (True, 0.5, "UV", 0.9, 1, torch.float16, torch.float16),
            (True, 0.5, "U", 0.6, 1, torch.float32, torch.float32),
            (True, 1.0, "V", 0.4, 8, torch.float16, torch.float32),
            (False, 1.0, "UV", 0.7, 1, torch.float32, torch.float32),
            (True, 0.5, "UV", 1.9, 1, torch.float32, torch.float32),       # requested rank > min(m,n): exact
        ]:
            g = torch.Generator().manual_seed(1000 + idx)
            lin = nn.Linear(n, m, bias=has_bias)
            lin.weight.data = (torch.randn(m, n, generator=g) * 0.05).to(wdtype)
            if has_bias:
                lin.bias.data = torch.randn(m, generator=g).to(wdtype)
            sdm = (torch.rand(n, generator=g) * 3 + 0.01).to(sdtype)
            if idx % 7 == 3:
                sdm[::9] = 0                                               # exact-zero channels hit the +1e-6
            lin.scaling_diag_matrix = sdm.clone()
            fisher = None
            if idx % 5 == 2:
                fisher = (torch.rand(n, generator=g) + 0.1).to(sdtype)
                lin.fisher_info = fisher.clone()
            torch.manual_seed(77 + idx)
            out = SVDLinear.from_linear(lin, ratio, act_aware=act_aware, alpha=alpha, sigma_fuse=fuse, rank_align=align)
            x = torch.randn(3, 5, n, generator=g).to(wdtype)
            with torch.no_grad():
                y = out(x)
            cases.append(dict(m=m, n=n, W=lin.weight.data.clone(), bias=None if not has_bias else lin.bias.data.clone(),
                              sdm=sdm, fisher=fisher, act_aware=act_aware, alpha=alpha, sigma_fuse=fuse, ratio=ratio,
                              rank_align=align, seed=77 + idx, A=out.ALinear.weight.data.clone(),
                              B=out.BLinear.weight.data.clone(), truncation_rank=out.truncation_rank,
                              state_dict_keys=list(out.state_dict().keys()), x=x,
                              y=y))
            idx += 1
    torch.save(cases, os.path.join(HERE, "from_linear_cases.pt"))
    print("from_linear cases:", len(cases))

    # ---------------------------------------------------------------- tiny-model pipeline
    cfg, model = tiny_opt()
    g = torch.Generator().manual_seed(233)
    loader = [dict(input_ids=torch.randint(0, 128, (1, 48), generator=g), attention_mask=torch.ones(1, 48, dtype=torch.long))
              for _ in range(3)]
    state = {k: v.clone() for k, v in model.state_dict().items()}
    pipe = dict(config=cfg.to_dict(), state_dict=state, loader=loader)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp); os.makedirs("cache")
        try:
            for method in ("abs_max", "abs_mean"):
                act_aware_utils.calib_input_distribution(model, loader, method, use_cache=False)
                pipe[f"sdm_{method}"] = {n: m.scaling_diag_matrix.clone() for n, m in model.named_modules()
                                         if isinstance(m, nn.Linear)}
                pipe[f"sdm_{method}_cache_file"] = sorted(os.listdir("cache"))
            ids = torch.cat([b["input_ids"] for b in loader], 0)
            pipe["ppl_raw"] = evaluate_utils.evaluate_perplexity(model, ids, 3)
            pipe["ppl_raw_limit2"] = evaluate_utils.evaluate_perplexity(model, ids, 2)
            args = argparse.Namespace(scaling_method="abs_mean", alpha=0.5, n_calib_samples=3, calib_dataset="synthetic",
                                      compress_kv_cache=False, rank_align=1, kv_cache_ratio_target=-1,
                                      param_ratio_target=0.8, ppl_target=-1, act_aware=True, sigma_fuse="UV")
            torch.manual_seed(4242)
            sens = sensitivity.calib_sensitivity_ppl(model, loader, args, use_cache=False)
            pipe["sensitivity_seed"] = 4242
            pipe["sensitivity"] = sens
            pipe["sensitivity_cache_files"] = sorted(os.listdir("cache"))
            pipe["sweep_order"] = list(sens.keys())
            torch.manual_seed(99)
            import io, contextlib
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                binary_search.binary_search_truncation_rank(model, sens, loader, args)
            pipe["binary_search_seed"] = 99
            pipe["binary_search_log"] = [l for l in buf.getvalue().splitlines() if l.startswith("low=") or l.startswith("===")]
            pipe["truncation_ranks"] = {n: m.truncation_rank for n, m in model.named_modules() if isinstance(m, SVDLinear)}
            pipe["ppl_decomposed"] = evaluate_utils.evaluate_perplexity(model, ids, 3)
            pipe["decomposed_state_dict_keys"] = list(model.state_dict().keys())
            # kv-cache mode allocation on a synthetic monotone table (no factorisation needed for the log)
            cfg2, model2 = tiny_opt()
            for n_, m_ in model2.named_modules():
                if isinstance(m_, nn.Linear):
                    m_.scaling_diag_matrix = pipe["sdm_abs_mean"][n_].clone()
            kv_sens = {}
            for li, (n_, m_) in enumerate([(n_, m_) for n_, m_ in model2.named_modules() if isinstance(m_, nn.Linear)]):
                kv_sens[n_] = {0.1 * i: 50.0 + (li % 5) * 0.37 + 9.0 / i for i in range(1, 20)}
            args_kv = argparse.Namespace(**{**vars(args), "compress_kv_cache": True, "kv_cache_ratio_target": 0.5,
                                             "param_ratio_target": -1})
            buf = io.StringIO()
            torch.manual_seed(5)
            with contextlib.redirect_stdout(buf):
                binary_search.binary_search_truncation_rank(model2, kv_sens, loader, args_kv)
            pipe["kv_sensitivity"] = kv_sens
            pipe["kv_log"] = [l for l in buf.getvalue().splitlines() if l.startswith("low=") or l.startswith("===")]
            pipe["kv_truncation_ranks"] = {n: m.truncation_rank for n, m in model2.named_modules() if isinstance(m, SVDLinear)}
        finally:
            os.chdir(cwd)
    torch.save(pipe, os.path.join(HERE, "tiny_opt_pipeline.pt"))
    print("pipeline golden written; ranks:", pipe["truncation_ranks"])
    print("kv ranks:", pipe["kv_truncation_ranks"])
    print(pipe["binary_search_log"][-3:])


if __name__ == "__main__":
    main()
