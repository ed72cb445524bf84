"""Pins oracle/asvd_oracle.py against outputs of the unmodified upstream code (tests/golden/*.pt)."""
import torch, torch.nn as nn, pytest
from oracle import asvd_oracle as O
from conftest import build_tiny_opt


def test_rank_formula_matches_survey_table():
    # SURVEY.md §8 a2 (values computed from modules/svd_linear.py:39-44)
    assert O.rank_for_ratio(768, 768, 0.9) == 345
    assert O.rank_for_ratio(3072, 768, 0.9) == 552
    assert O.rank_for_ratio(50272, 768, 0.9) == 680
    assert O.rank_for_ratio(4096, 4096, 0.9) == 1843
    assert O.rank_for_ratio(4096, 4096, 0.9, 128) == 1920
    assert O.rank_for_ratio(11008, 4096, 0.9) == 2686
    assert O.rank_for_ratio(32000, 4096, 0.9) == 3268
    assert O.rank_for_ratio(13824, 5120, 0.95) == 3549


def test_lowrank_restatement_reproduces_upstream_bitwise(golden_cases):
    """Same seed, same RNG consumption, same op sequence -> the restated svd_lowrank path reproduces the
    upstream factors (up to fp32 reassociation inside LAPACK: we allow 2e-4 abs on the products)."""
    for c in golden_cases:
        out = O.factorise_lowrank(c["W"], c["ratio"], sdm=c["sdm"], fisher=c["fisher"], alpha=c["alpha"],
                                  act_aware=c["act_aware"], sigma_fuse=c["sigma_fuse"], rank_align=c["rank_align"],
                                  seed=c["seed"])
        assert out["rank"] == c["truncation_rank"]
        assert tuple(out["A"].shape) == tuple(c["A"].shape) and tuple(out["B"].shape) == tuple(c["B"].shape)
        prod = out["A"].double() @ out["B"].double()
        ref = c["A"].double() @ c["B"].double()
        tol = 3e-3 if c["W"].dtype == torch.float16 else 2e-4       # upstream factors were rounded to fp16
        diff = prod - ref
        if c["act_aware"] and bool((c["sdm"] == 0).any()):
            # exact-zero channels are scaled by (0 + 1e-6)^alpha: after two power iterations their directions sit
            # below fp32 noise, so upstream's own output in those columns depends on the host's LAPACK kernels
            # (seen: identical on the box the fixtures were made on, O(|W|) apart on another CPU).  What upstream
            # determines there is the SCALED product, so compare in the metric the factorisation minimises.
            sv = O.scaling_vector(c["sdm"], c["fisher"], c["alpha"]).double()
            diff = diff * (sv / sv.max())
        assert diff.abs().max().item() < tol * max(1.0, ref.abs().max().item())


def test_exact_oracle_is_at_least_as_good_as_upstream(golden_cases):
    """Eckart-Young on the scaled matrix: the exact truncation can not reconstruct worse than svd_lowrank."""
    for c in golden_cases:
        ex = O.factorise_exact(c["W"], c["ratio"], sdm=c["sdm"], fisher=c["fisher"], alpha=c["alpha"],
                               act_aware=c["act_aware"], sigma_fuse=c["sigma_fuse"], rank_align=c["rank_align"],
                               compute_dtype=torch.float64)
        assert ex["rank"] == c["truncation_rank"]
        s = torch.ones(c["n"], dtype=torch.float64)
        if c["act_aware"]:
            s = O.scaling_vector(c["sdm"], c["fisher"], c["alpha"]).double()
        Wd = c["W"].double()
        err_exact = ((ex["A"] @ ex["B"] - Wd) * s).norm()
        err_ref = ((c["A"].double() @ c["B"].double() - Wd) * s).norm()
        slack = 2e-3 * (Wd * s).norm() if c["W"].dtype == torch.float16 else 1e-5 * (Wd * s).norm()
        assert err_exact <= err_ref + slack
        if O.rank_for_ratio(c["m"], c["n"], c["ratio"], c["rank_align"]) >= min(c["m"], c["n"]):   # full rank: both exact
            assert err_ref / (Wd * s).norm() < 1e-4 and err_exact / (Wd * s).norm() < 1e-10


def test_sigma_fuse_modes_agree():
    W, s = O.synthetic_weight(96, 64, seed=3)
    outs = [O.factorise_exact(W, 0.8, sdm=s, alpha=0.5, act_aware=True, sigma_fuse=f) for f in ("UV", "U", "V")]
    p0 = outs[0]["A"] @ outs[0]["B"]
    for o in outs[1:]:
        assert (o["A"] @ o["B"] - p0).abs().max() < 1e-5


def test_forward_matches_upstream(golden_cases):
    for c in golden_cases:
        y = O.lowrank_forward(c["x"], c["A"], c["B"], c["bias"])
        # same op sequence; the host BLAS may reassociate the dot products (bitwise on the box the fixtures came
        # from, last-ulp differences on another CPU): one ulp of the output dtype, relative to the largest entry
        ulp = 2e-3 if y.dtype == torch.float16 else 1e-6
        assert y.dtype == c["y"].dtype and y.shape == c["y"].shape
        assert (y.double() - c["y"].double()).abs().max().item() <= ulp * max(1.0, c["y"].double().abs().max().item())


def test_state_dict_keys(golden_cases):
    for c in golden_cases:
        mod = O.OracleSVDLinear(c["A"], c["B"], c["bias"], c["truncation_rank"])
        assert list(mod.state_dict().keys()) == c["state_dict_keys"]


def test_calibration_matches_upstream(golden_pipeline):
    for method in ("abs_max", "abs_mean"):
        model = build_tiny_opt(golden_pipeline)
        got = O.calib_input_distribution(model, golden_pipeline["loader"], method)
        want = golden_pipeline[f"sdm_{method}"]
        assert list(got.keys()) == list(want.keys())
        for k in want:
            assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
            assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-7), k     # host BLAS reassociation upstream of the hook


def test_perplexity_matches_upstream(golden_pipeline):
    model = build_tiny_opt(golden_pipeline)
    ids = torch.cat([b["input_ids"] for b in golden_pipeline["loader"]], 0)
    assert O.evaluate_perplexity(model, ids, 3) == pytest.approx(golden_pipeline["ppl_raw"], rel=1e-6)
    assert O.evaluate_perplexity(model, ids, 2) == pytest.approx(golden_pipeline["ppl_raw_limit2"], rel=1e-6)


def test_sweep_order_and_sensitivity_match_upstream(golden_pipeline):
    model = build_tiny_opt(golden_pipeline)
    for n, m in model.named_modules():
        if isinstance(m, nn.Linear):
            m.scaling_diag_matrix = golden_pipeline["sdm_abs_mean"][n].clone()
    assert [t[2] for t in O.enumerate_linears(model)] == golden_pipeline["sweep_order"]
    torch.manual_seed(golden_pipeline["sensitivity_seed"])
    sens = O.calib_sensitivity_ppl(model, golden_pipeline["loader"], alpha=0.5, n_calib_samples=3)
    want = golden_pipeline["sensitivity"]
    assert list(sens.keys()) == list(want.keys())
    for layer in want:
        assert list(sens[layer].keys()) == list(want[layer].keys())
        for ratio in want[layer]:
            assert sens[layer][ratio] == pytest.approx(want[layer][ratio], rel=2e-4), (layer, ratio)


def test_allocation_matches_upstream(golden_pipeline):
    model = build_tiny_opt(golden_pipeline)
    numel = {n: m.weight.numel() for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    chosen, mid = O.allocate_ratios(golden_pipeline["sensitivity"], numel, 0.8)
    got = {k: O.rank_for_ratio(*dict(model.named_modules())[k].weight.shape, r) for k, r in chosen.items() if r != 1}
    assert got == golden_pipeline["truncation_ranks"]
    chosen, mid = O.allocate_ratios(golden_pipeline["kv_sensitivity"], numel, 0.5, compress_kv_cache=True)
    mods = dict(model.named_modules())
    got = {k: min(O.rank_for_ratio(*mods[k].weight.shape, r), min(mods[k].weight.shape)) for k, r in chosen.items() if r != 2}
    assert got == golden_pipeline["kv_truncation_ranks"]


@pytest.fixture(scope="module")
def golden_fisher():
    import os
    from conftest import GOLDEN
    return torch.load(os.path.join(GOLDEN, "tiny_opt_fisher.pt"), weights_only=False)


def test_fisher_info_matches_upstream(golden_pipeline, golden_fisher):
    """act_aware_utils.py:8-44 (tests/golden/make_golden_fisher.py ran the upstream function on the same model)."""
    model = build_tiny_opt(golden_pipeline)
    got = O.calib_fisher_info(model, golden_pipeline["loader"])
    want = golden_fisher["fisher_info"]
    assert list(got.keys()) == list(want.keys()) == golden_fisher["cache_keys"]
    for k in want:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        assert torch.allclose(got[k], want[k], rtol=1e-4, atol=1e-9), k        # autograd + host BLAS reassociation


def test_fisher_and_abs_mean_scaling_matches_upstream(golden_pipeline, golden_fisher):
    """svd_linear.py:48-59 with both statistics present (scaling_method fisher_abs_mean), full rank: exact."""
    c = golden_fisher["from_linear"]
    model = build_tiny_opt(golden_pipeline)
    lin = dict(model.named_modules())[c["layer"]]
    out = O.factorise_exact(lin.weight.data, c["ratio"], sdm=golden_pipeline["sdm_abs_mean"][c["layer"]],
                            fisher=golden_fisher["fisher_info"][c["layer"]], alpha=c["alpha"], act_aware=True)
    assert out["rank"] == c["truncation_rank"]
    ref = c["A"].double() @ c["B"].double()
    assert (out["A"].double() @ out["B"].double() - ref).abs().max().item() < 1e-4 * ref.abs().max().item()


@pytest.fixture(scope="module")
def golden_ppl_target():
    import os
    from conftest import GOLDEN
    return torch.load(os.path.join(GOLDEN, "ppl_target_and_opt125m_shapes.pt"), weights_only=False)


def test_ppl_target_search_matches_upstream_log(golden_pipeline, golden_ppl_target):
    """binary_search.py:64-87 + final pass (tests/golden/make_golden_ppl_target.py ran the upstream function): the
    restatement reproduces every log line (same RNG consumption by svd_lowrank) and the final per-layer modules."""
    g = golden_ppl_target["ppl_target"]
    model = build_tiny_opt(golden_pipeline)
    for n, m in model.named_modules():
        if isinstance(m, nn.Linear):
            m.scaling_diag_matrix = golden_pipeline["sdm_abs_mean"][n].clone()
    torch.manual_seed(g["seed"])
    log = O.binary_search_truncation_rank(model, golden_pipeline["sensitivity"], golden_pipeline["loader"],
                                          ppl_target=g["target"], method="lowrank")
    assert len(log) == len(g["log"])
    for got, want in zip(log, g["log"]):
        if "ppl=" in got and not got.startswith("==="):
            head_g, ppl_g = got.split(", ppl=")[0], float(got.split("ppl=")[1].split(",")[0])
            head_w, ppl_w = want.split(", ppl=")[0], float(want.split("ppl=")[1].split(",")[0])
            assert head_g == head_w and got.split("param_ratio=")[1] == want.split("param_ratio=")[1]
            assert ppl_g == pytest.approx(ppl_w, rel=2e-4)
        else:
            assert got == want
    kinds = {n: (type(m).__name__.replace("Oracle", ""), getattr(m, "truncation_rank", None)) for n, m in model.named_modules()
             if n in g["kinds"]}
    assert kinds == g["kinds"]
    ids = torch.cat([b["input_ids"] for b in golden_pipeline["loader"]], 0)
    assert O.evaluate_perplexity(model, ids, 3) == pytest.approx(g["ppl_final"], rel=2e-3)


def test_opt125m_shape_cases_exact_truncation_beats_upstream(golden_ppl_target):
    """BASELINE config 1's real layer shapes: the exact-SVD oracle reconstructs at least as well as upstream's
    svd_lowrank factors did (Eckart-Young), at the same rank."""
    import importlib.util, os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLDEN, "make_golden_ppl_target.py"))
    src = open(os.path.join(GOLDEN, "make_golden_ppl_target.py")).read()
    ns = {}
    exec(src[src.index("def opt125m_case"):src.index("OPT125M_SHAPES")], {"torch": torch}, ns)
    for c in golden_ppl_target["opt125m_shapes"][:2]:
        W, sdm, x = ns["opt125m_case"](c["idx"], c["m"], c["n"])
        ex = O.factorise_exact(W, c["ratio"], sdm=sdm, alpha=0.5, act_aware=True)
        assert ex["rank"] == c["rank"]
        s = O.scaling_vector(sdm, None, 0.5).double()
        rec = (((ex["A"].double() @ ex["B"].double()) - W.double()) * s).norm() / (W.double() * s).norm()
        assert rec <= c["recon_scaled"] * (1 + 1e-3), (rec, c["recon_scaled"])
