"""Host-side mirror of the upstream search / module surface, checked on CPU against the golden fixtures
(no kernel is involved in ratio-target search, sweep order, module packaging or cache naming)."""
import argparse, contextlib, io
import torch, torch.nn as nn, pytest
from conftest import build_tiny_opt
from oracle import asvd_oracle as O


def _args(**kw):
    base = dict(scaling_method="abs_mean", alpha=0.5, n_calib_samples=3, calib_dataset="synthetic", compress_kv_cache=False,
                rank_align=1, kv_cache_ratio_target=-1, param_ratio_target=0.8, ppl_target=-1, act_aware=True, sigma_fuse="UV")
    base.update(kw)
    return argparse.Namespace(**base)


def test_sweep_order_matches_upstream(golden_pipeline):
    from asvd4llm_b200.sensitivity import enumerate_linears
    model = build_tiny_opt(golden_pipeline)
    assert [t[2] for t in enumerate_linears(model)] == golden_pipeline["sweep_order"]


def test_ratio_target_search_matches_upstream_log_and_ranks(golden_pipeline):
    from asvd4llm_b200.binary_search import search_allocation
    from asvd4llm_b200 import _lib
    model = build_tiny_opt(golden_pipeline)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        chosen, default = search_allocation(model, golden_pipeline["sensitivity"], golden_pipeline["loader"], _args())
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("low=") or l.startswith("===")]
    assert lines == golden_pipeline["binary_search_log"]
    mods = dict(model.named_modules())
    ranks = {k: _lib.rank_for_ratio(*mods[k].weight.shape, r) for k, r in chosen.items() if r != default}
    assert ranks == golden_pipeline["truncation_ranks"]


def test_kv_cache_search_matches_upstream(golden_pipeline):
    from asvd4llm_b200.binary_search import search_allocation
    from asvd4llm_b200 import _lib
    model = build_tiny_opt(golden_pipeline)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        chosen, default = search_allocation(model, golden_pipeline["kv_sensitivity"], golden_pipeline["loader"],
                                            _args(compress_kv_cache=True, kv_cache_ratio_target=0.5, param_ratio_target=-1))
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("low=") or l.startswith("===")]
    assert lines == golden_pipeline["kv_log"]
    assert default == 2 and all("k_proj" in k or "v_proj" in k for k in chosen)
    mods = dict(model.named_modules())
    ranks = {k: min(_lib.rank_for_ratio(*mods[k].weight.shape, r), min(mods[k].weight.shape)) for k, r in chosen.items() if r != 2}
    assert ranks == golden_pipeline["kv_truncation_ranks"]


def test_module_constructor_matches_upstream_fusion(golden_cases):
    from asvd4llm_b200 import SVDLinear
    for c in golden_cases[:6]:
        ex = O.factorise_exact(c["W"], c["ratio"], sdm=c["sdm"], fisher=c["fisher"], alpha=c["alpha"],
                               act_aware=c["act_aware"], sigma_fuse=c["sigma_fuse"], rank_align=c["rank_align"])
        mod = SVDLinear(ex["U"], ex["S"], ex["V"], c["bias"], c["sigma_fuse"])
        assert list(mod.state_dict().keys()) == c["state_dict_keys"]
        assert mod.truncation_rank == c["truncation_rank"]
        assert torch.allclose(mod.ALinear.weight.data, ex["A"]) and torch.allclose(mod.BLinear.weight.data, ex["B"])
        if c["bias"] is not None:
            assert mod.ALinear.bias.data_ptr() == c["bias"].data_ptr()        # shared storage, as upstream


def test_cache_file_names_match_upstream(golden_pipeline):
    from asvd4llm_b200.sensitivity import sensitivity_cache_file
    from asvd4llm_b200.act_aware_utils import _cache_file
    model = build_tiny_opt(golden_pipeline)
    got = {_cache_file(model, "abs_mean").split("/")[-1], _cache_file(model, "abs_max").split("/")[-1],
           sensitivity_cache_file(model, _args()).split("/")[-1]}
    assert got == set(golden_pipeline["sensitivity_cache_files"])


def test_calibration_cache_files_are_read_without_a_gpu(golden_pipeline, tmp_path, monkeypatch):
    """--use_cache: upstream's published cache files (README.md:110-114) are accepted as they are; reading them needs
    no kernel.  Covers calib_input_distribution (act_aware_utils.py:49-60) and calib_fisher_info (:9-16)."""
    import os
    from conftest import GOLDEN
    from asvd4llm_b200.act_aware_utils import calib_input_distribution, calib_fisher_info
    fisher = torch.load(os.path.join(GOLDEN, "tiny_opt_fisher.pt"), weights_only=False)
    monkeypatch.chdir(tmp_path); os.makedirs("cache")
    torch.save(golden_pipeline["sdm_abs_mean"], "cache/synthetic_tiny-opt_calib_input_distribution_abs_mean.pt")
    torch.save(fisher["fisher_info"], "cache/" + fisher["cache_files"][0])
    model = build_tiny_opt(golden_pipeline)
    calib_input_distribution(model, golden_pipeline["loader"], "abs_mean", use_cache=True)
    calib_fisher_info(model, golden_pipeline["loader"], use_cache=True)
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Linear):
            assert torch.equal(mod.scaling_diag_matrix, golden_pipeline["sdm_abs_mean"][name])
            assert torch.equal(mod.fisher_info, fisher["fisher_info"][name])
    # without the cache the product needs its CUDA library and a device: it must fail loudly, not fall back
    model2 = build_tiny_opt(golden_pipeline)
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            calib_input_distribution(model2, golden_pipeline["loader"], "abs_max", use_cache=False)
