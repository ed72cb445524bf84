"""HF export round trip (N2), batched perplexity (N1), stable-rank table (N4, GPU)."""
import argparse, json, os
import pytest, torch, torch.nn as nn
from conftest import build_tiny_opt
from oracle import asvd_oracle as O


def _decompose_with_oracle(model, names, ratio=0.6):
    """CPU stand-in for the kernels: exact factors from the oracle, packaged by the product's module class."""
    from asvd4llm_b200 import SVDLinear
    from asvd4llm_b200.sensitivity import enumerate_linears
    for father, name, full, lin in enumerate_linears(model):
        if full in names:
            ex = O.factorise_exact(lin.weight.data, ratio, sdm=getattr(lin, "scaling_diag_matrix", None), alpha=0.5, act_aware=False)
            bias = lin.bias.data if lin.bias is not None else None
            setattr(father, name, SVDLinear._from_factors(ex["A"], ex["B"], bias))


def test_hf_export_round_trip(golden_pipeline, tmp_path):
    from asvd4llm_b200 import hf_export
    model = build_tiny_opt(golden_pipeline)
    names = ["model.decoder.layers.0.fc1", "model.decoder.layers.1.self_attn.v_proj", "model.decoder.layers.1.fc2"]
    _decompose_with_oracle(model, names)
    ranks = hf_export.save_asvd_model(model, str(tmp_path / "repo"))
    assert set(ranks) == set(names)
    cfg = json.load(open(tmp_path / "repo" / "config.json"))
    assert cfg["truncation_ranks"] == ranks and cfg["architectures"] == ["ASVDOPTForCausalLM"]
    assert cfg["auto_map"] == {"AutoConfig": "configuration_asvd_opt.ASVDOPTConfig",
                               "AutoModelForCausalLM": "modeling_asvd_opt.ASVDOPTForCausalLM"}      # build_asvd_repo.py:71-76
    # (a) the repository's own remote code (what a hub consumer runs)
    from transformers import AutoModelForCausalLM
    remote = AutoModelForCausalLM.from_pretrained(str(tmp_path / "repo"), trust_remote_code=True)
    want = model.state_dict()
    got = remote.state_dict()
    assert set(got.keys()) == set(want.keys())
    for k in want:
        assert torch.equal(got[k], want[k]), k
    ids = golden_pipeline["loader"][0]["input_ids"]
    with torch.no_grad():
        ref_logits = remote(input_ids=ids)[0]
    # (b) the product loader: same checkpoint, SVDLinear modules
    ours = hf_export.load_asvd_model(str(tmp_path / "repo"))
    from asvd4llm_b200 import SVDLinear
    assert {n for n, m in ours.named_modules() if isinstance(m, SVDLinear)} == set(names)
    for k, v in ours.state_dict().items():
        assert torch.equal(v, want[k]), k
    if torch.cuda.is_available():
        with torch.no_grad():
            logits = ours.cuda()(input_ids=ids.cuda())[0].cpu()
        assert (logits - ref_logits).abs().max().item() < 1e-3


def test_batched_perplexity_equals_upstream_loop(golden_pipeline):
    from asvd4llm_b200.evaluate_utils import evaluate_perplexity
    model = build_tiny_opt(golden_pipeline)
    ids = torch.cat([b["input_ids"] for b in golden_pipeline["loader"]], 0)
    p1 = evaluate_perplexity(model, ids, 3)
    assert p1 == pytest.approx(golden_pipeline["ppl_raw"], rel=1e-6)
    assert evaluate_perplexity(model, ids, 2) == pytest.approx(golden_pipeline["ppl_raw_limit2"], rel=1e-6)
    for bs in (2, 3, 8):
        assert evaluate_perplexity(model, ids, 3, batch_size=bs) == pytest.approx(p1, rel=1e-5)


@pytest.mark.gpu
def test_stable_rank_sensitivity_matches_upstream_formula(golden_pipeline, tmp_path, monkeypatch):
    from asvd4llm_b200.sensitivity import calib_sensitivity_stable_rank
    monkeypatch.chdir(tmp_path); os.makedirs("cache")
    model = build_tiny_opt(golden_pipeline).cuda()
    args = argparse.Namespace(scaling_method="abs_mean", alpha=0.5, n_calib_samples=3, calib_dataset="synthetic")
    table = calib_sensitivity_stable_rank(model, golden_pipeline["loader"], args, use_cache=False)
    assert os.path.exists("cache/synthetic_tiny-opt_sensitivity_stable_rank_abs_mean_0.5_3_synthetic.pt")
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Linear):
            w = mod.weight.data.float().cpu()
            sr = (torch.norm(w, p="fro") ** 2 / torch.linalg.svdvals(w).max() ** 2) ** 0.5           # sensitivity.py:98-104
            assert list(table[name].keys()) == [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
            for ratio, v in table[name].items():
                assert float(v) == pytest.approx(float(-sr * ratio ** 0.1), rel=1e-4)
