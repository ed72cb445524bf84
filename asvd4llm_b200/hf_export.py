"""HF checkpoint format of a decomposed model — upstream huggingface_repos/build_asvd_repo.py:58-92
and modeling_asvd_{llama,opt}.py.

`save_asvd_model` writes what upstream's builder writes: `save_pretrained` weights whose decomposed layers have the
keys `<name>.BLinear.weight [r, n]`, `<name>.ALinear.weight [m, r]`, `<name>.ALinear.bias [m]`, and a config.json
carrying `truncation_ranks`, `auto_map` and `architectures`.  Two self-contained modeling files are copied next to
the weights so `AutoModelForCausalLM.from_pretrained(path, trust_remote_code=True)` works anywhere; they are
interchangeable with upstream's (same class names, same keys), and upstream-built repositories load into
`load_asvd_model` unchanged.
"""
import json
import os
import shutil

import torch
import torch.nn as nn

from .modules.svd_linear import SVDLinear
from .sensitivity import enumerate_linears

_HERE = os.path.dirname(os.path.abspath(__file__))
_FAMILIES = {
    "opt": ("modeling_asvd_opt.py", "configuration_asvd_opt.py", "ASVDOPTForCausalLM", "ASVDOPTConfig"),
    "llama": ("modeling_asvd_llama.py", "configuration_asvd_llama.py", "ASVDLlamaForCausalLM", "ASVDLlamaConfig"),
}


def _family(model):
    mt = getattr(model.config, "model_type", "")
    for key in _FAMILIES:
        if key in mt:
            return key
    raise ValueError(f"no ASVD modeling file for model_type {mt!r} (upstream ships opt and llama)")


def truncation_ranks(model):
    """{module name: rank} of every SVDLinear — build_asvd_repo.py:65-69."""
    return {name: mod.truncation_rank for name, mod in model.named_modules() if isinstance(mod, SVDLinear)}


def save_asvd_model(model, save_path, tokenizer=None):
    fam = _family(model)
    modeling, configuration, arch, cfg_cls = _FAMILIES[fam]
    os.makedirs(save_path, exist_ok=True)
    if tokenizer is not None:
        tokenizer.save_pretrained(save_path)
    model.save_pretrained(save_path)
    config = model.config.to_dict()
    config["truncation_ranks"] = truncation_ranks(model)
    config["auto_map"] = {"AutoConfig": f"{configuration[:-3]}.{cfg_cls}", "AutoModelForCausalLM": f"{modeling[:-3]}.{arch}"}
    config["architectures"] = [arch]
    for f in (modeling, configuration):
        shutil.copy(os.path.join(_HERE, "hf_files", f), os.path.join(save_path, f))
    json.dump(config, open(os.path.join(save_path, "config.json"), "w"), indent=2)
    return config["truncation_ranks"]


def swap_in_asvd_linears(model, ranks):
    """What the ASVD*ForCausalLM constructors do (modeling_asvd_llama.py:18-41): replace every linear named in
    `ranks` by an (uninitialised) SVDLinear-shaped module so the checkpoint keys line up."""
    where = {full: (father, name, lin) for father, name, full, lin in enumerate_linears(model)}
    for full, r in ranks.items():
        father, name, lin = where[full]
        A = torch.empty(lin.out_features, r, dtype=lin.weight.dtype, device=lin.weight.device)
        B = torch.empty(r, lin.in_features, dtype=lin.weight.dtype, device=lin.weight.device)
        bias = None if lin.bias is None else torch.empty_like(lin.bias.data)
        setattr(father, name, SVDLinear._from_factors(A, B, bias))
    return model


def load_asvd_model(path, torch_dtype=None, device=None):
    """Loads an ASVD repository (ours or upstream's) with every decomposed layer as an asvd4llm_b200 SVDLinear, i.e.
    with the sm_100a forward kernel, without executing the repository's remote code."""
    from transformers import AutoConfig, AutoModelForCausalLM
    from safetensors.torch import load_file
    cfg = json.load(open(os.path.join(path, "config.json")))
    ranks = cfg.pop("truncation_ranks")
    cfg.pop("auto_map", None)
    base_arch = {"ASVDOPTForCausalLM": "OPTForCausalLM", "ASVDLlamaForCausalLM": "LlamaForCausalLM"}.get(cfg["architectures"][0], cfg["architectures"][0])
    import transformers
    config = AutoConfig.for_model(cfg["model_type"], **{k: v for k, v in cfg.items() if k not in ("model_type", "architectures")})
    model = getattr(transformers, base_arch)(config)
    if torch_dtype is not None:
        model = model.to(torch_dtype)
    swap_in_asvd_linears(model, ranks)
    state = {}
    for f in sorted(os.listdir(path)):
        if f.endswith(".safetensors"):
            state.update(load_file(os.path.join(path, f)))
        elif f.endswith(".bin") and f.startswith("pytorch_model"):
            state.update(torch.load(os.path.join(path, f), map_location="cpu"))
    missing, unexpected = model.load_state_dict(state, strict=False)
    missing = [k for k in missing if "lm_head" not in k]                    # tied heads are not stored twice
    assert not unexpected and not missing, (missing, unexpected)
    if hasattr(model, "tie_weights"):
        model.tie_weights()
    return model.to(device) if device is not None else model
