"""ASVDOPTForCausalLM — OPTForCausalLM whose linears named in config.truncation_ranks are low-rank pairs
(BLinear [r, in], ALinear [out, r] + bias).  Same class names and state-dict keys as upstream
huggingface_repos/modeling_asvd_opt.py, so checkpoints are interchangeable.  Self-contained on purpose: a
repository consumer needs nothing but transformers (asvd4llm_b200.hf_export.load_asvd_model is the way to get the
sm_100a forward kernel under the same checkpoint)."""
import torch.nn as nn
from transformers import OPTForCausalLM

from .configuration_asvd_opt import ASVDOPTConfig


class ASVDLinear(nn.Module):
    def __init__(self, in_features, out_features, rank, bias=True):
        super().__init__()
        self.BLinear = nn.Linear(in_features, rank, bias=False)
        self.ALinear = nn.Linear(rank, out_features, bias=bias)

    def forward(self, input):
        return self.ALinear(self.BLinear(input))


class ASVDOPTForCausalLM(OPTForCausalLM):
    config_class = ASVDOPTConfig

    def __init__(self, config):
        super().__init__(config)
        self.truncation_ranks = config.truncation_ranks
        owners = {}
        for parent in self.modules():
            for child_name, child in parent.named_children():
                if isinstance(child, nn.Linear):
                    owners[child] = (parent, child_name)
        for name, module in list(self.named_modules()):
            if name in self.truncation_ranks and isinstance(module, nn.Linear):
                parent, child_name = owners[module]
                setattr(parent, child_name, ASVDLinear(module.in_features, module.out_features, self.truncation_ranks[name],
                                                       bias=module.bias is not None))
