"""ASVDOPTConfig — the stock OPT configuration plus `truncation_ranks` ({module name: rank}); same field and
class name as upstream huggingface_repos/configuration_asvd_opt.py, written as a subclass instead of a copy."""
from transformers import OPTConfig


class ASVDOPTConfig(OPTConfig):
    def __init__(self, truncation_ranks=None, **kwargs):
        super().__init__(**kwargs)
        self.truncation_ranks = truncation_ranks if truncation_ranks is not None else {}
