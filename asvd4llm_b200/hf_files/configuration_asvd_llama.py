"""ASVDLlamaConfig — the stock Llama configuration plus `truncation_ranks` ({module name: rank}); same field and
class name as upstream huggingface_repos/configuration_asvd_llama.py, written as a subclass instead of a copy."""
from transformers import LlamaConfig


class ASVDLlamaConfig(LlamaConfig):
    def __init__(self, truncation_ranks=None, **kwargs):
        super().__init__(**kwargs)
        self.truncation_ranks = truncation_ranks if truncation_ranks is not None else {}
