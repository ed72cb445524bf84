import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box through gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_cases():
    import torch
    return torch.load(os.path.join(GOLDEN, "from_linear_cases.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_pipeline():
    import torch
    return torch.load(os.path.join(GOLDEN, "tiny_opt_pipeline.pt"), weights_only=False)


def build_tiny_opt(pipe):
    import torch
    from transformers import OPTConfig, OPTForCausalLM
    cfg = OPTConfig(**{k: v for k, v in pipe["config"].items() if k not in ("architectures", "model_type", "transformers_version")})
    model = OPTForCausalLM(cfg).float().eval()
    model.load_state_dict(pipe["state_dict"])
    model.config._name_or_path = "synthetic/tiny-opt"
    return model
