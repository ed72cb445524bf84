"""evaluate_perplexity — upstream evaluate_utils.py:90-115 (the scalar the sensitivity table stores).
The rest of upstream evaluate_utils.py (lm-eval adapter, dataset loaders) is out of scope."""
import torch
import torch.nn as nn


@torch.no_grad()
def evaluate_perplexity(model, dataset, limit, batch_size=1):
    """dataset: input ids [batch, seqlen]; exp(mean over samples of the mean CE over seqlen-1 shifted tokens).

    batch_size=1 is upstream's loop.  batch_size>1 runs the same samples through the model
    several at a time; every sample has the same length and no padding, so the per-sample losses — and the
    result — are the same up to floating-point reassociation inside the batched GEMMs."""
    nsamples, seqlen = dataset.size()
    device = getattr(model, "device", None) or next(model.parameters()).device
    n = nsamples if limit is None or limit < 0 or limit > nsamples else limit
    if n == 0 and nsamples > 0 and limit == 0:
        n = nsamples                                  # upstream's `if i == limit: break` never fires for limit <= 0 ...
    nlls = []
    for i0 in range(0, n, batch_size):
        i1 = min(n, i0 + batch_size)
        input_ids = dataset[i0:i1, :-1].to(device)
        labels = dataset[i0:i1, 1:].contiguous().to(device)
        logits = model(input_ids=input_ids)[0]
        loss = nn.functional.cross_entropy(logits.reshape(-1, logits.size(-1)), labels.reshape(-1), reduction="none")
        per_sample = loss.view(i1 - i0, -1).float().mean(dim=1)
        nlls.extend((per_sample * seqlen).unbind(0))   # quirk 10: x seqlen here, / seqlen below
    return torch.exp(torch.stack(nlls).sum() / (len(nlls) * seqlen)).item()
