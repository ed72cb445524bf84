"""evaluate_perplexity — upstream evaluate_utils.py:90-115 (the scalar the sensitivity table stores).
The rest of upstream evaluate_utils.py (lm-eval adapter, dataset loaders) is out of scope."""
import torch
import torch.nn as nn


@torch.no_grad()
def evaluate_perplexity(model, dataset, limit):
    """dataset: input ids [batch, seqlen]; batch-1 forwards; exp(mean CE over seqlen-1 shifted tokens)."""
    nsamples, seqlen = dataset.size()
    device = getattr(model, "device", None) or next(model.parameters()).device
    nlls = []
    for i in range(nsamples):
        if i == limit:
            break
        input_ids = dataset[i:i + 1, :-1].to(device)
        labels = dataset[i:i + 1, 1:].contiguous().to(device)
        logits = model(input_ids=input_ids)[0]
        loss = nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), labels.view(-1))
        nlls.append(loss.float() * seqlen)          # quirk 10: x seqlen here, / seqlen below
    return torch.exp(torch.stack(nlls).sum() / (len(nlls) * seqlen)).item()
