"""Parity of the CUDA path (through the C ABI) against the oracle.  Tolerances are north_star's:
singular values within 1e-4 relative (all kept r), reconstructed W within 1e-3 Frobenius, forward within
1e-3 max-abs in fp16 under the |y|<1 normalisation."""
import argparse, contextlib, io, os
import pytest, torch, torch.nn as nn
from conftest import build_tiny_opt
from oracle import asvd_oracle as O

pytestmark = pytest.mark.gpu
SIGMA_RTOL = 1e-4
RECON_TOL = 1e-3


def _lib():
    from asvd4llm_b200 import _lib
    return _lib


def _scaled_norm(X, s):
    return (X * s).norm()


def check_factorisation(W, scale32, ratio, fuse="UV", out_dtype=torch.float32, rank_align=1, fact=None, b=0):
    L = _lib()
    m, n = W.shape
    if fact is None:
        fact = L.scaled_svd([W.cuda()], [None if scale32 is None else scale32.cuda()])
        assert fact.status == 0
    r = min(O.rank_for_ratio(m, n, ratio, rank_align), min(m, n))
    sd = torch.ones(n, dtype=torch.float64) if scale32 is None else scale32.double()
    Wd = W.double()
    U, S, Vh = torch.linalg.svd(Wd * sd, full_matrices=False)
    sig = fact.sigma(b).cpu().double()
    assert sig.shape[0] == min(m, n)
    assert torch.all(sig[:-1] >= sig[1:]), "sigma not sorted"
    # 1e-4 relative on every kept sigma, above the fp32 noise floor of the oracle itself (LAPACK fp32
    # has absolute error ~eps*sigma_max, so singular values below 1e-2*sigma_max get an absolute allowance)
    rel = ((sig[:r] - S[:r]).abs() / (S[:r] + 1e-2 * S[0])).max().item() * (1 + 1e-2)
    rel = max(rel, ((sig[:r] - S[:r]).abs() / S[:r])[S[:r] > 1e-2 * S[0]].max().item())
    assert rel < SIGMA_RTOL, f"sigma rel err {rel}"
    A, B = fact.extract(r, fuse, out_dtype, b)
    assert A.shape == (m, r) and B.shape == (r, n) and A.dtype == out_dtype
    A, B = A.cpu().double(), B.cpu().double()
    Wtr = (U[:, :r] * S[:r]) @ Vh[:r] / sd
    rec = _scaled_norm(A @ B - Wtr, sd) / _scaled_norm(Wd, sd)
    tol = RECON_TOL if out_dtype == torch.float32 else 2 * RECON_TOL
    assert rec < tol, f"reconstruction vs exact truncation {rec}"
    # balance of the fusion: column norms of A are sigma^a
    a_exp = {"UV": 0.5, "U": 1.0, "V": 0.0}[fuse]
    if out_dtype == torch.float32:
        an = A.norm(dim=0)
        assert torch.allclose(an, S[:r] ** a_exp, rtol=2e-4, atol=1e-7)
    return fact, rel, rec


def test_golden_cases_exact_parity(golden_cases):
    """Every upstream-generated case: our factors vs the exact-SVD oracle and vs the upstream result."""
    L = _lib()
    for c in golden_cases:
        s = None
        if c["act_aware"]:
            s = L.scaling_vector(c["sdm"].cuda(), None if c["fisher"] is None else c["fisher"].cuda(), c["alpha"], c["n"], "cuda").cpu()
            want = O.scaling_vector(c["sdm"], c["fisher"], c["alpha"]).float()
            assert torch.allclose(s, want, rtol=2e-3 if c["sdm"].dtype == torch.float16 else 1e-6, atol=0)
        fact, _, _ = check_factorisation(c["W"], s, c["ratio"], c["sigma_fuse"], torch.float32, c["rank_align"])
        # ours (exact) must reconstruct the scaled weight at least as well as upstream's svd_lowrank result
        r = c["truncation_rank"]
        A, B = fact.extract(r, c["sigma_fuse"], torch.float32)
        sd = torch.ones(c["n"], dtype=torch.float64) if s is None else s.double()
        Wd = c["W"].double()
        ours = _scaled_norm(A.cpu().double() @ B.cpu().double() - Wd, sd)
        ref = _scaled_norm(c["A"].double() @ c["B"].double() - Wd, sd)
        assert ours <= ref * (1 + 1e-4) + 2e-3 * _scaled_norm(Wd, sd) * (c["W"].dtype == torch.float16)


def test_from_linear_module_surface(golden_cases):
    from asvd4llm_b200 import SVDLinear
    for c in golden_cases[::3]:
        lin = nn.Linear(c["n"], c["m"], bias=c["bias"] is not None)
        lin.weight.data = c["W"].clone()
        if c["bias"] is not None:
            lin.bias.data = c["bias"].clone()
        lin = lin.cuda()
        lin.scaling_diag_matrix = c["sdm"].cuda()
        if c["fisher"] is not None:
            lin.fisher_info = c["fisher"].cuda()
        w_before = lin.weight.data.clone()
        mod = SVDLinear.from_linear(lin, c["ratio"], act_aware=c["act_aware"], alpha=c["alpha"], sigma_fuse=c["sigma_fuse"],
                                    rank_align=c["rank_align"])
        assert isinstance(mod, SVDLinear)
        assert list(mod.state_dict().keys()) == c["state_dict_keys"]
        assert mod.truncation_rank == c["truncation_rank"]
        assert mod.ALinear.weight.dtype == c["W"].dtype and mod.ALinear.weight.shape == c["A"].shape
        assert mod.BLinear.weight.shape == c["B"].shape
        assert torch.equal(lin.weight.data, w_before), "from_linear must not mutate the layer"
        if c["bias"] is not None:
            assert mod.ALinear.bias.data_ptr() == lin.bias.data_ptr()
        x = c["x"].cuda()
        y = mod(x)
        yref = O.lowrank_forward(c["x"], mod.ALinear.weight.data.cpu(), mod.BLinear.weight.data.cpu(),
                                 None if c["bias"] is None else c["bias"], compute_dtype=torch.float64)
        tol = 2e-3 * max(1.0, yref.abs().max().item()) if c["W"].dtype == torch.float16 else 1e-4
        assert (y.cpu().double() - yref).abs().max().item() < tol


@pytest.mark.parametrize("m,n,kind", [(1024, 1024, "gauss"), (1024, 1024, "power"), (1536, 512, "gauss"), (512, 1536, "power"),
                                      (300, 700, "gauss"), (129, 257, "gauss")])
def test_synthetic_sigma_and_reconstruction(m, n, kind):
    W, s = O.synthetic_weight(m, n, seed=233, kind=kind)
    scale = (s ** 0.5 + 1e-6).float()
    for fuse in ("UV", "U", "V"):
        check_factorisation(W, scale, 0.9, fuse)


def test_batched_equals_single_bitwise():
    """Sharding invariant: a weight's factors do not depend on what else is in the launch."""
    L = _lib()
    Ws, Ss = [], []
    for b in range(3):
        W, s = O.synthetic_weight(512, 384, seed=10 + b)
        Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
    batched = L.scaled_svd(Ws, Ss)
    for b in range(3):
        single = L.scaled_svd([Ws[b]], [Ss[b]])
        assert torch.equal(single.sigma(0), batched.sigma(b))
        A1, B1 = single.extract(200, "UV", torch.float16, 0)
        A2, B2 = batched.extract(200, "UV", torch.float16, b)
        assert torch.equal(A1, A2) and torch.equal(B1, B2)


def test_edge_cases():
    L = _lib()
    # rank-deficient weight and exact-zero scale channels
    g = torch.Generator().manual_seed(5)
    W = (torch.randn(200, 40, generator=g) @ torch.randn(40, 160, generator=g) * 0.01).half()
    sdm = torch.rand(160, generator=g).half(); sdm[::5] = 0
    s = L.scaling_vector(sdm.cuda(), None, 0.5, 160, "cuda")
    assert torch.allclose(s.cpu(), O.scaling_vector(sdm, None, 0.5).float(), rtol=2e-3)
    fact = L.scaled_svd([W.cuda()], [s])
    A, B = fact.extract(100, "UV", torch.float32)
    assert torch.isfinite(A).all() and torch.isfinite(B).all()
    sd = s.cpu().double()
    rec = _scaled_norm(A.cpu().double() @ B.cpu().double() - W.double(), sd) / _scaled_norm(W.double(), sd)
    assert rec < 2e-3            # rank 40 < 100: exact up to the fp16 rounding of W's low-rank structure
    # all-zero weight
    fact = L.scaled_svd([torch.zeros(64, 96, dtype=torch.half, device="cuda")], [None])
    A, B = fact.extract(30, "UV", torch.float16)
    assert (A == 0).all() and (B == 0).all() and (fact.sigma() == 0).all()
    # NaN -> status 4 -> module-level fallback like upstream (fresh nn.Linear, "nan in S")
    from asvd4llm_b200 import SVDLinear
    lin = nn.Linear(64, 48).half().cuda()
    lin.weight.data[3, 5] = float("nan")
    out = SVDLinear.from_linear(lin, 0.9)
    assert isinstance(out, nn.Linear) and out.weight.shape == lin.weight.shape and out.weight.dtype == torch.float16
    # ... and PER LAYER, as upstream's loop would: the healthy layers of the same kernel batch are decomposed normally,
    # bitwise as if the bad layer had not been there (an Inf in the statistics counts as a bad layer too)
    from asvd4llm_b200.modules.svd_linear import from_linear_batch
    gb = torch.Generator().manual_seed(8)
    lins = []
    for i in range(4):
        l = nn.Linear(96, 80, bias=False)
        l.weight.data = (torch.randn(80, 96, generator=gb) * 0.05).half()
        l = l.cuda()
        l.scaling_diag_matrix = (torch.rand(96, generator=gb) + 0.1).half().cuda()
        lins.append(l)
    lins[1].weight.data[7, 7] = float("nan")
    lins[3].scaling_diag_matrix[5] = float("inf")
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        mods = from_linear_batch(lins, [0.8] * 4, act_aware=True, alpha=0.5)
    assert buf.getvalue().count("nan in S") == 2
    assert [isinstance(m, SVDLinear) for m in mods] == [True, False, True, False]
    for i in (0, 2):
        alone = SVDLinear.from_linear(lins[i], 0.8, act_aware=True, alpha=0.5)
        assert torch.equal(alone.ALinear.weight.data, mods[i].ALinear.weight.data)
        assert torch.equal(alone.BLinear.weight.data, mods[i].BLinear.weight.data)
    # act_aware without statistics raises like upstream
    lin2 = nn.Linear(64, 48).cuda()
    with pytest.raises(AttributeError):
        SVDLinear.from_linear(lin2, 0.9, act_aware=True)
    # requested rank above min(m, n) yields min(m, n) and an exact reconstruction (quirk 6)
    lin3 = nn.Linear(96, 32).cuda()
    mod = SVDLinear.from_linear(lin3, 1.9)
    assert mod.truncation_rank == 32
    rec = (mod.ALinear.weight.data @ mod.BLinear.weight.data - lin3.weight.data).norm() / lin3.weight.data.norm()
    assert rec < 1e-5


def test_sensitivity_cache_reuses_one_svd():
    from asvd4llm_b200 import SVDLinear, _lib as L
    W, s = O.synthetic_weight(256, 256, seed=3)
    lin = nn.Linear(256, 256, bias=False)
    lin.weight.data = W
    lin = lin.cuda(); lin.scaling_diag_matrix = s.half().cuda()
    before = L.profile_read()["gram"][1]
    mods = [SVDLinear.from_linear(lin, r, act_aware=True, alpha=0.5) for r in O.RATIO_CANDIDATES]
    mid = L.profile_read()["gram"][1]
    assert mid > before
    SVDLinear.from_linear(lin, 0.5, act_aware=True, alpha=0.5)
    assert L.profile_read()["gram"][1] == mid, "second visit of the same layer must re-slice the cached SVD"
    assert [m.truncation_rank for m in mods] == [O.rank_for_ratio(256, 256, r) for r in O.RATIO_CANDIDATES]
    lin.weight.data.mul_(2.0)          # in-place change bumps the version -> new SVD
    SVDLinear.from_linear(lin, 0.5, act_aware=True, alpha=0.5)
    assert L.profile_read()["gram"][1] > mid


@pytest.mark.parametrize("r", [256, 512, 1024, 1843, 345])
def test_forward_parity_fp16(r):
    """config 4 shapes at reduced M: 1e-3 max-abs with |y| < 1, and error vs fp64 no worse than the reference's."""
    L = _lib()
    g = torch.Generator().manual_seed(233)
    n = m = 4096
    x = (torch.randn(2, 512, n, generator=g) * 0.125).half()
    B = (torch.randn(r, n, generator=g) / n ** 0.5).half()
    A = (torch.randn(m, r, generator=g) / r ** 0.5 * 0.5).half()
    bias = (torch.randn(m, generator=g) * 0.1).half()
    for bi in (None, bias):
        y = L.lowrank_forward(x.cuda(), A.cuda(), B.cuda(), None if bi is None else bi.cuda()).cpu()
        y64 = O.lowrank_forward(x, A, B, bi, compute_dtype=torch.float64)
        assert y64.abs().max() < 1.0
        yref16 = O.lowrank_forward(x.float(), A.float(), B.float(), None if bi is None else bi.float())   # fp32 math of the same module
        err = (y.double() - y64).abs().max().item()
        assert err < 1e-3, err
        assert (y.float() - yref16).abs().max().item() < 1e-3


@pytest.mark.parametrize("r,with_bias", [(256, True), (1024, False), (1843, True)])
def test_forward_full_size_config4(r, with_bias):
    """BASELINE config 4 at FULL size (B=32, L=2048, d=4096 -> 65 536 tokens): the C-ABI forward against an fp64
    reference on a strided row sample (every 128-row tile position and both CTAs of a pair are hit), 1e-3 max-abs with
    |y| < 1.  The unsampled rows are covered by a checksum against the module-dtype torch product."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(7)
    n = m = 4096
    x = (torch.randn(32, 2048, n, device="cuda", generator=g) * 0.125).half()
    B = (torch.randn(r, n, device="cuda", generator=g) / n ** 0.5).half()
    A = (torch.randn(m, r, device="cuda", generator=g) / r ** 0.5 * 0.5).half()
    bias = (torch.randn(m, device="cuda", generator=g) * 0.1).half() if with_bias else None
    before = L.profile_read()["forward"][1]
    y = L.lowrank_forward(x, A, B, bias)
    assert L.profile_read()["forward"][1] > before
    assert y.shape == (32, 2048, m)
    y2 = y.reshape(-1, m)
    rows = torch.arange(0, 65536, 97, device="cuda")
    rows = torch.cat([rows, torch.tensor([127, 128, 255, 256, 65535 - 128, 65535], device="cuda")])
    xs = x.reshape(-1, n)[rows].double()
    t = (xs @ B.double().t()).half().double()                   # BLinear output is materialised in fp16 upstream too
    y64 = t @ A.double().t() + (0 if bias is None else bias.double())
    assert y64.abs().max() < 1.0
    err = (y2[rows].double() - y64).abs().max().item()
    assert err < 1e-3, err
    # every row: column sums of y against the same sums of a torch fp16 product of the same factors (fp32 accumulate)
    ref = torch.nn.functional.linear(torch.nn.functional.linear(x.reshape(-1, n), B), A, bias)
    d = (y2.float() - ref.float()).abs().max().item()
    assert d < 2e-3, d


def test_forward_rejects_mixed_dtypes_and_supports_autograd():
    L = _lib()
    from asvd4llm_b200 import SVDLinear
    g = torch.Generator().manual_seed(3)
    n, r, m = 64, 13, 48
    B = (torch.randn(r, n, generator=g) / n ** 0.5).half().cuda()
    A = (torch.randn(m, r, generator=g) / r ** 0.5).half().cuda()
    with pytest.raises(RuntimeError, match="same dtype"):
        L.lowrank_forward(torch.randn(4, n, device="cuda"), A, B, None)            # fp32 activations, fp16 module
    with pytest.raises(RuntimeError, match="same device|CUDA"):
        L.lowrank_forward(torch.randn(4, n).half().cuda(), A.cpu(), B, None)
    # differentiable like upstream's two nn.Linear children
    mod = SVDLinear._from_factors(A.float().clone(), B.float().clone(), torch.zeros(m, device="cuda"))
    x = torch.randn(5, n, device="cuda", requires_grad=True)
    y = mod(x)
    y.square().sum().backward()
    ref_A, ref_B = A.float().clone().requires_grad_(), B.float().clone().requires_grad_()
    x2 = x.detach().clone().requires_grad_()
    (torch.nn.functional.linear(torch.nn.functional.linear(x2, ref_B), ref_A)).square().sum().backward()
    assert torch.allclose(x.grad, x2.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(mod.ALinear.weight.grad, ref_A.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(mod.BLinear.weight.grad, ref_B.grad, rtol=1e-3, atol=1e-4)
    # a rank that is not a multiple of 8 keeps a padded kernel copy, refreshed when the weight is replaced
    h = SVDLinear._from_factors(A, B, None)
    xh = torch.randn(3, n, device="cuda").half()
    y1 = h(xh)
    assert h._kernel_A().stride(0) == 64
    h.ALinear.weight.data = (A * 2).contiguous()
    y2 = h(xh)
    assert torch.allclose(y2.float(), 2 * y1.float(), rtol=2e-3, atol=2e-3)


def test_forward_fused_kernel_matches_pair_kernels(monkeypatch):
    """ASVD_B200_FWD=fused (one kernel, intermediate in shared memory, ranks <= 256) against the default two-GEMM path:
    both round the intermediate to the module dtype, so the results agree to fp16 output rounding; ragged M / m / r and
    the split tail tiles (M = 19 000: 75 tiles on 74 clusters) included."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(11)
    for (M, n, r, m) in [(300, 512, 100, 384), (19000, 1024, 256, 1024), (2048, 4096, 256, 4096), (6221, 768, 153, 3072)]:
        x = (torch.randn(M, n, device="cuda", generator=g) * 0.125).half()
        B = (torch.randn(r, n, device="cuda", generator=g) / n ** 0.5).half()
        A = (torch.randn(m, r, device="cuda", generator=g) / r ** 0.5 * 0.5).half()
        bias = (torch.randn(m, device="cuda", generator=g) * 0.1).half()
        monkeypatch.delenv("ASVD_B200_FWD", raising=False)
        y0 = L.lowrank_forward(x, A, B, bias)
        monkeypatch.setenv("ASVD_B200_FWD", "fused")
        y1 = L.lowrank_forward(x, A, B, bias)
        monkeypatch.delenv("ASVD_B200_FWD", raising=False)
        assert (y0.float() - y1.float()).abs().max().item() < 1e-3, (M, n, r, m)


def test_forward_bf16_and_ragged_shapes():
    L = _lib()
    g = torch.Generator().manual_seed(1)
    for (M, n, r, m) in [(1, 64, 8, 32), (37, 100, 13, 77), (130, 257, 129, 131)]:
        x = torch.randn(M, n, generator=g).bfloat16(); B = (torch.randn(r, n, generator=g) / n ** 0.5).bfloat16()
        A = (torch.randn(m, r, generator=g) / r ** 0.5).bfloat16()
        y = L.lowrank_forward(x.cuda(), A.cuda(), B.cuda(), None).cpu()
        t = (x.double() @ B.double().t()).bfloat16().double()
        y64 = t @ A.double().t()
        assert (y.double() - y64).abs().max().item() < 3e-2 * max(1.0, y64.abs().max().item())
    y = L.lowrank_forward(torch.empty(0, 64, dtype=torch.half, device="cuda"), torch.randn(32, 8).half().cuda(),
                          torch.randn(8, 64).half().cuda(), None)
    assert y.shape == (0, 32)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("method", ["abs_mean", "abs_max"])
def test_absstat_matches_hook_math(dtype, method):
    L = _lib()
    g = torch.Generator().manual_seed(9)
    for (Lr, n) in [(2048, 768), (5, 3), (333, 1001)]:
        acc = torch.zeros(n, dtype=dtype, device="cuda")
        want = 0
        for k in range(3):
            x = (torch.randn(1, Lr, n, generator=g) * (k + 1)).to(dtype).cuda()
            L.absstat_accum(x, acc, method)
            want = O.abs_stat_update(want, x, method)
        if method == "abs_max":
            assert torch.equal(acc, want)
        else:
            ulp = {torch.float16: 1e-3, torch.bfloat16: 8e-3, torch.float32: 2e-6}[dtype]
            assert torch.allclose(acc.float(), want.float(), rtol=2 * ulp, atol=0)


def _args(**kw):
    base = dict(scaling_method="abs_mean", alpha=0.5, n_calib_samples=3, calib_dataset="synthetic", compress_kv_cache=False,
                rank_align=1, kv_cache_ratio_target=-1, param_ratio_target=0.8, ppl_target=-1, act_aware=True, sigma_fuse="UV")
    base.update(kw)
    return argparse.Namespace(**base)


def test_tiny_opt_pipeline_against_upstream_golden(golden_pipeline, tmp_path, monkeypatch):
    """calib -> sensitivity -> search -> decomposition on the tiny OPT, every stage against the upstream run."""
    from asvd4llm_b200 import SVDLinear
    from asvd4llm_b200.act_aware_utils import calib_input_distribution
    from asvd4llm_b200.sensitivity import calib_sensitivity_ppl
    from asvd4llm_b200.binary_search import binary_search_truncation_rank
    from asvd4llm_b200.evaluate_utils import evaluate_perplexity
    monkeypatch.chdir(tmp_path)
    os.makedirs("cache")
    loader = golden_pipeline["loader"]
    ids = torch.cat([b["input_ids"] for b in loader], 0)
    for method in ("abs_max", "abs_mean"):
        model = build_tiny_opt(golden_pipeline).cuda()
        calib_input_distribution(model, loader, method, use_cache=False)
        want = golden_pipeline[f"sdm_{method}"]
        for name, mod in model.named_modules():
            if isinstance(mod, nn.Linear):
                assert torch.allclose(mod.scaling_diag_matrix.cpu(), want[name], rtol=1e-4, atol=1e-6), (method, name)
        saved = torch.load(f"cache/synthetic_tiny-opt_calib_input_distribution_{method}.pt")
        assert list(saved.keys()) == list(want.keys())
    assert evaluate_perplexity(model, ids, 3) == pytest.approx(golden_pipeline["ppl_raw"], rel=1e-3)
    with contextlib.redirect_stdout(io.StringIO()):
        sens = calib_sensitivity_ppl(model, loader, _args(), use_cache=False)
    want = golden_pipeline["sensitivity"]
    assert list(sens.keys()) == list(want.keys())
    for layer in want:
        assert list(sens[layer].keys()) == list(want[layer].keys())
        for ratio in want[layer]:
            # exact truncation vs svd_lowrank: same table up to the truncation-quality difference
            assert sens[layer][ratio] == pytest.approx(want[layer][ratio], rel=2e-2), (layer, ratio)
    assert os.path.exists("cache/synthetic_tiny-opt_sensitivity_abs_mean_0.5_3_synthetic.pt")
    # the search on the SAME table must make the same decisions
    with contextlib.redirect_stdout(io.StringIO()):
        binary_search_truncation_rank(model, want, loader, _args())
    ranks = {n: m.truncation_rank for n, m in model.named_modules() if isinstance(m, SVDLinear)}
    assert ranks == golden_pipeline["truncation_ranks"]
    assert list(model.state_dict().keys()) == golden_pipeline["decomposed_state_dict_keys"]
    ppl = evaluate_perplexity(model, ids, 3)
    assert ppl <= golden_pipeline["ppl_decomposed"] * 1.01, (ppl, golden_pipeline["ppl_decomposed"])


def test_full_size_4096_properties():
    """configs[1] at full size: sigma vs LAPACK on the host, reconstruction vs the projector property."""
    L = _lib()
    W, s = O.synthetic_weight(4096, 4096, seed=233)
    scale = (s ** 0.5 + 1e-6).float()
    fact = L.scaled_svd([W.cuda()], [scale.cuda()])
    assert fact.status == 0
    r = O.rank_for_ratio(4096, 4096, 0.9)
    assert r == 1843
    ref = torch.linalg.svdvals(W.float() * scale)          # fp32 LAPACK: the north_star oracle
    sig = fact.sigma().cpu()
    assert ((sig[:r] - ref[:r]).abs() / ref[:r]).max().item() < SIGMA_RTOL
    A, B = fact.extract(r, "UV", torch.float32)
    # A B = P W with P an orthogonal projector of rank r: idempotence and energy identities, size independent
    Wc = W.float().cuda()
    AB = A @ B
    An = A / A.norm(dim=0, keepdim=True)
    gram_err = (An.t() @ An - torch.eye(r, device="cuda")).abs().max().item()
    assert gram_err < 5e-5
    assert ((An @ (An.t() @ Wc)) - AB).norm().item() / AB.norm().item() < 1e-4
    sc = scale.cuda()
    kept = ((AB * sc).norm() ** 2).item()
    assert kept == pytest.approx((ref[:r].double() ** 2).sum().item(), rel=1e-4)


def test_tensorcore_path_matches_simt_reference(monkeypatch):
    """The tcgen05 Gram / update kernels (TF32 3-term split) against the fp32 SIMT kernels kept in the library for
    this purpose (ASVD_B200_SIMT=1): same singular values, same truncated reconstruction."""
    L = _lib()
    W, s = O.synthetic_weight(768, 640, seed=21)
    scale = (s ** 0.5 + 1e-6).float().cuda()
    tc = L.scaled_svd([W.cuda()], [scale])
    monkeypatch.setenv("ASVD_B200_SIMT", "1")
    simt = L.scaled_svd([W.cuda()], [scale])
    monkeypatch.delenv("ASVD_B200_SIMT")
    assert tc.status == 0 and simt.status == 0
    s1, s2 = tc.sigma().double(), simt.sigma().double()
    assert ((s1 - s2).abs() / (s2 + 1e-3 * s2[0])).max().item() < 2e-5
    A1, B1 = tc.extract(300, "UV", torch.float32)
    A2, B2 = simt.extract(300, "UV", torch.float32)
    rel = ((A1 @ B1 - A2 @ B2) * scale).norm().item() / ((A2 @ B2) * scale).norm().item()
    assert rel < 5e-4, rel


def test_rectangular_llama_shapes_reduced():
    """Shapes with the aspect ratios of gate/up (tall), down (wide) and lm_head (very tall), scaled down 8x."""
    for (m, n) in [(1376, 512), (512, 1376), (4000, 512)]:
        W, s = O.synthetic_weight(m, n, seed=5)
        check_factorisation(W, (s ** 0.5 + 1e-6).float(), 0.9, "UV")


def test_inner_orderings_and_tails_agree(monkeypatch):
    """The inner eigen-solvers (triangular with a parameter warp -- the default --, quad round-robin, odd-even) and the
    two tails (unit column norms / Newton-Schulz) are different routes to the same factorisation: singular values and the
    truncated product agree."""
    L = _lib()
    W, s = O.synthetic_weight(1024, 1024, seed=11)
    scale = (s ** 0.5 + 1e-6).float().cuda()
    ref = torch.linalg.svdvals(W.double().cuda() * scale.double())
    results = {}
    for solve, polish in (("quad", "norm"), ("oddeven", "norm"), ("quad", "NS"), ("oddeven", "NS"), ("tri", "norm")):
        monkeypatch.setenv("ASVD_B200_SOLVE", solve)
        monkeypatch.setenv("ASVD_B200_POLISH", polish)
        f = L.scaled_svd([W.cuda()], [scale])
        assert f.status == 0
        sig = f.sigma().double()
        assert ((sig[:460] - ref[:460]).abs() / ref[:460]).max().item() < 2e-5, (solve, polish)
        A, B = f.extract(460, "UV", torch.float32)
        results[(solve, polish)] = (A @ B) * scale
    monkeypatch.delenv("ASVD_B200_SOLVE")
    monkeypatch.delenv("ASVD_B200_POLISH")
    base = results[("quad", "norm")]
    for k, v in results.items():
        assert (v - base).norm().item() / base.norm().item() < 5e-4, k


def test_llama13b_block_counts_reduced():
    """5120-wide layers give 80 blocks of 64 vectors (not a power of two) and 40 pairs per round; same structure at
    1/8 scale: 640 vectors = 10 blocks, plus the 2.7:1 rectangles of the 13B MLP."""
    for (m, n) in [(640, 640), (1728, 640), (640, 1728)]:
        W, s = O.synthetic_weight(m, n, seed=9)
        check_factorisation(W, (s ** 0.5 + 1e-6).float(), 0.95, "UV")


def test_suggest_batch_fills_whole_waves(monkeypatch):
    L = _lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for waves in (1, 2, 4):
        monkeypatch.setenv("ASVD_B200_WAVES", str(waves))
        for (m, n) in [(4096, 4096), (11008, 4096), (768, 768), (5120, 5120)]:
            b = L.suggest_batch(m, n)
            pairs = (min(m, n) + 127) // 128
            assert 1 <= b <= 32 and b * pairs <= max(waves * L.SOLVE_CTAS_PER_SM * sms, pairs)
    monkeypatch.delenv("ASVD_B200_WAVES")
    assert L.suggest_batch(4096, 4096) == min(32, 2 * L.SOLVE_CTAS_PER_SM * sms // 32)     # the batch bench.py times
    assert L.balanced_batches(128, 18) == [16] * 8 and L.balanced_batches(5, 18) == [5] and L.balanced_batches(19, 18) == [10, 9]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gram_preconditioned_rectangles(dtype, monkeypatch):
    """Shapes that take the Gram pre-conditioner (>= 2:1, >= 1024 vectors, 16-bit weights), both orientations: same
    contract as every other shape, and the same singular values as the direct path (ASVD_B200_GRAMPRE=0)."""
    L = _lib()
    for (m, n) in [(2304, 1024), (1024, 2304)]:
        W, s = O.synthetic_weight(m, n, seed=31)
        W = W.to(dtype)
        scale = (s ** 0.5 + 1e-6).float()
        fact, rel, rec = check_factorisation(W, scale, 0.9, "UV")
        monkeypatch.setenv("ASVD_B200_GRAMPRE", "0")
        direct = L.scaled_svd([W.cuda()], [scale.cuda()])
        monkeypatch.delenv("ASVD_B200_GRAMPRE")
        s1, s2 = fact.sigma().double(), direct.sigma().double()
        assert ((s1 - s2).abs() / (s2 + 1e-3 * s2[0])).max().item() < 2e-5


def test_recovery_tensor_core_matches_simt(monkeypatch):
    """The bf16-plane recovery GEMM against the fp32 SIMT GEMM it replaces: sigma and factors."""
    L = _lib()
    W, s = O.synthetic_weight(1280, 1024, seed=41)
    scale = (s ** 0.5 + 1e-6).float().cuda()
    a = L.scaled_svd([W.cuda()], [scale])
    monkeypatch.setenv("ASVD_B200_RECOVER", "simt")
    b = L.scaled_svd([W.cuda()], [scale])
    monkeypatch.delenv("ASVD_B200_RECOVER")
    s1, s2 = a.sigma().double(), b.sigma().double()
    assert ((s1 - s2).abs() / (s2 + 1e-3 * s2[0])).max().item() < 1e-5
    A1, B1 = a.extract(460, "UV", torch.float32)
    A2, B2 = b.extract(460, "UV", torch.float32)
    assert ((A1 @ B1 - A2 @ B2) * scale).norm().item() / ((A2 @ B2) * scale).norm().item() < 1e-4


@pytest.mark.parametrize("m,n,batch", [(1024, 1024, 4), (768, 1280, 3), (2560, 1024, 2)])
def test_overlapped_half_batches_bitwise(m, n, batch, monkeypatch):
    """ASVD_B200_OVERLAP=1 runs the two halves of a batch on two streams (the solve of one half beside the streaming
    passes of the other).  Same kernels on offset pointers: the factors must be bitwise those of the one-stream
    schedule, for even and odd batches, square, wide and Gram-pre-conditioned tall shapes."""
    L = _lib()
    Ws, Ss = [], []
    for b in range(batch):
        W, s = O.synthetic_weight(m, n, seed=40 + b)
        Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
    monkeypatch.setenv("ASVD_B200_OVERLAP", "0")
    ref = L.scaled_svd(Ws, Ss)
    monkeypatch.setenv("ASVD_B200_OVERLAP", "1")
    for _ in range(2):                      # twice: the second run reuses the side streams and events
        got = L.scaled_svd(Ws, Ss)
        r = min(m, n) // 2
        for b in range(batch):
            assert torch.equal(ref.sigma(b), got.sigma(b))
            A1, B1 = ref.extract(r, "UV", torch.float16, b)
            A2, B2 = got.extract(r, "UV", torch.float16, b)
            assert torch.equal(A1, A2) and torch.equal(B1, B2)
    assert ref.sweeps == got.sweeps


def test_fisher_info_against_upstream_golden(golden_pipeline, tmp_path, monkeypatch):
    """calib_fisher_info (upstream act_aware_utils.py:8-44): the weight-gradient reduction runs in
    asvd_absstat_accum(SQ_MEAN); values, cache file name and keys are upstream's; a second call reads the cache."""
    import os
    from conftest import GOLDEN
    from asvd4llm_b200.act_aware_utils import calib_fisher_info
    gold = torch.load(os.path.join(GOLDEN, "tiny_opt_fisher.pt"), weights_only=False)
    monkeypatch.chdir(tmp_path); os.makedirs("cache")
    model = build_tiny_opt(golden_pipeline).cuda()
    n0 = _lib().launch_count()
    calib_fisher_info(model, golden_pipeline["loader"], use_cache=False)
    assert _lib().launch_count() > n0
    assert sorted(os.listdir("cache")) == gold["cache_files"]
    table = torch.load(os.path.join("cache", gold["cache_files"][0]), map_location="cpu")
    assert list(table.keys()) == gold["cache_keys"]
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Linear):
            want = gold["fisher_info"][name]
            assert mod.fisher_info.dtype == want.dtype and mod.fisher_info.is_cuda
            assert torch.allclose(mod.fisher_info.cpu(), want, rtol=2e-3, atol=1e-9), name    # GPU backward vs CPU backward
            assert torch.equal(table[name], mod.fisher_info.cpu())
    model2 = build_tiny_opt(golden_pipeline).cuda()
    n1 = _lib().launch_count()
    calib_fisher_info(model2, golden_pipeline["loader"], use_cache=True)
    assert _lib().launch_count() == n1
    assert torch.equal(dict(model2.named_modules())["lm_head"].fisher_info.cpu(), table["lm_head"])
    # the factorisation honours both statistics (scaling_method fisher_abs_mean)
    c = gold["from_linear"]
    lin = dict(model.named_modules())[c["layer"]]
    lin.scaling_diag_matrix = golden_pipeline["sdm_abs_mean"][c["layer"]].cuda()
    from asvd4llm_b200 import SVDLinear
    mod = SVDLinear.from_linear(lin, c["ratio"], act_aware=True, alpha=c["alpha"])
    assert mod.truncation_rank == c["truncation_rank"]
    ref = c["A"].double() @ c["B"].double()
    got = mod.ALinear.weight.data.double().cpu() @ mod.BLinear.weight.data.double().cpu()
    assert (got - ref).abs().max().item() < 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_sq_mean_matches_torch_expression(dtype):
    """SQ_MEAN restates `acc += g.pow(2).mean(0)` with upstream's roundings (square in the gradient dtype, fp32 sum,
    mean and accumulator rounded to the dtype), on a ragged shape with tiny values that flush in fp16."""
    L = _lib()
    g = torch.Generator().manual_seed(11)
    G = (torch.randn(333, 1001, generator=g) * torch.logspace(-5, 0, 1001)).to(dtype)
    acc0 = torch.rand(1001, generator=g).to(dtype) * 1e-2
    want = O.fisher_stat_update(acc0.clone(), G)
    acc = acc0.clone().cuda()
    L.absstat_accum(G.cuda(), acc, "sq_mean")
    tol = {torch.float16: 1e-3, torch.bfloat16: 8e-3, torch.float32: 2e-6}[dtype]
    assert torch.allclose(acc.cpu().float(), want.float(), rtol=tol, atol=1e-12)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_layers_on_two_devices_in_one_process():
    """Upstream loads models with device_map="auto": one process, linears spread over several GPUs.  Function
    attributes, the quad schedule table and the side streams are per-device state; factors and the forward must be
    the same on every device (cuda:1 first, so that nothing was initialised by an earlier test on that device)."""
    from asvd4llm_b200 import SVDLinear
    outs = []
    for dev in ("cuda:1", "cuda:0"):
        torch.manual_seed(0)
        W, s = O.synthetic_weight(640, 512, seed=77)
        lin = nn.Linear(512, 640, bias=True).half()
        lin.weight.data = W
        lin.scaling_diag_matrix = s.half()
        lin = lin.to(dev)
        lin.scaling_diag_matrix = lin.scaling_diag_matrix.to(dev)
        mod = SVDLinear.from_linear(lin, 0.8, act_aware=True, alpha=0.5)
        assert mod.ALinear.weight.device == torch.device(dev)
        x = (torch.randn(3, 40, 512, generator=torch.Generator().manual_seed(1)) * 0.125).half().to(dev)
        y = mod(x)
        acc = torch.zeros(512, dtype=torch.float16, device=dev)
        _lib().absstat_accum(x, acc, "abs_mean")
        torch.cuda.synchronize(dev)
        outs.append((mod.ALinear.weight.data.cpu(), mod.BLinear.weight.data.cpu(), y.cpu(), acc.cpu()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("m,n", [(1, 1), (1, 5), (5, 1), (2, 3), (7, 13), (13, 7), (65, 129), (129, 65), (127, 128), (1, 300), (300, 1)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_tiny_and_odd_shapes(m, n, dtype):
    """Shapes far below one 64-vector block, odd, and degenerate (a single row / column), scaled: sigma against
    torch.linalg.svd in fp64 and an exact reconstruction at full rank; a truncated extraction stays finite."""
    L = _lib()
    g = torch.Generator().manual_seed(1000 * m + n)
    W = (torch.randn(m, n, generator=g) * 0.05).to(dtype)
    s = (torch.rand(n, generator=g) + 0.25).float()
    fact = L.scaled_svd([W.cuda()], [s.cuda()])
    assert fact.status == 0
    k = min(m, n)
    S = torch.linalg.svdvals(W.double() * s.double())
    sig = fact.sigma(0).cpu().double()
    assert sig.shape == (k,) and torch.all(sig[:-1] >= sig[1:])
    assert ((sig - S).abs() / (S + 1e-2 * S[0])).max().item() < SIGMA_RTOL
    for fuse in ("UV", "U", "V"):
        A, B = fact.extract(k, fuse, torch.float32)
        rec = (A.cpu().double() @ B.cpu().double() - W.double()).norm() / W.double().norm()
        assert rec < 1e-5, (fuse, rec)
    if k > 1:
        A, B = fact.extract(k - 1, "UV", dtype)
        assert A.shape == (m, k - 1) and B.shape == (k - 1, n) and torch.isfinite(A).all() and torch.isfinite(B).all()
        err = (A.double().cpu() @ B.double().cpu() - W.double()) * s.double()
        # Eckart-Young: dropping the smallest singular value leaves exactly sigma_k (plus factor rounding in fp16)
        assert abs(err.norm().item() - S[-1].item()) < 2e-3 * S[0].item()


def test_rank_zero_mirrors_upstream():
    """A ratio so small that the rank formula (svd_linear.py:39-44) gives 0: upstream returns an SVDLinear with empty
    factors whose forward is the bias alone (probe on the upstream code: type SVDLinear, truncation_rank 0)."""
    from asvd4llm_b200 import SVDLinear
    lin = nn.Linear(64, 48).half().cuda()
    mod = SVDLinear.from_linear(lin, 0.001)
    assert isinstance(mod, SVDLinear) and mod.truncation_rank == 0
    assert mod.ALinear.weight.shape == (48, 0) and mod.BLinear.weight.shape == (0, 64)
    assert list(mod.state_dict().keys()) == ["ALinear.weight", "ALinear.bias", "BLinear.weight"]
    x = torch.randn(2, 5, 64, device="cuda").half()
    y = mod(x)
    assert y.shape == (2, 5, 48) and torch.equal(y, lin.bias.data.expand(2, 5, 48))
    lin2 = nn.Linear(64, 48, bias=False).half().cuda()
    y2 = SVDLinear.from_linear(lin2, 0.001)(x)
    assert y2.shape == (2, 5, 48) and (y2 == 0).all()


@pytest.mark.parametrize("m,n,batch", [(1024, 1024, 2), (768, 1280, 1), (2048, 2048, 4), (2304, 1024, 2)])
def test_lean_solve_matches_quad_bitwise(m, n, batch, monkeypatch):
    """ASVD_B200_SOLVE=lean splits the inner sweep into a G-only kernel (two CTAs per SM) and a replay kernel that
    rebuilds R from the streamed rotation history.  Same operations in the same order: bitwise the quad kernel's result."""
    L = _lib()
    Ws, Ss = [], []
    for b in range(batch):
        W, s = O.synthetic_weight(m, n, seed=60 + b)
        Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
    monkeypatch.setenv("ASVD_B200_SOLVE", "quad")
    ref = L.scaled_svd(Ws, Ss)
    # "leanr": the same G-only kernel followed by the DEFAULT solve's replay kernel (solve_tri_r_kernel: shuffles instead
    # of a staging area, bulk-copied record) -- every element sees the same operations in the same order, so the default
    # path's replay is tied bit for bit to the monolithic quad kernel
    for mode in ("lean", "leanr"):
        monkeypatch.setenv("ASVD_B200_SOLVE", mode)
        got = L.scaled_svd(Ws, Ss)
        assert ref.sweeps == got.sweeps, mode
        for b in range(batch):
            assert torch.equal(ref.sigma(b), got.sigma(b)), mode
            A1, B1 = ref.extract(min(m, n) // 2, "UV", torch.float16, b)
            A2, B2 = got.extract(min(m, n) // 2, "UV", torch.float16, b)
            assert torch.equal(A1, A2) and torch.equal(B1, B2), mode


@pytest.mark.parametrize("m,n,batch", [(1024, 1024, 2), (768, 1280, 1), (2048, 2048, 4), (2304, 1024, 2), (200, 136, 3)])
def test_tri_solve_default(m, n, batch, monkeypatch):
    """The default inner solve (solve_tri_g_kernel: upper triangle of the Gram matrix, rotation parameters on a warp of
    their own; solve_tri_r_kernel: barrier-free replay on R) against the quad kernel it replaces and against fp64:
      * the default IS the triangular solve (same bits as ASVD_B200_SOLVE=tri), run-to-run bitwise reproducible;
      * a weight's factors do not depend on its batch-mates (bitwise, where the Gram chunking coincides);
      * singular values within 2e-5 of fp64 (and of the quad kernel's) on the kept half, truncated product on the
        Eckart-Young floor like the quad kernel's."""
    L = _lib()
    Ws, Ss = [], []
    for b in range(batch):
        W, s = O.synthetic_weight(m, n, seed=80 + b)
        Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
    r = min(m, n) // 2
    monkeypatch.delenv("ASVD_B200_SOLVE", raising=False)
    dflt = L.scaled_svd(Ws, Ss)
    monkeypatch.setenv("ASVD_B200_SOLVE", "tri")
    tri = L.scaled_svd(Ws, Ss)
    alone = L.scaled_svd(Ws[-1:], Ss[-1:])
    monkeypatch.setenv("ASVD_B200_SOLVE", "quad")
    quad = L.scaled_svd(Ws, Ss)
    monkeypatch.delenv("ASVD_B200_SOLVE")
    assert dflt.status == 0 and tri.status == 0 and dflt.sweeps == tri.sweeps
    for b in range(batch):
        assert torch.equal(dflt.sigma(b), tri.sigma(b))
        A0, B0 = dflt.extract(r, "UV", torch.float16, b)
        A1, B1 = tri.extract(r, "UV", torch.float16, b)
        assert torch.equal(A0, A1) and torch.equal(B0, B1)
        ref = torch.linalg.svdvals(Ws[b].double() * Ss[b].double())
        sig = tri.sigma(b).double()
        assert ((sig[:r] - ref[:r]).abs() / ref[:r]).max().item() < 2e-5
        # both kernels' truncated products sit on the Eckart-Young floor (the products themselves may differ by more:
        # the subspace at the cut of a clustered spectrum is ill-conditioned, the error of the best approximation is not)
        floor = (ref[r:] ** 2).sum().sqrt().item()
        for fact in (quad, tri):
            Af, Bf = fact.extract(r, "UV", torch.float32, b)
            rec = ((Af.double() @ Bf.double() - Ws[b].double()) * Ss[b].double()).norm().item()
            assert rec <= floor * (1 + 1e-4) + 1e-6 * ref[0].item(), (rec, floor)
        assert ((quad.sigma(b).double()[:r] - sig[:r]).abs() / ref[:r]).max().item() < 2e-5
    # batch independence is bitwise whenever both batch sizes cut the long dimension into the same Gram chunks
    # (make_plan: just enough chunks to give every SM an item, at least 512 columns each)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    pairs, max_chunks = (min(m, n) + 127) // 128, (max(m, n) + 127) // 128 * 128 // 512 + (1 if (max(m, n) + 127) // 128 * 128 % 512 else 0)
    if sms // (batch * pairs) >= max_chunks:
        assert torch.equal(alone.sigma(0), tri.sigma(batch - 1))
        Aa, Ba = alone.extract(r, "UV", torch.float16, 0)
        Ab, Bb = tri.extract(r, "UV", torch.float16, batch - 1)
        assert torch.equal(Aa, Ab) and torch.equal(Ba, Bb)


def test_pinned_gram_chunks_make_every_batch_size_bitwise_equal(monkeypatch):
    """The Gram pass cuts the long dimension into just enough chunks to give every SM an item, so the chunking -- hence
    the summation order of a Gram entry -- depends on how many weights share the launch: 2048^2 alone runs four chunks,
    in a batch of four two.  ASVD_B200_GRAM_CHUNKS pins the count; with it a weight's factors are bitwise the same alone
    and in a batch (without it they agree to rounding: checked here too)."""
    L = _lib()
    m = n = 2048
    Ws, Ss = [], []
    for b in range(4):
        W, s = O.synthetic_weight(m, n, seed=90 + b)
        Ws.append(W.cuda()); Ss.append((s ** 0.5 + 1e-6).float().cuda())
    loose_b, loose_a = L.scaled_svd(Ws, Ss), L.scaled_svd(Ws[2:3], Ss[2:3])
    assert ((loose_b.sigma(2)[:900] - loose_a.sigma(0)[:900]).abs() / loose_a.sigma(0)[:900]).max().item() < 2e-5
    monkeypatch.setenv("ASVD_B200_GRAM_CHUNKS", "2")
    batch, alone = L.scaled_svd(Ws, Ss), L.scaled_svd(Ws[2:3], Ss[2:3])
    monkeypatch.delenv("ASVD_B200_GRAM_CHUNKS")
    assert torch.equal(batch.sigma(2), alone.sigma(0))
    A1, B1 = batch.extract(900, "UV", torch.float16, 2)
    A2, B2 = alone.extract(900, "UV", torch.float16, 0)
    assert torch.equal(A1, A2) and torch.equal(B1, B2)


# ------------------------------------------------------------------------------------------------ full-size shapes (§8d configs 3/5)
FULL_SIZE_SHAPES = [(11008, 4096, 0.9), (4096, 11008, 0.9), (32000, 4096, 0.9), (13824, 5120, 0.95), (50272, 768, 0.9)]


@pytest.mark.parametrize("m,n,ratio", FULL_SIZE_SHAPES)
def test_full_size_rectangles_vs_fp64(m, n, ratio):
    """Every rectangular shape of Llama-2-7B / 13B / OPT-125m at FULL size, through the default path (Gram
    pre-conditioner where it applies), against an fp64 reference computed on the device:
      * sigma: 1e-4 relative on every kept value (strict, no allowance) vs fp64 singular values;
      * reconstruction: the scaled error of the fp32 factors sits on the Eckart-Young floor sqrt(sum_{j>r} sigma_j^2);
      * the factor product is an orthogonal projection of W (size-independent property)."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(1000 + m % 97)
    W = (torch.randn(m, n, device="cuda", generator=g) * 0.02).half()
    sdm = torch.exp(torch.randn(n, device="cuda", generator=g)).half()
    scale = L.scaling_vector(sdm, None, 0.5, n, "cuda")
    fact = L.scaled_svd([W], [scale])
    assert fact.status == 0
    r = min(L.rank_for_ratio(m, n, ratio, 1), min(m, n))
    Xd = W.double() * scale.double()
    ref = torch.linalg.svdvals(Xd)                               # fp64 on the device
    sig = fact.sigma(0).double()
    rel = ((sig[:r] - ref[:r]).abs() / ref[:r]).max().item()
    assert rel < SIGMA_RTOL, f"{m}x{n}: kept-sigma rel err {rel}"
    A, B = fact.extract(r, "UV", torch.float32, 0)
    rec = ((A.double() @ B.double() - W.double()) * scale.double()).norm().item()
    floor = (ref[r:] ** 2).sum().sqrt().item()
    assert rec <= floor * (1 + 1e-4) + 1e-6 * ref[0].item(), (rec, floor)
    assert rec / Xd.norm().item() < 1.0


def test_ill_conditioned_sigma_strict_and_with_allowance():
    """Outlier-heavy / power-law inputs (activation-scaled LLM weights are like this).  Reported separately:
    (a) the STRICT 1e-4 relative error on kept sigma above 1e-3 sigma_max -- must hold;
    (b) below that, fp32 arithmetic (ours and LAPACK's) resolves sigma to ~eps32 * sigma_max ABSOLUTE, so
        the bound is 1e-4 * (sigma + 1e-2 sigma_max)/... as in check_factorisation; the measured strict relative error
        there (up to ~9e-4 at kappa_kept 2e5, profiles/r01_illcond_check.jsonl) is a documented deviation."""
    L = _lib()
    for kind, seed in (("power", 5), ("gauss", 6)):
        W, s = O.synthetic_weight(1024, 1024, seed=seed, kind=kind)
        if kind == "gauss":
            s[::17] *= 300.0                                       # outlier channels
        scale = (s.half().float() ** 0.5 + 1e-6)
        fact = L.scaled_svd([W.cuda()], [scale.cuda()])
        ref = torch.linalg.svdvals(W.double() * scale.double())
        sig = fact.sigma(0).cpu().double()
        r = O.rank_for_ratio(1024, 1024, 0.9)
        big = ref[:r] > 1e-3 * ref[0]
        strict = ((sig[:r] - ref[:r]).abs() / ref[:r])[big].max().item()
        assert strict < SIGMA_RTOL, (kind, strict)
        allowance = ((sig[:r] - ref[:r]).abs() / (ref[:r] + 1e-2 * ref[0])).max().item()
        assert allowance < SIGMA_RTOL, (kind, allowance)
        absolute = (sig[:r] - ref[:r]).abs().max().item() / ref[0].item()
        assert absolute < 5e-6, (kind, absolute)                   # ~40 eps32 of sigma_max


# ------------------------------------------------------------------------------------------------ --ppl_target (binary_search.py:64-87)
@pytest.fixture(scope="module")
def golden_ppl_target():
    from conftest import GOLDEN
    return torch.load(os.path.join(GOLDEN, "ppl_target_and_opt125m_shapes.pt"), weights_only=False)


def test_ppl_target_search_against_exact_oracle(golden_pipeline, golden_ppl_target):
    """The ppl-target branch end to end on the GPU against the oracle's restatement run with the EXACT factorisation
    (the restatement itself is pinned to upstream's log in tests/test_oracle_golden.py).  Same bisection decisions,
    same final modules -- raw nn.Linear restored where the final allocation says default -- same perplexity."""
    from asvd4llm_b200 import SVDLinear
    from asvd4llm_b200.binary_search import binary_search_truncation_rank
    from asvd4llm_b200.evaluate_utils import evaluate_perplexity
    g = golden_ppl_target["ppl_target"]
    loader = golden_pipeline["loader"]
    ids = torch.cat([b["input_ids"] for b in loader], 0)
    for target in (g["target"], 0.5 * (golden_pipeline["ppl_raw"] + g["target"])):
        ref_model = build_tiny_opt(golden_pipeline)
        model = build_tiny_opt(golden_pipeline).cuda()
        for mdl in (ref_model, model):
            for n_, m_ in mdl.named_modules():
                if isinstance(m_, nn.Linear):
                    m_.scaling_diag_matrix = golden_pipeline["sdm_abs_mean"][n_].clone().to(m_.weight.device)
        want = O.binary_search_truncation_rank(ref_model, golden_pipeline["sensitivity"], loader, ppl_target=target, method="exact")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            binary_search_truncation_rank(model, golden_pipeline["sensitivity"], loader, _args(ppl_target=target, param_ratio_target=-1))
        got = [l for l in buf.getvalue().splitlines() if l.startswith("low=") or l.startswith("===")]
        ambiguous = False
        for lg, lw in zip(got, want):
            if "ppl=" in lw and not lw.startswith("==="):
                pw = float(lw.split("ppl=")[1].split(",")[0]); pg = float(lg.split("ppl=")[1].split(",")[0])
                assert lg.split(", ppl=")[0] == lw.split(", ppl=")[0], (lg, lw)
                assert pg == pytest.approx(pw, rel=2e-3), (lg, lw)
                if abs(pw - target) < 4e-3 * target:               # a decision inside the fp tolerance: stop comparing the path
                    ambiguous = True
                    break
        if ambiguous:
            continue
        assert len(got) == len(want)
        kinds_ref = {n: (type(m).__name__.replace("Oracle", ""), getattr(m, "truncation_rank", None))
                     for n, m in ref_model.named_modules() if n in g["kinds"]}
        kinds = {n: (type(m).__name__, getattr(m, "truncation_rank", None)) for n, m in model.named_modules() if n in g["kinds"]}
        assert kinds == kinds_ref
        assert evaluate_perplexity(model, ids, 3) == pytest.approx(O.evaluate_perplexity(ref_model, ids, 3), rel=2e-3)


def test_opt125m_shapes_against_upstream(golden_ppl_target):
    """BASELINE config 1's real layer shapes (768x768, 3072x768, 768x3072) against what the UNMODIFIED upstream
    from_linear produced for the same seeded inputs (tests/golden/make_golden_ppl_target.py): same rank, reconstruction
    no worse than upstream's (Eckart-Young), leading singular values and module output in agreement."""
    from asvd4llm_b200 import SVDLinear
    from conftest import GOLDEN
    src = open(os.path.join(GOLDEN, "make_golden_ppl_target.py")).read()
    ns = {}
    exec(src[src.index("def opt125m_case"):src.index("OPT125M_SHAPES")], {"torch": torch}, ns)
    for c in golden_ppl_target["opt125m_shapes"]:
        W, sdm, x = ns["opt125m_case"](c["idx"], c["m"], c["n"])
        lin = nn.Linear(c["n"], c["m"], bias=False)
        lin.weight.data = W.clone()
        lin = lin.cuda()
        lin.scaling_diag_matrix = sdm.cuda()
        mod = SVDLinear.from_linear(lin, c["ratio"], act_aware=True, alpha=0.5, sigma_fuse="UV")
        assert mod.truncation_rank == c["rank"]
        A, B = mod.ALinear.weight.data.double().cpu(), mod.BLinear.weight.data.double().cpu()
        s = O.scaling_vector(sdm, None, 0.5).double()
        rec = (((A @ B) - W.double()) * s).norm() / (W.double() * s).norm()
        assert rec <= c["recon_scaled"] * (1 + 2e-3), (c["m"], c["n"], float(rec), c["recon_scaled"])
        lead = max(8, c["rank"] // 8)
        an = A.norm(dim=0)[:lead]
        assert torch.allclose(an, c["a_col_norms"][:lead].double(), rtol=2e-2), (c["m"], c["n"])
        # module output: the two rank-r approximations span different tail subspaces (svd_lowrank vs exact), so the outputs
        # agree to within the truncation error itself (both are that far from x W^T), not to rounding
        y = mod(x.cuda()).float().cpu()
        yu = c["y"].float()
        yfull = x.float() @ W.float().t()
        assert (y - yu).norm().item() < 1.5 * c["recon_scaled"] * yfull.norm().item() + 1e-3
        assert (y - yfull).norm().item() < 1.5 * (yu - yfull).norm().item() + 1e-3


# ------------------------------------------------------------------------------------------------ N1 / multi-GPU host paths on the device
def test_sensitivity_sweep_units_and_batched_evaluation(golden_pipeline, tmp_path, monkeypatch):
    """the sweep sharded by (layer, ratio) units over two 'ranks' and merged equals the unsharded table,
    and evaluating several calibration samples per forward (--eval_batch_size) reproduces upstream's batch-1 table."""
    from asvd4llm_b200.sensitivity import calib_sensitivity_ppl
    from asvd4llm_b200 import sharding
    monkeypatch.chdir(tmp_path)
    os.makedirs("cache")
    loader = golden_pipeline["loader"]
    model = build_tiny_opt(golden_pipeline).cuda()
    for n_, m_ in model.named_modules():
        if isinstance(m_, nn.Linear):
            m_.scaling_diag_matrix = golden_pipeline["sdm_abs_mean"][n_].clone().cuda()
    with contextlib.redirect_stdout(io.StringIO()):
        full = calib_sensitivity_ppl(model, loader, _args(), use_cache=False)
        parts = [calib_sensitivity_ppl(model, loader, _args(), use_cache=False, unit_filter=lambda u, r=r: u % 2 == r) for r in (0, 1)]
        batched = calib_sensitivity_ppl(model, loader, _args(eval_batch_size=3), use_cache=False)
    assert sum(len(row) for p in parts for row in p.values()) == sum(len(row) for row in full.values())
    merged = {}
    for p in parts:
        for layer, row in p.items():
            merged.setdefault(layer, {}).update(row)
    merged = {k: dict(sorted(merged[k].items())) for k in full}
    assert merged == full                                            # same kernels, same inputs: bitwise
    for layer in full:
        for ratio in full[layer]:
            assert batched[layer][ratio] == pytest.approx(full[layer][ratio], rel=1e-5), (layer, ratio)


def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        from asvd4llm_b200 import sharding
        from asvd4llm_b200.modules.svd_linear import SVDLinear
        from asvd4llm_b200.sensitivity import enumerate_linears
        torch.cuda.set_device(0)
        model = bench.build_llama_like(torch.device("cuda", 0), n_blocks=2, hidden=256, inter=640, vocab=1000)
        chosen = {full: 0.9 for _, _, full, _ in enumerate_linears(model)}
        chosen[list(chosen)[3]] = 1                               # one layer stays raw
        ns = argparse.Namespace(alpha=0.5, act_aware=True, sigma_fuse="UV", rank_align=1)
        stats = sharding.decompose_sharded(model, chosen, 1, ns)
        digest = {}
        for name, mod in model.named_modules():
            if isinstance(mod, SVDLinear):
                digest[name] = (mod.truncation_rank, mod.ALinear.weight.data.cpu().view(torch.int16).long().sum().item(),
                                mod.BLinear.weight.data.cpu().view(torch.int16).long().sum().item())
        q.put((rank, "ok", digest, stats["collectives"]))
    except Exception:  # noqa
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


def test_sharded_final_pass_is_bitwise_independent_of_world_size():
    """config 3's path at a small size: sharding.decompose_sharded with two ranks (gloo, both on cuda:0 -- the GPU box
    of the test tier has one device) installs, on every rank, bit-identical factors to the single-rank run."""
    import socket
    import torch.multiprocessing as mp
    results = {}
    for world in (1, 2):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = [q.get(timeout=600) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        for rank, msg, digest, ncoll in got:
            assert msg == "ok", f"world {world} rank {rank}:\n{msg}"
            results[(world, rank)] = digest
            assert ncoll == (0 if world == 1 else 2)
    assert len(results[(1, 0)]) == 14                               # 15 linears, one kept raw
    assert results[(2, 0)] == results[(1, 0)] and results[(2, 1)] == results[(1, 0)]


def test_mixed_convergence_batch_is_bitwise_independent():
    """Weights that reach the precise Gram mode in different sweeps (a Gaussian bulk, a decaying spectrum, an almost
    orthogonal matrix, a rank-deficient one) share one batch: every matrix must come out bitwise as when it is factorised
    alone -- the Gram mode is per-matrix state, a round of a batch in transition runs one Gram launch per mode."""
    L = _lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 1024
    Ws = []
    Ws.append((torch.randn(n, n, device="cuda", generator=g) * 0.02).half())
    u = torch.randn(n, 64, device="cuda", generator=g); v = torch.randn(64, n, device="cuda", generator=g)
    Ws.append(((torch.randn(n, n, device="cuda", generator=g) + (u * torch.logspace(0, -2, 64, device="cuda") * 8) @ v) * 0.02).half())
    Q = torch.linalg.qr(torch.randn(n, n, device="cuda", generator=g)).Q
    Ws.append((Q * torch.linspace(1.0, 0.1, n, device="cuda")).half())
    Ws.append((torch.randn(n, 100, device="cuda", generator=g) @ torch.randn(100, n, device="cuda", generator=g) * 0.01).half())
    Ws.append((torch.randn(n, n, device="cuda", generator=g) * 0.02).half())
    scales = [L.scaling_vector(torch.exp(torch.randn(n, device="cuda", generator=g)).half(), None, 0.5, n, "cuda") for _ in Ws]
    fact = L.scaled_svd(Ws, scales)
    assert len(set(fact.sweeps)) > 1, fact.sweeps                      # they really do converge at different times
    for b in range(len(Ws)):
        single = L.scaled_svd([Ws[b]], [scales[b]])
        assert torch.equal(single.sigma(0), fact.sigma(b)), b
        A1, B1 = single.extract(400, "UV", torch.float16, 0)
        A2, B2 = fact.extract(400, "UV", torch.float16, b)
        assert torch.equal(A1, A2) and torch.equal(B1, B2), b


# ------------------------------------------------------------------------------------------------ N3: statistic fused into its producer
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("method", ["abs_mean", "abs_max"])
def test_linear_forward_stat_matches_linear_and_hook_math(dtype, method):
    """asvd_linear_forward_stat: the layer output equals the library linear to output rounding, and the accumulator gets
    the hook's update (act_aware_utils.py:64-74, oracle abs_stat_update) -- over two calls, ragged M, with and without bias."""
    L = _lib()
    g = torch.Generator().manual_seed(17)
    for (M, n, m, with_bias) in [(2048, 768, 3072, True), (333, 64, 136, False), (1, 8, 8, True), (700, 1024, 6280, True)]:
        W = (torch.randn(m, n, generator=g) / n ** 0.5).to(dtype)
        bias = (torch.randn(m, generator=g) * 0.1).to(dtype) if with_bias else None
        acc = torch.zeros(n, dtype=dtype, device="cuda")
        ref_acc = 0
        for call in range(2):
            x = (torch.randn(1, M, n, generator=g) * (1.0 + call)).to(dtype)
            assert L.linear_stat_eligible(x.cuda(), W.cuda(), None if bias is None else bias.cuda())
            y = L.linear_forward_stat(x.cuda(), W.cuda(), None if bias is None else bias.cuda(), acc, method)
            yref = torch.nn.functional.linear(x.cuda(), W.cuda(), None if bias is None else bias.cuda())
            tol = (2e-2 if dtype == torch.bfloat16 else 3e-3) * max(1.0, yref.float().abs().max().item())
            assert y.shape == yref.shape and (y.float() - yref.float()).abs().max().item() < tol
            ref_acc = O.abs_stat_update(ref_acc, x, method)
        rtol = 2e-2 if dtype == torch.bfloat16 else 2e-3                 # one ulp of the accumulator dtype
        assert torch.allclose(acc.float().cpu(), ref_acc.float(), rtol=rtol, atol=1e-6), (M, n, m, method)
    # NaN: the mean propagates it (`+= abs_mean`), upstream's `torch.where(abs_max > acc, abs_max, acc)` drops it
    x = torch.ones(1, 16, 8, dtype=dtype); x[0, 3, 2] = float("nan")
    acc = torch.zeros(8, dtype=dtype, device="cuda")
    L.linear_forward_stat(x.cuda(), torch.eye(8, dtype=dtype).cuda(), None, acc, method)
    ref = O.abs_stat_update(torch.zeros(8, dtype=dtype), x, method)
    assert torch.equal(torch.isnan(acc.cpu()), torch.isnan(ref)) and torch.isfinite(acc[[0, 1, 3, 4, 5, 6, 7]]).all()


def test_fused_calibration_matches_hook_path(golden_pipeline, tmp_path, monkeypatch):
    """calib_input_distribution on a 16-bit model: the fused path (statistic as a side output of every layer's GEMM) against
    upstream's structure (torch forward + hook, ASVD_B200_CALIB=hook).  The two forwards differ by GEMM rounding, so the
    statistics agree to a few ulps of fp16, not bitwise."""
    from asvd4llm_b200.act_aware_utils import calib_input_distribution
    monkeypatch.chdir(tmp_path)
    os.makedirs("cache")
    loader = golden_pipeline["loader"]
    out = {}
    for mode in ("hook", "fused"):
        monkeypatch.setenv("ASVD_B200_CALIB", mode)
        for method in ("abs_mean", "abs_max"):
            model = build_tiny_opt(golden_pipeline).half().cuda()
            before = _lib().profile_read()["forward"][1]
            calib_input_distribution(model, loader, method, use_cache=False)
            used_gemm = _lib().profile_read()["forward"][1] > before
            assert used_gemm == (mode == "fused")
            assert all("forward" not in vars(mod) for mod in model.modules())       # patched forwards are gone again
            out[(mode, method)] = {n: mod.scaling_diag_matrix.float().cpu() for n, mod in model.named_modules() if isinstance(mod, nn.Linear)}
    for method in ("abs_mean", "abs_max"):
        for name, ref in out[("hook", method)].items():
            got = out[("fused", method)][name]
            assert torch.allclose(got, ref, rtol=1e-2, atol=1e-4), (method, name, (got - ref).abs().max().item())
