"""GPU: kernel-level parity of the tensor-core / recurrence kernels against the live oracle (fp64 on CPU) and against the
exact-fp32 companion path of the same library, at the shapes and edge cases the model-level golden tests do not reach."""
import ctypes

import numpy as np
import pytest
import torch

from _util import rel_err, poison_cuda_cache
from oracle import eegclip_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3   # north star: outputs and gradients <= 1e-3 relative


@pytest.fixture(scope="module")
def lib():
    import transformer_clip_eeg_b200 as p  # noqa: F401
    from transformer_clip_eeg_b200 import _lib
    _lib.load()
    return _lib


@pytest.fixture(scope="module")
def cm(lib):
    from transformer_clip_eeg_b200 import clip_model
    return clip_model


# ---------------------------------------------------------------------------------------------------------------
# bidirectional LSTM (clip_model.py:267-268,322-323) vs the oracle's restated recurrence, incl. ragged batches
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("In,H,B,T", [(64, 128, 5, 24), (64, 128, 8, 64), (64, 128, 20, 320), (256, 4, 7, 40), (256, 4, 32, 64)])
def test_bilstm_vs_oracle(cm, In, H, B, T):
    torch.manual_seed(In + H + B)
    mod = torch.nn.LSTM(In, H, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, In)
    w = torch.randn(B, T, 2 * H)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in mod.state_dict().items()}
    xr = x.double().requires_grad_(True)
    yr = O.bilstm(sd, "", xr)
    (yr * w.double()).sum().backward()
    mg = mod.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    yg = cm._bilstm(mg, xg)
    assert yg.shape == (B, T, 2 * H)
    (yg * w.to(DEV)).sum().backward()
    assert rel_err(yg, yr) < TOL
    assert rel_err(xg.grad, xr.grad) < TOL
    gall = sum(float(v.grad.norm()) ** 2 for v in sd.values()) ** 0.5
    for k, p in mg.named_parameters():
        assert rel_err(p.grad, sd[k].grad, floor=1e-3 * gall) < TOL, k


def test_bilstm_uncovered_shape_raises(cm, lib):
    """LSTM shapes outside the two of the default speech tower raise: there is no library (cuDNN) fallback on this path."""
    from transformer_clip_eeg_b200 import _lib
    mod = torch.nn.LSTM(32, 16, batch_first=True, bidirectional=True).to(DEV)
    with pytest.raises(_lib.EegclipError):
        cm._bilstm(mod, torch.randn(2, 8, 32, device=DEV))


# ---------------------------------------------------------------------------------------------------------------
# tensor-core attention vs the exact-fp32 attention kernels of the same library (identical Philox masks)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,B,train,p", [(64, 3, False, 0.5), (192, 2, True, 0.5), (320, 2, True, 0.5), (128, 2, True, 0.3), (448, 1, True, 0.5), (512, 1, True, 0.5)])
def test_attention_tc_vs_fp32(cm, lib, T, B, train, p):
    torch.manual_seed(T + B)
    blk = cm.TransformerEncoderBlock(64, drop_p=p, forward_drop_p=p).to(DEV)
    blk.train(train)
    x = torch.randn(B, T, 64, device=DEV)
    w = torch.randn(B, T, 64, device=DEV)
    res = {}
    seed = 1234567
    for math in ("fp32", "bf16x3"):
        lib.set_default_math(math)
        poison_cuda_cache()               # the second run must not inherit the first run's (correct) output buffers
        try:
            torch.manual_seed(seed)       # same Philox key for both runs
            xx = x.clone().requires_grad_(True)
            blk.zero_grad()
            y = blk(xx)
            (y * w).sum().backward()
            res[math] = [y.detach().clone(), xx.grad.detach().clone()] + [q.grad.detach().clone() for q in blk.parameters()]
        finally:
            lib.set_default_math("bf16x3")
    gall = sum(float(t.norm()) ** 2 for t in res["fp32"][2:]) ** 0.5
    for a, b in zip(res["bf16x3"], res["fp32"]):
        assert rel_err(a, b, floor=1e-3 * gall) < TOL


@pytest.mark.parametrize("T", [320, 192, 128])
def test_attention_backward_is_bitwise_reproducible(cm, T):
    """dQ is summed over the key warps in warp order through per-warp shared-memory slots (no atomics): three runs of the attention
    forward + backward on the same inputs and dropout key give bit-identical, finite outputs and gradients (T = 128 / 192 run
    with 64 / 96 threads, fewer than the 128 elements of a 16-query dQ block; free memory is NaN before every run)."""
    torch.manual_seed(5)
    qkv = torch.randn(3, T, 192, device=DEV)
    w = torch.randn(3, T, 64, device=DEV)
    runs = []
    for _ in range(3):
        poison_cuda_cache()
        q = qkv.clone().requires_grad_(True)
        y = cm._AttentionFn.apply(q, 0.5, True, 3, 4242)
        (y * w).sum().backward()
        runs.append((y.detach().clone(), q.grad.detach().clone()))
    assert bool(torch.isfinite(runs[0][0]).all()) and bool(torch.isfinite(runs[0][1]).all())
    for y, g in runs[1:]:
        assert torch.equal(y, runs[0][0]) and torch.equal(g, runs[0][1])


# ---------------------------------------------------------------------------------------------------------------
# token GEMMs through the C ABI: every (N, K) family, ragged M, against fp64
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (4100, 192, 64), (777, 256, 64), (2048, 64, 256), (1300, 64, 1024), (520, 64, 192),
                                   (640, 128, 128), (5, 64, 64), (33, 192, 64), (31, 64, 256)])
def test_linear_tc_vs_fp64(cm, M, N, K):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, requires_grad=True)
    w = (torch.randn(N, K, device=DEV) * 0.2).requires_grad_(True)
    b = torch.randn(N, device=DEV, requires_grad=True)
    g = torch.randn(M, N, device=DEV)
    y = cm._LinearFn.apply(x, w, b)
    (y * g).sum().backward()
    xd, wd, bd, gd = (t.detach().double().cpu() for t in (x, w, b, g))
    yr = xd @ wd.t() + bd
    assert rel_err(y, yr) < 2e-5
    assert rel_err(x.grad, gd @ wd) < 2e-5
    assert rel_err(w.grad, gd.t() @ xd) < 2e-5
    assert rel_err(b.grad, gd.sum(0)) < 2e-5


# ---------------------------------------------------------------------------------------------------------------
# tcgen05 head at the benchmark size and sharded (two row blocks emulated on one GPU through the C ABI)
# ---------------------------------------------------------------------------------------------------------------
def _head_ref(E, S, tau):
    Ed, Sd = E.detach().double().cpu().requires_grad_(True), S.detach().double().cpu().requires_grad_(True)
    td = tau.detach().double().cpu().requires_grad_(True)
    loss = O.symmetric_infonce(Ed, Sd, td)
    loss.backward()
    return loss, Ed.grad, Sd.grad, td.grad


@pytest.mark.parametrize("B,D", [(256, 2560), (128, 1536), (320, 512)])
def test_head_tc_vs_fp64(lib, B, D):
    from transformer_clip_eeg_b200.parallel import infonce_loss
    torch.manual_seed(B + D)
    E = torch.randn(B, D, device=DEV, requires_grad=True)
    S = (0.7 * E.detach() + torch.randn(B, D, device=DEV)).requires_grad_(True)   # correlated pairs: a peaked softmax
    tau = torch.tensor(1.3, device=DEV, requires_grad=True)
    loss = infonce_loss(E, S, tau)
    loss.backward()
    lr, dE, dS, dt = _head_ref(E, S, tau)
    assert abs(float(loss) - float(lr)) <= 1e-5 * max(1.0, abs(float(lr)))
    assert rel_err(E.grad, dE) < TOL and rel_err(S.grad, dS) < TOL
    assert abs(float(tau.grad) - float(dt)) <= TOL * max(abs(float(dt)), 1e-3)


def test_head_sharded_rows_match_unsharded(lib):
    """Rank r of 2 scores rows [128r, 128r+128) against the gathered 256: LSE / diag / gradients equal the world-1 result."""
    from transformer_clip_eeg_b200.parallel import CudaHeadOps
    ops = CudaHeadOps()
    torch.manual_seed(5)
    Bg, b, D = 256, 128, 1536
    En, _ = ops.l2norm_fwd(torch.randn(Bg, D, device=DEV))
    Sn, _ = ops.l2norm_fwd(torch.randn(Bg, D, device=DEV))
    tau = torch.tensor([0.4], device=DEV)
    full = ops.lse(Sn, En, tau, Bg, 0, False)
    parts = [ops.lse(Sn, En, tau, b, r * b, False) for r in range(2)]
    vec = torch.cat(parts, dim=1)
    assert rel_err(vec, full) < 1e-5
    dl = torch.ones(1, device=DEV)
    dS0, dE0, dt0 = ops.backward(Sn, En, tau, full, Bg, 0, dl, False)
    outs = [ops.backward(Sn, En, tau, vec.contiguous(), b, r * b, dl, False) for r in range(2)]
    assert rel_err(torch.cat([o[0] for o in outs]), dS0) < 1e-4
    assert rel_err(torch.cat([o[1] for o in outs]), dE0) < 1e-4
    assert abs(float(outs[0][2] + outs[1][2]) - float(dt0)) <= 1e-4 * max(abs(float(dt0)), 1e-3)


# ---------------------------------------------------------------------------------------------------------------
# gradient sinks: backward kernels writing straight into the optimizer's arena == ordinary autograd accumulation
# ---------------------------------------------------------------------------------------------------------------
def test_grad_sinks_match_autograd(cm, lib):
    from transformer_clip_eeg_b200.optim import AdamW
    from transformer_clip_eeg_b200 import train_clip_final as tcf
    torch.manual_seed(11)
    T, B = 64, 8
    args = tcf.build_parser().parse_args(["--attention_depth", "2"])
    model = tcf.build_model(args, T, 100, torch.device(DEV)).eval()
    for m in model.modules():
        if isinstance(m, torch.nn.LSTM):
            m.train()
    eeg, sp = torch.randn(B, T, 64, device=DEV), torch.randn(B, T, 1024, device=DEV)
    ids = torch.arange(1, B + 1, device=DEV)
    mem0 = model.eegMemoryBank.memory.clone()
    # ordinary path: no arena registered yet
    _, _, tot = model(eeg, sp, ids)
    tot.backward()
    ref = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    # sink path: arena views claimed after zero_grad
    model.eegMemoryBank.memory.copy_(mem0)
    opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    _, _, tot = model(eeg, sp, ids)
    opt.zero_grad()
    tot.backward()
    flat = opt.flat_grads()[0]
    gall = sum(float(v.norm()) ** 2 for v in ref.values()) ** 0.5
    for k, p in model.named_parameters():
        if not k.startswith("temperature"):   # (the two log-scales' gradients come through autograd; step() folds them into the arena)
            assert p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr(), k   # a view of the arena
        assert rel_err(p.grad, ref[k], floor=1e-3 * gall) < 1e-5, k
    # a second backward without zero_grad must accumulate (falls back to autograd's add)
    _, _, tot = model(eeg, sp, ids)
    model.eegMemoryBank.memory.copy_(mem0)
    tot.backward()
    w = dict(model.named_parameters())["eegModel.conv_0.conv.weight"]
    assert rel_err(w.grad, 2 * ref["eegModel.conv_0.conv.weight"]) < 1e-3


# ---------------------------------------------------------------------------------------------------------------
# match-mismatch scoring (train_clip_helper_functions.py:153-163,176-187): candidate row-dots and the N x M bank
# similarity on the tcgen05 kernel; decisions (argmax / top-k sets) must equal fp64 decisions on continuous inputs
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,M,D", [(64, 100, 1536), (300, 1000, 2560), (129, 257, 2560), (16, 2053, 512)])
def test_bank_logits_tc_vs_fp64(lib, N, M, D):
    from transformer_clip_eeg_b200 import train_clip_helper_functions as H
    torch.manual_seed(N + M)
    E = torch.randn(N, D, device=DEV)
    Bk = torch.randn(M, D, device=DEV)
    got = H.bank_logits(E, Bk)
    ref = E.double() @ Bk.double().T
    assert got.shape == (N, M)
    assert rel_err(got, ref) < 1e-5
    k = min(100, M)
    ti = torch.topk(got, k, dim=1).indices
    tr = torch.topk(ref, k, dim=1).indices
    assert torch.equal(ti[:, 0], tr[:, 0])
    # same top-k SET per row (order inside the set may flip only where fp64 gaps are below fp32 resolution)
    same = (torch.sort(ti, dim=1).values == torch.sort(tr, dim=1).values).all(dim=1).float().mean()
    assert float(same) > 0.99


@pytest.mark.parametrize("K", [2, 5, 100])
def test_mm_rowdots_decisions(lib, K):
    from transformer_clip_eeg_b200 import train_clip_helper_functions as H
    torch.manual_seed(K)
    N, D = 257, 2560
    E = torch.randn(N, D, device=DEV)
    C = torch.randn(N, K, D, device=DEV)
    scores, choice = H.mm_scores(E, C)
    ref = torch.einsum("nd,nkd->kn", E.double(), C.double())
    assert rel_err(scores, ref) < 1e-5
    assert torch.equal(choice, ref.argmax(dim=0))


def test_bank_topk_world1_matches_reference_topk(lib):
    from transformer_clip_eeg_b200 import train_clip_helper_functions as H
    torch.manual_seed(3)
    E, Bk = torch.randn(96, 2560, device=DEV), torch.randn(1500, 2560, device=DEV)
    v, i = H.bank_topk(E, Bk, 100)
    rv, ri = torch.topk(E.double() @ Bk.double().T, 100, dim=1)
    assert torch.equal(i[:, :10], ri[:, :10])
    assert rel_err(v, rv) < 1e-5


def test_loss_reader_returns_every_step_in_order(lib):
    """The e2e loop's deferred device->host loss read (train_clip_final.LossReader) hands back every pushed value, in order."""
    from transformer_clip_eeg_b200.train_clip_final import LossReader
    rd = LossReader(torch.device(DEV), depth=4)
    got = []
    vals = [float(i) * 0.5 - 3.0 for i in range(11)]
    for v in vals:
        got += rd.push(torch.tensor(v, device=DEV))
    assert len(got) == len(vals) - 1          # lag of exactly one step
    got += rd.drain()
    assert got == vals


def test_train_step_is_bitwise_reproducible_and_stream_independent(cm, lib, monkeypatch):
    """Every reduction of the default train step runs in a fixed order (weight-gradient partials, LayerNorm affine gradients,
    attention dQ, conv bias column sums, the H = 4 LSTM's recurrent weight gradient, the head's d tau: no float atomics), and the
    speech tower on its side stream (clip_model.run_towers, the default) runs the same kernels in the same per-tower order as the
    single-stream step.  So three train steps (train mode, same seeds) are BITWISE equal -- losses, every step's gradient arena and
    the final parameters / bank -- between two runs and between the two stream modes; a missing cross-stream edge (forward join,
    end-of-backward join before optimizer.step, next step's side stream waiting for the parameter update) would show up here."""
    from transformer_clip_eeg_b200.optim import AdamW
    from transformer_clip_eeg_b200 import train_clip_final as tcf
    T, B, lr = 128, 16, 1e-3
    args = tcf.build_parser().parse_args(["--attention_depth", "2"])
    res = []
    for mode in ("0", "1", "1"):
        monkeypatch.setenv("EEGCLIP_TWO_STREAMS", mode)
        assert cm.two_streams_enabled() == (mode == "1")
        poison_cuda_cache()
        torch.manual_seed(5)
        model = tcf.build_model(args, T, 100, torch.device(DEV)).train()
        opt = AdamW(model.parameters(), lr=lr, weight_decay=0.01)
        g = torch.Generator().manual_seed(9)
        losses, grads = [], []
        for step in range(3):
            eeg, sp = torch.randn(B, T, 64, generator=g).to(DEV), torch.randn(B, T, 1024, generator=g).to(DEV)
            ids = torch.arange(1, B + 1, device=DEV)
            loss_ce, _, _ = tcf.train_step(model, opt, eeg, sp, ids)
            losses.append(loss_ce.detach().clone())
            grads.append(opt.flat_grads()[0].detach().clone())
        torch.cuda.synchronize()
        res.append((torch.stack(losses).cpu(), grads, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    l0, g0, p0 = res[0]
    assert float(g0[0].norm()) > 0 and bool(torch.isfinite(l0).all())
    for l1, g1, p1 in res[1:]:
        assert torch.equal(l0, l1)
        for a, b in zip(g0, g1):
            assert torch.equal(a, b), float((a - b).abs().max())
        for k, v in p0.items():
            assert torch.equal(v, p1[k]), k


def test_pdl_off_matches_pdl_on(cm, lib):
    """Programmatic dependent launch only overlaps launch latency: the forward is bit-identical with the attribute off
    (g_tune[7]); so are the gradients (every reduction runs in a fixed order)."""
    torch.manual_seed(11)
    model = cm.EEGConformerInterleaved(output_dim=8, time_dimension=192, depth=2).to(DEV).eval()
    x = torch.randn(4, 192, 64, device=DEV)
    outs = []
    for off in (0, 1):
        lib.call("eegclip_tune_set", 7, off)
        try:
            xx = x.clone().requires_grad_(True)
            model.zero_grad()
            y = model(xx)
            y.square().sum().backward()
            outs.append((y.detach().clone(), xx.grad.detach().clone(), model.conv_0.conv.weight.grad.detach().clone()))
        finally:
            lib.call("eegclip_tune_set", 7, 0)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_device_prefetcher_fp32_and_fp16_staging(lib):
    """The input pipeline hands back exactly the host batches (fp32 staging) or their fp16 rounding (opt-in staging), in order."""
    from transformer_clip_eeg_b200.train_clip_final import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(3, 64, 64, generator=g), [torch.randn(3, 64, 1024, generator=g)], torch.arange(3) + 10 * i, None) for i in range(5)]
    for dt in (None, torch.float16):
        got = [(e.clone(), s.clone(), i.clone()) for e, s, i in DevicePrefetcher(batches, torch.device(DEV), speech_stage_dtype=dt)]
        assert len(got) == len(batches)
        for (e, s, i), (he, hs, hi, _) in zip(got, batches):
            assert torch.equal(e.cpu(), he) and torch.equal(i.cpu(), hi) and s.dtype == torch.float32
            ref = hs[0] if dt is None else hs[0].half().float()
            assert torch.equal(s.cpu(), ref)
