"""GPU tests of mvae_gemm_chain: a Linear+Swish stack and its autograd chain as ONE persistent launch whose tiles wait
on row-block completion counters.  Checked against an fp64 torch reference of the same stack and, bit for bit where the
arithmetic is order-independent, against the one-launch-per-layer path (mvae_gemm_batch)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {0: 3e-3, 1: 2e-5}   # same statement as tests/test_kernels_gpu.py: max|out-ref| / max|ref| vs fp64


@pytest.fixture(scope="module")
def ops():
    from multimodal_vae_public_b200 import ops as _ops
    return _ops


def _rel(a, ref):
    return (a.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def _stack(M, widths, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, widths[0], device="cuda", generator=g)
    ws = [torch.randn(widths[i + 1], widths[i], device="cuda", generator=g) / widths[i] ** 0.5 for i in range(len(widths) - 1)]
    bs = [torch.randn(widths[i + 1], device="cuda", generator=g) for i in range(len(widths) - 1)]
    return x, ws, bs


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M", [8, 300, 1024, 8192])
def test_forward_stack_chain(ops, prec, M):
    """4-layer Linear+Swish stack (the MNIST image decoder shape) in one launch; every layer waits on the previous."""
    widths = [64, 512, 512, 512, 784]
    x, ws, bs = _stack(M, widths, seed=M)
    a = [torch.full((M, w), float("nan"), device="cuda") for w in widths[1:]]
    h = [torch.full((M, w), float("nan"), device="cuda") for w in widths[1:-1]]
    wsb = ops.chain_workspace("cuda")
    for rep in range(3):     # the kernel must leave the workspace ready for the next launch
        descs, xin = [], x
        for l in range(4):
            last = l == 3
            descs.append(ops.gemm_desc(xin, ws[l], a[l], M, widths[l + 1], widths[l], bias=bs[l],
                                       out2=None if last else h[l], epilogue=ops.EPI_STORE if last else ops.EPI_BIAS_SWISH))
            if not last:
                xin = h[l]
        ops.gemm_chain(descs, [-1, 0, 1, 2], wsb, prec)
        torch.cuda.synchronize()
        assert int(wsb[1]) == 0, "a dependency wait timed out"
        assert int(wsb.abs().sum()) == 0, "counters not reset"
    ref = x.double()
    for l in range(4):
        ref = ref @ ws[l].double().t() + bs[l].double()
        assert _rel(a[l], ref) < TOL[prec] * (l + 1)
        if l < 3:
            ref = ref * torch.sigmoid(ref)
    # the layer-per-launch path computes the very same tiles: identical bits
    a2 = [torch.empty_like(t) for t in a]; h2 = [torch.empty_like(t) for t in h]
    xin = x
    for l in range(4):
        last = l == 3
        ops.gemm_batch([ops.gemm_desc(xin, ws[l], a2[l], M, widths[l + 1], widths[l], bias=bs[l],
                                      out2=None if last else h2[l],
                                      epilogue=ops.EPI_STORE if last else ops.EPI_BIAS_SWISH)], prec)
        if not last:
            xin = h2[l]
    for l in range(4):
        assert torch.equal(a[l], a2[l])


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M", [64, 1000, 4096])
def test_backward_chain(ops, prec, M):
    """dgrad chain (K-major A dependencies) + wgrads (MN-major A: k-range dependencies) + fused bias-gradient sums."""
    widths = [64, 512, 512, 784]
    x, ws, bs = _stack(M, widths, seed=7 * M)
    g = torch.Generator(device="cuda").manual_seed(M + 1)
    # forward in fp64 to get the saved activations
    acts_a, acts_h, cur = [], [x], x.double()
    for l in range(3):
        cur = cur @ ws[l].double().t() + bs[l].double()
        acts_a.append(cur.float())
        if l < 2:
            cur = cur * torch.sigmoid(cur)
            acts_h.append(cur.float())
    dy = torch.randn(M, 784, device="cuda", generator=g)
    dA = [torch.full((M, 512), float("nan"), device="cuda") for _ in range(2)]     # dA[1] = d a_2, dA[0] = d a_1
    dX = torch.zeros(M, 64, device="cuda")
    dW = [torch.zeros_like(w) for w in ws]
    db = [torch.zeros(512, device="cuda") for _ in range(2)]
    split = max(1, min((M // 32) // 16, 32))
    D = ops.gemm_desc
    descs = [
        D(dy, ws[2], dA[1], M, 512, 784, b_mn=True, aux=acts_a[1], epilogue=ops.EPI_MUL_DSWISH, colsum=db[1]),   # 0
        D(dy, acts_h[2], dW[2], 784, 512, M, a_mn=True, b_mn=True, split_k=split, accumulate=True),              # 1
        D(dA[1], ws[1], dA[0], M, 512, 512, b_mn=True, aux=acts_a[0], epilogue=ops.EPI_MUL_DSWISH, colsum=db[0]),  # 2 <- 0
        D(dA[1], acts_h[1], dW[1], 512, 512, M, a_mn=True, b_mn=True, split_k=split, accumulate=True),            # 3 <- 0
        D(dA[0], ws[0], dX, M, 64, 512, b_mn=True, accumulate=True),                                              # 4 <- 2
        D(dA[0], acts_h[0], dW[0], 512, 64, M, a_mn=True, b_mn=True, split_k=split, accumulate=True),             # 5 <- 2
    ]
    wsb = ops.chain_workspace("cuda")
    ops.gemm_chain(descs, [-1, -1, 0, 0, 2, 2], wsb, prec)
    torch.cuda.synchronize()
    assert int(wsb[1]) == 0 and int(wsb.abs().sum()) == 0

    def dswish(a):
        s = torch.sigmoid(a.double())
        return s * (1 + a.double() * (1 - s))
    r_dA1 = (dy.double() @ ws[2].double()) * dswish(acts_a[1])
    r_dA0 = (r_dA1 @ ws[1].double()) * dswish(acts_a[0])
    tol = TOL[prec]
    assert _rel(dA[1], r_dA1) < tol
    assert _rel(dA[0], r_dA0) < 2 * tol
    assert _rel(dX, r_dA0 @ ws[0].double()) < 3 * tol
    assert _rel(dW[2], dy.double().t() @ acts_h[2].double()) < tol
    assert _rel(dW[1], r_dA1.t() @ acts_h[1].double()) < 2 * tol
    assert _rel(dW[0], r_dA0.t() @ acts_h[0].double()) < 3 * tol
    assert _rel(db[1], r_dA1.sum(0)) < 5 * tol
    assert _rel(db[0], r_dA0.sum(0)) < 5 * tol


def test_chain_argument_errors(ops):
    from multimodal_vae_public_b200._lib import MvaeError
    x = torch.randn(256, 64, device="cuda"); w = torch.randn(128, 64, device="cuda")
    y = torch.empty(256, 128, device="cuda"); w2 = torch.randn(32, 128, device="cuda"); y2 = torch.empty(256, 32, device="cuda")
    wsb = ops.chain_workspace("cuda")
    d0 = ops.gemm_desc(x, w, y, 256, 128, 64)
    with pytest.raises(MvaeError):      # dependency must point backwards
        ops.gemm_chain([d0, ops.gemm_desc(y, w2, y2, 256, 32, 128)], [1, -1], wsb, 1)
    with pytest.raises(MvaeError):      # A is not the producer's output
        ops.gemm_chain([d0, ops.gemm_desc(x, w, y, 256, 128, 64)], [-1, 0], wsb, 1)
    with pytest.raises(MvaeError):      # slice not aligned to a row block
        ops.gemm_chain([d0, ops.gemm_desc(y[64:], w2, y2, 192, 32, 128)], [-1, 0], wsb, 1)
    with pytest.raises(MvaeError):      # workspace too small
        ops.gemm_chain([d0], [-1], torch.zeros(2, dtype=torch.int32, device="cuda"), 1)
    # an aligned row slice of the producer is fine
    ops.gemm_chain([d0, ops.gemm_desc(y[128:], w2, y2[128:], 128, 32, 128)], [-1, 0], wsb, 1)
    torch.cuda.synchronize()
    assert _rel(y2[128:], (x.double() @ w.double().t())[128:] @ w2.double().t()) < 2e-5
    assert int(wsb.abs().sum()) == 0


@pytest.mark.parametrize("B", [32, 512])
def test_trainer_chain_matches_layerwise(B):
    """The chained MNIST step computes the same tiles as the layer-per-launch step: forward outputs identical, gradients
    equal up to the order of the fp32 atomics (split-K / bias sums)."""
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    rs = np.random.RandomState(3)
    image = torch.from_numpy(rs.uniform(0, 1, (B, 784)).astype(np.float32)).cuda()
    text = torch.from_numpy(rs.randint(0, 10, B).astype(np.int64)).cuda()
    noise = torch.from_numpy(rs.standard_normal((3, B, 64)).astype(np.float32)).cuda()
    out = {}
    for chain in (False, True):
        tr = MnistMVAETrainer(64, B, use_graph=False, chain=chain, seed=5)
        loss = tr.step(image, text, annealing_factor=0.7, noise=noise, update=False)
        out[chain] = (loss, tr.logit_t.clone(), tr.enc_i.clone(), tr.flat_grads.clone(), tr.launches_per_step)
        assert int(tr.chain_ws[1]) == 0
    assert out[True][4] < out[False][4] - 8           # fewer launches
    assert torch.equal(out[True][2], out[False][2])   # encoder outputs: same tiles, same bits
    assert abs(out[True][0] - out[False][0]) <= 1e-6 * abs(out[False][0])
    g0, g1 = out[False][3].double(), out[True][3].double()
    assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item()
    # graph replay of the chained step: workspace is self-resetting, so the captured graph can be replayed
    tr = MnistMVAETrainer(64, B, use_graph=True, chain=True, seed=5)
    losses = [tr.step(image, text, annealing_factor=0.7, noise=noise, update=False) for _ in range(4)]
    assert max(losses) - min(losses) <= 1e-6 * abs(losses[0])
    assert abs(losses[0] - out[False][0]) <= 1e-5 * abs(out[False][0])
    assert int(tr.chain_ws[1]) == 0
