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


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M,N,K,split", [(512, 512, 6272, 9), (1024, 512, 784, 4), (300, 128, 512, 4), (128, 64, 2048, 16),
                                         (1000, 784, 512, 2), (8, 512, 512, 4)])
def test_fused_split_k_forward_stack(ops, prec, M, N, K, split):
    """Fused split-K (last-arriver epilogue): layer 1 = bias + Swish with its k range split over `split` CTAs per output
    tile, layer 2 consumes its out2 inside the same chained launch (dependency counters count ONE publication per tile
    and warp), layer 2 itself split as well with a plain bias epilogue.  Scratch and counters must come back zeroed."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w1 = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5; b1 = torch.randn(N, device="cuda", generator=g)
    w2 = torch.randn(64, N, device="cuda", generator=g) / N ** 0.5; b2 = torch.randn(64, device="cuda", generator=g)
    a1 = torch.full((M, N), float("nan"), device="cuda"); h1 = torch.full((M, N), float("nan"), device="cuda")
    y2 = torch.full((M, 64), float("nan"), device="cuda")
    s1 = torch.zeros(M, N, device="cuda"); s2 = torch.zeros(M, 64, device="cuda")
    wsb = ops.chain_workspace("cuda")
    for rep in range(3):
        ops.gemm_chain([ops.gemm_desc(x, w1, a1, M, N, K, bias=b1, out2=h1, epilogue=ops.EPI_BIAS_SWISH, split_k=split, split_ws=s1),
                        ops.gemm_desc(h1, w2, y2, M, 64, N, bias=b2, split_k=2, split_ws=s2)], [-1, 0], wsb, prec)
        torch.cuda.synchronize()
        assert int(wsb[1]) == 0 and int(wsb.abs().sum()) == 0
        assert float(s1.abs().max()) == 0.0 and float(s2.abs().max()) == 0.0, "scratch not re-zeroed"
    r1 = x.double() @ w1.double().t() + b1.double()
    assert _rel(a1, r1) < TOL[prec]
    rh = r1 * torch.sigmoid(r1)
    assert _rel(h1, rh) < TOL[prec]
    assert _rel(y2, rh @ w2.double().t() + b2.double()) < 2 * TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("M,N,K,split", [(512, 512, 6272, 8), (1024, 64, 512, 4), (200, 512, 128, 2)])
def test_fused_split_k_dgrad_with_dswish_and_colsum(ops, prec, M, N, K, split):
    """dgrad form (B MN-major) with the MUL_DSWISH epilogue and the fused bias-gradient column sums, k range split."""
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    dy = torch.randn(M, K, device="cuda", generator=g); w = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    aux = torch.randn(M, N, device="cuda", generator=g)
    dx = torch.full((M, N), float("nan"), device="cuda"); cs = torch.zeros(N, device="cuda")
    s = torch.zeros(M, N, device="cuda")
    wsb = ops.chain_workspace("cuda")
    ops.gemm_chain([ops.gemm_desc(dy, w, dx, M, N, K, b_mn=True, aux=aux, epilogue=ops.EPI_MUL_DSWISH, colsum=cs,
                                  split_k=split, split_ws=s)], [-1], wsb, prec)
    torch.cuda.synchronize()
    sg = torch.sigmoid(aux.double())
    ref = (dy.double() @ w.double()) * (sg * (1 + aux.double() * (1 - sg)))
    assert _rel(dx, ref) < TOL[prec]
    assert _rel(cs, ref.sum(0)) < TOL[prec] * 4
    assert float(s.abs().max()) == 0.0 and int(wsb.abs().sum()) == 0


def test_fused_split_k_needs_scratch_and_chain(ops):
    from multimodal_vae_public_b200 import _lib
    x = torch.randn(128, 256, device="cuda"); w = torch.randn(64, 256, device="cuda"); b = torch.randn(64, device="cuda")
    y = torch.empty(128, 64, device="cuda")
    with pytest.raises(_lib.MvaeError):      # bias epilogue + split_k without scratch
        ops.gemm_chain([ops.gemm_desc(x, w, y, 128, 64, 256, bias=b, split_k=2)], [-1], ops.chain_workspace("cuda"), 1)
    with pytest.raises(_lib.MvaeError):      # gemm_batch has no counter workspace
        ops.gemm_batch([ops.gemm_desc(x, w, y, 128, 64, 256, bias=b, split_k=2, split_ws=torch.zeros(128, 64, device="cuda"))], 1)


@pytest.mark.parametrize("kind", ["fwd", "dgrad"])
def test_presplit_b_lo_is_bit_identical_to_in_loop_split(ops, kind):
    """3xTF32 with B_lo supplied (mvae_split_lo: weights split once per step, low halves fetched by TMA) computes exactly
    what the in-loop split computes: same lo values, same MMAs."""
    g = torch.Generator(device="cuda").manual_seed(17)
    M, N, K = 1000, 512, 784
    x = torch.randn(M, K, device="cuda", generator=g)
    arena = torch.randn(N * K + 64, device="cuda", generator=g) / K ** 0.5        # a "parameter arena" holding W at offset 64
    lo = torch.empty_like(arena)
    if kind == "fwd":
        w = arena[64:].view(N, K); b = torch.randn(N, device="cuda", generator=g)
        mk = lambda y, h: ops.gemm_desc(x, w, y, M, N, K, bias=b, out2=h, epilogue=ops.EPI_BIAS_SWISH)   # noqa: E731
    else:
        w = arena[64:].view(K, N)                                                  # dx[M,N] = dy[M,K] W[K,N]: B MN-major
        aux = torch.randn(M, N, device="cuda", generator=g)
        mk = lambda y, h: ops.gemm_desc(x, w, y, M, N, K, b_mn=True, aux=aux, epilogue=ops.EPI_MUL_DSWISH)   # noqa: E731
    y0 = torch.empty(M, N, device="cuda"); h0 = torch.empty(M, N, device="cuda")
    d0 = mk(y0, h0)
    assert not d0.B_lo
    ops.gemm_batch([d0], 1)
    ops.split_lo(arena, lo)
    ops.register_lo_arena(arena, lo)
    try:
        y1 = torch.empty(M, N, device="cuda"); h1 = torch.empty(M, N, device="cuda")
        d1 = mk(y1, h1)
        assert d1.B_lo == lo.data_ptr() + 64 * 4
        ops.gemm_batch([d1], 1)
    finally:
        ops.unregister_lo_arena(arena)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)
    if kind == "fwd":
        assert torch.equal(h0, h1)
    ref = arena.double()
    assert torch.equal(lo.double(), ref - (arena.view(torch.int32) & -8192).view(torch.float32).double())
