"""Parity at BASELINE.json's FULL sizes through size-independent properties.

The CPU oracle cannot finish the full configurations in seconds, so these tests use
properties that hold for the exact result whatever the size: bit-exact comparison with NumPy where
the host can afford one pass (config 1), sampled blocks against a host recomputation plus
associativity (config 2), ``||Q^T Q - I||`` / ``||Q R - X||`` / invariance of R under row-block
permutation (config 3), and stationarity of the Newton fixed point plus agreement of the fused kernel
with the unfused interface path (config 4).  Inputs are generated on the device (seeded) to keep the
suite short; every arithmetic kernel under test is ours.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_blockarray(system, shape, block_shape, fill):
    from nums_b200 import blocks
    from nums_b200.grid import ArrayGrid
    ba = blocks.BlockArray(ArrayGrid(shape, block_shape, "float64"), system)
    for entry in ba.grid.get_entry_iterator():
        ba.blocks[entry].oid = fill(entry, ba.grid.get_block_shape(entry))
    return ba


def test_config1_bop_full_size_bit_exact(cuda_system):
    """u + v and u * v on two 1e8-element float64 arrays in 8 blocks: bit-identical to NumPy."""
    from nums_b200 import blocks
    n = 100_000_000
    u = np.random.default_rng(1).random(n)
    v = np.random.default_rng(2).random(n)
    app = blocks.ArrayApp(cuda_system)
    U, V = app.array(u, (n // 8,)), app.array(v, (n // 8,))
    assert np.array_equal((U + V).get(), u + v)
    assert np.array_equal((U * V).get(), u * v)
    # checksum of checksums: block sums folded on the device == NumPy's block sums folded on the host
    s = app.sum(U + V).get()
    ref = sum(np.sum(u[i * (n // 8):(i + 1) * (n // 8)] + v[i * (n // 8):(i + 1) * (n // 8)]) for i in range(8))
    assert abs(s - ref) <= 1e-12 * abs(ref)


def test_config2_matmul_full_size_properties(cuda_system):
    """16384^2 blocked matmul: sampled C blocks vs a host recomputation, and (A B) x == A (B x)."""
    import torch
    N, bs, g = 16384, 2048, 8
    dev = torch.device("cuda", torch.cuda.current_device())

    def fill(seed):
        def make(entry, shape):
            gen = torch.Generator(device=dev)
            gen.manual_seed(seed * 100 + entry[0] * 8 + entry[1])
            return torch.randn(shape, dtype=torch.float64, device=dev, generator=gen)
        return make
    A = _device_blockarray(cuda_system, (N, N), (bs, bs), fill(3))
    B = _device_blockarray(cuda_system, (N, N), (bs, bs), fill(4))
    C = A @ B
    for (i, j) in [(0, 0), (3, 5), (7, 7)]:
        got = cuda_system.get(C.blocks[i, j].oid)[:256, :256]
        ref = np.zeros((256, 256))
        for k in range(g):
            a = cuda_system.get(A.blocks[i, k].oid)[:256]
            b = cuda_system.get(B.blocks[k, j].oid)[:, :256]
            ref += a @ b
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-10
    x = _device_blockarray(cuda_system, (N,), (bs,), lambda e, s: torch.randn(s, dtype=torch.float64, device=dev))
    left = (C @ x).get()
    right = (A @ (B @ x)).get()
    assert np.linalg.norm(left - right) / np.linalg.norm(right) <= 1e-10


def test_config3_tsqr_full_size_properties(cuda_system):
    """16M x 128 TSQR: Q^T Q = I, Q R = X (checked per block on the device with our own kernels),
    R upper triangular, and R invariant (up to row signs) under a permutation of the row blocks."""
    import torch
    from nums_b200 import blocks
    from nums_b200 import cuda_compute as cc
    from tests.helpers import canon_r, rel_fro
    m, n, G = 16_777_216, 128, 8
    dev = torch.device("cuda", torch.cuda.current_device())

    def make(entry, shape):
        gen = torch.Generator(device=dev)
        gen.manual_seed(500 + entry[0])
        return torch.randn(shape, dtype=torch.float64, device=dev, generator=gen)
    X = _device_blockarray(cuda_system, (m, n), (m // G, n), make)
    app = blocks.ArrayApp(cuda_system)
    Q, R = app.indirect_tsqr(X)
    r = R.get()
    assert np.allclose(np.tril(r, -1), 0.0)
    # Q^T Q accumulated over the row blocks with the skinny / DMMA kernels
    gram = None
    resid2, norm2 = 0.0, 0.0
    r_dev = cuda_system.put(r)
    for i in range(G):
        q = cuda_system.contractions.resolve(Q.blocks[i, 0].oid)
        part = cc.tensordot(cc._transpose_view(q), q, 1)
        gram = part if gram is None else cc.elementwise("add", gram, part)
        back = cc.tensordot(q, r_dev, 1)
        diff = cc.elementwise("subtract", back, X.blocks[i, 0].oid)
        resid2 += float(cuda_system.get(cc.sum_of_squares(diff)))
        norm2 += float(cuda_system.get(cc.sum_of_squares(X.blocks[i, 0].oid)))
        del q, part, back, diff
    gram = cuda_system.get(gram)
    assert np.linalg.norm(gram - np.eye(n)) <= 1e-10
    assert np.sqrt(resid2 / norm2) <= 1e-10
    # permutation invariance of R
    Xp = blocks.BlockArray(X.grid.copy(), cuda_system)
    perm = [3, 0, 7, 1, 6, 2, 5, 4]
    for i in range(G):
        Xp.blocks[i, 0].oid = X.blocks[perm[i], 0].oid
    rp = app.indirect_tsr(Xp).get()
    assert rel_fro(canon_r(rp), canon_r(r)) <= 1e-10


def test_config4_newton_full_size_properties(cuda_system):
    """11M x 28 Newton LR: the fused kernel agrees with the unfused interface path on g and H, Newton
    converges, and the gradient at the returned beta is (numerically) zero."""
    import torch
    from nums_b200 import blocks, multi_gpu
    from nums_b200 import cuda_compute as cc
    N, d, G = 11_000_000, 28, 8
    dev = torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=dev)
    gen.manual_seed(6)
    X = _device_blockarray(cuda_system, (N, d), (N // G, d),
                           lambda e, s: torch.randn(s, dtype=torch.float64, device=dev, generator=gen))
    theta = torch.randn(d, dtype=torch.float64, device=dev, generator=gen) / np.sqrt(d)
    app = blocks.ArrayApp(cuda_system)
    y = blocks.BlockArray(blocks.ArrayGrid((N,), (N // G,), "float64"), cuda_system)
    for i in range(G):
        xb = X.blocks[i, 0].oid
        p = torch.sigmoid(xb @ theta)
        y.blocks[i].oid = (torch.rand(xb.shape[0], dtype=torch.float64, device=dev, generator=gen) < p).to(torch.float64)
    xs = [X.blocks[i, 0].oid for i in range(G)]
    ys = [y.blocks[i].oid for i in range(G)]
    # fused vs interface path at a non-trivial beta
    beta0 = app.array(np.linspace(-0.3, 0.3, d), (d,))
    model = blocks.LogisticRegression(app)
    mu = model.forward(X, beta0)
    g_ref = model.gradient(X, y, mu).get()
    h_ref = model.hessian(X, y, mu).get()
    fused = cuda_system.get(cc.lr_grad_hess_blocks(xs, ys, beta0.blocks[0].oid))
    assert np.linalg.norm(fused[:d] - g_ref) / np.linalg.norm(g_ref) <= 1e-10
    assert np.linalg.norm(fused[d:].reshape(d, d) - h_ref) / np.linalg.norm(h_ref) <= 1e-10
    # Newton to convergence, then stationarity
    beta, iters = multi_gpu.newton_lr(cuda_system, multi_gpu.Comm(), xs, ys, d, 1e-6, 25, cc.lr_grad_hess_blocks)
    assert iters < 25
    # the fused update kernel takes the same path: same iteration count, same beta to rounding
    beta_f, iters_f = multi_gpu.newton_lr(cuda_system, multi_gpu.Comm(), xs, ys, d, 1e-6, 25, cc.lr_grad_hess_blocks,
                                          step=cc.newton_step)
    assert iters_f == iters
    assert np.linalg.norm(cuda_system.get(beta_f) - cuda_system.get(beta)) <= 1e-10 * np.linalg.norm(cuda_system.get(beta))
    final = cuda_system.get(cc.lr_grad_hess_blocks(xs, ys, beta))
    assert np.abs(final[:d]).max() <= 1e-6
    b = cuda_system.get(beta)
    assert np.linalg.norm(b - cuda_system.get(theta)) <= 0.02     # recovers the generating parameters
