"""GPU tests of the int8 tensor-core contraction engine (csrc/ozaki.cuh: fp64 operands split exactly into 7 x 8-bit slices,
tcgen05.mma kind::i8 with int32 accumulation in TMEM, fp64 recombination) against float64 references."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from onoffgpf import _cabi, synth


def _gemm(ctx, engine, In, W, tri, stats=True):
    rows, K = In.shape
    cols = W.shape[0]
    Out = torch.empty((rows, cols), dtype=torch.float64, device=In.device)
    st = torch.zeros((cols // 64, rows), dtype=torch.float64, device=In.device) if stats else None
    ctx.call("zigp_debug_gemm", engine, In.data_ptr(), K, W.data_ptr(), K, Out.data_ptr(), cols, rows, cols, K, tri,
             st.data_ptr() if stats else None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return Out, st


@pytest.mark.parametrize("rows,cols,K", [(128, 64, 128), (384, 192, 256), (1024, 1024, 1024), (512, 128, 4096)])
def test_int8_engine_matches_float64_matmul(rows, cols, K):
    """error relative to |In| |W|^T: the slicing keeps 56 bits below each row's largest entry, so the bound is that of an fp64
    dot product scaled by (row max / typical entry); operands with a 1e3 dynamic range inside a row stay below 1e-12"""
    dev = torch.device("cuda:0")
    ctx = _cabi.Context(0, 64, 64, 1, 1e-6, 512)
    g = torch.Generator().manual_seed(rows + K)
    for tri in (0, 1, 2):
        if tri and cols != K:
            continue
        In = (torch.randn(rows, K, generator=g, dtype=torch.float64) * torch.exp(3 * torch.randn(rows, 1, generator=g, dtype=torch.float64))).to(dev)
        W = torch.randn(cols, K, generator=g, dtype=torch.float64) * torch.exp(torch.randn(cols, K, generator=g, dtype=torch.float64))
        W = (torch.tril(W) if tri == 1 else torch.triu(W) if tri == 2 else W).to(dev)
        ref = In @ W.T
        bound = In.abs() @ W.abs().T
        out, st = _gemm(ctx, 1, In, W, tri)
        assert ((out - ref).abs() / bound.clamp_min(1e-300)).max().item() < 1e-12
        sref = (ref * ref).reshape(rows, cols // 64, 64).sum(-1).T
        assert ((st - sref).abs() / sref.clamp_min(1e-300)).max().item() < 1e-11
        # exact zeros, exact powers of two and a zero row survive the slicing exactly
    In = torch.zeros(128, 128, dtype=torch.float64, device=dev); In[3, 5] = 2.0 ** -40; In[7] = 1.0
    W = torch.eye(128, dtype=torch.float64, device=dev)[:64].contiguous()
    out, _ = _gemm(ctx, 1, In, W, 0, stats=False)
    assert torch.equal(out, In[:, :64])


@pytest.mark.parametrize("rows,cols,K,ns", [(128, 64, 64, 7), (256, 128, 192, 7), (1024, 1024, 4096, 7), (384, 1024, 1024, 4)])
def test_mn_major_contraction_matches_float64_matmul(rows, cols, K, ns):
    """the rank update's operand form: planes [p][k][row] (the product's row index contiguous, k strided), both operands sliced under one
    power-of-two scale -- Out = InT^T WT against fp64, error relative to K * bound_in * bound_w (the absolute error model of a fixed scale)"""
    dev = torch.device("cuda:0")
    ctx = _cabi.Context(0, 64, 64, 1, 1e-6, 512)
    if ns == 4:
        ctx.set_precision("fp32")
    g = torch.Generator().manual_seed(rows + K + ns)
    InT = torch.randn(K, rows, generator=g, dtype=torch.float64).to(dev)
    WT = (torch.randn(K, cols, generator=g, dtype=torch.float64) * torch.exp(torch.randn(K, cols, generator=g, dtype=torch.float64))).to(dev)
    Out = torch.full((rows, cols), float("nan"), dtype=torch.float64, device=dev)
    bi, bw = float(InT.abs().max()), float(WT.abs().max())
    ctx.call("zigp_debug_gemm_mn", InT.data_ptr(), rows, WT.data_ptr(), cols, Out.data_ptr(), cols, rows, cols, K, bi, bw, 0, torch.cuda.current_stream().cuda_stream)
    ref = InT.T @ WT
    err = float((Out - ref).abs().max()) / (K * bi * bw)
    assert err < (1e-15 if ns == 7 else 2e-8), err
    # exactness on integers: small integer operands survive slicing exactly, so the product is exact
    A = torch.randint(-7, 8, (K, rows), generator=g).to(torch.float64).to(dev); B = torch.randint(-7, 8, (K, cols), generator=g).to(torch.float64).to(dev)
    ctx.call("zigp_debug_gemm_mn", A.data_ptr(), rows, B.data_ptr(), cols, Out.data_ptr(), cols, rows, cols, K, 8.0, 8.0, 0, torch.cuda.current_stream().cuda_stream)
    assert torch.equal(Out, A.T @ B)


def test_planes_pipeline_equals_the_r1_pipeline(monkeypatch):
    """ZIGP_OZ_PIPE=1 (default: A only as digit planes, gradient kernel over the planes, gv * A sliced point-major under one scale, MN-major rank
    update) against ZIGP_OZ_PIPE=0 (fp64 A, transposed slicing with per-column scales): same ELBO to 1e-13 (the forward contractions are the
    same launches) and the packed gradient to 1e-10 of its norm, on a ragged two-chunk problem with far-away points and an outlier row that
    makes max |gv| 1e3 x the typical value"""
    N, M, D = 70_001, 192, 2
    X, Y = synth.make_data(N, D, 21)
    X[:500] += 3.0
    Y[1000] = 40.0
    p = synth.make_params(M, M, D, 21)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ZIGP_OZ_PIPE", mode)
        m = synth.build_model(X, Y, p, engine="int8", chunk_rows=65536)
        m.build_likelihood()
        outs[mode] = m._last_out.cpu().numpy()
        del m
    a, b = outs["1"], outs["0"]
    assert np.all(np.isfinite(a)) and a[3] == 0
    assert abs(a[0] - b[0]) <= 1e-13 * abs(b[0])
    ga, gb = synth.unpack_grad(a[4:], M, M, D), synth.unpack_grad(b[4:], M, M, D)
    full = np.linalg.norm(b[4:])
    for k in gb:
        assert np.linalg.norm(ga[k] - gb[k]) <= 1e-10 * np.linalg.norm(gb[k]) + 1e-13 * full, (k, np.linalg.norm(ga[k] - gb[k]) / np.linalg.norm(gb[k]))


@pytest.mark.parametrize("M,chunk", [(192, 65536), (512, 65536), (1024, 65536), (320, 147456)])
def test_balanced_gradient_launch_equals_the_plain_tiling(monkeypatch, M, chunk):
    """ZIGP_KG_BALANCED=1 (default: kgrad_planes_kernel as ONE wave of CTAs, the chunk's rows dealt evenly to (resident CTAs) / (column groups)
    row groups, sub-tiles of 128 rows inside a CTA) against =0 (128-row blocks): only the order of the per-column sums changes.  Same ELBO bit
    for bit (forward path untouched), packed gradient to 1e-12 of its norm.  Ragged problems: a full chunk plus a last chunk that is above the
    small-chunk limit (16384 rows) but far from full, so the balanced row count is not a multiple of anything; 1, 2, 4 and 2 column groups."""
    N, D = chunk + 23_457, 3
    X, Y = synth.make_data(N, D, 33)
    X[:300] += 2.5
    p = synth.make_params(M, M, D, 33)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ZIGP_KG_BALANCED", mode)
        m = synth.build_model(X, Y, p, engine="int8", chunk_rows=chunk)
        m.build_likelihood()
        outs[mode] = m._last_out.cpu().numpy()
        del m
    a, b = outs["1"], outs["0"]
    assert np.all(np.isfinite(a)) and a[3] == 0
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    ga, gb = synth.unpack_grad(a[4:], M, M, D), synth.unpack_grad(b[4:], M, M, D)
    full = np.linalg.norm(b[4:])
    for k in gb:
        assert np.linalg.norm(ga[k] - gb[k]) <= 1e-12 * np.linalg.norm(gb[k]) + 1e-14 * full, (k, np.linalg.norm(ga[k] - gb[k]) / np.linalg.norm(gb[k]))
    # the q_sqrt gradients come from the rank update over the digit planes the kernel writes: those planes must not depend on the launch shape
    for k in ("u_fs_sqrt", "u_gs_sqrt"):
        assert np.array_equal(ga[k], gb[k]), k


def test_stream_schedules_are_bit_identical(monkeypatch):
    """ZIGP_OZ_SCHED=2 (default: chunk-pipelined -- contractions on a high-priority stream, Kmn tiles / likelihood / gradient sums on a second
    stream one chunk ahead / behind, doubled hand-over buffers, event edges) and =1 (same two roles without the doubling) only change WHEN kernels
    run: the packed result must be bit-identical to ZIGP_OZ_SCHED=0 (one stream per latent), on a ragged five-chunk problem, three times in a row
    (a missing dependency shows up as a changed or irreproducible result)"""
    N, M, D = 300_001, 256, 3
    X, Y = synth.make_data(N, D, 23)
    p = synth.make_params(M, M, D, 23)
    outs = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("ZIGP_OZ_SCHED", mode)
        m = synth.build_model(X, Y, p, engine="int8", chunk_rows=65536)
        seq = []
        for _ in range(3):
            m.build_likelihood()
            seq.append(m._last_out.cpu().numpy().copy())
        outs[mode] = seq
        del m
        torch.cuda.empty_cache()
    for mode in ("1", "2"):
        for a in outs[mode]:
            assert np.array_equal(a, outs["0"][0]), mode
    # one chunk and two chunks (prologue / epilogue of the pipeline), with graph replay as the mirrors use it for small batches
    for n in (3000, 70_000):
        res = {}
        for mode in ("0", "2"):
            monkeypatch.setenv("ZIGP_OZ_SCHED", mode)
            m = synth.build_model(X[:n], Y[:n], p, engine="int8", chunk_rows=65536)
            m.build_likelihood(); m.build_likelihood()
            res[mode] = m._last_out.cpu().numpy().copy()
            del m
        assert np.array_equal(res["0"], res["2"]), n


def test_engines_agree_at_scale():
    """the two engines on the same 200 000-point problem: ELBO to 1e-12, every gradient block to 1e-10 of its norm"""
    N, M, D = 200_000, 256, 3
    X, Y = synth.make_data(N, D, 10)
    p = synth.make_params(M, M, D, 10)
    outs = {}
    for eng in ("dmma", "int8"):
        m = synth.build_model(X, Y, p, engine=eng)
        m.build_likelihood()
        outs[eng] = m._last_out.cpu().numpy()
    a, b = outs["int8"], outs["dmma"]
    assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and abs(a[1] - b[1]) <= 1e-12 * abs(b[1])
    ga, gb = synth.unpack_grad(a[4:], M, M, D), synth.unpack_grad(b[4:], M, M, D)
    full = np.linalg.norm(b[4:])
    for k in gb:
        assert np.linalg.norm(ga[k] - gb[k]) <= 1e-10 * np.linalg.norm(gb[k]) + 1e-13 * full, k


@pytest.mark.parametrize("name", ["cfg3_small", "cfg4_small", "cfg2_M512", "cfg3_M1024"])
def test_reduced_precision_mode_within_1e_4(name):
    """BASELINE.json's fp32 mode (configs[2]: "fp32 and fp64"): contractions with 4 x 8-bit slices (31 bits below each row's scale),
    everything else fp64; tolerance rel 1e-4 on the ELBO and on every gradient block (norm-wise), against the float64 goldens produced by the reference's own source
    (cfg2_M512 is BASELINE.json configs[2] -- "fp32 and fp64" -- at its own M = 512, D = 2)"""
    import os
    from oracle import onoff_oracle as orc
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_%s.npz" % name))
    N, Mf, Mg, D, seed, nd = (int(z[k]) for k in ("N", "Mf", "Mg", "D", "seed", "num_data"))
    X, Y = synth.make_data(N, D, seed)
    p = {k: z["p_" + k] for k in orc.PARAM_NAMES}
    m = synth.build_model(X, Y, p, engine="int8", precision="fp32")
    m.num_data = nd
    assert m._ctx.int8_config() == (4, 8, 10)
    m.build_likelihood()
    out = m._last_out.cpu().numpy()
    assert abs(out[0] - float(z["elbo"])) <= 1e-4 * abs(float(z["elbo"]))
    g = synth.unpack_grad(out[4:], Mf, Mg, D)
    full = np.sqrt(sum(float(np.sum(np.square(z["g_" + k]))) for k in orc.PARAM_NAMES))
    worst = 0.0
    for k in orc.PARAM_NAMES:
        a, b = np.asarray(g[k]).reshape(-1), z["g_" + k].reshape(-1)
        err = np.linalg.norm(a - b)
        assert err <= 1e-4 * np.linalg.norm(b) + 1e-7 * full, (k, err / np.linalg.norm(b))
        worst = max(worst, err / (np.linalg.norm(b) + 1e-300))
    assert abs(out[0] - float(z["elbo"])) > 0          # it IS the reduced mode
    gall = np.concatenate([np.asarray(g[k]).reshape(-1) for k in orc.PARAM_NAMES]); rall = np.concatenate([z["g_" + k].reshape(-1) for k in orc.PARAM_NAMES])
    whole = np.linalg.norm(gall - rall) / np.linalg.norm(rall)
    assert whole <= 1e-4, whole                        # the packed gradient as a whole; single small blocks may sit a few 1e-4 of their own norm
    print("fp32-class mode, %s: packed gradient rel err %.2e" % (name, whole))
    print("fp32-class mode, %s: ELBO rel err %.2e, worst gradient block rel err %.2e" % (name, abs(out[0] - float(z["elbo"])) / abs(float(z["elbo"])), worst))


def test_reduced_precision_mode_where_the_1e_4_claim_ends():
    """a 31-bit operand representation cannot hold 1e-4 once cond(Kmm) approaches 1e6-1e7.  On the ill-conditioned fixture
    (cond 4.6e6 / 7.9e6) the ELBO still meets 1e-4 but the gradient does not have to: the test states the measured class (<= 5e-2 norm-wise)
    so that a regression OR a silent widening of the claim shows up; the fp64 mode on the same fixture is test_golden_ill_conditioned."""
    import os
    from oracle import onoff_oracle as orc
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_illcond.npz"))
    N, Mf, Mg, D, seed, nd = (int(z[k]) for k in ("N", "Mf", "Mg", "D", "seed", "num_data"))
    X, Y = synth.make_data(N, D, seed)
    p = {k: z["p_" + k] for k in orc.PARAM_NAMES}
    m = synth.build_model(X, Y, p, engine="int8", precision="fp32")
    m.num_data = nd
    m.build_likelihood()
    out = m._last_out.cpu().numpy()
    assert out[3] == 0 and np.all(np.isfinite(out))
    e_err = abs(out[0] - float(z["elbo"])) / abs(float(z["elbo"]))
    gref = np.concatenate([z["g_" + k].reshape(-1) for k in orc.PARAM_NAMES])
    g = synth.unpack_grad(out[4:], Mf, Mg, D)
    gerr = np.linalg.norm(np.concatenate([np.asarray(g[k]).reshape(-1) for k in orc.PARAM_NAMES]) - gref) / np.linalg.norm(gref)
    print("fp32-class mode at cond ~6e6: ELBO rel err %.2e, gradient rel err %.2e" % (e_err, gerr))
    assert e_err <= 1e-4 and gerr <= 5e-2


def test_rank_update_split_guard_on_a_short_last_chunk():
    """chunk of 147 456 rows -> 9 k-splits of the rank update (int32 exactness); the ragged last chunk has only 8 k blocks"""
    N, M, D = 147456 + 100, 16, 2
    X, Y = synth.make_data(N, D, 5)
    p = synth.make_params(M, M, D, 5)
    outs = {}
    for eng in ("dmma", "int8"):
        m = synth.build_model(X, Y, p, engine=eng, chunk_rows=147456)
        m.build_likelihood()
        outs[eng] = m._last_out.cpu().numpy()
    a, b = outs["int8"], outs["dmma"]
    assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0])
    assert np.linalg.norm(a[4:] - b[4:]) <= 1e-10 * np.linalg.norm(b[4:])


def test_reslicing_modes_agree(monkeypatch):
    """ZIGP_OZ_FUSE = 0 (A re-sliced by separate kernels, per-row maxima), 1 (digit planes of A from the A product's epilogue, fixed
    scale from |A_nm| <= sqrt(var); the default) and 2 (also the transposed planes for the rank update): same ELBO to 1e-12 and the same
    packed gradient to 1e-10 of its norm, on a ragged two-chunk problem whose far-away points exercise rows with tiny |A|"""
    N, M, D = 70_001, 192, 2
    X, Y = synth.make_data(N, D, 21)
    X[:500] += 3.0                                   # points far from every inducing point: K and A rows ~ 1e-200
    p = synth.make_params(M, M, D, 21)
    outs = {}
    monkeypatch.setenv("ZIGP_OZ_PIPE", "0")           # the fuse modes belong to the r1 pipeline (the planes pipeline always slices in the epilogue)
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("ZIGP_OZ_FUSE", mode)
        m = synth.build_model(X, Y, p, engine="int8", chunk_rows=65536)
        m.build_likelihood()
        outs[mode] = m._last_out.cpu().numpy()
        del m
    ref = outs["0"]
    for mode in ("1", "2"):
        a = outs[mode]
        assert np.all(np.isfinite(a))
        assert abs(a[0] - ref[0]) <= 1e-12 * abs(ref[0])
        assert np.linalg.norm(a[2:] - ref[2:]) <= 1e-10 * np.linalg.norm(ref[2:])


@pytest.mark.parametrize("M,D", [(1536, 3), (2048, 5)])
def test_engines_agree_for_larger_inducing_sets(M, D):
    """beyond the benchmark's M = 1024: 24 and 32 column tiles, split N > 256 instructions, D = 5 kernel tiles"""
    N = 40_000
    X, Y = synth.make_data(N, D, 40 + D)
    p = synth.make_params(M, M, D, 40 + D)
    outs = {}
    for eng in ("dmma", "int8"):
        m = synth.build_model(X, Y, p, engine=eng)
        m.build_likelihood()
        outs[eng] = m._last_out.cpu().numpy()
        del m
        torch.cuda.empty_cache()
    a, b = outs["int8"], outs["dmma"]
    assert a[3] == 0 and b[3] == 0
    assert abs(a[0] - b[0]) <= 1e-11 * abs(b[0]) and abs(a[1] - b[1]) <= 1e-11 * abs(b[1])
    assert np.linalg.norm(a[4:] - b[4:]) <= 1e-9 * np.linalg.norm(b[4:])


def test_mixed_mode_keeps_the_elbo_and_bounds_the_gradient_error():
    """precision='fp64-grad6': the ELBO-path contractions stay at 7 slices (ELBO, data term and KL bit-identical to the fp64 mode), the
    gradient-only contractions use the top 6 planes: the packed gradient stays within 1e-11 of the fp64 mode's on a well-conditioned problem"""
    N, M, D = 70_001, 256, 3
    X, Y = synth.make_data(N, D, 22)
    p = synth.make_params(M, M, D, 22)
    outs = {}
    for prec in ("fp64", "fp64-grad6"):
        m = synth.build_model(X, Y, p, engine="int8", precision=prec)
        m.build_likelihood()
        outs[prec] = m._last_out.cpu().numpy()
        del m
    a, b = outs["fp64-grad6"], outs["fp64"]
    assert np.array_equal(a[:4], b[:4])
    err = np.linalg.norm(a[4:] - b[4:]) / np.linalg.norm(b[4:])
    assert 0 < err <= 1e-11, err
