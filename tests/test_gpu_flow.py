"""GPU parity of the flow forward (csrc/flow.cu through the C ABI) against the float64 oracle.

Tolerance (BASELINE.json north_star): max |err| <= 1e-3 x per-feature std, written below as TOL.
"""
import numpy as np
import pytest

from oracle import nodes as onodes

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _check(gflow, flow, x, std=None, tol=TOL):
    y_ref = onodes.flow_execute(flow, x.astype(np.float64))
    y = gflow.execute(x)
    assert y.dtype == np.float64 and y.shape == y_ref.shape
    if std is None:
        std = y_ref.std(axis=0)
    err = np.abs(y - y_ref) / std
    assert np.isfinite(y).all()
    assert err.max() <= tol, "max err/std %.3g" % err.max()
    return err.max()


@pytest.mark.parametrize("mode", ["auto", "fold", "two_pass"])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 700])
def test_tiny_flow_parity(tiny_flow, mode, n):
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    g = GpuFlow(tiny_flow, igsfa_mode=mode)
    x = synthetic.synthetic_patches(n, (16, 16), 100 + n)
    ref_std = onodes.flow_execute(tiny_flow, synthetic.synthetic_patches(500, (16, 16), 7).astype(np.float64)).std(axis=0)
    for dt in (np.uint8, np.float32, np.float64):
        _check(g, tiny_flow, x.astype(dt), std=ref_std)
    g.close()


def test_empty_and_bad_dim(tiny_flow):
    from pyfaceanalysis_b200 import GpuFlow
    g = GpuFlow(tiny_flow)
    assert g.execute(np.zeros((0, 256))).shape == (0, 16)
    with pytest.raises(ValueError):
        g.execute(np.zeros((3, 255)))
    g.close()


def test_u11l_parity_in_distribution(u11l_flow):
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    g = GpuFlow(u11l_flow)
    x = synthetic.synthetic_patches(600, (64, 64), 11)
    e = _check(g, u11l_flow, x, std=u11l_flow._train_output_std)
    print("U11L_64 in-distribution max err/std", e)
    # feature slice (the caller's sl[:, 0:D]) and float32 output
    y9 = g.execute(x[:50], n_features=9, out_dtype=np.float32)
    yall = g.execute(x[:50])
    assert y9.shape == (50, 9) and np.allclose(y9, yall[:, :9], rtol=1e-6, atol=1e-4)
    g.close()


@pytest.mark.parametrize("engine", ["ffma", "tc"])
def test_u11l_both_engines(u11l_flow, engine, monkeypatch):
    """The packed-FP32 kernel and the tcgen05 3xTF32 kernel both meet the tolerance, and agree with each other."""
    from pyfaceanalysis_b200 import GpuFlow, plan, synthetic
    monkeypatch.setattr(plan, "ENGINE", engine)
    g = GpuFlow(u11l_flow)
    assert {op.engine for op in g.spec.ops} == {engine}
    assert g.fused_front == (engine == "tc")          # the FP32 engine never fuses: it is the exact-FP32 option
    x = synthetic.synthetic_patches(384, (64, 64), 21)
    e = _check(g, u11l_flow, x, std=u11l_flow._train_output_std, tol=(2e-5 if engine == "ffma" else TOL))
    print("U11L_64 engine", engine, "max err/std", e)
    g.profile(True)
    g.execute(x)
    st = g.op_stats()
    assert len(st) == 11
    if engine == "tc":      # uint8 windows: layers 0-2 run as one fused launch whose time is booked on op 0
        assert [s["engine"] for s in st] == ["front"] * 3 + ["f16"] * 8      # layers 3-10: layer_tc_kernel on FP16 pieces
        assert st[0]["ms"] > 0 and st[1]["ms"] == 0 and st[2]["ms"] == 0 and all(s["ms"] > 0 for s in st[3:])
    else:
        assert all(s["engine"] == engine and s["ms"] > 0 for s in st)
    g.close()


def test_fused_front_matches_per_layer_path(u11l_flow, monkeypatch):
    """Layers 0-2 fused in one lane-resident kernel (csrc/front_tc.cuh, FP16-split tcgen05) against the same flow
    run layer by layer (HGSFA_FRONT=0), for row-major and window-minor uint8 input, ragged window counts, the host
    entry point, and float inputs (which keep the per-layer path)."""
    import ctypes as C
    import torch
    from pyfaceanalysis_b200 import GpuFlow, _lib, synthetic
    fused = GpuFlow(u11l_flow)
    monkeypatch.setenv("HGSFA_FRONT", "0")          # read by hgsfa_plan_create
    plain = GpuFlow(u11l_flow)
    monkeypatch.delenv("HGSFA_FRONT")
    assert fused.fused_front and fused.front_reason == ""
    fused.profile(True)
    plain.profile(True)
    std = u11l_flow._train_output_std
    rng = np.random.default_rng(5)
    for n in (1, 127, 129, 1000):
        x = np.concatenate([synthetic.synthetic_patches(n // 2 + 1, (64, 64), 40 + n),
                            rng.integers(0, 256, (n, 4096), dtype=np.uint8)])[:n]
        xt = torch.as_tensor(x, device="cuda")
        y_p = plain.execute_torch(xt).cpu().numpy().astype(np.float64)
        y_f = fused.execute_torch(xt).cpu().numpy().astype(np.float64)
        tiled = torch.zeros((n + 127) // 128 * 128 * 4096, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.load().hgsfa_tile_windows_device(C.c_void_p(xt.data_ptr()), _lib.U8, n, 4096, 4096,
                                                         C.c_void_p(tiled.data_ptr()), _lib.U8, None))
        y_t = fused.execute_torch(tiled, layout=_lib.TILED, n=n).cpu().numpy().astype(np.float64)
        assert np.array_equal(y_t, y_f)                     # both input layouts feed the same arithmetic
        sd = np.maximum(std, y_p.std(axis=0)) if n > 1 else std
        assert (np.abs(y_f - y_p) / sd).max() <= TOL, n
        if n <= 129:
            ref = onodes.flow_execute(u11l_flow, x.astype(np.float64))
            assert (np.abs(y_f - ref) / sd).max() <= TOL, n
        # rows whose pitch is not 16-byte aligned cannot go through the tensor map: tiled first, same kernel
        if n > 1:
            wide = torch.zeros((n, 4099), dtype=torch.uint8, device="cuda")
            wide[:, :4096] = xt
            y_w = fused.execute_torch(wide[:, :4096]).cpu().numpy().astype(np.float64)
            assert np.array_equal(y_w, y_f)
    assert [s["engine"] for s in fused.op_stats()][:4] == ["front", "front", "front", "f16"]
    # per-layer path: the first op (unbounded caller input) on 3xTF32, bounded ones on FP16 pieces
    assert [s["engine"] for s in plain.op_stats()][:3] == ["tc", "f16", "f16"] and plain.op_stats()[1]["ms"] > 0
    # host entry point (pieces through the staging buffers) and float input (per-layer path, exact same values as `plain`)
    xh = rng.integers(0, 256, (70000, 4096), dtype=np.uint8)
    y_f, y_p = fused.execute(xh, out_dtype=np.float32), plain.execute(xh, out_dtype=np.float32)
    assert (np.abs(y_f.astype(np.float64) - y_p) / np.maximum(std, y_p.std(axis=0))).max() <= TOL
    xf = xh[:300].astype(np.float32)
    assert np.array_equal(fused.execute(xf), plain.execute(xf))
    fused.close()
    plain.close()


def test_u11l_parity_uniform_noise(u11l_flow):
    """BASELINE config 2 input: uniform integer pixels 0..255 (far outside the fitted distribution; the
    inter-layer saturation keeps the network finite)."""
    from pyfaceanalysis_b200 import GpuFlow
    g = GpuFlow(u11l_flow)
    rng = np.random.default_rng(12345600)
    x = rng.integers(0, 256, (512, 4096), dtype=np.uint8)
    _check(g, u11l_flow, x)
    g.close()


def test_chunking_is_invisible(u11l_flow):
    """Front/back chunk sizes (workspace tiling) must not change results; exercises several chunks."""
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    x = synthetic.synthetic_patches(1500, (64, 64), 21)
    g = GpuFlow(u11l_flow)
    y0 = g.execute(x)
    g.set_chunks(front=512, back=1024)
    y1 = g.execute(x)
    assert np.array_equal(y0, y1)
    g.close()


def test_node_facades(tiny_flow):
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    g = GpuFlow(tiny_flow)
    x = synthetic.synthetic_patches(40, (16, 16), 5).astype(np.float64)
    cur_ref = x
    cur = x
    for k, node in enumerate(g):
        cur_ref = onodes.execute(onodes.flow_nodes(tiny_flow)[k], cur_ref)
        cur = node.execute(cur)
        assert cur.shape == cur_ref.shape
        scale = max(1.0, np.abs(cur_ref).max())
        assert np.abs(cur - cur_ref).max() <= 2e-5 * scale, (k, type(node.node).__name__)
    part = g.execute(x, nodenr=1)
    assert np.abs(part - onodes.flow_execute(tiny_flow, x, nodenr=1)).max() < 1e-2
    g.close()


@pytest.mark.parametrize("n,dim", [(1, 16), (128, 128), (300, 144), (1000, 4096), (77, 200)])
def test_tile_windows_layout(n, dim):
    """row-major (n, dim) uint8 -> TILED [tile][feature][128] (include/hgsfa.h), padding windows zero."""
    import ctypes as C
    import torch
    from pyfaceanalysis_b200 import _lib
    rng = np.random.default_rng(n * 7 + dim)
    x = rng.integers(0, 256, (n, dim), dtype=np.uint8)
    tiles = (n + 127) // 128
    ref = np.zeros((tiles, dim, 128), dtype=np.uint8)
    for t in range(tiles):
        blk = x[t * 128:(t + 1) * 128]
        ref[t, :, :len(blk)] = blk.T
    for ld_pad in (0, 16, 3):                       # aligned fast path, aligned with padding, unaligned fallback
        xp = np.zeros((n, dim + ld_pad), dtype=np.uint8)
        xp[:, :dim] = x
        d_x = torch.as_tensor(xp, device="cuda:0")
        d_y = torch.full((tiles * dim * 128,), 255, dtype=torch.uint8, device="cuda:0")
        _lib.check(_lib.load().hgsfa_tile_windows_device(C.c_void_p(d_x.data_ptr()), _lib.U8, n, dim, dim + ld_pad,
                                                         C.c_void_p(d_y.data_ptr()), _lib.U8, None))
        torch.cuda.synchronize()
        assert np.array_equal(d_y.cpu().numpy().reshape(tiles, dim, 128), ref), (n, dim, ld_pad)


def test_age_like_96x96_flow_and_real_heads(u11l96_flow, classifiers):
    """BASELINE config 4 shape: 96x96 crops (9216 inputs), overlapping 6x6 / stride-3 fields, 3-wide joins,
    receptive fields up to 180 inputs (forces the two-pass iGSFA form), then the REAL shipped age / race /
    gender heads (4x39, 5x2, 5x2) on the first features."""
    from pyfaceanalysis_b200 import GpuFlow, GpuGaussianClassifier, synthetic
    from oracle import gauss as ogauss
    g = GpuFlow(u11l96_flow)
    assert g.input_dim == 9216
    x = synthetic.synthetic_patches(260, (96, 96), 31)
    y = g.execute(x)
    ref = onodes.flow_execute(u11l96_flow, x.astype(np.float64))
    err = np.abs(y - ref) / u11l96_flow._train_output_std
    assert err.max() <= TOL, err.max()
    heads = [c for c in classifiers if "Generalize" in c.name]
    assert sorted((c.input_dim, len(c.p)) for c in heads) == [(4, 39), (5, 2), (5, 2)]
    for clf in heads:
        h = GpuGaussianClassifier(clf)
        D = h.input_dim
        # the synthetic features are not the distribution these heads were fitted on: map them onto the head's
        # class means so that the comparison exercises finite posteriors as well as the NaN regime
        feats = np.asarray(clf.means)[np.arange(len(y)) % len(clf.means)] + y[:, :D] * 0.05
        got = h.regression(feats.astype(np.float32), clf.avg_labels, estimate_std=(D == 4))
        exp = ogauss.regression(clf, feats.astype(np.float32).astype(np.float64), clf.avg_labels, estimate_std=(D == 4))
        if D == 4:
            assert np.allclose(got[0], exp[0], rtol=1e-9, atol=1e-9, equal_nan=True)
            assert np.allclose(got[1], exp[1], rtol=1e-7, atol=1e-7, equal_nan=True)
        else:
            assert np.allclose(got, exp, rtol=1e-9, atol=1e-9, equal_nan=True)
        h.close()
    g.close()


@pytest.mark.parametrize("engine", ["ffma", "tc"])
def test_exotic_expansions_both_engines(engine, monkeypatch):
    """Term kinds the synthetic networks do not use: table products that are not a full triangle, triple
    products, signed powers, |x| -- the generic paths of both kernels."""
    import test_plan as tp
    from pyfaceanalysis_b200 import GpuFlow, expansions as ex, plan
    monkeypatch.setattr(plan, "ENGINE", engine)
    rng = np.random.default_rng(12)
    d, n_nodes, m = 12, 5, 7
    funcs = ["identity", "signed_08expo", "pair_prod_adj2_ex", "s3CT", "abs", "s4QT", "unsigned_06expo"]
    D = ex.expanded_dim(funcs, m)
    conn = rng.permutation(n_nodes * d)
    flow = [tp._sb(conn, n_nodes * d), tp._layer([tp._pca(d, m, rng) for _ in range(n_nodes)]),
            tp._layer([tp._exp(m, funcs) for _ in range(n_nodes)]), tp._layer([tp._sfa(D, 6, rng) for _ in range(n_nodes)])]
    g = GpuFlow(flow, input_dim=n_nodes * d)
    assert g.spec.ops[-1].engine == engine
    x = (rng.standard_normal((333, n_nodes * d)) * 2).astype(np.float32)
    ref = onodes.flow_execute(flow, x.astype(np.float64))
    y = g.execute(x)
    assert np.abs(y - ref).max() <= 1e-3 * max(1.0, np.abs(ref).std()), np.abs(y - ref).max()
    g.close()


def test_attribute_estimator(u11l96_flow, classifiers):
    """estimate_age_race_gender on normalised crops (face_analysis.py:1170-1306): one batched flow + the three REAL
    shipped heads == the oracle evaluated the reference's way (per face)."""
    from pyfaceanalysis_b200 import AttributeEstimator, GpuFlow, GpuGaussianClassifier, synthetic
    from pyfaceanalysis_b200.attributes import map_real_gender_labels_to_strings, map_real_race_labels_to_strings
    from oracle import gauss as ogauss
    heads = {("Age" if "Age" in c.name else "Race" if "Race" in c.name else "Gender"): c
             for c in classifiers if "Generalize" in c.name}
    g = GpuFlow(u11l96_flow)
    est = AttributeEstimator(g, *[GpuGaussianClassifier(heads[k]) for k in ("Age", "Race", "Gender")])
    x = synthetic.synthetic_patches(37, (96, 96), 77)
    age, age_std, race, gender, conf = est.estimate(x)
    sl = g.execute(x)                                  # the heads are compared on the same (GPU) features
    a_ref, s_ref = ogauss.regression(heads["Age"], sl[:, :4], heads["Age"].avg_labels, estimate_std=True)
    assert np.allclose(age, a_ref, rtol=1e-9, atol=1e-9, equal_nan=True)
    assert np.allclose(age_std, s_ref, rtol=1e-7, atol=1e-7, equal_nan=True)
    r_ref = ogauss.regression(heads["Race"], sl[:, :5], heads["Race"].avg_labels)
    g_ref = ogauss.regression(heads["Gender"], sl[:, :5], heads["Gender"].avg_labels)
    ok = ~(np.isnan(r_ref) | np.isnan(g_ref))
    assert np.allclose(conf["race_confidences"], np.abs(r_ref) / 2.0, rtol=1e-9, atol=1e-9, equal_nan=True)
    assert [r for r, k in zip(race, ok) if k] == [r for r, k in zip(map_real_race_labels_to_strings(r_ref), ok) if k]
    assert [r for r, k in zip(gender, ok) if k] == [r for r, k in zip(map_real_gender_labels_to_strings(g_ref), ok) if k]
    assert len(est.estimate(np.zeros((0, 9216)))[0]) == 0
    g.close()


def test_plan_creation_order_does_not_matter(u11l96_flow, tiny_flow):
    """The dynamic shared-memory limit of the layer kernels is per device, shared by all plans: a small plan created
    after a large one must not shrink it (regression: the large plan's launches failed with 'invalid argument')."""
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    big = GpuFlow(u11l96_flow)
    small = GpuFlow(tiny_flow)
    x = synthetic.synthetic_patches(130, (96, 96), 3).astype(np.float32)
    y = big.execute(x)
    assert np.isfinite(y).all() and small.execute(synthetic.synthetic_patches(5, (16, 16), 1)).shape == (5, 16)
    big.close()
    small.close()


def test_layer_kernel_tf32_pieces_opt_out(u11l_flow, monkeypatch):
    """Default: layer_tc_kernel takes 2-piece FP16 operands (tcgen05 kind::f16, 64-term chunks) for every op whose inputs
    are bounded by the previous op's saturation -- layers 3-10 behind the fused front, layers 1-10 of the per-layer path
    (the first op sees unbounded caller input and stays on 3xTF32).  HGSFA_TC_F16=0: 3xTF32 pieces everywhere (round 1's
    arithmetic).  Both within TOL of the float64 oracle."""
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    x = synthetic.synthetic_patches(300, (64, 64), 33)
    monkeypatch.setenv("HGSFA_TC_F16", "0")
    g0 = GpuFlow(u11l_flow)
    monkeypatch.delenv("HGSFA_TC_F16")
    g1 = GpuFlow(u11l_flow)
    errs = []
    for g, want in ((g0, "tc"), (g1, "f16")):
        errs.append(_check(g, u11l_flow, x, std=u11l_flow._train_output_std))
        g.profile(True)
        g.execute(x)
        assert [s["engine"] for s in g.op_stats()] == ["front"] * 3 + [want] * 8
        xf = x[:130].astype(np.float32)                                    # float input: per-layer ops only
        _check(g, u11l_flow, xf, std=u11l_flow._train_output_std)
        g.close()
    print("U11L_64 max err/std: 3xTF32 layers %.3g, FP16-piece layers %.3g" % tuple(errs))


def test_single_layer_fp16_kernel_opt_in(u11l_flow, monkeypatch):
    """HGSFA_BACK=1: the layers behind the fused front run on the single-layer FP16-split kernel (csrc/back_tc.cuh) --
    kept because it is the more accurate of the two tensor-core layer kernels, off by default because it is slower."""
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    monkeypatch.setenv("HGSFA_BACK", "1")
    g = GpuFlow(u11l_flow)
    monkeypatch.delenv("HGSFA_BACK")
    x = synthetic.synthetic_patches(300, (64, 64), 31)
    e = _check(g, u11l_flow, x, std=u11l_flow._train_output_std)
    g.profile(True)
    g.execute(x)
    assert [s["engine"] for s in g.op_stats()] == ["front"] * 3 + ["f16"] * 8
    print("U11L_64 front + f16 layers max err/std", e)
    xf = x[:130].astype(np.float32)                 # float input: every layer but the first (no input bound) on the FP16 kernel
    _check(g, u11l_flow, xf, std=u11l_flow._train_output_std)
    g.close()


@pytest.mark.parametrize("spec", ["F4L_32", "F4L_32_mid", "F4L_32_wide"])
def test_fused_front_other_widths(spec, monkeypatch):
    """The fused front is instantiated per pair of padded child widths: (8, 16), (16, 16) and (16, 32) on small 32x32
    networks (one to four chunks per join, 16- and 32-column accumulators, subtrees split over gridDim.y), against the
    per-layer path and the float64 oracle."""
    import torch
    from pyfaceanalysis_b200 import GpuFlow, synthetic
    flow = synthetic.make_flow(spec, seed=2)
    fused = GpuFlow(flow)
    assert fused.fused_front, fused.front_reason
    monkeypatch.setenv("HGSFA_FRONT", "0")
    plain = GpuFlow(flow)
    monkeypatch.delenv("HGSFA_FRONT")
    std = flow._train_output_std
    for n in (1, 130, 1500):
        x = synthetic.synthetic_patches(n, (32, 32), 60 + n)
        ref = onodes.flow_execute(flow, x.astype(np.float64))
        xt = torch.as_tensor(x, device="cuda")
        y_f = fused.execute_torch(xt).cpu().numpy().astype(np.float64)
        y_p = plain.execute_torch(xt).cpu().numpy().astype(np.float64)
        assert (np.abs(y_f - ref) / std).max() <= TOL, (spec, n, (np.abs(y_f - ref) / std).max())
        assert (np.abs(y_p - ref) / std).max() <= TOL
        assert np.array_equal(fused.execute(x), fused.execute_torch(xt).cpu().numpy().astype(np.float64))
    fused.profile(True)
    fused.execute(x)
    assert [s["engine"] for s in fused.op_stats()][:3] == ["front"] * 3
    fused.close()
    plain.close()


def test_full_size_batch_properties(u11l_flow):
    """configs[1] at BASELINE.json's full size (1 048 576 windows x 4096 uint8, resident in HBM), through properties that
    do not need the slow oracle at that size:
      * position independence: the batch is 8 192 distinct windows replicated 128 times in a random order -- every copy
        of a window must give bit-identical features whatever its tile, lane, chunk or CTA;
      * shard invariance (shard.window_shard): the features of a tile-aligned slice computed alone are
        bit-identical to the same rows of the full batch;
      * parity of a sample of the distinct windows against the float64 oracle within TOL x std."""
    import torch
    from pyfaceanalysis_b200 import GpuFlow, shard, synthetic
    n_distinct, n = 8192, 1 << 20
    base = torch.from_numpy(np.ascontiguousarray(synthetic.synthetic_patches(n_distinct, (64, 64), 77).astype(np.uint8))).cuda()
    gen = torch.Generator(device="cuda").manual_seed(3)
    src = torch.randperm(n, device="cuda", generator=gen) % n_distinct      # window i is a copy of base[src[i]]
    x = base[src]
    assert x.shape == (n, 4096) and x.dtype == torch.uint8
    g = GpuFlow(u11l_flow)
    assert g.fused_front
    y = g.execute_torch(x).clone()
    assert y.shape == (n, 60) and bool(torch.isfinite(y).all())
    # one representative per distinct window (its first occurrence), then every row against its representative
    first = torch.full((n_distinct,), n, dtype=torch.int64, device="cuda").scatter_reduce(
        0, src, torch.arange(n, device="cuda"), reduce="amin")
    rep = y[first]
    assert bool((y.view(torch.int32) == rep[src].view(torch.int32)).all())
    for rank, world in ((3, 8), (1, 2)):
        a, b = shard.window_shard(n, rank, world)
        part = g.execute_torch(x[a:b]).clone()
        assert bool((part.view(torch.int32) == y[a:b].view(torch.int32)).all())
    sample = np.arange(0, n_distinct, 64)
    ref = onodes.flow_execute(u11l_flow, base[sample].cpu().numpy().astype(np.float64))
    err = np.abs(rep[sample].double().cpu().numpy() - ref) / u11l_flow._train_output_std
    assert err.max() <= TOL, "max err/std %.3g" % err.max()
    g.close()
