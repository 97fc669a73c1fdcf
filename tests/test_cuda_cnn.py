"""GPU parity: CNN forward through the C-ABI vs the fp32 CPU restatement (oracle/cnn_oracle.py) with the
released weights.  Tolerance 1e-4 on output probabilities (BASELINE.json north_star)."""
import numpy as np
import pytest

from tests.golden_util import golden_chunk, load_case

pytestmark = pytest.mark.gpu
TOL = 1e-4
IMPLS = [1, 0]


def _ctx():
    from nanocaller_b200.host import snp_pileups
    return snp_pileups.context(0)


def _golden_x(name, limit=None):
    rs, dct, chunks, bed, g = load_case(name)
    xs, refs, dps, depths = [], [], [], []
    for ci in range(len(chunks)):
        w = golden_chunk(g, ci)
        if len(w["pos"]):
            xs.append(w["mat"].astype(np.float32)); refs.append(w["ref"].astype(np.float32))
            dps.append(w["dp"]); depths.append(np.full(len(w["pos"]), float(w["depth"])))
    x, ref = np.concatenate(xs), np.concatenate(refs)
    return x[:limit], ref[:limit], np.concatenate(dps)[:limit], np.concatenate(depths)[:limit]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("model", ["ONT-HG002", "CCS-HG002", "NanoCaller1"])
def test_snp_model_forward_matches_oracle(model, impl):
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle, snp_oracle
    tensors, meta = W.load_model("snp", model)
    ctx = _ctx()
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
    x, ref, dp, depth = _golden_x("ont_diploid", 1500)
    x = snp_oracle.scale_counts(x, meta["train_coverage"] or 30.0, coverage=float(depth[0]))
    want = np.concatenate(cnn_oracle.snp_model(tensors, x, ref), 1)           # [n,10]
    got = ctx.snp_model_forward(x, ref, haploid=False, impl=impl)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 127, 128, 129, 443, 444, 445, 1333])
def test_tensor_core_path_ragged_batches(n):
    """Batch sizes around every tiling boundary of the tensor-core path: 3 sites per conv3 group, 128 sites per fc1 tile,
    444 sites per wave of the conv1/conv2 kernel (148 SMs x 3 warpgroups), a few groups per TB ring."""
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle, snp_oracle
    tensors, meta = W.load_model("snp", "ONT-HG002")
    ctx = _ctx()
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
    x, ref, dp, depth = _golden_x("ont_diploid", n)
    assert len(x) == n
    x = snp_oracle.scale_counts(x, meta["train_coverage"], coverage=float(depth[0]))
    want = np.concatenate(cnn_oracle.snp_model(tensors, x, ref), 1)
    got = ctx.snp_model_forward(x, ref, haploid=False, impl=0)
    assert got.shape == want.shape and np.abs(got - want).max() < TOL, np.abs(got - want).max()


@pytest.mark.parametrize("impl", IMPLS)
def test_haploid_snp_model_forward_matches_oracle(impl):
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle, snp_oracle
    tensors, meta = W.load_model("snp", "haploid")
    ctx = _ctx()
    ctx.load_snp_weights(W.pack_snp_blob(tensors, True), 30.0, True)
    x, ref, dp, depth = _golden_x("haploid", 1500)
    x = snp_oracle.scale_counts(x, 30.0, coverage=float(depth[0]))
    want = cnn_oracle.haploid_snp_model(tensors, x, ref)
    got = ctx.snp_model_forward(x, ref, haploid=True, impl=impl)
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("normalize", [True, False])
def test_fused_scan_forward_matches_oracle(normalize, impl):
    """nc_snp_scan + nc_snp_forward (coverage scaling fused into the first layer's operand load, tensors
    never leave the device) vs oracle tensors -> scale_counts -> fp32 CNN."""
    from nanocaller_b200.host import snp_pileups, weights as W
    from oracle import cnn_oracle, snp_oracle
    rs, dct, chunks, bed, g = load_case("ont_diploid")
    tensors, meta = W.load_model("snp", "ONT-HG002")
    ctx = _ctx()
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
    snp_pileups.scan_chunks(ctx, rs, dct, chunks, "diploid")
    got = ctx.snp_forward(normalize=normalize, impl=impl)
    want = []
    for ci in range(len(chunks)):
        w = golden_chunk(g, ci)
        if normalize:
            x = snp_oracle.scale_counts(w["mat"], meta["train_coverage"], coverage=float(w["depth"]))
        else:
            x = snp_oracle.scale_counts(w["mat"], meta["train_coverage"], dp=w["dp"])
        want.append(cnn_oracle.snp_probs(tensors, x, w["ref"].astype(np.float32)))
    want = np.concatenate(want)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


@pytest.mark.parametrize("impl", IMPLS)
def test_fused_haploid_forward_matches_oracle(impl):
    from nanocaller_b200.host import snp_pileups, weights as W
    from oracle import cnn_oracle, snp_oracle
    rs, dct, chunks, bed, g = load_case("haploid")
    tensors, meta = W.load_model("snp", "haploid")
    ctx = _ctx()
    ctx.load_snp_weights(W.pack_snp_blob(tensors, True), 30.0, True)       # hap_train_coverage, snpCaller.py:73
    snp_pileups.scan_chunks(ctx, rs, dct, chunks, "haploid")
    got = ctx.snp_forward(normalize=True, impl=impl)
    want = []
    for ci in range(len(chunks)):
        w = golden_chunk(g, ci)
        x = snp_oracle.scale_counts(w["mat"], 30.0, coverage=float(w["depth"]))
        want.append(cnn_oracle.haploid_snp_model(tensors, x, w["ref"].astype(np.float32)))
    want = np.concatenate(want)
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


def _indel_like(n, rows, seed):
    """Tensors shaped like generate_indel_pileups.msa output: chan0 = column frequency - ref one-hot,
    chan1 = ref one-hot (rows A,G,T,C,-), 128 columns with a zero-padded tail."""
    rng = np.random.RandomState(seed)
    x = np.zeros((n, rows, 128, 2), np.float32)
    for b in range(rows // 5):
        width = rng.randint(60, 129, n)
        freq = rng.dirichlet([0.3] * 5, size=(n, 128)).astype(np.float32)       # [n,128,5]
        refi = rng.randint(0, 5, (n, 128))
        onehot = np.eye(5, dtype=np.float32)[refi]
        live = (np.arange(128)[None, :] < width[:, None])[:, :, None]
        x[:, 5 * b:5 * b + 5, :, 0] = np.transpose((freq - onehot) * live, (0, 2, 1))
        x[:, 5 * b:5 * b + 5, :, 1] = np.transpose(onehot * live, (0, 2, 1))
    return x


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("model", ["ONT-HG002", "CCS-HG002"])
def test_indel_model_forward_matches_oracle(model, impl):
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle
    tensors, _ = W.load_model("indel", model)
    ctx = _ctx()
    ctx.load_indel_weights(W.pack_indel_blob(tensors), False)
    x = _indel_like(300, 15, 3)
    want = cnn_oracle.indel_model(tensors, x)
    got = ctx.indel_model_forward(x, haploid=False, impl=impl)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


@pytest.mark.parametrize("impl", IMPLS)
def test_haploid_indel_model_forward_matches_oracle(impl):
    from nanocaller_b200.host import weights as W
    from oracle import cnn_oracle
    tensors, _ = W.load_model("indel", "haploid")
    ctx = _ctx()
    ctx.load_indel_weights(W.pack_indel_blob(tensors), True)
    x = _indel_like(300, 5, 4)
    want = cnn_oracle.haploid_indel_model(tensors, x)
    got = ctx.indel_model_forward(x, haploid=True, impl=impl)
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()


def test_forward_requires_weights_and_scan():
    from nanocaller_b200.host import capi
    ctx = capi.Context(0)
    with pytest.raises(capi.NcError) as e:
        ctx.snp_model_forward(np.zeros((1, 5, 41, 5), np.float32), np.zeros((1, 4), np.float32))
    assert e.value.code == capi.NC_ESTATE
    with pytest.raises(capi.NcError):
        ctx.load_snp_weights(np.zeros(10, np.float32), 30.0, False)
    ctx.close()
