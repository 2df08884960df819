"""INTEGRATION.md §1, compiled and run: the UNMODIFIED reference host — feather::Net, ConvLayer, the ncnn loader — with
tests/integration/cuda_booster.cpp standing where src/booster/avx/booster.cpp stood, so that booster::ConvBooster's function
table (booster.h:151-170) points into libfcuda.so.  Its Forward must match the reference's AVX build on the same model and
input: the drop-in claim, exercised from the reference's side of the boundary."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_pair(oracle, cuda):
    if not (oracle.reference_available() and oracle.reference_cuda_available()):
        pytest.skip("oracle/_ref/libfeather_ref[_cuda].so not built (needs /root/reference at build time)")
    return oracle.Reference(), oracle.reference_on_cuda()


@pytest.mark.parametrize("name", ["single_conv", "mini", "mobilenet_v1"])
def test_reference_net_on_libfcuda_matches_its_avx_build(oracle, ref_pair, model_dir, name):
    from feathercnn_b200 import booster
    from feathercnn_b200.tools import modelgen
    avx, gpu = ref_pair
    booster.set_precision(booster.PRECISION_TF32X3)
    m = modelgen.ZOO[name]()
    param, binf = m.save(model_dir / f"{name}_integration")
    a = oracle.ReferenceNet(param, binf, ref=avx)
    g = oracle.ReferenceNet(param, binf, ref=gpu)
    before = booster.fcuda().fcuda_launch_count()
    worst = 0.0
    for i in range(2):
        x = modelgen.synthetic_input(m.shape["data"], i)
        a.forward(x)
        g.forward(x)
        for b in sorted(m.blobs):
            try:
                want = a.extract(b)
            except KeyError:
                continue
            e = rel_err(g.extract(b), want)
            worst = max(worst, e)
            assert e < 1e-3, (name, b, i, e)
    print(f"{name}: reference host on libfcuda, worst blob rel err {worst:.2e}")
    assert worst < 2e-4


def test_reference_conv_booster_table_is_bound_to_libfcuda(oracle, ref_pair):
    """The reference's ConvBooster protocol (SelectAlgo -> GetBufferSize -> Init -> Forward, conv_layer.h:92-172) through the
    shim's layer-level entry point, AVX table vs CUDA table, on BASELINE.json configs[0]."""
    avx, gpu = ref_pair
    rng = np.random.default_rng(0)
    p = oracle.ConvParam.make(64, 64, 56, 56, 3, pad=1, bias=True)
    x = rng.uniform(-0.5, 0.5, (64, 56, 56)).astype(np.float32)
    w = (rng.standard_normal((64, 64, 3, 3)) * np.sqrt(2.0 / 576)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, 64).astype(np.float32)
    assert rel_err(gpu.conv(p, x, w, b), avx.conv(p, x, w, b)) < 2e-4
