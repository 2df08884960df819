"""SURVEY.md §8f rank 4 — layer / algorithm coverage the reference rejects (dilation: conv_layer.h:43-47, partial groups:
avx/booster.cpp:304-308, Eltwise PROD / MAX / coefficients: eltwise_layer.h:57-66, Concat on other axes:
concat_layer.h:50-54), plus the two `group == input_channels` quirks ADVICE r1 flagged.  The reference cannot serve as
the oracle for operations it refuses, so the checker is an independent fp64 convolution (torch on the CPU); the bar is the
same max|d|/max|ref| the parity tests use."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _conv_ref(x, w, b, stride, pad, dilation, groups, relu=False):
    import torch
    y = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(),
                                   None if b is None else torch.from_numpy(b).double(), stride=stride, padding=pad,
                                   dilation=dilation, groups=groups)
    if relu:
        y = y.clamp_min(0)
    return y.numpy()


GROUPED = [
    # oc, ic, h, w, k, stride, pad, group
    (16, 8, 12, 12, 3, 1, 1, 4),      # partial groups, table path (ICg = 2)
    (64, 64, 14, 14, 3, 1, 1, 2),     # ICg = 32: fast path on channel slices
    (48, 32, 9, 11, 1, 1, 0, 4),      # grouped pointwise (flattened H*W addressing)
    (24, 8, 10, 10, 3, 2, 1, 8),      # depthwise with channel multiplier 3 (the reference truncates it to OC = IC)
]


@pytest.mark.parametrize("geom", GROUPED)
@pytest.mark.parametrize("batch", [1, 3])
def test_grouped_convolution(cuda, geom, batch):
    from feathercnn_b200 import booster
    oc, ic, h, w, k, stride, pad, group = geom
    rng = np.random.default_rng(hash(geom) % 2**31)
    x = rng.uniform(-0.5, 0.5, (batch, ic, h, w)).astype(np.float32)
    wt = (rng.standard_normal((oc, ic // group, k, k)) * np.sqrt(2.0 / (ic // group * k * k))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, oc).astype(np.float32)
    p = booster.ConvParam.make(oc, ic, h, w, k, stride=stride, pad=pad, group=group, relu=True)
    p.output_channels = oc  # total channels (AssignOutputDim applies the reference's group == IC rule)
    out, algo = booster.conv_forward(p, cuda.from_numpy(x).cuda(), cuda.from_numpy(wt).cuda(), cuda.from_numpy(b).cuda(),
                                     tuned=True)
    cuda.cuda.synchronize()
    assert algo == booster.SGECONV
    want = _conv_ref(x, wt, b, stride, pad, 1, group, relu=True)
    assert rel_err(out.cpu().numpy(), want) < 2e-4, geom


@pytest.mark.parametrize("geom", [
    (32, 32, 20, 20, 3, 1, 2, 2),     # dilation 2, "same" padding, fast path
    (16, 6, 17, 13, 3, 1, 0, 3),      # dilation 3, no padding, table path
    (8, 16, 15, 15, 3, 2, 2, 2),      # strided + dilated
])
def test_dilated_convolution(cuda, geom):
    from feathercnn_b200 import booster
    oc, ic, h, w, k, stride, pad, dil = geom
    rng = np.random.default_rng(hash(geom) % 2**31)
    x = rng.uniform(-0.5, 0.5, (2, ic, h, w)).astype(np.float32)
    wt = (rng.standard_normal((oc, ic, k, k)) * np.sqrt(2.0 / (ic * k * k))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, oc).astype(np.float32)
    p = booster.ConvParam.make(oc, ic, h, w, k, stride=stride, pad=pad)
    p.output_h = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    p.output_w = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    out, _ = booster.conv_forward(p, cuda.from_numpy(x).cuda(), cuda.from_numpy(wt).cuda(), cuda.from_numpy(b).cuda(),
                                  algo=booster.SGECONV, dilation=dil)
    cuda.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), _conv_ref(x, wt, b, stride, pad, dil, 1)) < 2e-4, geom
    # every algorithm that cannot space its taps refuses (-200) instead of computing the dense convolution
    with pytest.raises(booster.FcudaError) as e:
        booster.conv_forward(p, cuda.from_numpy(x).cuda(), cuda.from_numpy(wt).cuda(), cuda.from_numpy(b).cuda(),
                             algo=booster.IM2COL, dilation=dil)
    assert e.value.code == -200


def test_single_input_channel_is_an_ordinary_convolution(cuda):
    """group == 1, IC == 1 (LeNet conv1): the reference's `group == input_channels` test makes it a 1-output depthwise
    (booster.h:121, avx/booster.cpp:285); the tuned policy keeps all 20 output channels."""
    from feathercnn_b200 import booster
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, (2, 1, 28, 28)).astype(np.float32)
    wt = (rng.standard_normal((20, 1, 5, 5)) * 0.2).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, 20).astype(np.float32)
    p = booster.ConvParam.make(20, 1, 28, 28, 5)
    p.output_channels = 20
    out, algo = booster.conv_forward(p, cuda.from_numpy(x).cuda(), cuda.from_numpy(wt).cuda(), cuda.from_numpy(b).cuda(),
                                     tuned=True)
    cuda.cuda.synchronize()
    assert algo == booster.SGECONV and out.shape == (2, 20, 24, 24)
    assert rel_err(out.cpu().numpy(), _conv_ref(x, wt, b, 1, 0, 1, 1)) < 2e-4


@pytest.mark.parametrize("n", [1000, 4096 + 3])
def test_eltwise_prod_max_coeffs(cuda, n):
    from feathercnn_b200 import booster
    rng = np.random.default_rng(n)
    a = rng.uniform(-1, 1, n).astype(np.float32)
    b = rng.uniform(-1, 1, n).astype(np.float32)
    da, db = cuda.from_numpy(a).cuda(), cuda.from_numpy(b).cuda()
    np.testing.assert_array_equal(booster.eltwise(da, db, 0).cpu().numpy(), a * b)
    np.testing.assert_array_equal(booster.eltwise(da, db, 2).cpu().numpy(), np.maximum(a, b))
    np.testing.assert_allclose(booster.eltwise(da, db, 1, 0.5, -2.0).cpu().numpy(), 0.5 * a - 2.0 * b, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(booster.eltwise(da, db, 1, 0.5, -2.0, relu=True).cpu().numpy(),
                               np.maximum(0.5 * a - 2.0 * b, 0), rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(booster.eltwise(da, db, 1).cpu().numpy(), a + b)


def _write_model(tmp_path, lines, blobs):
    import struct
    param = tmp_path / "ext.param"
    binf = tmp_path / "ext.bin"
    n_blobs = sum(int(l.split()[3]) for l in lines)
    param.write_text("7767517\n%d %d\n%s\n" % (len(lines), n_blobs, "\n".join(lines)))
    raw = bytearray()
    for arr, flagged in blobs:
        if flagged:
            raw += struct.pack("<I", 0)
        raw += np.ascontiguousarray(arr, np.float32).tobytes()
    binf.write_bytes(bytes(raw))
    return str(param), str(binf)


@pytest.mark.parametrize("fusion", [False, True])
def test_net_with_extension_layers(cuda, tmp_path, fusion):
    """feather::Net end to end over the §8f layers: dilated conv -> grouped conv -> Eltwise MAX / PROD / weighted SUM ->
    Concat along width and height, against the same graph evaluated in fp64."""
    from feathercnn_b200.net import Net
    rng = np.random.default_rng(11)
    w1 = (rng.standard_normal((16, 8, 3, 3)) * 0.15).astype(np.float32)
    b1 = rng.uniform(-0.1, 0.1, 16).astype(np.float32)
    w2 = (rng.standard_normal((16, 4, 3, 3)) * 0.2).astype(np.float32)  # 4 groups of 4 -> 4
    lines = [
        "Input data 0 1 data 0=12 1=12 2=8",
        "Convolution dil 1 1 data dil 0=16 1=3 2=2 3=1 4=2 5=1 6=%d" % w1.size,
        "ReLU dil_relu 1 1 dil dil_relu",
        "Split sp 1 3 dil_relu s0 s1 s2",
        "ConvolutionDepthWise grp 1 1 s0 grp 0=16 1=3 3=1 4=1 5=0 6=%d 7=4" % w2.size,
        "Split sp2 1 3 grp g0 g1 g2",
        "Eltwise mx 2 1 s1 g0 mx 0=2",
        "Eltwise pr 2 1 s2 g1 pr 0=0",
        "Eltwise ws 2 1 mx g2 ws 0=1 -23301=2,0.5,-1.5",
        "Concat cw 2 1 ws pr cw 0=2",
        "Split sp3 1 2 cw c0 c1",
        "Concat ch 2 1 c0 c1 ch 0=1",
    ]
    param, binf = _write_model(tmp_path, lines, [(w1, True), (b1, False), (w2, True)])
    x = rng.uniform(-0.5, 0.5, (3, 8, 12, 12)).astype(np.float32)
    net = Net(fusion=fusion)
    net.LoadParam(param)
    net.LoadWeights(binf)
    net.Forward(x)
    d = _conv_ref(x, w1, b1, 1, 2, 2, 1, relu=True)
    g = _conv_ref(d.astype(np.float32), w2, None, 1, 1, 1, 4)
    mx = np.maximum(d, g)
    pr = d * g
    ws = 0.5 * mx - 1.5 * g
    cw = np.concatenate([ws, pr], axis=3)
    ch = np.concatenate([cw, cw], axis=2)
    assert net.BlobShape("ch") == ch.shape
    assert rel_err(net.Extract("ch"), ch) < 2e-4
    assert rel_err(net.Extract("grp"), g) < 2e-4
