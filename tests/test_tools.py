"""Host tools (CPU): the Caffe -> ncnn/.feathermodel converter (the reference's tools/ is absent,
/root/reference/CMakeLists.txt:84-89).  A caffemodel is assembled here with an independent protobuf writer, converted, and
the result must be exactly the ncnn grammar the loaders read, load in the product's host parser and run in the oracle."""
import struct

import numpy as np

from feathercnn_b200.tools import caffe2feather


# ---- minimal protobuf writer (test side; the tool has its own reader) ------------------------------------------------
def _vi(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _f(field: int, wt: int, payload: bytes) -> bytes:
    return _vi((field << 3) | wt) + payload


def fint(field, v):
    return _f(field, 0, _vi(int(v)))


def fbytes(field, b):
    if isinstance(b, str):
        b = b.encode()
    return _f(field, 2, _vi(len(b)) + b)


def ffloat(field, v):
    return _f(field, 5, struct.pack("<f", v))


def blob(a: np.ndarray) -> bytes:
    a = np.ascontiguousarray(a, np.float32)
    shape = fbytes(7, b"".join(fint(1, d) for d in a.shape))
    return shape + fbytes(5, a.tobytes())      # BlobProto: shape = 7, packed float data = 5


def layer(name, type_, bottoms, tops, blobs=(), **params) -> bytes:
    msg = fbytes(1, name) + fbytes(2, type_)
    msg += b"".join(fbytes(3, b) for b in bottoms) + b"".join(fbytes(4, t) for t in tops)
    msg += b"".join(fbytes(7, blob(b)) for b in blobs)
    for field, payload in params.items():
        msg += fbytes(int(field[1:]), payload)
    return fbytes(100, msg)                    # NetParameter.layer = 100


def test_caffe_converter_writes_the_grammar_the_loaders_read(tmp_path):
    rng = np.random.default_rng(0)
    w1 = rng.standard_normal((8, 4, 3, 3)).astype(np.float32) * 0.2
    b1 = rng.uniform(-0.1, 0.1, 8).astype(np.float32)
    mean, var, sf = rng.uniform(-0.2, 0.2, 8).astype(np.float32), rng.uniform(0.5, 1.5, 8).astype(np.float32), np.array([2.0], np.float32)
    gamma, beta = rng.uniform(0.9, 1.1, 8).astype(np.float32), rng.uniform(-0.1, 0.1, 8).astype(np.float32)
    wdw = rng.standard_normal((8, 1, 3, 3)).astype(np.float32) * 0.3
    wfc = rng.standard_normal((5, 8)).astype(np.float32) * 0.3
    bfc = rng.uniform(-0.1, 0.1, 5).astype(np.float32)
    conv_p = fint(1, 8) + fint(2, 1) + fint(3, 1) + fint(4, 3) + fint(6, 1)            # num_output, bias, pad, kernel, stride
    dw_p = fint(1, 8) + fint(2, 0) + fint(3, 1) + fint(4, 3) + fint(5, 8) + fint(6, 2)  # group 8, stride 2, no bias
    net = (fbytes(1, "tiny") + fbytes(3, "data") + b"".join(fint(4, d) for d in (1, 4, 12, 12))
           + layer("conv1", "Convolution", ["data"], ["conv1"], [w1, b1], f106=conv_p)
           + layer("bn1", "BatchNorm", ["conv1"], ["conv1"], [mean, var, sf], f139=ffloat(3, 1e-5))   # in place
           + layer("scale1", "Scale", ["conv1"], ["conv1"], [gamma, beta], f142=fint(4, 1))           # in place
           + layer("relu1", "ReLU", ["conv1"], ["conv1"])                                             # in place
           + layer("dw", "Convolution", ["conv1"], ["dw"], [wdw], f106=dw_p)                          # reader 1 of relu output
           + layer("pool", "Pooling", ["conv1"], ["pool"], f121=fint(1, 0) + fint(2, 2) + fint(3, 2))  # reader 2 -> Split
           + layer("sum", "Eltwise", ["dw", "pool"], ["sum"], f110=fint(1, 1))
           + layer("gap", "Pooling", ["sum"], ["gap"], f121=fint(1, 1) + fint(12, 1))
           + layer("drop", "Dropout", ["gap"], ["gap"], f108=ffloat(1, 0.5))
           + layer("fc", "InnerProduct", ["gap"], ["fc"], [wfc, bfc], f117=fint(1, 5) + fint(2, 1))
           + layer("train_only", "Dropout", ["fc"], ["fc"], f8=fint(1, 0))                            # include { phase: TRAIN }
           + layer("prob", "Softmax", ["fc"], ["prob"]))
    text, blob_bytes = caffe2feather.convert(net)
    want = "\n".join([
        "7767517",
        "13 14",
        "Input data 0 1 data 0=12 1=12 2=4",
        "Convolution conv1 1 1 data conv1 0=8 1=3 11=3 2=1 3=1 13=1 4=1 14=1 5=1 6=288",
        "BatchNorm bn1 1 1 conv1 conv1_bn1 0=8 1=1.000000e-05",
        "Scale scale1 1 1 conv1_bn1 conv1_scale1 0=8 1=1",
        "ReLU relu1 1 1 conv1_scale1 conv1_relu1",
        "Split splitncnn_conv1_relu1 1 2 conv1_relu1 conv1_relu1_splitncnn_0 conv1_relu1_splitncnn_1",
        "ConvolutionDepthWise dw 1 1 conv1_relu1_splitncnn_0 dw 0=8 1=3 11=3 2=1 3=2 13=2 4=1 14=1 5=0 6=72 7=8",
        "Pooling pool 1 1 conv1_relu1_splitncnn_1 pool 0=0 1=2 11=2 2=2 12=2 3=0 13=0 4=0",
        "Eltwise sum 2 1 dw pool sum 0=1",
        "Pooling gap 1 1 sum gap 0=1 1=1 11=1 2=1 12=1 3=0 13=0 4=1",
        "Dropout drop 1 1 gap gap_drop",
        "InnerProduct fc 1 1 gap_drop fc 0=5 1=1 2=40",
        "Softmax prob 1 1 fc prob",
    ]) + "\n"
    assert text == want
    flag = struct.pack("<I", 0)
    want_bin = (flag + w1.tobytes() + b1.tobytes() + np.ones(8, np.float32).tobytes() + (mean / 2).tobytes() + (var / 2).tobytes()
                + np.zeros(8, np.float32).tobytes() + gamma.tobytes() + beta.tobytes() + flag + wdw.tobytes() + flag + wfc.tobytes()
                + bfc.tobytes())
    assert blob_bytes == want_bin

    # the product's host-side parser accepts it (no GPU involved) ...
    from feathercnn_b200.net import Net
    net_h = Net()
    net_h.LoadParamFromText(text)
    assert "prob" in net_h.BlobNames() and net_h.input_shape == (4, 12, 12)
    # ... and so does the oracle, which also runs it
    from oracle import oracle as O
    param, binf = caffe2feather.convert_file(_write(tmp_path / "tiny.caffemodel", net), tmp_path / "tiny", feathermodel=True)
    cpu = O.OracleNet(param, binf)
    x = rng.uniform(-0.5, 0.5, (4, 12, 12)).astype(np.float32)
    cpu.forward(x)
    prob = cpu.extract("prob")
    assert prob.shape[0] == 5 and np.isfinite(prob).all() and abs(float(prob.sum()) - 1.0) < 1e-5
    assert (tmp_path / "tiny.feathermodel").read_bytes()[:8] == b"FTHRB200"
    if O.reference_available():  # the unmodified reference loads and runs the converted model too
        ref = O.ReferenceNet(param, binf)
        ref.forward(x)
        np.testing.assert_allclose(ref.extract("prob").ravel(), prob.ravel(), rtol=1e-4, atol=1e-6)


def _write(path, data: bytes):
    path.write_bytes(data)
    return path
