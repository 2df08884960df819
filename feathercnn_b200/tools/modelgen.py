"""Synthetic ncnn-format (.param/.bin) model writer.

The reference snapshot loads ncnn ``.param`` (text, magic 7767517) + ``.bin`` (4-byte flag + fp32 for
weight blobs, raw fp32 for bias/BN/Scale blobs) — /root/reference/src/net.cpp:68-230,
/root/reference/src/ncnn/modelbin.cpp:47-197.  Its own converter (``tools/``) is absent from the
snapshot, and there is no network for pretrained checkpoints, so benchmarks and tests use random-init
models of the architectures BASELINE.json names: a single 3x3 conv, VGG-16, ResNet-50 (Caffe topology:
BN + Scale + ReLU as separate layers, Split/Eltwise residuals), MobileNet-v1, plus a small "mini" net
that exercises every layer type.  Weight statistics follow SURVEY.md §8(d).
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np


class ModelWriter:
    def __init__(self, seed: int = 0):
        self.rng = np.random.default_rng(seed)
        self.lines: list[str] = []
        self.blobs: set[str] = set()
        self.bin = bytearray()
        self.flops = 0.0  # direct-conv FLOPs per image, booster.h:145-148 convention

    # ---- low level ---------------------------------------------------------------------------
    def _layer(self, type_: str, name: str, bottoms: list[str], tops: list[str], params: dict | None = None):
        kv = " ".join(f"{k}={v}" for k, v in (params or {}).items())
        self.lines.append(f"{type_} {name} {len(bottoms)} {len(tops)} {' '.join(bottoms + tops)} {kv}".rstrip())
        self.blobs.update(tops)

    def _weights(self, a: np.ndarray, flagged: bool):
        if flagged:
            self.bin += struct.pack("<I", 0)  # raw fp32 tag
        self.bin += np.ascontiguousarray(a, np.float32).tobytes()

    # ---- layers ------------------------------------------------------------------------------
    def input(self, name="data", c=3, h=224, w=224):
        self._layer("Input", name, [], [name], {0: w, 1: h, 2: c})
        self.shape = {name: (c, h, w)}
        return name

    def conv(self, name, bottom, oc, k, stride=1, pad=0, bias=True, group=1, top=None, gain=None):
        top = top or name
        ic, h, w = self.shape[bottom]
        icg = ic // group
        fan_in = icg * k * k
        std = np.sqrt(2.0 / fan_in) if gain is None else gain
        wts = (self.rng.standard_normal((oc, icg, k, k)) * std).astype(np.float32)
        ltype = "ConvolutionDepthWise" if group > 1 else "Convolution"
        params = {0: oc, 1: k, 3: stride, 4: pad, 5: int(bias), 6: wts.size}
        if group > 1:
            params[7] = group
        self._layer(ltype, name, [bottom], [top], params)
        self._weights(wts, True)
        if bias:
            self._weights(self.rng.uniform(-0.1, 0.1, oc).astype(np.float32), False)
        oh = (h + 2 * pad - k) // stride + 1
        ow = (w + 2 * pad - k) // stride + 1
        self.shape[top] = (oc, oh, ow)
        self.flops += 2.0 * oc * icg * oh * ow * k * k
        return top

    def relu(self, name, bottom, top=None):
        top = top or name
        self._layer("ReLU", name, [bottom], [top])
        self.shape[top] = self.shape[bottom]
        return top

    def pool(self, name, bottom, type_=0, k=2, stride=2, pad=0, global_pooling=False, top=None):
        top = top or name
        c, h, w = self.shape[bottom]
        params = {0: type_, 1: k, 2: stride, 3: pad, 4: int(global_pooling)}
        self._layer("Pooling", name, [bottom], [top], params)
        if global_pooling:
            self.shape[top] = (c, 1, 1)
        else:  # ceil mode, pooling_layer.h:129-130
            oh = int(np.ceil((h + 2 * pad - k) / stride)) + 1
            ow = int(np.ceil((w + 2 * pad - k) / stride)) + 1
            self.shape[top] = (c, oh, ow)
        return top

    def fc(self, name, bottom, out, bias=True, top=None):
        top = top or name
        n_in = int(np.prod(self.shape[bottom]))
        wts = (self.rng.standard_normal((out, n_in)) * np.sqrt(2.0 / n_in)).astype(np.float32)
        self._layer("InnerProduct", name, [bottom], [top], {0: out, 1: int(bias), 2: wts.size})
        self._weights(wts, True)
        if bias:
            self._weights(self.rng.uniform(-0.1, 0.1, out).astype(np.float32), False)
        self.shape[top] = (out, 1, 1)
        self.flops += 2.0 * out * n_in
        return top

    def batchnorm(self, name, bottom, top=None, eps=1e-5):
        top = top or name
        c = self.shape[bottom][0]
        self._layer("BatchNorm", name, [bottom], [top], {0: c, 1: f"{eps:e}"})
        r = self.rng
        self._weights(r.uniform(0.9, 1.1, c).astype(np.float32), False)   # slope
        self._weights(r.uniform(-0.1, 0.1, c).astype(np.float32), False)  # mean
        self._weights(r.uniform(0.5, 1.5, c).astype(np.float32), False)   # var
        self._weights(r.uniform(-0.1, 0.1, c).astype(np.float32), False)  # bias
        self.shape[top] = self.shape[bottom]
        return top

    def scale(self, name, bottom, bias=True, top=None, center=1.0):
        top = top or name
        c = self.shape[bottom][0]
        self._layer("Scale", name, [bottom], [top], {0: c, 1: int(bias)})
        self._weights((center * self.rng.uniform(0.9, 1.1, c)).astype(np.float32), False)
        if bias:
            self._weights(self.rng.uniform(-0.1, 0.1, c).astype(np.float32), False)
        self.shape[top] = self.shape[bottom]
        return top

    def split(self, name, bottom, tops):
        self._layer("Split", name, [bottom], tops)
        for t in tops:
            self.shape[t] = self.shape[bottom]
        return tops

    def eltwise(self, name, bottoms, top=None):
        top = top or name
        self._layer("Eltwise", name, bottoms, [top], {0: 1})
        self.shape[top] = self.shape[bottoms[0]]
        return top

    def concat(self, name, bottoms, top=None):
        top = top or name
        self._layer("Concat", name, bottoms, [top], {0: 0})
        c = sum(self.shape[b][0] for b in bottoms)
        self.shape[top] = (c,) + self.shape[bottoms[0]][1:]
        return top

    def softmax(self, name, bottom, top=None):
        top = top or name
        self._layer("Softmax", name, [bottom], [top])
        self.shape[top] = self.shape[bottom]
        return top

    def dropout(self, name, bottom, top=None, scale=None):
        top = top or name
        self._layer("Dropout", name, [bottom], [top], {} if scale is None else {0: f"{scale:e}"})
        self.shape[top] = self.shape[bottom]
        return top

    # ---- output ------------------------------------------------------------------------------
    def save(self, prefix) -> tuple[str, str]:
        prefix = str(prefix)
        Path(prefix).parent.mkdir(parents=True, exist_ok=True)
        text = "7767517\n" + f"{len(self.lines)} {len(self.blobs)}\n" + "\n".join(self.lines) + "\n"
        Path(prefix + ".param").write_text(text)
        Path(prefix + ".bin").write_bytes(bytes(self.bin))
        return prefix + ".param", prefix + ".bin"


# --------------------------------------------------------------------------------------------------
# model zoo
# --------------------------------------------------------------------------------------------------
def single_conv(ic=64, oc=64, h=56, w=56, k=3, stride=1, pad=1, bias=True, group=1, seed=0) -> ModelWriter:
    """BASELINE.json configs[0]: one 3x3 conv 64->64 on 56x56."""
    m = ModelWriter(seed)
    m.input("data", ic, h, w)
    m.conv("conv", "data", oc, k, stride, pad, bias, group)
    return m


def vgg16(seed=0, num_classes=1000, size=224) -> ModelWriter:
    m = ModelWriter(seed)
    x = m.input("data", 3, size, size)
    cfg = [(64, 2), (128, 2), (256, 3), (512, 3), (512, 3)]
    for si, (ch, n) in enumerate(cfg, 1):
        for li in range(1, n + 1):
            x = m.conv(f"conv{si}_{li}", x, ch, 3, 1, 1, True)
            x = m.relu(f"relu{si}_{li}", x)
        x = m.pool(f"pool{si}", x, 0, 2, 2)
    x = m.fc("fc6", x, 4096); x = m.relu("relu6", x); x = m.dropout("drop6", x)
    x = m.fc("fc7", x, 4096); x = m.relu("relu7", x); x = m.dropout("drop7", x)
    x = m.fc("fc8", x, num_classes)
    m.softmax("prob", x)
    return m


def _bn_scale_relu(m: ModelWriter, tag: str, x: str, relu=True, center=1.0):
    x = m.batchnorm("bn" + tag, x)
    x = m.scale("scale" + tag, x, True, center=center)
    if relu:
        x = m.relu(tag.lstrip("_") + "_relu", x)
    return x


def resnet50(seed=0, num_classes=1000, size=224) -> ModelWriter:
    """Caffe ResNet-50 layout: stride on the first 1x1 of each stage, BN/Scale/ReLU layers, Split/Eltwise."""
    m = ModelWriter(seed)
    x = m.input("data", 3, size, size)
    x = m.conv("conv1", x, 64, 7, 2, 3, True)
    x = _bn_scale_relu(m, "_conv1", x)
    x = m.pool("pool1", x, 0, 3, 2, 0)
    stages = [(2, 3, 64, 1), (3, 4, 128, 2), (4, 6, 256, 2), (5, 3, 512, 2)]
    for sid, nblocks, width, stride in stages:
        for b in range(nblocks):
            blk = f"{sid}{chr(ord('a') + b)}"
            s = stride if b == 0 else 1
            a, c = m.split(f"res{blk}_split", x, [f"res{blk}_in_a", f"res{blk}_in_b"])
            if b == 0:
                sc = m.conv(f"res{blk}_branch1", a, width * 4, 1, s, 0, False)
                sc = _bn_scale_relu(m, f"{blk}_branch1", sc, relu=False)
            else:
                sc = a
            y = m.conv(f"res{blk}_branch2a", c, width, 1, s, 0, False)
            y = _bn_scale_relu(m, f"{blk}_branch2a", y)
            y = m.conv(f"res{blk}_branch2b", y, width, 3, 1, 1, False)
            y = _bn_scale_relu(m, f"{blk}_branch2b", y)
            y = m.conv(f"res{blk}_branch2c", y, width * 4, 1, 1, 0, False)
            y = _bn_scale_relu(m, f"{blk}_branch2c", y, relu=False, center=0.3)  # keep residual sums O(1)
            x = m.eltwise(f"res{blk}", [sc, y])
            x = m.relu(f"res{blk}_relu", x)
    x = m.pool("pool5", x, 1, 7, 1, 0, global_pooling=True)
    x = m.fc("fc1000", x, num_classes)
    m.softmax("prob", x)
    return m


def mobilenet_v1(seed=0, num_classes=1000, size=224) -> ModelWriter:
    m = ModelWriter(seed)
    x = m.input("data", 3, size, size)
    x = m.conv("conv1", x, 32, 3, 2, 1, False)
    x = _bn_scale_relu(m, "_conv1", x)
    cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1), (512, 1),
           (1024, 2), (1024, 1)]
    for i, (oc, s) in enumerate(cfg, 2):
        ic = m.shape[x][0]
        x = m.conv(f"conv{i}_dw", x, ic, 3, s, 1, False, group=ic)
        x = _bn_scale_relu(m, f"_conv{i}_dw", x)
        x = m.conv(f"conv{i}_pw", x, oc, 1, 1, 0, False)
        x = _bn_scale_relu(m, f"_conv{i}_pw", x)
    x = m.pool("pool6", x, 1, 7, 1, 0, global_pooling=True)
    x = m.conv("fc7", x, num_classes, 1, 1, 0, True)
    m.softmax("prob", x)
    return m


def mini(seed=0, size=32, ch=16) -> ModelWriter:
    """Small net touching every layer type (SURVEY.md §8c's 12-layer probe net, plus concat/dropout/depthwise)."""
    m = ModelWriter(seed)
    x = m.input("data", 4, size, size)
    x = m.conv("conv1", x, ch, 3, 1, 1, True)               # Winograd F63 eligible (size > 8, ch % 4 == 0)
    x = m.batchnorm("bn1", x); x = m.scale("scale1", x, True); x = m.relu("relu1", x)
    a, b = m.split("split1", x, ["s1a", "s1b"])
    y = m.conv("conv2", b, ch, 1, 1, 0, False)               # 1x1 -> IM2COL
    y = m.batchnorm("bn2", y); y = m.scale("scale2", y, False)
    x = m.eltwise("sum1", [a, y]); x = m.relu("relu2", x)
    x = m.conv("conv3_dw", x, ch, 3, 2, 1, False, group=ch)  # depthwise s2
    x = m.relu("relu3", x)
    p, q = m.split("split2", x, ["s2a", "s2b"])
    q = m.conv("conv4", q, ch, 3, 2, 1, True)                # 3x3 s2 -> IM2COL
    p = m.pool("pool1", p, 0, 3, 2, 0)                       # max, ceil mode
    x = m.concat("cat1", [p, q])
    x = m.conv("conv5", x, ch * 2, 3, 1, 1, True)            # Winograd if spatial > 8 else IM2COL
    x = m.relu("relu5", x)
    x = m.pool("pool2", x, 1, 2, 2, 0, global_pooling=True)  # global average
    x = m.dropout("drop1", x, scale=0.5)
    x = m.fc("fc1", x, 24, True)
    x = m.relu("relu6", x)
    x = m.fc("fc2", x, 10, True)
    m.softmax("prob", x)
    return m


ZOO = {"single_conv": single_conv, "vgg16": vgg16, "resnet50": resnet50, "mobilenet_v1": mobilenet_v1, "mini": mini}


def synthetic_input(shape, index: int = 0) -> np.ndarray:
    """uniform(-0.5, 0.5), seed 1234 + image index."""
    return np.random.default_rng(1234 + index).uniform(-0.5, 0.5, shape).astype(np.float32)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("model", choices=sorted(ZOO))
    ap.add_argument("prefix")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    mw = ZOO[a.model](seed=a.seed)
    print(mw.save(a.prefix), f"{mw.flops / 1e9:.3f} GFLOP/image")
