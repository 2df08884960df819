"""Python mirror of ``booster::ConvBooster`` / ``booster::ConvParam`` over the fcuda C ABI.

Same names and call protocol as the reference (/root/reference/src/booster/include/booster/booster.h:42-170,
used as in src/layers/conv_layer.h:92-172): fill a ConvParam -> AssignOutputDim -> SelectAlgo ->
GetBufferSize -> Init (once) -> Forward.  Tensors are torch CUDA tensors used purely as device memory;
every call goes through ``libfcuda.so`` (no torch ops on the data path).
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import FcudaConvParam, fcuda

NAIVE, IM2COL, SGECONV, DEPTHWISE, WINOGRADF63, WINOGRADF63FUSED, WINOGRADF23 = range(7)
ALGO_NAMES = ["NAIVE", "IM2COL", "SGECONV", "DEPTHWISE", "WINOGRADF63", "WINOGRADF63FUSED", "WINOGRADF23"]
PRECISION_TF32X3, PRECISION_TF32, PRECISION_FP32_SPLIT = 0, 1, 2


class FcudaError(RuntimeError):
    def __init__(self, fn: str, code: int):
        super().__init__(f"{fn} returned {code}")
        self.code = code


def _check(fn: str, rc: int) -> None:
    if rc != 0:
        raise FcudaError(fn, rc)


def _ptr(t: torch.Tensor | None):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_precision(mode: int) -> None:
    _check("fcuda_set_precision", fcuda().fcuda_set_precision(mode))


def get_precision() -> int:
    return fcuda().fcuda_get_precision()


def set_l2_chunk_bytes(n: int) -> None:
    fcuda().fcuda_set_l2_chunk_bytes(n)


class ConvParam(FcudaConvParam):
    @classmethod
    def make(cls, oc, ic, h, w, kh, kw=None, stride=1, pad=0, group=1, bias=True, relu=False, stride_w=None,
             pad_lbrt=None):
        kw = kh if kw is None else kw
        sw = stride if stride_w is None else stride_w
        pl, pb, pr, pt = pad_lbrt if pad_lbrt is not None else (pad, pad, pad, pad)
        p = cls(oc, ic, h, w, kh, kw, 0, 0, stride, sw, pl, pb, pr, pt, group, int(bool(bias)), int(bool(relu)))
        p.AssignOutputDim()
        return p

    def AssignOutputDim(self):
        _check("fcuda_conv_assign_output_dim", fcuda().fcuda_conv_assign_output_dim(ctypes.byref(self)))

    def GetFLOPS(self) -> float:  # booster.h:145-148
        return (2.0 * self.output_channels * self.input_channels * self.output_h * self.output_w * self.kernel_h *
                self.kernel_w / self.group)

    @property
    def weight_shape(self):
        if self.group == self.input_channels and self.group > 1:
            return (self.input_channels, 1, self.kernel_h, self.kernel_w)
        return (self.output_channels, self.input_channels, self.kernel_h, self.kernel_w)


class ConvBooster:
    """ConvBooster doesn't allocate any memory (booster.h:155): the caller passes every buffer."""

    def __init__(self):
        self.algo = -1

    def SelectAlgo(self, param: ConvParam) -> int:
        a = ctypes.c_int(-1)
        rc = fcuda().fcuda_conv_select_algo(ctypes.byref(param), ctypes.byref(a))
        self.algo = a.value
        return rc

    def SelectAlgoTuned(self, param: ConvParam) -> int:
        a = ctypes.c_int(-1)
        rc = fcuda().fcuda_conv_select_algo_tuned(ctypes.byref(param), ctypes.byref(a))
        self.algo = a.value
        return rc

    def ForceSelectAlgo(self, algo: int) -> int:
        self.algo = int(algo)
        return 0

    def GetBufferSize(self, param: ConvParam, batch: int = 1) -> tuple[int, int]:
        s, k = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _check("fcuda_conv_get_buffer_size",
               fcuda().fcuda_conv_get_buffer_size(ctypes.byref(param), self.algo, batch, ctypes.byref(s), ctypes.byref(k)))
        return s.value, k.value

    def Init(self, param: ConvParam, processed_kernel: torch.Tensor, kernel: torch.Tensor) -> None:
        kp = ctypes.c_void_p(kernel.data_ptr())  # host or device pointer
        _check("fcuda_conv_init",
               fcuda().fcuda_conv_init(ctypes.byref(param), self.algo, _ptr(processed_kernel), kp, _stream()))

    def ForwardResidual(self, param: ConvParam, output, input, kernel, buffer, bias_arr, residual, relu_after_add,
                        batch: int = 1) -> None:
        _check("fcuda_conv_forward_residual",
               fcuda().fcuda_conv_forward_residual(ctypes.byref(param), self.algo, _ptr(output), _ptr(input), _ptr(kernel),
                                                   _ptr(buffer), _ptr(bias_arr), _ptr(residual), int(relu_after_add), batch,
                                                   _stream()))

    def Forward(self, param: ConvParam, output, input, kernel, buffer, bias_arr, batch: int = 1) -> None:
        _check("fcuda_conv_forward",
               fcuda().fcuda_conv_forward(ctypes.byref(param), self.algo, _ptr(output), _ptr(input), _ptr(kernel),
                                          _ptr(buffer), _ptr(bias_arr), batch, _stream()))


def conv_forward(param: ConvParam, x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None = None,
                 algo: int | None = None, residual: torch.Tensor | None = None,
                 relu_after_add: bool = False, dilation: int = 1, tuned: bool = False,
                 pool: bool = False) -> tuple[torch.Tensor, int]:
    """Whole ConvBooster protocol for a batch x (N, IC, H, W) -> (N, OC, OH, OW).  Returns (output, algo).
    With `residual` (shaped like the output) the fused Eltwise-SUM entry point is used instead of Forward.
    `dilation` > 1 goes through fcuda_conv_forward_ext (the caller sets param.output_h/w for the dilated extent)."""
    cb = ConvBooster()
    if algo is None:
        rc = cb.SelectAlgoTuned(param) if tuned else cb.SelectAlgo(param)
        if rc != 0:
            raise FcudaError("fcuda_conv_select_algo", rc)
    else:
        cb.ForceSelectAlgo(algo)
    n = x.shape[0]
    scratch_n, packed_n = cb.GetBufferSize(param, n)
    packed = torch.empty(max(packed_n, 1), device=x.device, dtype=torch.float32)
    scratch = torch.empty(max(scratch_n, 1), device=x.device, dtype=torch.float32)
    cb.Init(param, packed, w.contiguous())
    if pool:  # fused trailing 2x2 / stride-2 max pooling: the output is the pooled blob
        out = torch.empty((n, param.output_channels, (param.output_h + 1) // 2, (param.output_w + 1) // 2), device=x.device,
                          dtype=torch.float32)
        _check("fcuda_conv_forward_pool",
               fcuda().fcuda_conv_forward_pool(ctypes.byref(param), cb.algo, _ptr(out), _ptr(x.contiguous()), _ptr(packed),
                                               _ptr(scratch), _ptr(b), n, _stream()))
        return out, cb.algo
    out = torch.empty((n, param.output_channels, param.output_h, param.output_w), device=x.device, dtype=torch.float32)
    if dilation > 1:
        _check("fcuda_conv_forward_ext",
               fcuda().fcuda_conv_forward_ext(ctypes.byref(param), cb.algo, _ptr(out), _ptr(x.contiguous()), _ptr(packed),
                                              _ptr(scratch), _ptr(b), _ptr(residual.contiguous()) if residual is not None else None,
                                              int(relu_after_add), dilation, dilation, n, _stream()))
    elif residual is None:
        cb.Forward(param, out, x.contiguous(), packed, scratch, b, n)
    else:
        cb.ForwardResidual(param, out, x.contiguous(), packed, scratch, b, residual.contiguous(), relu_after_add, n)
    return out, cb.algo


# ---- other layer entry points (thin wrappers used by the parity tests) ------------------------------
def inner_product(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, relu: bool = False) -> torch.Tensor:
    lib = fcuda()
    n = x.shape[0]
    in_size = x[0].numel()
    out_size = w.shape[0]
    s, k = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _check("fcuda_inner_product_get_buffer_size",
           lib.fcuda_inner_product_get_buffer_size(in_size, out_size, n, ctypes.byref(s), ctypes.byref(k)))
    packed = torch.empty(max(k.value, 1), device=x.device)
    scratch = torch.empty(max(s.value, 1), device=x.device)
    _check("fcuda_inner_product_init", lib.fcuda_inner_product_init(in_size, out_size, _ptr(packed), _ptr(w.contiguous()), _stream()))
    out = torch.empty((n, out_size), device=x.device)
    _check("fcuda_inner_product_forward",
           lib.fcuda_inner_product_forward(in_size, out_size, _ptr(out), _ptr(x.contiguous()), _ptr(packed), _ptr(b),
                                           _ptr(scratch), int(relu), n, _stream()))
    return out


def eltwise(a: torch.Tensor, b: torch.Tensor, op: int, ca: float = 1.0, cb: float = 1.0, relu: bool = False) -> torch.Tensor:
    """op 0 PROD, 1 SUM (with coefficients), 2 MAX — ncnn Eltwise semantics (the reference only has plain SUM)."""
    out = torch.empty_like(a)
    _check("fcuda_eltwise_forward",
           fcuda().fcuda_eltwise_forward(_ptr(out), _ptr(a.contiguous()), _ptr(b.contiguous()), a.numel(), op, ca, cb,
                                         int(relu), _stream()))
    return out


PIXEL_RGB, PIXEL_BGR, PIXEL_GRAY, PIXEL_RGBA = 1, 2, 4, 8  # ncnn pixel types (mat.h:126-129); conversion = from | (to << 16)


def from_pixels(pixels: torch.Tensor, type_: int, target_w: int = 0, target_h: int = 0, mean=None, norm=None) -> torch.Tensor:
    """(N, h, w[, c]) uint8 CUDA tensor -> (N, C, th, tw) fp32: from_pixels[_resize] + substract_mean_normalize on the GPU."""
    import numpy as np
    lib = fcuda()
    assert pixels.is_cuda and pixels.dtype == torch.uint8 and pixels.is_contiguous()
    n, h, w = pixels.shape[:3]
    sc, oc = ctypes.c_int(), ctypes.c_int()
    _check("fcuda_pixel_channels", lib.fcuda_pixel_channels(type_, ctypes.byref(sc), ctypes.byref(oc)))
    assert pixels[0].numel() == h * w * sc.value
    tw, th = target_w or w, target_h or h
    out = torch.empty((n, oc.value, th, tw), device=pixels.device, dtype=torch.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    m = None if mean is None else np.ascontiguousarray(mean, np.float32)
    s = None if norm is None else np.ascontiguousarray(norm, np.float32)
    _check("fcuda_from_pixels",
           lib.fcuda_from_pixels(_ptr(out), ctypes.c_void_p(pixels.data_ptr()), type_, w, h, tw, th,
                                 None if m is None else m.ctypes.data_as(fp), None if s is None else s.ctypes.data_as(fp), n,
                                 _stream()))
    return out


def pooling(x: torch.Tensor, type_: int, kh, kw, sh, sw, pl, pr, pt, pb, global_pooling=False) -> torch.Tensor:
    lib = fcuda()
    n, c, h, w = x.shape
    if global_pooling:
        oh = ow = 1
    else:
        oh = lib.fcuda_pooling_out_dim(h, pt, pb, kh, sh)
        ow = lib.fcuda_pooling_out_dim(w, pl, pr, kw, sw)
    out = torch.empty((n, c, oh, ow), device=x.device)
    _check("fcuda_pooling_forward",
           lib.fcuda_pooling_forward(_ptr(out), _ptr(x.contiguous()), c, h, w, type_, kh, kw, sh, sw, pl, pr, pt, pb,
                                     int(global_pooling), n, _stream()))
    return out


def batchnorm(x, alpha, beta, scale=None, scale_bias=None, relu=False):
    n, c = x.shape[:2]
    out = torch.empty_like(x)
    _check("fcuda_batchnorm_forward",
           fcuda().fcuda_batchnorm_forward(_ptr(out), _ptr(x.contiguous()), c, x[0, 0].numel(), _ptr(alpha), _ptr(beta),
                                           _ptr(scale), _ptr(scale_bias), int(relu), n, _stream()))
    return out


def scale(x, s, b=None):
    n, c = x.shape[:2]
    out = torch.empty_like(x)
    _check("fcuda_scale_forward",
           fcuda().fcuda_scale_forward(_ptr(out), _ptr(x.contiguous()), c, x[0, 0].numel(), _ptr(s), _ptr(b), n, _stream()))
    return out


def eltwise_add(a, b, relu=False):
    out = torch.empty_like(a)
    _check("fcuda_eltwise_add_forward",
           fcuda().fcuda_eltwise_add_forward(_ptr(out), _ptr(a.contiguous()), _ptr(b.contiguous()), a.numel(), int(relu), _stream()))
    return out


def relu(x):
    out = torch.empty_like(x)
    _check("fcuda_relu_forward", fcuda().fcuda_relu_forward(_ptr(out), _ptr(x.contiguous()), x.numel(), _stream()))
    return out


def softmax(x):
    out = torch.empty_like(x)
    _check("fcuda_softmax_forward",
           fcuda().fcuda_softmax_forward(_ptr(out), _ptr(x.contiguous()), x[0].numel(), x.shape[0], _stream()))
    return out


def dropout(x, s):
    out = torch.empty_like(x)
    _check("fcuda_dropout_forward", fcuda().fcuda_dropout_forward(_ptr(out), _ptr(x.contiguous()), x.numel(), float(s), _stream()))
    return out


def tensor_gemm(a: torch.Tensor, b: torch.Tensor, x3: bool = True) -> torch.Tensor:
    """D[g] = A[g] @ B[g]^T with A (G, M, K), B (G, N, K) — the raw TensorGEMM (a8 in SURVEY.md §8)."""
    lib = fcuda()
    g, m, k = a.shape
    n = b.shape[1]
    a = a.contiguous(); b = b.contiguous()
    d = torch.empty((g, m, n), device=a.device)
    if x3:  # A is split inside the kernel (tensor memory); B carries its hi / lo planes
        bh, bl = torch.empty_like(b), torch.empty_like(b)
        _check("fcuda_split_tf32", lib.fcuda_split_tf32(_ptr(bh), _ptr(bl), _ptr(b), b.numel(), _stream()))
        _check("fcuda_tensor_gemm", lib.fcuda_tensor_gemm(_ptr(d), _ptr(a), _ptr(bh), _ptr(bl), m, n, k, g, _stream()))
    else:
        _check("fcuda_tensor_gemm", lib.fcuda_tensor_gemm(_ptr(d), _ptr(a), _ptr(b), None, m, n, k, g, _stream()))
    return d
