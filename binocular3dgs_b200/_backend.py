"""ctypes binding of the C-ABI in ``include/b3gs.h`` — the ``_C`` module of the operator surface.

The reference binds its rasterizer with pybind11
(``submodules/diff-gaussian-rasterization/ext.cpp:15-18``) and does the tensor
allocation in C++ (``rasterize_points.cu:35-229``).  Here the shared library has no
torch types in its signatures, so this file is the host side above the C-ABI: it
allocates outputs and the three opaque state blobs as torch tensors (the blobs through
the resize callback, like ``resizeFunctional`` at ``rasterize_points.cu:27-33``),
unwraps ``data_ptr()``/current stream, and exposes the three functions with the exact
names, argument order and return tuples of the reference's ``_C`` module:

    rasterize_gaussians(...)           -> (int, color, depth, alpha, radii, geom, binning, img)
    rasterize_gaussians_backward(...)  -> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D,
                                           dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
    mark_visible(means3D, viewmatrix, projmatrix) -> bool tensor

The class is parameterised by symbol prefix so that the test infrastructure
(``oracle/refbackend.py``) can bind the reference's own kernels — compiled behind an
identical C-ABI — and drive them through identical host code.  Nothing in this
package loads anything from ``oracle/``.

There is deliberately NO fallback: if the CUDA library cannot be loaded the import
fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_RESIZE_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


class _Buffer(ctypes.Structure):
    _fields_ = [("resize", _RESIZE_FN), ("user", ctypes.c_void_p)]


_F = ctypes.c_void_p  # every device pointer travels as void*
_FWD_ARGTYPES = (
    [_Buffer, _Buffer, _Buffer, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int]
    + [_F, _F, _F, _F, _F, ctypes.c_float, _F, _F, _F, _F, _F, ctypes.c_float, ctypes.c_float, ctypes.c_int]
    + [_F, _F, _F, _F, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
)
_BWD_ARGTYPES = (
    [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F, ctypes.c_int, ctypes.c_int]
    + [_F, _F, _F, _F, _F, ctypes.c_float, _F, _F, _F, _F, _F, ctypes.c_float, ctypes.c_float, _F]
    + [_F, _F, _F, _F, _F, _F]
    + [_F] * 10
    + [ctypes.c_int, ctypes.c_void_p]
)


def _ptr(t: torch.Tensor):
    """Reference null convention (rasterize_points.cu:96-115): empty tensor -> nullptr."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t: torch.Tensor, name: str) -> torch.Tensor:
    if t is None:
        return torch.empty(0)
    if t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:       # a contiguous view at an odd offset: the kernels load float4s
        t = t.clone()
    return t


class Backend:
    """One loaded rasterizer library (ours or the reference veneer)."""

    def __init__(self, path: str, prefix: str, needs_zeroed_outputs: bool, name: str):
        if not os.path.exists(path):
            raise ImportError(
                f"{name}: CUDA library not found at {path}. Build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback."
            )
        self.path = path
        self.name = name
        self.prefix = prefix
        self.needs_zeroed_outputs = needs_zeroed_outputs
        self.lib = ctypes.CDLL(path)
        g = lambda s: getattr(self.lib, prefix + s)
        self._forward = g("forward")
        self._forward.argtypes = _FWD_ARGTYPES
        self._forward.restype = ctypes.c_int
        self._backward = g("backward")
        self._backward.argtypes = _BWD_ARGTYPES
        self._backward.restype = ctypes.c_int
        # b3gs_backward_flags (B3GS_BWD_ACCUMULATE): ours only
        self._backward_flags = getattr(self.lib, prefix + "backward_flags", None)
        if self._backward_flags is not None:
            self._backward_flags.argtypes = [ctypes.c_uint] + _BWD_ARGTYPES
            self._backward_flags.restype = ctypes.c_int
        # backward fused with the data-parallel exchange (ours only)
        self._backward_exchange = getattr(self.lib, prefix + "backward_exchange", None)
        if self._backward_exchange is not None:
            self._backward_exchange.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_uint] + _BWD_ARGTYPES
            self._backward_exchange.restype = ctypes.c_int
        # the no-sync forward (ours only): same arguments minus `debug` and `num_rendered`, plus capacity, ticket
        self._forward_nosync = getattr(self.lib, prefix + "forward_nosync", None)
        if self._forward_nosync is not None:
            self._forward_nosync.argtypes = _FWD_ARGTYPES[:-3] + [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
            self._forward_nosync.restype = ctypes.c_int
            self._nosync_supported = g("forward_nosync_supported")
            self._nosync_supported.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
            self._nosync_supported.restype = ctypes.c_int
            self._count_wait = g("count_wait")
            self._count_wait.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
            self._count_wait.restype = ctypes.c_int
        # the raw-parameter entry (ours only)
        self._forward_raw = getattr(self.lib, prefix + "forward_raw", None)
        if self._forward_raw is not None:
            self._forward_raw.argtypes = ([_Buffer, _Buffer, _Buffer, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F,
                                           ctypes.c_int, ctypes.c_int] + [_F] * 5 + [ctypes.c_float] + [_F] * 4 +
                                          [ctypes.c_float, ctypes.c_float] + [_F] * 4 +
                                          [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)])
            self._forward_raw.restype = ctypes.c_int
            self._backward_raw = g("backward_raw")
            self._backward_raw.argtypes = ([ctypes.c_uint] + [ctypes.c_int] * 4 + [_F, ctypes.c_int, ctypes.c_int] +
                                           [_F] * 5 + [ctypes.c_float] + [_F] * 5 + [ctypes.c_float, ctypes.c_float] +
                                           [_F] * 4 + [_F] * 3 + [_F] * 7 + [ctypes.c_void_p])
            self._backward_raw.restype = ctypes.c_int
        self._mark_visible = g("mark_visible")
        self._mark_visible.argtypes = [ctypes.c_int, _F, _F, _F, _F, ctypes.c_void_p]
        self._mark_visible.restype = ctypes.c_int
        self._last_error = g("last_error")
        self._last_error.restype = ctypes.c_char_p
        self._version = g("version")
        self._version.restype = ctypes.c_char_p
        self._launch_count = g("launch_count")
        self._launch_count.restype = ctypes.c_ulonglong
        for kind in ("geometry", "binning", "image"):
            f = g(kind + "_offset")
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p] if kind == "image" else [ctypes.c_int, ctypes.c_char_p]
            setattr(self, "_" + kind + "_offset", f)
            f = g(kind + "_bytes")
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_int, ctypes.c_int] if kind == "image" else [ctypes.c_int]
            setattr(self, "_" + kind + "_bytes", f)
        # per-stage device timing (only our library implements it)
        self.has_profile = hasattr(self.lib, prefix + "profile_read")
        if self.has_profile:
            g("profile_enable").argtypes = [ctypes.c_int]
            g("profile_stage_name").restype = ctypes.c_char_p
            g("profile_stage_name").argtypes = [ctypes.c_int]
            g("profile_read").argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
        # Optional gradient sink: the backward writes its five parameter gradients ("means3D",
        # "shs", "opacities", "scales", "rotations") INTO preallocated memory instead of
        # allocating them — how dp.GradientBucket receives gradients without a pack copy.
        # Either a dict name -> tensor (every backward overwrites it; for callers that drive
        # `_C` directly, one backward per step) or an object with
        # ``acquire(autograd: bool) -> (dict | None, accumulate: bool)`` (dp.GradientBucket):
        # the first backward of a step gets the views; later ones get fresh tensors under
        # autograd (which adds them into the stolen views itself) or accumulate in the kernel
        # (B3GS_BWD_ACCUMULATE) when `_C` is driven directly.
        self.grad_sink = None
        self.in_autograd = False     # set by rasterizer._RasterizeGaussians.backward around its call
        # our C-ABI treats NULL dL/ddepth, dL/dalpha as zeros; the reference veneer does not
        self.accepts_null_grads = prefix == "b3gs_"
        self.supports_skip_unobservable = prefix == "b3gs_"
        # One persistent C callback; `user` is the slot index (0 geom, 1 binning, 2 image).
        self._tls = threading.local()
        self._cb = _RESIZE_FN(self._resize)

    # ------------------------------------------------------------------ helpers
    def version(self) -> str:
        return self._version().decode()

    def launch_count(self) -> int:
        return int(self._launch_count())

    def profile_enable(self, on: bool):
        if self.has_profile:
            getattr(self.lib, self.prefix + "profile_enable")(int(bool(on)))

    def profile_read(self):
        """{stage: (device_ms_total, calls)} since the last read; waits for the events."""
        if not self.has_profile:
            return {}
        n = getattr(self.lib, self.prefix + "profile_num_stages")()
        ms = (ctypes.c_double * n)()
        calls = (ctypes.c_ulonglong * n)()
        getattr(self.lib, self.prefix + "profile_read")(ms, calls, n)
        name = getattr(self.lib, self.prefix + "profile_stage_name")
        return {name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}

    def _err(self, what: str, code: int) -> RuntimeError:
        return RuntimeError(f"{self.name}.{what} failed ({code}): {self._last_error().decode()}")

    def _resize(self, user, nbytes):
        slot = int(user or 0)
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=self._tls.device)
        self._tls.slots[slot] = t
        return t.data_ptr()

    def _buffers(self):
        return (_Buffer(self._cb, 0), _Buffer(self._cb, 1), _Buffer(self._cb, 2))

    def blob_view(self, blob: torch.Tensor, kind: str, name: str, dtype: torch.dtype, count: int, *dims):
        """Slice a named internal array out of an opaque blob (parity tests only)."""
        if kind == "image":
            off = self._image_offset(dims[0], dims[1], name.encode())
        else:
            off = getattr(self, "_" + kind + "_offset")(dims[0], name.encode())
        if off == ctypes.c_size_t(-1).value:
            raise KeyError(name)
        base = blob.data_ptr()
        aligned = (base + 255) // 256 * 256 if self.prefix == "b3gs_" else (base + 127) // 128 * 128
        off += aligned - base
        nbytes = count * torch.empty(0, dtype=dtype).element_size()
        return blob[off : off + nbytes].view(dtype)

    # ------------------------------------------------------------------ _C surface
    def nosync_supported(self, P: int, W: int, H: int) -> bool:
        return self._forward_nosync is not None and bool(self._nosync_supported(int(P), int(W), int(H)))

    def count_wait(self, ticket: int) -> int:
        """Block until the instance count of a no-sync forward has landed; return it."""
        r = ctypes.c_int(0)
        rc = self._count_wait(int(ticket), ctypes.byref(r))
        if rc != 0:
            raise self._err("count_wait", rc)
        return r.value

    def rasterize_gaussians_nosync(self, capacity, *args):
        """`rasterize_gaussians` without the host wait: the binning blob holds `capacity`
        instances and the first element of the result is a TICKET for :meth:`count_wait`
        (include/b3gs.h: b3gs_forward_nosync).  `args` as for rasterize_gaussians."""
        return self.rasterize_gaussians(*args, _capacity=int(capacity))

    def rasterize_gaussians(
        self, background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
        prefiltered, debug, _capacity=None,
    ):
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
        dev = means3D.device
        if dev.type != "cuda":
            raise RuntimeError("means3D must be a CUDA tensor (no CPU path exists)")
        alloc = torch.zeros if self.needs_zeroed_outputs else torch.empty
        with torch.cuda.device(dev):
            out_color = alloc((3, H, W), dtype=torch.float32, device=dev)
            out_depth = alloc((1, H, W), dtype=torch.float32, device=dev)
            out_alpha = alloc((1, H, W), dtype=torch.float32, device=dev)
            radii = alloc((P,), dtype=torch.int32, device=dev)
            self._tls.device = dev
            self._tls.slots = [torch.empty(0, dtype=torch.uint8, device=dev) for _ in range(3)]
            M = int(sh.size(1)) if sh.dim() >= 2 and sh.size(0) != 0 else 0
            bg, m3, col, opa, sca, rot, cov, vm, pm, shs, cam = (
                _prep(background, "background"), _prep(means3D, "means3D"), _prep(colors, "colors_precomp"),
                _prep(opacity, "opacities"), _prep(scales, "scales"), _prep(rotations, "rotations"),
                _prep(cov3D_precomp, "cov3D_precomp"), _prep(viewmatrix, "viewmatrix"),
                _prep(projmatrix, "projmatrix"), _prep(sh, "sh"), _prep(campos, "campos"),
            )
            rendered = ctypes.c_int(0)
            stream = torch.cuda.current_stream(dev).cuda_stream
            g, b, i = self._buffers()
            if _capacity is None:
                rc = self._forward(
                    g, b, i, P, int(degree), M, _ptr(bg), W, H, _ptr(m3), _ptr(shs), _ptr(col), _ptr(opa), _ptr(sca),
                    float(scale_modifier), _ptr(rot), _ptr(cov), _ptr(vm), _ptr(pm), _ptr(cam), float(tan_fovx),
                    float(tan_fovy), int(bool(prefiltered)), out_color.data_ptr(), out_depth.data_ptr(),
                    out_alpha.data_ptr(), _ptr(radii), int(bool(debug)), stream, ctypes.byref(rendered),
                )
            else:   # `rendered` receives the ticket
                rc = self._forward_nosync(
                    g, b, i, P, int(degree), M, _ptr(bg), W, H, _ptr(m3), _ptr(shs), _ptr(col), _ptr(opa), _ptr(sca),
                    float(scale_modifier), _ptr(rot), _ptr(cov), _ptr(vm), _ptr(pm), _ptr(cam), float(tan_fovx),
                    float(tan_fovy), int(bool(prefiltered)), out_color.data_ptr(), out_depth.data_ptr(),
                    out_alpha.data_ptr(), _ptr(radii), stream, int(_capacity), ctypes.byref(rendered),
                )
            if rc != 0:
                raise self._err("rasterize_gaussians", rc)
            geom, binning, img = self._tls.slots
            self._tls.slots = None
        return rendered.value, out_color, out_depth, out_alpha, radii, geom, binning, img

    def rasterize_gaussians_backward(
        self, background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
        projmatrix, tan_fovx, tan_fovy, dL_dout_color, dL_dout_depth, dL_dout_alpha, sh, degree, campos,
        geomBuffer, R, binningBuffer, imageBuffer, alphas, debug, skip_unobservable=False,
    ):
        """``skip_unobservable`` (an addition, used by the autograd surface): do not materialise
        dL_dcolors when colours came from SH and dL_dcov3D when covariances came from scale/rotation
        — the reference fills both (rasterize_points.cu:160,164) although autograd discards them;
        they are returned as None then.  dL_dconic and dL_ddepth never leave this function."""
        P = int(means3D.size(0))
        H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
        if not self.accepts_null_grads:
            dL_dout_depth = torch.zeros_like(alphas) if dL_dout_depth is None else dL_dout_depth
            dL_dout_alpha = torch.zeros_like(alphas) if dL_dout_alpha is None else dL_dout_alpha
        dev = means3D.device
        M = int(sh.size(1)) if sh.dim() >= 2 and sh.size(0) != 0 else 0
        # The reference zero-fills all ten (rasterize_points.cu:158-167); our library
        # writes every element itself.
        alloc = torch.zeros if (self.needs_zeroed_outputs or P == 0) else torch.empty
        with torch.cuda.device(dev):
            o = dict(dtype=torch.float32, device=dev)
            dL_dmeans3D = alloc((P, 3), **o)
            dL_dmeans2D = alloc((P, 3), **o)
            own = self.prefix == "b3gs_"            # the reference veneer needs all ten, zero-filled
            want_colors = not (own and skip_unobservable and colors.numel() == 0)
            want_cov3D = not (own and skip_unobservable and cov3D_precomp.numel() == 0)
            dL_dcolors = alloc((P, 3), **o) if want_colors else None
            dL_ddepths = None if own else alloc((P, 1), **o)
            dL_dconic = None if own else alloc((P, 2, 2), **o)
            dL_dopacity = alloc((P, 1), **o)
            dL_dcov3D = alloc((P, 6), **o) if want_cov3D else None
            dL_dsh = alloc((P, M, 3), **o)
            dL_dscales = alloc((P, 3), **o)
            dL_drotations = alloc((P, 4), **o)
            sink, accumulate, exchange = self.grad_sink, False, None
            if sink is not None and hasattr(sink, "acquire"):
                owner = sink
                sink, accumulate = owner.acquire(self.in_autograd)
                # last backward of the step on a symmetric-memory bucket: compute and exchange overlapped
                if sink is not None and self._backward_exchange is not None and hasattr(owner, "exchange_if_last"):
                    exchange = owner.exchange_if_last()
            if accumulate and self._backward_flags is None:
                raise RuntimeError(f"{self.name}: this library cannot accumulate into a gradient sink")
            if sink is not None and P != 0:
                def take(name, t):
                    v = sink.get(name)
                    if (v is None or v.numel() != t.numel() or v.device != t.device or not v.is_contiguous()
                            or v.data_ptr() % 16 != 0):
                        return t
                    if self.needs_zeroed_outputs:
                        v.zero_()
                    return v.view(t.shape)
                dL_dmeans3D = take("means3D", dL_dmeans3D)
                dL_dsh = take("shs", dL_dsh)
                dL_dopacity = take("opacities", dL_dopacity)
                dL_dscales = take("scales", dL_dscales)
                dL_drotations = take("rotations", dL_drotations)
            if P != 0:
                bg, m3, col, sca, rot, cov, vm, pm, shs, cam, alp, gc, gd, ga = (
                    _prep(background, "background"), _prep(means3D, "means3D"), _prep(colors, "colors_precomp"),
                    _prep(scales, "scales"), _prep(rotations, "rotations"), _prep(cov3D_precomp, "cov3D_precomp"),
                    _prep(viewmatrix, "viewmatrix"), _prep(projmatrix, "projmatrix"), _prep(sh, "sh"),
                    _prep(campos, "campos"), _prep(alphas, "alphas"), _prep(dL_dout_color, "dL_dout_color"),
                    _prep(dL_dout_depth, "dL_dout_depth"), _prep(dL_dout_alpha, "dL_dout_alpha"),
                )
                rad = radii.contiguous()
                stream = torch.cuda.current_stream(dev).cuda_stream
                fn = self._backward if not accumulate else (lambda *a: self._backward_flags(1, *a))
                if exchange is not None:
                    handle, scale = exchange
                    fn = lambda *a: self._backward_exchange(handle, scale, 1 if accumulate else 0, *a)
                rc = fn(
                    P, int(degree), M, int(R), _ptr(bg), W, H, _ptr(m3), _ptr(shs), _ptr(col), _ptr(alp), _ptr(sca),
                    float(scale_modifier), _ptr(rot), _ptr(cov), _ptr(vm), _ptr(pm), _ptr(cam), float(tan_fovx),
                    float(tan_fovy), _ptr(rad), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                    _ptr(gc), _ptr(gd), _ptr(ga), dL_dmeans2D.data_ptr(), _ptr(dL_dconic),
                    dL_dopacity.data_ptr(), _ptr(dL_dcolors), _ptr(dL_ddepths), dL_dmeans3D.data_ptr(),
                    _ptr(dL_dcov3D), _ptr(dL_dsh), dL_dscales.data_ptr(), dL_drotations.data_ptr(),
                    int(bool(debug)), stream,
                )
                if rc != 0:
                    raise self._err("rasterize_gaussians_backward", rc)
        return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations

    # ------------------------------------------------------------------ raw-parameter entry
    def rasterize_raw(self, background, xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, scale_modifier,
                      viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, degree, campos,
                      capacity=None):
        """b3gs_forward_raw: the reference's RAW parameters in, activations fused into the
        preprocess.  Returns (R or ticket, color, depth, alpha, radii, geom, binning, img);
        with ``capacity`` the forward does not wait and the first element is a ticket."""
        if xyz.dim() != 2 or xyz.size(1) != 3:
            raise RuntimeError("xyz must have dimensions (num_points, 3)")
        if not xyz.is_cuda:
            raise RuntimeError("xyz must be a CUDA tensor (no CPU path exists)")
        P, H, W, dev = int(xyz.size(0)), int(image_height), int(image_width), xyz.device
        M = 1 + (int(f_rest.size(1)) if f_rest is not None and f_rest.numel() else 0)
        with torch.cuda.device(dev):
            out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            out_depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
            out_alpha = torch.empty((1, H, W), dtype=torch.float32, device=dev)
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            self._tls.device = dev
            self._tls.slots = [torch.empty(0, dtype=torch.uint8, device=dev) for _ in range(3)]
            t = [_prep(x, n) for x, n in ((background, "background"), (xyz, "xyz"), (f_dc, "f_dc"), (f_rest, "f_rest"),
                                          (opacity_raw, "opacity"), (scaling_raw, "scaling"), (rotation_raw, "rotation"),
                                          (viewmatrix, "viewmatrix"), (projmatrix, "projmatrix"), (campos, "campos"))]
            bg, m3, dc, rest, opa, sca, rot, vm, pm, cam = t
            res = ctypes.c_int(0)
            g, b, i = self._buffers()
            if P == 0:
                out_color.zero_(); out_depth.zero_(); out_alpha.zero_()
                rc = 0
            else:
                rc = self._forward_raw(
                    g, b, i, P, int(degree), M, _ptr(bg), W, H, _ptr(m3), _ptr(dc), _ptr(rest), _ptr(opa), _ptr(sca),
                    float(scale_modifier), _ptr(rot), _ptr(vm), _ptr(pm), _ptr(cam), float(tan_fovx), float(tan_fovy),
                    out_color.data_ptr(), out_depth.data_ptr(), out_alpha.data_ptr(), _ptr(radii),
                    torch.cuda.current_stream(dev).cuda_stream, -1 if capacity is None else int(capacity),
                    ctypes.byref(res))
            if rc != 0:
                raise self._err("rasterize_raw", rc)
            geom, binning, img = self._tls.slots
            self._tls.slots = None
        return res.value, out_color, out_depth, out_alpha, radii, geom, binning, img

    def rasterize_raw_backward(self, background, xyz, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw,
                               scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                               dL_dout_depth, dL_dout_alpha, degree, campos, radii, geomBuffer, R, binningBuffer,
                               imageBuffer, alphas, outputs=None, accumulate=False):
        """b3gs_backward_raw -> (dL_dmeans2D, dL_dxyz, dL_df_dc, dL_df_rest, dL_dopacity,
        dL_dscaling, dL_drotation), gradients of the RAW parameters.  ``outputs``: optional dict
        name -> preallocated tensor ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
        written in place (``accumulate``: added to), e.g. slices of a dp.ParameterBucket."""
        P, dev = int(xyz.size(0)), xyz.device
        H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
        M = 1 + (int(f_rest.size(1)) if f_rest is not None and f_rest.numel() else 0)
        o = dict(dtype=torch.float32, device=dev)
        alloc = torch.zeros if P == 0 else torch.empty
        outputs = outputs or {}

        def out(name, shape):
            v = outputs.get(name)
            if v is not None and v.numel() == int(torch.Size(shape).numel()) and v.is_contiguous() and v.data_ptr() % 16 == 0:
                return v.view(shape)
            if accumulate:
                raise RuntimeError("accumulate=True needs preallocated outputs")
            return alloc(shape, **o)
        with torch.cuda.device(dev):
            g2d = alloc((P, 3), **o)
            g_xyz, g_dc, g_rest = out("xyz", (P, 3)), out("f_dc", (P, 1, 3)), out("f_rest", (P, M - 1, 3))
            g_op, g_sc, g_rot = out("opacity", (P, 1)), out("scaling", (P, 3)), out("rotation", (P, 4))
            if P != 0:
                t = [_prep(x, n) for x, n in ((background, "background"), (xyz, "xyz"), (f_dc, "f_dc"),
                                              (f_rest, "f_rest"), (opacity_raw, "opacity"), (scaling_raw, "scaling"),
                                              (rotation_raw, "rotation"), (alphas, "alphas"), (viewmatrix, "viewmatrix"),
                                              (projmatrix, "projmatrix"), (campos, "campos"),
                                              (dL_dout_color, "dL_dout_color"), (dL_dout_depth, "dL_dout_depth"),
                                              (dL_dout_alpha, "dL_dout_alpha"))]
                bg, m3, dc, rest, opa, sca, rot, alp, vm, pm, cam, gc, gd, ga = t
                rc = self._backward_raw(
                    1 if accumulate else 0, P, int(degree), M, int(R), _ptr(bg), W, H, _ptr(m3), _ptr(dc), _ptr(rest),
                    _ptr(opa), _ptr(sca), float(scale_modifier), _ptr(rot), _ptr(alp), _ptr(vm), _ptr(pm), _ptr(cam),
                    float(tan_fovx), float(tan_fovy), _ptr(radii.contiguous()), _ptr(geomBuffer), _ptr(binningBuffer),
                    _ptr(imageBuffer), _ptr(gc), _ptr(gd), _ptr(ga), g2d.data_ptr(), g_xyz.data_ptr(), g_dc.data_ptr(),
                    _ptr(g_rest), g_op.data_ptr(), g_sc.data_ptr(), g_rot.data_ptr(),
                    torch.cuda.current_stream(dev).cuda_stream)
                if rc != 0:
                    raise self._err("rasterize_raw_backward", rc)
        return g2d, g_xyz, g_dc, g_rest, g_op, g_sc, g_rot

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        P = int(means3D.size(0))
        dev = means3D.device
        present = torch.zeros((P,), dtype=torch.bool, device=dev)
        if P != 0:
            with torch.cuda.device(dev):
                m3, vm, pm = _prep(means3D, "means3D"), _prep(viewmatrix, "viewmatrix"), _prep(projmatrix, "projmatrix")
                rc = self._mark_visible(P, _ptr(m3), _ptr(vm), _ptr(pm), present.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream)
                if rc != 0:
                    raise self._err("mark_visible", rc)
        return present


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb3gs.so")

_native = None
_compiled = None
COMPILED_PATH = os.path.join(_HERE, "_b3gs_torch.so")


class CompiledBackend:
    """The same ``_C`` surface served by the compiled host side (csrc/torch_binding.cpp, a
    pybind11 module over the identical C-ABI calls): ~4x less host time per call than the
    ctypes path, which matters because the host is on the critical path after the forward's
    synchronisation.  Everything that is not one of the three operator entry points
    (profiling, blob slicing for the parity tests, the DP gradient sink) is delegated to the
    ctypes ``Backend`` bound to the same ``libb3gs.so``."""

    accepts_null_grads = True
    supports_skip_unobservable = True

    def __init__(self, module, ctypes_backend: Backend):
        self._m = module
        self._ct = ctypes_backend
        self.name, self.prefix, self.path = "b3gs(compiled host)", ctypes_backend.prefix, COMPILED_PATH
        self.needs_zeroed_outputs = False
        self.rasterize_gaussians = module.rasterize_gaussians
        self.mark_visible = module.mark_visible
        if hasattr(module, "rasterize_gaussians_nosync"):
            self.rasterize_gaussians_nosync = module.rasterize_gaussians_nosync
            self.count_wait = module.count_wait

    def rasterize_gaussians_backward(self, *args, skip_unobservable=False):
        if self._ct.grad_sink is not None:      # gradients go straight into the DP bucket
            return self._ct.rasterize_gaussians_backward(*args, skip_unobservable=skip_unobservable)
        return self._m.rasterize_gaussians_backward(*args, skip_unobservable)

    # state that lives on the ctypes backend must be SET there too (a plain attribute
    # assignment would land on this wrapper and be ignored by the backward)
    @property
    def grad_sink(self):
        return self._ct.grad_sink

    @grad_sink.setter
    def grad_sink(self, value):
        self._ct.grad_sink = value

    @property
    def in_autograd(self):
        return self._ct.in_autograd

    @in_autograd.setter
    def in_autograd(self, value):
        self._ct.in_autograd = value

    def __getattr__(self, name):                # lib, launch_count, profile_*, blob_view, ...
        return getattr(self._ct, name)


def preferred():
    """The compiled host side if it has been built (``__graft_entry__.build()``), else the
    ctypes one.  Both call the same kernels in the same ``libb3gs.so``."""
    global _compiled
    if _compiled is None:
        ct = native()
        _compiled = ct
        if os.path.exists(COMPILED_PATH) and os.environ.get("B3GS_HOST", "compiled") != "ctypes":
            import importlib.util
            try:
                spec = importlib.util.spec_from_file_location("_b3gs_torch", COMPILED_PATH)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _compiled = CompiledBackend(mod, ct)
            except (ImportError, OSError) as ex:      # stale build against another torch: keep ctypes, say so
                import warnings
                warnings.warn(f"binocular3dgs_b200: compiled host side not loadable ({ex}); using the ctypes one")
    return _compiled


def native() -> Backend:
    """Our sm_100a library.  Raises ImportError if it has not been built."""
    global _native
    if _native is None:
        _native = Backend(LIB_PATH, "b3gs_", needs_zeroed_outputs=False, name="b3gs")
    return _native
