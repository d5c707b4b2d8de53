"""ctypes binding of the C ABI in include/dyffusion_b200.h.

The shared library is the product: if it is missing this module raises at import time (there is no CPU or
PyTorch fallback for the arithmetic).  PyTorch is used only for device memory, streams and (elsewhere)
torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DYF_LIB") or os.path.join(_HERE, "libdyffusion_b200.so")  # DYF_LIB: A/B builds of the same ABI

ARCH_UNET_SIMPLE, ARCH_UNET_RESNET, ARCH_CONVNET = 0, 1, 2
ABI_VERSION = 2


class NetDesc(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("dim", C.c_int32), ("in_channels", C.c_int32), ("cond_channels", C.c_int32),
        ("out_channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("with_time_emb", C.c_int32),
        ("upsample_h", C.c_int32), ("upsample_w", C.c_int32), ("dropout", C.c_float), ("input_dropout", C.c_float),
        ("n_mults", C.c_int32), ("dim_mults", C.c_int32 * 8), ("groups", C.c_int32), ("block_dropout", C.c_float),
        ("block_dropout1", C.c_float), ("attn_dropout", C.c_float), ("keep_spatial_dims", C.c_int32),
        ("init_kernel", C.c_int32), ("init_padding", C.c_int32), ("init_stride", C.c_int32),
        ("n_kernels", C.c_int32), ("kernel_sizes", C.c_int32 * 8), ("residual", C.c_int32),
    ]


class Dropout(C.Structure):
    _fields_ = [("mode", C.c_int32), ("seed", C.c_uint64), ("stream", C.c_uint64), ("row_offset", C.c_uint64)]


class SamplerDesc(C.Structure):
    _fields_ = [
        ("num_timesteps", C.c_int32), ("n_schedule", C.c_int32), ("schedule", C.POINTER(C.c_double)),
        ("tau", C.POINTER(C.c_double)), ("time_forecaster", C.POINTER(C.c_double)),
        ("forward_conditioning", C.c_int32), ("sampling_type", C.c_int32),
        ("use_cold_sampling_for_last_step", C.c_int32), ("n_refine", C.c_int32),
        ("refine_times", C.POINTER(C.c_double)), ("enable_interpolator_dropout", C.c_int32),
        ("channels", C.c_int32), ("window_channels", C.c_int32), ("static_channels", C.c_int32),
        ("interpolator_horizon", C.c_int32), ("max_rows_per_call", C.c_int32), ("forecaster_dropout", C.c_int32),
        ("cuda_graph", C.c_int32),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dyffusion_b200.build` (nvcc, sm_100a). "
            "dyffusion_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64, sz = C.c_void_p, C.c_int32, C.c_uint64, C.c_size_t
    sig = {
        "dyf_abi_version": (C.c_int, []),
        "dyf_last_error": (C.c_char_p, []),
        "dyf_act_dtype": (C.c_char_p, []),
        "dyf_nvtx_enable": (C.c_int, [i32]),
        "dyf_launch_count": (u64, []),
        "dyf_net_create": (C.c_int, [C.POINTER(NetDesc), C.POINTER(vp)]),
        "dyf_net_destroy": (None, [vp]),
        "dyf_net_set_param": (C.c_int, [vp, C.c_char_p, vp, C.POINTER(C.c_int64), i32]),
        "dyf_net_finalize": (C.c_int, [vp, vp]),
        "dyf_net_num_params": (C.c_int, [vp]),
        "dyf_net_param_key": (C.c_char_p, [vp, i32]),
        "dyf_net_param_shape": (C.c_int, [vp, i32, C.POINTER(C.c_int64), C.POINTER(i32)]),
        "dyf_net_workspace_bytes": (C.c_int, [vp, i32, C.POINTER(sz)]),
        "dyf_net_forward": (C.c_int, [vp, i32, vp, vp, vp, vp, C.POINTER(Dropout), vp, sz, vp]),
        "dyf_net_forward_srcs": (C.c_int, [vp, i32, C.POINTER(vp), C.POINTER(i32), i32, vp, vp, C.POINTER(Dropout),
                                           vp, sz, vp]),
        "dyf_sampler_create": (C.c_int, [vp, vp, C.POINTER(SamplerDesc), C.POINTER(vp)]),
        "dyf_sampler_destroy": (None, [vp]),
        "dyf_sampler_workspace_bytes": (C.c_int, [vp, i32, C.POINTER(sz)]),
        "dyf_sampler_num_outputs": (C.c_int, [vp, C.POINTER(i32), C.POINTER(C.c_double), i32]),
        "dyf_sampler_run": (C.c_int, [vp, i32, vp, vp, vp, vp, u64, u64, vp, sz, vp]),
        "dyf_sampler_graph_replays": (C.c_int, [vp, C.POINTER(u64)]),
        "dyf_debug_dropout_mask": (C.c_int, [u64, u64, C.c_uint32, C.c_float, C.c_int64, vp, vp]),
        "dyf_ensemble_metrics_workspace_bytes": (C.c_int, [i32, C.c_int64, C.c_int64, C.POINTER(sz)]),
        "dyf_ensemble_metrics": (C.c_int, [vp, vp, i32, C.c_int64, C.c_int64, vp, vp, vp, sz, vp]),
        "dyf_boundary_conditions_navier_stokes": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
        "dyf_boundary_conditions_spring_mesh": (C.c_int, [vp, vp, vp, C.c_int64, i32, i32, i32, vp]),
        "dyf_window_gather": (C.c_int, [vp, C.c_int64, C.c_int64, C.POINTER(C.c_int64), i32, i32, vp, vp]),
        "dyf_adamw_workspace_bytes": (C.c_int, [C.c_int64, C.POINTER(sz)]),
        "dyf_grad_sq_norm": (C.c_int, [vp, C.c_int64, vp, vp, sz, vp]),
        "dyf_adamw_step": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_int64, C.c_double, vp, sz, vp]),
        "dyf_profile_enable": (C.c_int, [i32]),
        "dyf_profile_filter": (C.c_int, [i32]),
        "dyf_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(u64), i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.dyf_abi_version() != ABI_VERSION:
        raise ImportError("libdyffusion_b200.so ABI version mismatch")
    return lib


LIB = _load()
EXPORTED = ["dyf_abi_version", "dyf_last_error", "dyf_act_dtype", "dyf_nvtx_enable", "dyf_launch_count", "dyf_net_create", "dyf_net_destroy",
            "dyf_net_set_param", "dyf_net_finalize", "dyf_net_num_params", "dyf_net_param_key", "dyf_net_param_shape",
            "dyf_net_workspace_bytes", "dyf_net_forward", "dyf_net_forward_srcs", "dyf_sampler_create",
            "dyf_sampler_destroy", "dyf_sampler_workspace_bytes", "dyf_sampler_num_outputs", "dyf_sampler_run",
            "dyf_sampler_graph_replays",
            "dyf_debug_dropout_mask", "dyf_profile_enable", "dyf_profile_filter", "dyf_profile_read",
            "dyf_ensemble_metrics_workspace_bytes", "dyf_ensemble_metrics", "dyf_boundary_conditions_navier_stokes",
            "dyf_boundary_conditions_spring_mesh", "dyf_window_gather", "dyf_adamw_workspace_bytes",
            "dyf_grad_sq_norm", "dyf_adamw_step"]
KERNEL_CLASSES = ["conv_mma", "conv_umma", "pack", "upsample", "groupnorm", "readout", "time_tables", "elementwise",
                  "attention", "conv_up", "conv_flat"]


class EngineError(RuntimeError):
    pass


def _check(rc: int) -> None:
    if rc != 0:
        msg = LIB.dyf_last_error().decode(errors="replace")
        if rc == -4:
            raise NotImplementedError(f"dyffusion_b200: {msg}")
        if rc == -1:
            raise ValueError(f"dyffusion_b200: {msg}")
        raise EngineError(f"dyffusion_b200 (code {rc}): {msg}")


def launch_count() -> int:
    return int(LIB.dyf_launch_count())


def act_dtype() -> str:
    """Storage type of activations / tensor-core operands of the loaded library ("fp16" or "bf16")."""
    return LIB.dyf_act_dtype().decode()


def nvtx_enable(on: bool) -> None:
    _check(LIB.dyf_nvtx_enable(int(bool(on))))


def profile_enable(on: bool) -> None:
    _check(LIB.dyf_profile_enable(int(bool(on))))


def profile_filter(kernel_class: Optional[str]) -> None:
    """Bracket only launches of `kernel_class` with events (None = every class)."""
    _check(LIB.dyf_profile_filter(-1 if kernel_class is None else KERNEL_CLASSES.index(kernel_class)))


def profile_read() -> Dict[str, Dict[str, float]]:
    """Per kernel class: device ms (CUDA events around every launch), algorithmic FLOPs / bytes, launch count."""
    n = len(KERNEL_CLASSES)
    ms, fl, by, la = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_uint64 * n)()
    _check(LIB.dyf_profile_read(ms, fl, by, la, n))
    return {k: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(la[i])) for i, k in enumerate(KERNEL_CLASSES)}


def ensemble_metrics(preds: torch.Tensor, targets: torch.Tensor, per_member: bool = False):
    """preds [members, samples, inner] / targets [samples, inner], fp32 CUDA -> (per_sample [samples, 3] float64 sums of
    crps / squared error of the ensemble mean / member variance over `inner`, per_member_mse [members] float64 or None)."""
    preds = _require_cuda(preds, "predictions").contiguous()
    targets = _require_cuda(targets, "targets").contiguous()
    n, s, d = preds.shape
    nbytes = C.c_size_t()
    _check(LIB.dyf_ensemble_metrics_workspace_bytes(n, s, d, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=preds.device)
    per_sample = torch.empty((s, 3), dtype=torch.float64, device=preds.device)
    member = torch.empty((n,), dtype=torch.float64, device=preds.device) if per_member else None
    _check(LIB.dyf_ensemble_metrics(preds.data_ptr(), targets.data_ptr(), n, s, d, per_sample.data_ptr(),
                                    member.data_ptr() if per_member else None, ws.data_ptr(), nbytes.value, _stream_ptr()))
    return per_sample, member


def _stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise EngineError(f"dyffusion_b200 has no CPU path: `{name}` must be a CUDA tensor (got {t.device})")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Workspace:
    """Grow-only scratch per (device, CUDA stream), owned by torch's caching allocator: engine calls enqueued on different
    streams never share a workspace (the epilogue-table cache of a net is still per net: issue a net's first call for a
    new time tuple on one stream before using it from others)."""

    def __init__(self):
        self.buf: Dict[Tuple[int, int], torch.Tensor] = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, int(torch.cuda.current_stream(idx).cuda_stream))
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            self.buf[key] = b = None  # release before growing
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


WORKSPACE = _Workspace()


class NetHandle:
    """Owns one `dyf_net` (re-packed weights + layer plan)."""

    def __init__(self, desc: NetDesc):
        self.desc = desc
        h = C.c_void_p()
        _check(LIB.dyf_net_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.finalized = False

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and LIB is not None:  # LIB may already be torn down at interpreter exit
            LIB.dyf_net_destroy(h)

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def param_specs(self) -> List[Tuple[str, Tuple[int, ...], bool]]:
        out = []
        shape = (C.c_int64 * 4)()
        isbuf = C.c_int32()
        for i in range(LIB.dyf_net_num_params(self._h)):
            key = LIB.dyf_net_param_key(self._h, i).decode()
            nd = LIB.dyf_net_param_shape(self._h, i, shape, C.byref(isbuf))
            if nd < 0:
                _check(nd)
            out.append((key, tuple(int(shape[k]) for k in range(nd)), bool(isbuf.value)))
        return out

    def load(self, state: Dict[str, torch.Tensor]) -> None:
        """Strict load of a reference-keyed state dict (device tensors) followed by finalize."""
        keys = {k for k, _, _ in self.param_specs()}
        missing, unexpected = keys - set(state), set(state) - keys
        if missing or unexpected:
            raise KeyError(f"state dict mismatch: missing={sorted(missing)[:5]} unexpected={sorted(unexpected)[:5]}")
        keep = []
        for k, v in state.items():
            if v.dtype in (torch.int64, torch.int32):  # num_batches_tracked
                v = v.to(torch.float32)
            v = _require_cuda(v.detach(), k)
            keep.append(v)
            shp = (C.c_int64 * max(1, v.dim()))(*v.shape)
            _check(LIB.dyf_net_set_param(self._h, k.encode(), C.c_void_p(v.data_ptr()), shp, v.dim()))
        _check(LIB.dyf_net_finalize(self._h, C.c_void_p(_stream_ptr())))
        self.finalized = True

    def workspace_bytes(self, rows: int) -> int:
        n = C.c_size_t()
        _check(LIB.dyf_net_workspace_bytes(self._h, rows, C.byref(n)))
        return int(n.value)

    def forward(self, x: torch.Tensor, time: Optional[torch.Tensor], cond: Optional[torch.Tensor],
                dropout: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        x = _require_cuda(x, "inputs")
        rows = x.shape[0]
        d = self.desc
        if x.dim() != 4 or x.shape[1] != d.in_channels or tuple(x.shape[2:]) != (d.height, d.width):
            raise ValueError(f"inputs must be [rows, {d.in_channels}, {d.height}, {d.width}], got {tuple(x.shape)}")
        if cond is not None:
            cond = _require_cuda(cond, "condition")
            if tuple(cond.shape) != (rows, d.cond_channels, d.height, d.width):
                raise ValueError(f"condition must be [rows, {d.cond_channels}, {d.height}, {d.width}], got {tuple(cond.shape)}")
        if time is not None:
            time = _require_cuda(torch.as_tensor(time, device=x.device).reshape(-1), "time")
            if time.numel() == 1 and rows > 1:
                time = time.expand(rows).contiguous()
            if time.numel() != rows:
                raise ValueError(f"time must have one entry per row ({rows}), got {time.numel()}")
        y = torch.empty((rows, d.out_channels, d.height, d.width), dtype=torch.float32, device=x.device)
        ws = WORKSPACE.get(self.workspace_bytes(rows), x.device)
        dr = Dropout(1, dropout[0], dropout[1], dropout[2] if len(dropout) > 2 else 0) if dropout is not None \
            else Dropout(0, 0, 0, 0)
        with torch.cuda.device(x.device):
            _check(LIB.dyf_net_forward(
                self._h, rows, C.c_void_p(x.data_ptr()), C.c_void_p(cond.data_ptr()) if cond is not None else None,
                C.c_void_p(time.data_ptr()) if time is not None else None, C.c_void_p(y.data_ptr()), C.byref(dr),
                C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(_stream_ptr())))
        return y


class SamplerHandle:
    def __init__(self, forecaster: NetHandle, interpolator: NetHandle, *, num_timesteps: int,
                 schedule: Sequence[float], tau: Sequence[float], time_forecaster: Sequence[float],
                 forward_conditioning: str, sampling_type: str, use_cold_sampling_for_last_step: bool,
                 refine_times: Sequence[float], enable_interpolator_dropout: bool, channels: int,
                 window_channels: int, static_channels: int, interpolator_horizon: int, max_rows_per_call: int = 0,
                 forecaster_dropout: bool = False, cuda_graph: Optional[bool] = None):
        n = len(schedule)
        if cuda_graph is None:  # default on; DYF_CUDA_GRAPH=0 keeps plain launches (A/B, debugging)
            cuda_graph = os.environ.get("DYF_CUDA_GRAPH", "1") != "0"
        self.cuda_graph = bool(cuda_graph)
        arr = lambda v: (C.c_double * max(1, len(v)))(*[float(x) for x in v])
        self._keep = (arr(schedule), arr(tau), arr(time_forecaster), arr(refine_times), forecaster, interpolator)
        fc = {"none": 0, "data": 1, "data+noise": 2}
        if forward_conditioning not in fc:
            raise ValueError(f"Invalid forward conditioning type: {forward_conditioning}")
        st = {"cold": 0, "naive": 1}
        if sampling_type not in st:
            raise ValueError(f"unknown sampling type {sampling_type}")
        d = SamplerDesc(num_timesteps, n, self._keep[0], self._keep[1], self._keep[2], fc[forward_conditioning],
                        st[sampling_type], int(bool(use_cold_sampling_for_last_step)), len(refine_times),
                        self._keep[3], int(bool(enable_interpolator_dropout)), channels, window_channels,
                        static_channels, interpolator_horizon, max_rows_per_call, int(bool(forecaster_dropout)),
                        int(bool(cuda_graph)))
        h = C.c_void_p()
        _check(LIB.dyf_sampler_create(forecaster.handle, interpolator.handle, C.byref(d), C.byref(h)))
        self._h = h
        self.channels, self.window_channels, self.static_channels = channels, window_channels, static_channels
        self.hw = (forecaster.desc.height, forecaster.desc.width)
        nout = C.c_int32()
        _check(LIB.dyf_sampler_num_outputs(self._h, C.byref(nout), None, 0))
        keys = (C.c_double * max(1, nout.value))()
        _check(LIB.dyf_sampler_num_outputs(self._h, C.byref(nout), keys, nout.value))
        self.keys = [float(keys[i]) for i in range(nout.value)]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and LIB is not None:
            LIB.dyf_sampler_destroy(h)

    def workspace_bytes(self, rows: int) -> int:
        n = C.c_size_t()
        _check(LIB.dyf_sampler_workspace_bytes(self._h, rows, C.byref(n)))
        return int(n.value)

    def run(self, ic: torch.Tensor, static: Optional[torch.Tensor], seed: int,
            want_x0: bool = False, row_offset: int = 0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        ic = _require_cuda(ic, "initial_condition")
        rows = ic.shape[0]
        if tuple(ic.shape[1:]) != (self.window_channels, *self.hw):
            raise ValueError(f"initial_condition must be [rows, {self.window_channels}, {self.hw[0]}, {self.hw[1]}]")
        if static is not None:
            static = _require_cuda(static, "static_condition")
            if tuple(static.shape) != (rows, self.static_channels, *self.hw):
                raise ValueError("static_condition has the wrong shape")
        preds = torch.empty((len(self.keys), rows, self.channels, *self.hw), dtype=torch.float32, device=ic.device)
        x0 = torch.empty((rows, self.channels, *self.hw), dtype=torch.float32, device=ic.device) if want_x0 else None
        ws = WORKSPACE.get(self.workspace_bytes(rows), ic.device)
        with torch.cuda.device(ic.device):
            _check(LIB.dyf_sampler_run(
                self._h, rows, C.c_void_p(ic.data_ptr()), C.c_void_p(static.data_ptr()) if static is not None else None,
                C.c_void_p(preds.data_ptr()), C.c_void_p(x0.data_ptr()) if x0 is not None else None,
                C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.c_uint64(int(row_offset)), C.c_void_p(ws.data_ptr()), ws.numel(),
                C.c_void_p(_stream_ptr())))
        return preds, x0

    def graph_replays(self) -> int:
        """Runs of this sampler that were served by replaying a captured CUDA graph (measurement hook)."""
        n = C.c_uint64(0)
        _check(LIB.dyf_sampler_graph_replays(self._h, C.byref(n)))
        return int(n.value)


def debug_dropout_mask(seed: int, stream: int, site: int, p: float, n: int, device="cuda") -> torch.Tensor:
    m = torch.empty(n, dtype=torch.uint8, device=device)
    _check(LIB.dyf_debug_dropout_mask(seed, stream, site, p, n, C.c_void_p(m.data_ptr()), C.c_void_p(_stream_ptr())))
    return m
