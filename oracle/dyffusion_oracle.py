"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch, fp32) restatement of the reference's DYffusion sampling hot path, written functionally over
plain state-dicts.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / the timed CPU baseline -- never as the thing shipped.
The product path (`dyffusion_b200`) never imports it and fails loudly without its CUDA extension.

Parity pinning: the reference ships NO tests / golden vectors for this path, so this oracle is
pinned against outputs of the reference itself, imported unmodified in the build container through
`oracle/ref_shims.py` (see `tests/golden/make_golden.py`, which writes `tests/golden/*.pt`, and
`tests/test_oracle_vs_reference.py`, which compares live when /root/reference is present).

Every function cites the reference file:line (relative to /root/reference) it restates.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]
# A dropout hook receives (site_name, tensor, p) and returns the tensor after dropout.  `None` means dropout off.
DropFn = Optional[Callable[[str, Tensor, float], Tensor]]


def torch_dropout(site: str, x: Tensor, p: float) -> Tensor:
    """nn.Dropout in train mode (what the reference's inference-dropout scope enables, src/utilities/utils.py:560-567)."""
    return F.dropout(x, p, training=True) if p > 0 else x


def _drop(fn: DropFn, site: str, x: Tensor, p: float) -> Tensor:
    return x if (fn is None or p <= 0) else fn(site, x, p)


# ----------------------------------------------------------------------------------------------------------------
# time embedding -- src/models/modules/misc.py:20-32 (SinusoidalPosEmb), :54-67 (Linear -> GELU -> Linear)
# ----------------------------------------------------------------------------------------------------------------
def sinusoidal_embedding(t: Tensor, dim: int) -> Tensor:
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
    arg = t.float()[:, None] * freqs[None, :]
    return torch.cat([arg.sin(), arg.cos()], dim=-1)


def time_embedding(sd: SD, t: Tensor, dim: int, prefix: str = "time_emb_mlp") -> Tensor:
    e = sinusoidal_embedding(t, dim)
    e = F.linear(e, sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"])
    e = F.gelu(e)  # exact erf GELU (nn.GELU() default)
    return F.linear(e, sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"])


def _scale_shift(sd: SD, key: str, temb: Tensor) -> Tuple[Tensor, Tensor]:
    """SiLU -> Linear(time_dim, 2C) -> chunk (unet_simple.py:70-77, unet.py:100-103, simple_conv_net.py:42-49)."""
    ss = F.linear(F.silu(temb), sd[f"{key}.weight"], sd[f"{key}.bias"])
    scale, shift = ss[:, :, None, None].chunk(2, dim=1)
    return scale, shift


def _bn_eval(sd: SD, key: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[f"{key}.running_mean"], sd[f"{key}.running_var"], sd[f"{key}.weight"],
                        sd[f"{key}.bias"], training=False, eps=1e-5)


def _bn(sd: SD, key: str, x: Tensor, train_stats: Optional[Dict[str, Tensor]]) -> Tensor:
    """nn.BatchNorm2d.  `train_stats is None`: eval mode (running statistics; what sampling uses).  Otherwise train mode
    (SURVEY.md 8f-1): normalise with the batch mean / biased batch variance and write the running statistics after the
    momentum-0.1 update (unbiased variance) into `train_stats` -- chained through it when a net is called more than once, as
    in `p_losses`."""
    if train_stats is None:
        return _bn_eval(sd, key, x)
    rm = train_stats.get(f"{key}.running_mean", sd[f"{key}.running_mean"]).clone()
    rv = train_stats.get(f"{key}.running_var", sd[f"{key}.running_var"]).clone()
    y = F.batch_norm(x, rm, rv, sd[f"{key}.weight"], sd[f"{key}.bias"], training=True, momentum=0.1, eps=1e-5)
    train_stats[f"{key}.running_mean"], train_stats[f"{key}.running_var"] = rm, rv
    return y


# ----------------------------------------------------------------------------------------------------------------
# Navier-Stokes backbone -- src/models/unet_simple.py:13-197
# ----------------------------------------------------------------------------------------------------------------
def unet_simple_forward(sd: SD, inputs: Tensor, time: Optional[Tensor], condition: Optional[Tensor], *,
                        dim: int = 64, upsample_dims: Optional[Sequence[int]] = (256, 256),
                        outer_sample_mode: str = "bilinear", dropout: float = 0.0, input_dropout: float = 0.0,
                        drop: DropFn = None, train_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    x = inputs if condition is None else torch.cat([inputs, condition], dim=1)  # inputs first (:184)
    temb = time_embedding(sd, time, dim) if "time_emb_mlp.1.weight" in sd else None
    hw = x.shape[-2:]
    if upsample_dims is not None:  # nn.Upsample(size=..., mode=...) (:100-101, :193)
        x = F.interpolate(x, size=tuple(upsample_dims), mode=outer_sample_mode)
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"])  # 1x1 stem (:113-115)
    x = _drop(drop, "dropout_input", x, input_dropout)
    skips: List[Tensor] = []
    enc = [(4, 1), (4, 1), (4, 1), (4, 1), (2, 0), (2, 0)]  # (kernel, pad) of the six stride-2 encoder blocks (:120-129)
    for i, (_, pad) in enumerate(enc):
        p = f"input_ops.{i}"
        x = F.conv2d(x, sd[f"{p}.ops.0.weight"], sd[f"{p}.ops.0.bias"], stride=2, padding=pad)
        if f"{p}.ops.1.running_mean" in sd:
            x = _bn(sd, f"{p}.ops.1", x, train_stats)
        else:  # last encoder block uses GroupNorm(8) (:56, :128)
            x = F.group_norm(x, 8, sd[f"{p}.ops.1.weight"], sd[f"{p}.ops.1.bias"], eps=1e-5)
        if temb is not None:
            sc, sh = _scale_shift(sd, f"{p}.time_mlp.1", temb)
            x = x * (sc + 1) + sh
        x = F.leaky_relu(x, 0.2)
        x = _drop(drop, p, x, dropout)
        skips.append(x)
    x = skips.pop()
    dec_pad = [0, 0, 1, 1, 1, 1]  # decoder conv kernel = size-1 -> 1,1,3,3,3,3 (:133-140, :41-51)
    for i, pad in enumerate(dec_pad):
        p = f"output_ops.{i}"
        x = F.interpolate(x, scale_factor=2, mode="bilinear")
        x = F.conv2d(x, sd[f"{p}.ops.1.weight"], sd[f"{p}.ops.1.bias"], stride=1, padding=pad)
        x = _bn(sd, f"{p}.ops.2", x, train_stats)
        if temb is not None:
            sc, sh = _scale_shift(sd, f"{p}.time_mlp.1", temb)
            x = x * (sc + 1) + sh
        x = F.relu(x)
        x = _drop(drop, p, x, dropout)
        if skips:
            x = torch.cat([x, skips.pop()], dim=1)
    x = F.conv_transpose2d(x, sd["readout.0.weight"], sd["readout.0.bias"], stride=2, padding=1)  # (:143-150)
    if upsample_dims is None:
        return F.interpolate(x, size=hw, mode=outer_sample_mode)
    return F.interpolate(x, size=hw, mode=outer_sample_mode)  # (:195)


# ----------------------------------------------------------------------------------------------------------------
# spring-mesh backbone -- src/models/simple_conv_net.py:12-131
# ----------------------------------------------------------------------------------------------------------------
def simple_conv_net_forward(sd: SD, inputs: Tensor, time: Optional[Tensor], condition: Optional[Tensor], *,
                            dim: int = 64, kernel_sizes: Sequence[int] = (9, 7, 5, 3), residual: bool = True,
                            dropout: float = 0.0, drop: DropFn = None,
                            train_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    x = inputs if condition is None else torch.cat([inputs, condition], dim=1)  # (:121)
    temb = time_embedding(sd, time, dim) if "time_emb_mlp.1.weight" in sd else None
    for i, k in enumerate(kernel_sizes):
        p = f"convs.{i}"
        res = x
        w = sd[f"{p}.conv.weight"]
        x = F.conv2d(x, w, sd[f"{p}.conv.bias"], padding=(k - 1) // 2)
        x = _bn(sd, f"{p}.norm", x, train_stats)
        if temb is not None:
            sc, sh = _scale_shift(sd, f"{p}.time_mlp.1", temb)
            x = x * (sc + 1) + sh
        x = F.gelu(x)
        x = _drop(drop, p, x, dropout)
        if residual and w.shape[0] == w.shape[1]:  # residual only when C_in == C_out (:31, :53-54)
            x = x + res
    return F.conv2d(x, sd["head.weight"], sd["head.bias"])


# ----------------------------------------------------------------------------------------------------------------
# SST backbone -- src/models/unet.py:26-315 and src/models/modules/attention.py:7-73
# ----------------------------------------------------------------------------------------------------------------
def _ws_conv3x3(sd: SD, key: str, x: Tensor) -> Tensor:
    """WeightStandardizedConv2d (unet.py:32-40): per-output-channel mean / biased var, eps 1e-5 (fp32)."""
    w = sd[f"{key}.weight"]
    mu = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return F.conv2d(x, (w - mu) * torch.rsqrt(var + 1e-5), sd[f"{key}.bias"], padding=1)


def _unet_block(sd: SD, key: str, x: Tensor, ss, p: float, groups: int, drop: DropFn) -> Tensor:
    """Block (unet.py:66-76): WS-conv -> GroupNorm -> x*(scale+1)+shift -> SiLU -> Dropout."""
    x = _ws_conv3x3(sd, f"{key}.proj", x)
    x = F.group_norm(x, groups, sd[f"{key}.norm.weight"], sd[f"{key}.norm.bias"], eps=1e-5)
    if ss is not None:
        x = x * (ss[0] + 1) + ss[1]
    return _drop(drop, key, F.silu(x), p)


def _resnet_block(sd: SD, key: str, x: Tensor, temb, p1: float, p2: float, groups: int, drop: DropFn) -> Tensor:
    """ResnetBlock (unet.py:98-109): only block1 receives scale/shift; 1x1 residual conv iff C_in != C_out."""
    ss = _scale_shift(sd, f"{key}.mlp.1", temb) if (temb is not None and f"{key}.mlp.1.weight" in sd) else None
    h = _unet_block(sd, f"{key}.block1", x, ss, p1, groups, drop)
    h = _unet_block(sd, f"{key}.block2", h, None, p2, groups, drop)
    if f"{key}.residual_conv.weight" in sd:
        x = F.conv2d(x, sd[f"{key}.residual_conv.weight"], sd[f"{key}.residual_conv.bias"])
    return h + x


def _channel_layernorm(g: Tensor, x: Tensor) -> Tensor:
    """LayerNorm over dim=1, gain only (unet.py:43-52)."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mu = x.mean(dim=1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + 1e-5) * g


def _linear_attention(sd: SD, key: str, x: Tensor, p: float, drop: DropFn, heads: int = 4, dh: int = 32) -> Tensor:
    """Residual(PreNorm(LinearAttention(rescale='qkv'))) (unet.py:187-191; attention.py:22-44)."""
    b, c, h, w = x.shape
    n = h * w
    y = _channel_layernorm(sd[f"{key}.fn.norm.g"], x)
    y = _drop(drop, f"{key}.to_qkv", y, p)  # Dropout sits on the *input* of the qkv conv (attention.py:13)
    qkv = F.conv2d(y, sd[f"{key}.fn.fn.to_qkv.1.weight"]).reshape(b, 3, heads, dh, n)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    q = q.softmax(dim=-2) * dh ** -0.5
    k = k.softmax(dim=-1)
    v = v / n
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, heads * dh, h, w)
    out = F.conv2d(out, sd[f"{key}.fn.fn.to_out.weight"], sd[f"{key}.fn.fn.to_out.bias"])
    return out + x


def _full_attention(sd: SD, key: str, x: Tensor, p: float, drop: DropFn, heads: int = 4, dh: int = 32) -> Tensor:
    """Residual(PreNorm(Attention)) (unet.py:209; attention.py:61-73)."""
    b, c, h, w = x.shape
    n = h * w
    y = _channel_layernorm(sd[f"{key}.fn.norm.g"], x)
    qkv = F.conv2d(y, sd[f"{key}.fn.fn.to_qkv.weight"]).reshape(b, 3, heads, dh, n)
    q, k, v = qkv[:, 0] * dh ** -0.5, qkv[:, 1], qkv[:, 2]
    sim = torch.einsum("bhdi,bhdj->bhij", q, k)
    attn = _drop(drop, f"{key}.attn", sim.softmax(dim=-1), p)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)  # [b, heads, n, dh]
    out = out.permute(0, 1, 3, 2).reshape(b, heads * dh, h, w)  # "b h (x y) d -> b (h d) x y"
    out = F.conv2d(out, sd[f"{key}.fn.fn.to_out.weight"], sd[f"{key}.fn.fn.to_out.bias"])
    return out + x


def unet_resnet_forward(sd: SD, x: Tensor, time: Optional[Tensor], condition: Optional[Tensor], *,
                        dim: int = 64, dim_mults: Sequence[int] = (1, 2, 4), groups: int = 8,
                        block_dropout: float = 0.0, block_dropout1: float = 0.0, attn_dropout: float = 0.0,
                        input_dropout: float = 0.0, keep_spatial_dims: bool = False, init_padding: int = 3,
                        init_stride: int = 1, drop: DropFn = None,
                        train_stats: Optional[Dict[str, Tensor]] = None) -> Tensor:
    # `train_stats` is accepted for a uniform signature and unused: GroupNorm / weight standardisation / channel LayerNorm
    # carry no batch statistics, so train mode differs from eval mode only through dropout
    if condition is not None:
        x = torch.cat([condition, x], dim=1)  # condition FIRST (unet.py:269)
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], stride=init_stride, padding=init_padding)
    r = _drop(drop, "dropout_input_for_residual", x, input_dropout)
    x = _drop(drop, "dropout_input", x, input_dropout)
    temb = time_embedding(sd, time, dim) if "time_emb_mlp.1.weight" in sd else None
    nres = len(dim_mults)
    rb = dict(p1=block_dropout1, p2=block_dropout, groups=groups, drop=drop)
    hs: List[Tensor] = []
    for l in range(nres):
        x = _resnet_block(sd, f"downs.{l}.0", x, temb, **rb)
        hs.append(x)
        x = _resnet_block(sd, f"downs.{l}.1", x, temb, **rb)
        x = _linear_attention(sd, f"downs.{l}.2", x, attn_dropout, drop)
        hs.append(x)
        w = sd[f"downs.{l}.3.weight"]
        if w.shape[-1] == 4:  # Downsample = Conv2d(k4, s2, p1) (unet.py:22-23)
            x = F.conv2d(x, w, sd[f"downs.{l}.3.bias"], stride=2, padding=1)
        else:
            x = F.conv2d(x, w, sd[f"downs.{l}.3.bias"], padding=1)
    x = _resnet_block(sd, "mid_block1", x, temb, **rb)
    x = _full_attention(sd, "mid_attn", x, attn_dropout, drop)
    x = _resnet_block(sd, "mid_block2", x, temb, **rb)
    for l in range(nres):
        x = torch.cat([x, hs.pop()], dim=1)
        x = _resnet_block(sd, f"ups.{l}.0", x, temb, **rb)
        x = torch.cat([x, hs.pop()], dim=1)
        x = _resnet_block(sd, f"ups.{l}.1", x, temb, **rb)
        x = _linear_attention(sd, f"ups.{l}.2", x, attn_dropout, drop)
        if f"ups.{l}.3.1.weight" in sd:  # Sequential(Upsample(nearest x2), Conv3x3) (unet.py:16-19)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[f"ups.{l}.3.1.weight"], sd[f"ups.{l}.3.1.bias"], padding=1)
        else:
            x = F.conv2d(x, sd[f"ups.{l}.3.weight"], sd[f"ups.{l}.3.bias"], padding=1)
    x = torch.cat([x, r], dim=1)
    x = _resnet_block(sd, "final_res_block", x, temb, **rb)
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])


BACKBONES = {"unet_simple": unet_simple_forward, "unet_resnet": unet_resnet_forward,
             "simple_conv_net": simple_conv_net_forward}


# ----------------------------------------------------------------------------------------------------------------
# host logic -- src/diffusion/dyffusion.py:44-138 (step<->time map), :245-333 (sampling-schedule parser)
# ----------------------------------------------------------------------------------------------------------------
class Schedule:
    """Step->time map and sampling-schedule parser of BaseDYffusion, restated without torch/Lightning."""

    def __init__(self, timesteps: int, schedule: str = "before_t1_only", additional_interpolation_steps: int = 0,
                 additional_interpolation_steps_factor: int = 0, interpolate_before_t1: bool = True,
                 sampling_schedule: Union[None, str, Sequence[float]] = None):
        if timesteps <= 1:
            raise AssertionError("horizon must be > 1")
        self.kind = schedule
        self.k = 0
        self.fac = 0
        self.add = 0
        if schedule == "linear":
            assert additional_interpolation_steps == 0
            self.fac = additional_interpolation_steps_factor
            interpolated = timesteps - 1 if interpolate_before_t1 else timesteps - 2
            self.add = 0 if interpolate_before_t1 else additional_interpolation_steps_factor
            extra = additional_interpolation_steps_factor * interpolated
        elif schedule == "before_t1_only":
            assert additional_interpolation_steps_factor == 0
            assert interpolate_before_t1
            extra = self.k = additional_interpolation_steps
        else:
            raise ValueError(f"Invalid schedule: {schedule}")
        self.num_timesteps = timesteps + extra
        d2i = {d: self.tau(d) for d in range(1, self.num_timesteps)}
        self.dynamical_steps = {d: i for d, i in d2i.items() if float(i).is_integer()}
        self.artificial_steps = {d: i for d, i in d2i.items() if not float(i).is_integer()}
        self.sampling_schedule = self.parse(sampling_schedule if sampling_schedule not in (None, "None")
                                            else list(range(self.num_timesteps)))

    def tau(self, d):
        """diffusion_step_to_interpolation_step (dyffusion.py:101-138) for python scalars."""
        assert 0 <= d <= self.num_timesteps - 1
        if self.kind == "linear":
            return (d + self.add) / (self.fac + 1)
        return d - self.k if d >= self.k + 1 else d / (self.k + 1)

    def parse(self, spec) -> list:
        import numpy as np
        name = spec
        if isinstance(spec, str):
            base = [0] + list(self.dynamical_steps.keys())
            art = list(self.artificial_steps.keys())
            if "only_dynamics" in name:
                out = []
                if "only_dynamics_plus" in name:
                    n = int(name.replace("only_dynamics_plus", "").replace("_discrete", ""))
                    out = list(np.linspace(0, base[1], n + 1, endpoint=False))
                    if "_discrete" in name:
                        out = [int(np.floor(s)) for s in out]
                else:
                    assert name == "only_dynamics"
            elif name.startswith("every"):
                n = int(name.replace("every", "").replace("th", "").replace("nd", "").replace("rd", ""))
                assert 1 <= n <= self.num_timesteps
                out = art[::n]
            elif name.startswith("first"):
                f = float(name.replace("first", "").replace("v2", ""))
                if f < 1:
                    assert 0 < f < 1
                    out = art[: int(np.ceil(f * len(art)))]
                else:
                    assert f.is_integer() and 1 <= f <= self.num_timesteps
                    out = art[: int(f)]
            else:
                raise ValueError(f"Invalid sampling schedule: ``{name}``. ")
            spec = sorted(set(out + base))
        spec = list(spec)
        assert 1 <= spec[-1] <= self.num_timesteps
        if spec[0] != 0:
            spec = [0] + spec
        for a, b in zip(spec, spec[1:]):
            assert b > a, f"Invalid sampling schedule not monotonically increasing: {spec}"
        if all(float(s).is_integer() for s in spec):
            spec = [int(s) for s in spec]
        return spec


# ----------------------------------------------------------------------------------------------------------------
# the sampler -- src/diffusion/dyffusion.py:335-426 (sample_loop), :140-163 (q_sample), :205-239 (predict_x_last)
# ----------------------------------------------------------------------------------------------------------------
NetFn = Callable[[Tensor, Tensor, Optional[Tensor]], Tensor]  # (x, time[R], condition) -> y


def sample_loop(forecaster: NetFn, interpolator: NetFn, sched: Schedule, initial_condition: Tensor,
                static_condition: Optional[Tensor] = None, *, num_input_channels: int,
                forward_conditioning: str = "data", sampling_type: str = "cold", time_encoding: str = "dynamics",
                refine_intermediate_predictions: bool = False, use_cold_sampling_for_last_step: bool = False,
                prediction_timesteps: Optional[Sequence[float]] = None,
                noise_fn: Callable[[Tensor], Tensor] = torch.randn_like) -> Dict[str, Tensor]:
    R = initial_condition.shape[0]
    N = sched.num_timesteps
    S = sched.sampling_schedule
    full = lambda v: torch.full((R,), float(v), dtype=torch.float32)

    def F_(x_s, s):  # predict_x_last (:205-239) + _predict_last_dynamics (:192-203) + get_condition (:177-190)
        if forward_conditioning == "data":
            c = initial_condition
        elif forward_conditioning == "none":
            c = None
        elif "data+noise" in forward_conditioning:
            w = full(s).view(R, 1, 1, 1) / (N - 1)
            c = w * initial_condition + (1 - w) * noise_fn(initial_condition)
        else:
            raise ValueError(forward_conditioning)
        if static_condition is not None:
            c = static_condition if c is None else torch.cat([c, static_condition], dim=1)
        t = {"discrete": s, "normalized": s / N, "dynamics": sched.tau(s)}[time_encoding]
        return forecaster(x_s, full(t), c)

    def I_(x0_hat, t):  # q_sample (:140-163) + DYffusion._interpolate (:480-494)
        assert 0 < t < sched.tau(N - 1) + 1
        return interpolator(torch.cat([initial_condition, x0_hat], dim=1), full(t), static_condition)

    x_s = initial_condition[:, -num_input_channels:]  # (:348)
    out: Dict[str, Tensor] = {}
    last_plus = S[-1] + 1
    step_key = 0
    x0_hat = None
    for s, s_next in zip(S, S[1:] + [last_plus]):
        is_last = s == N - 1
        x0_hat = F_(x_s, s)
        t_next = sched.tau(s_next) if not is_last else math.inf
        is_dyn = float(t_next).is_integer() or is_last
        x_next = I_(x0_hat, sched.tau(s_next)) if s_next <= N - 1 else x0_hat  # (:374-379)
        if sampling_type == "cold":
            if is_last and not use_cold_sampling_for_last_step:
                x_s = x0_hat
            else:
                x_cur = I_(x0_hat, sched.tau(s)) if s > 0 else x_s
                x_s = x_s - x_cur + x_next  # (:386-388)
        elif sampling_type == "naive":
            x_s = x_next
        else:
            raise ValueError(sampling_type)
        step_key = int(t_next) if s < N - 1 else step_key + 1  # (:395-397)
        if is_dyn:
            out[f"t{step_key}_preds"] = x_s
    if refine_intermediate_predictions:  # (:408-422)
        steps = prediction_timesteps or list(sched.dynamical_steps.values())
        for i in [i for i in steps if i < N]:
            key = int(i) if float(i).is_integer() else i
            out[f"t{key}_preds"] = I_(x0_hat, i)
    return out


def count_calls(sched: Schedule, refine: bool, sampling_type: str = "cold",
                use_cold_sampling_for_last_step: bool = False) -> Tuple[int, int]:
    """(forecaster calls, interpolator calls) of one sample_loop -- SURVEY.md F6 / Appendix B."""
    S, N = sched.sampling_schedule, sched.num_timesteps
    f = len(S)
    i = 0
    for s, s_next in zip(S, S[1:] + [S[-1] + 1]):
        if s_next <= N - 1:
            i += 1
        if sampling_type == "cold" and not (s == N - 1 and not use_cold_sampling_for_last_step) and s > 0:
            i += 1
    if refine:
        i += len([v for v in sched.dynamical_steps.values() if v < N])
    return f, i


# ----------------------------------------------------------------------------------------------------------------
# the training / validation objective -- src/diffusion/dyffusion.py:496-567 (p_losses); SURVEY.md 8f-1
# ----------------------------------------------------------------------------------------------------------------
def p_losses(forecaster: NetFn, interpolator: NetFn, sched: Schedule, xt_last: Tensor, condition: Tensor, t: Tensor,
             static_condition: Optional[Tensor] = None, *, forward_conditioning: str = "data",
             time_encoding: str = "dynamics", lambda_reconstruction: float = 0.5, lambda_reconstruction2: float = 0.5,
             criterion: Callable[[Tensor, Tensor], Tensor] = torch.nn.functional.l1_loss,
             noise_fn: Callable[[Tensor], Tensor] = torch.randn_like) -> Dict[str, Tensor]:
    """`t` holds one integer diffusion step per row.  Returns {"loss", "loss_forward", "loss_forward2"}.
    The interpolator / forecaster callables decide about dropout (the reference forces interpolator dropout on while
    training or when `enable_interpolator_dropout`, :154-160)."""
    N = sched.num_timesteps
    tau = lambda steps: torch.tensor([float(sched.tau(int(s))) for s in steps], dtype=torch.float32)

    def F_(x_t, cond_rows, static_rows, steps):  # predict_x_last (:205-239)
        if forward_conditioning == "data":
            c = cond_rows
        elif forward_conditioning == "none":
            c = None
        else:
            w = (steps / (N - 1)).view(-1, 1, 1, 1)
            c = w * cond_rows + (1 - w) * noise_fn(cond_rows)
        if static_rows is not None:
            c = static_rows if c is None else torch.cat([c, static_rows], dim=1)
        tt = {"discrete": steps.float(), "normalized": steps / N, "dynamics": tau(steps)}[time_encoding]
        return forecaster(x_t, tt, c)

    def I_(cond_rows, x_last_rows, static_rows, steps):  # q_sample (:140-163) -> _interpolate (:480-494)
        times = tau(steps)
        assert (0 < times).all() and (times < sched.tau(N - 1) + 1).all()
        return interpolator(torch.cat([cond_rows, x_last_rows], dim=1), times, static_rows)

    pick = lambda v, m: None if v is None else v[m]
    x_t = condition.clone()                                                   # :512
    nz = t > 0                                                                # :515-526
    if nz.any():
        x_t[nz] = I_(condition[nz], xt_last[nz], pick(static_condition, nz), t[nz])
    pred = F_(x_t, condition, static_condition, t)                            # :530-532
    loss1 = criterion(pred, xt_last)
    nl = t <= N - 2                                                           # :535-557
    loss2 = torch.zeros(())
    if lambda_reconstruction2 > 0 and nl.any():
        t2 = t[nl] + 1
        x_i2 = I_(condition[nl], pred[nl], pick(static_condition, nl), t2)
        pred2 = F_(x_i2, condition[nl], pick(static_condition, nl), t2)
        loss2 = criterion(pred2, xt_last[nl])
    return {"loss": lambda_reconstruction * loss1 + lambda_reconstruction2 * loss2, "loss_forward": loss1,
            "loss_forward2": loss2}
