"""Developer tool (CPU): error budget of moving the SST attention blocks to bf16 tensor-core operands (DESIGN.md section 10,
item 2).  The oracle `Unet` is run twice on the synthetic SST weights: as is (fp32), and with the MMA operands of the attention
blocks rounded to bf16 (q / k / v as stored by the qkv conv, softmax(q), softmax(k), the 32 x 32 context, the softmax
probabilities of the bottleneck attention) while accumulating in fp32 -- what a tcgen05 / mma.sync implementation would do.
Prints the rel-L2 error this rounding alone adds to one forward and to a short sampling trajectory."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import configs as C  # noqa: E402
from oracle import dyffusion_oracle as O  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402
from tests import helpers as H  # noqa: E402

r = lambda t: t.to(torch.bfloat16).float()


def linear_attention_bf16(sd, key, x, p, drop, heads=4, dh=32):
    b, c, h, w = x.shape
    n = h * w
    y = O._channel_layernorm(sd[f"{key}.fn.norm.g"], x)
    y = O._drop(drop, f"{key}.to_qkv", y, p)
    qkv = r(F.conv2d(y, sd[f"{key}.fn.fn.to_qkv.1.weight"])).reshape(b, 3, heads, dh, n)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    q = r(q.softmax(dim=-2) * dh ** -0.5)
    k = r(k.softmax(dim=-1))
    v = r(v / n)
    ctx = r(torch.einsum("bhdn,bhen->bhde", k, v))
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, heads * dh, h, w)
    out = F.conv2d(out, sd[f"{key}.fn.fn.to_out.weight"], sd[f"{key}.fn.fn.to_out.bias"])
    return out + x


def full_attention_bf16(sd, key, x, p, drop, heads=4, dh=32):
    b, c, h, w = x.shape
    n = h * w
    y = O._channel_layernorm(sd[f"{key}.fn.norm.g"], x)
    qkv = r(F.conv2d(y, sd[f"{key}.fn.fn.to_qkv.weight"])).reshape(b, 3, heads, dh, n)
    q, k, v = r(qkv[:, 0] * dh ** -0.5), qkv[:, 1], qkv[:, 2]
    sim = torch.einsum("bhdi,bhdj->bhij", q, k)
    attn = r(O._drop(drop, f"{key}.attn", sim.softmax(dim=-1), p))
    out = torch.einsum("bhij,bhdj->bhid", attn, v).permute(0, 1, 3, 2).reshape(b, heads * dh, h, w)
    out = F.conv2d(out, sd[f"{key}.fn.fn.to_out.weight"], sd[f"{key}.fn.fn.to_out.bias"])
    return out + x


def main():
    shapes = H.golden_json("state_shapes.json")
    sdF = synth_state_dict({k: tuple(v) for k, v in shapes["sst_F"].items()}, seed=3)
    sdI = synth_state_dict({k: tuple(v) for k, v in shapes["sst_I"].items()}, seed=2)
    x, cond = H.forward_inputs("sst", "I", rows=2)
    t = torch.tensor([1.0, 2.5])
    lin, full = O._linear_attention, O._full_attention
    with torch.no_grad():
        y32 = H.oracle_net("sst", "I", sdI)(x, t, cond)
        dk = C.diffusion_kwargs("sst", horizon=3, additional_interpolation_steps=2, forward_conditioning="data")
        ic, _ = H.sampler_case_inputs("budget", "sst", 2)
        run = lambda: O.sample_loop(H.oracle_net("sst", "F", sdF), H.oracle_net("sst", "I", sdI), H.oracle_schedule(dk), ic, None,
                                    num_input_channels=1, forward_conditioning="data")
        traj32 = run()
        O._linear_attention, O._full_attention = linear_attention_bf16, full_attention_bf16
        try:
            y16 = H.oracle_net("sst", "I", sdI)(x, t, cond)
            traj16 = run()
        finally:
            O._linear_attention, O._full_attention = lin, full
    print(f"one forward: rel-L2 added by bf16 attention operands = {H.rel_l2(y16, y32):.3e}")
    for k in traj32:
        print(f"trajectory {k}: {H.rel_l2(traj16[k], traj32[k]):.3e}")


if __name__ == "__main__":
    main()
